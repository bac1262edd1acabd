"""aces_b200 -- B200-native ACES estimation hot path (K1 parity counts, LS fit, covariance).

The directory is named `quantumaces.jl_b200` (not importable by that name); import it as
`aces_b200` through the loader module at the repository root.
"""
from . import _lib  # noqa: F401
from ._lib import AcesError, Context, NotPositiveDefinite, PauliTables  # noqa: F401
from .design import (  # noqa: F401
    FlatDesign, build_pauli_tables, shots_divide, synthetic_design, synthetic_estimates,
)
from .estimate import DesignEstimator, estimate_experiment  # noqa: F401
from . import merit  # noqa: F401
from . import shard  # noqa: F401
from . import project  # noqa: F401
from . import optimise  # noqa: F401
from .optimise import (  # noqa: F401
    OptimOptions, calc_gls_merit_grad_log, calc_ols_merit_grad_log, calc_wls_merit_grad_log, get_transform, optimise_weights,
)
from .merit import (  # noqa: F401
    calc_gls_covariance_from_factor, calc_precision_matrix_from_factor, estimate_gate_eigenvalues,
    get_clipped_indices, get_clipped_mapping_lengths, DeviceDesign, estimate_gate_noise_core, mean_ensemble,
    nrmse_moments, estimate_gate_noise_projected, calc_merit, eigvals, nrmse_moments_eigenvalues,
)
