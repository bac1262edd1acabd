import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def parity_cases():
    z = np.load(os.path.join(GOLDEN, "parity_golden.npz"))
    for ci in range(int(z["n_cases"])):
        ptr = z[f"c{ci}_ptr"]
        idx = z[f"c{ci}_idx"]
        supports = [idx[ptr[i]:ptr[i + 1]].tolist() for i in range(len(ptr) - 1)]
        yield dict(
            n=int(z[f"c{ci}_n"]), counts=z[f"c{ci}_counts"], supports=supports,
            signs=z[f"c{ci}_signs"], prefixes=z[f"c{ci}_prefixes"], est=z[f"c{ci}_est"],
            pauli_shots=z[f"c{ci}_pauli_shots"],
        )
