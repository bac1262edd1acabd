"""Generates a real design on the host and pickles it under build/ (git-ignored, shipped to the GPU box), so that
multi-GPU drivers do not spend GPU-box minutes in the numpy design generator.
python tools/make_design_pickle.py rotated 17"""
import os
import pickle
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aces_b200  # noqa: E402

kind, d = sys.argv[1], int(sys.argv[2])
gen = aces_b200.rotated_design if kind == "rotated" else aces_b200.unrotated_design
fd = gen(d, full_covariance=True, noise="lognormal", seed=9)
os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
path = os.path.join(ROOT, "build", f"aces_design_{kind}_{d}.pkl")
with open(path, "wb") as f:
    pickle.dump(fd, f, protocol=4)
print(path, os.path.getsize(path) >> 20, "MiB")
