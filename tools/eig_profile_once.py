import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
import aces_b200
n = int(sys.argv[1])
ctx = aces_b200.Context(0)
rng = np.random.default_rng(n)
a = rng.standard_normal((n, n)); a = np.asfortranarray((a + a.T) / 2)
aces_b200.eigvals(ctx, a)
ctx.close()
