"""Deterministic-mode PCS assignment of the C2 catalogue twice (for ncu captures of
k_assign_gather_warp)."""
import sys
sys.path.insert(0, '.')
import numpy as np
from triumvirate_b200 import core
pos = np.random.default_rng(42).uniform(0., 1000., size=(3, 10**7))
for it in range(2):
    m = core.mesh(pos, 1000., 512, "pcs", stage=0, deterministic=True)
print(m.real.sum())
