import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
from tests.test_gpu_parity import _kp_group, _sorted_c
from alpha_eigen_modules.time_unit import TimeUnit
g0 = np.load('/root/repo/tests/golden/dmd_march_answers_50_16_200.npz')
s, g = _kp_group(50, 16)
tu = TimeUnit(slab=s, group=g, num_steps=200, dt=0.1)
tu.calculate_time_dependent_flux()
al, info = tu.compute_alpha_eigenvalues(return_info=True)
ref = g0["alphas"]
# match each reference mode to the nearest device mode
d = [(abs(al[np.argmin(np.abs(al - r))] - r) / abs(r), r) for r in ref]
for e, r in sorted(d, key=lambda x: x[0]): print(f"{e:.3e}  {r:.6f}")
print("sigma/sigma0:", (info["sigma"][:24] / info["sigma"][0]))
