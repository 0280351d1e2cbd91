import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import torch
from alpha_eigen_b200.ensemble import Ensemble, synthetic_ensemble, solve_member
members = synthetic_ensemble(1024, 2048, 16, 12)[::64][:16]
Ensemble(members[:2], dt=0.1, num_steps=20, concurrent=2).run()
# breakdown for one member, serial
from alpha_eigen_b200.device import SlabDevice
from alpha_eigen_b200.time_unit import dmd_window
p = members[0]
def sync(): torch.cuda.synchronize()
for rep in range(2):
    sync(); t0=time.perf_counter()
    dev = SlabDevice(p["Z"], p["hx"], p["mu"], p["weight"], p["mat"], p["total"], p["scat"], p["chi"], p["nusf"], p["invspeed"], dt=0.1)
    psi0 = np.repeat(0.5 * p["phi0"][:, None, :], p["A"], axis=1)
    psi = dev.psi_to_device(psi0)
    n_psi = dev.G*dev.A*dev.Zp
    snap = torch.zeros(200*n_psi, dtype=torch.float64, device=dev.dev)
    sync(); t1=time.perf_counter()
    st, outer, inner = dev.time_march(psi, 200, 1e-10, snap, None)
    sync(); t2=time.perf_counter()
    first,K = dmd_window(200)
    R = dev.tsqr(snap, n_psi, first, K+1)
    sync(); t3=time.perf_counter()
    al = dev.dmd_from_r(R, K, 0.1)
    sync(); t4=time.perf_counter()
    print(f"setup {t1-t0:.3f} march {t2-t1:.3f} (sweeps {int(inner.sum())+200}) tsqr {t3-t2:.3f} dense {t4-t3:.3f}")
for c in (1, 4, 12, 24):
    sync(); t0=time.perf_counter()
    res = Ensemble(members, dt=0.1, num_steps=200, concurrent=c).run()
    sync(); t1=time.perf_counter()
    print("concurrent", c, "wall", round(t1-t0,3), "members/s", round(len(members)/(t1-t0),2))
