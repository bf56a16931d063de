import time, numpy as np, torch, sys
sys.path.insert(0, "/root/repo")
import biem_helmholtz_sphere_b200 as bhs
c = bhs.create_from_branching_types("ba")
def run(n_end, B):
    cen = np.zeros((B, 3)); cen[:, 1] = 4.0 * np.arange(B) - 2.0 * (B - 1)
    k = np.asarray(1.0)
    uin = bhs.plane_wave(k=k, direction=np.array([1.0, 0, 0]))[0]
    calc = bhs.biem(c, uin=uin, k=k, n_end=n_end, eta=np.asarray(1.0), centers=cen, radii=np.ones(B), kind="outer")
    return calc.uscat(np.zeros(3))
for n_end, B in ((6, 2), (10, 4), (16, 4), (16, 16)):
    run(n_end, B); run(n_end, B)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): run(n_end, B)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print(f"n_end={n_end} B={B} N={B*n_end*n_end}: {dt*1e3:.1f} ms per biem()+uscat() call (NumPy in/out, matrix returned)")
