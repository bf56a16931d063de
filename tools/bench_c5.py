"""Config C5 end to end on one B200: 3-D 'ba' 8x8 grid of 64 unit spheres, n_end = 24 (N = 36 864 unknowns,
21.7 GB complex128 matrix), k = 1, plane wave e0; then u_scat on an n x n field grid (default 2048 x 2048).

Prints one JSON line with the stage times and two size-independent correctness figures:
  * the relative residual |A phi - f| / |f| of the dense solve (A re-assembled after the in-place LU),
  * the sound-soft boundary condition |u_in + u_scat| on points ON the sphere surfaces (evaluated just outside),
    which is zero up to the truncation error of the n_end = 24 expansion.
Used by tests/test_gpu_c5.py (reduced sizes) and for the numbers quoted in DESIGN.md.
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biem_helmholtz_sphere_b200 import _ops  # noqa: E402
from biem_helmholtz_sphere_b200.geometry import field_grid, grid_centers  # noqa: E402


def run(half=4, n_end=24, k=1.0, grid=2048, check_residual=True):
    dev = torch.device("cuda", torch.cuda.current_device())
    cen = torch.as_tensor(grid_centers(half, 3), device=dev)
    B = cen.shape[0]
    rad = torch.ones(B, dtype=torch.float64, device=dev)
    H = n_end * n_end
    N = B * H
    kk = torch.tensor([k], dtype=torch.float64, device=dev)
    eta = torch.ones(1, dtype=torch.float64, device=dev)
    dirv = torch.tensor([1.0, 0.0, 0.0], dtype=torch.float64, device=dev)

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    t_plan0 = time.perf_counter()
    _ops.get_plan(3, n_end)
    t_plan = time.perf_counter() - t_plan0
    A = torch.empty((1, N, N), dtype=torch.complex128, device=dev)
    bufs = _ops.SolveBuffers(N, 1)
    torch.cuda.synchronize()
    e0 = ev()
    f = _ops.rhs_expand(3, n_end, centers=cen, radii=rad, k_in=kk, direction=dirv)
    e1 = ev()
    _ops.assemble(3, n_end, cen, rad, kk, eta, out=A)
    e2 = ev()
    dens = f.reshape(N).clone()
    _ops.zgesv_(A[0], dens, bufs)
    e3 = ev()
    torch.cuda.synchronize()
    info = int(bufs.info.item())
    out = {"config": f"{B} spheres, n_end={n_end}, N={N}, k={k}", "plan_s": t_plan, "rhs_ms": e0.elapsed_time(e1),
           "assemble_ms": e1.elapsed_time(e2), "solve_ms": e2.elapsed_time(e3), "lu_info": info,
           "assemble_gbs": 16.0 * N * N / (e1.elapsed_time(e2) * 1e-3) * 1e-9,
           "lu_tflops": (8.0 / 3.0) * N ** 3 / (e2.elapsed_time(e3) * 1e-3) * 1e-12}
    if check_residual:
        _ops.assemble(3, n_end, cen, rad, kk, eta, out=A)  # LU overwrote A
        r = torch.mv(A[0], dens) - f.reshape(N)
        out["solve_rel_residual"] = float(torch.linalg.vector_norm(r) / torch.linalg.vector_norm(f))
    del A
    # boundary condition on the spheres: random directions on each of 8 spheres, radius rho (1 + 1e-9)
    rng = np.random.default_rng(1)
    v = rng.standard_normal((3, 64))
    v /= np.linalg.norm(v, axis=0, keepdims=True)
    pts = []
    for b in rng.choice(B, size=min(8, B), replace=False):
        pts.append(cen[b].cpu().numpy()[:, None] + (1.0 + 1e-9) * v)
    xb = torch.as_tensor(np.concatenate(pts, axis=1), device=dev)
    us = _ops.uscat(3, n_end, cen, rad, k, 1.0, dens.reshape(B, H), xb)
    uin = torch.exp(1j * k * xb[0])
    out["bc_residual_max"] = float(torch.max(torch.abs(us + uin)))
    # field grid
    if grid > 0:
        x = torch.as_tensor(field_grid(grid, 20.0, 3).reshape(3, -1).copy(), device=dev)
        _ops.uscat(3, n_end, cen, rad, k, 1.0, dens.reshape(B, H), x[:, : 128 * 1024].contiguous())
        torch.cuda.synchronize()
        g0 = ev()
        u = _ops.uscat(3, n_end, cen, rad, k, 1.0, dens.reshape(B, H), x)
        g1 = ev()
        torch.cuda.synchronize()
        ms = g0.elapsed_time(g1)
        out.update(uscat_ms=ms, uscat_points_per_s=grid * grid / (ms * 1e-3), nan_fraction=float(torch.isnan(u.real).double().mean()),
                   uscat_counted_tflops=8.0 * grid * grid * B * H / (ms * 1e-3) * 1e-12)
        fin = u[~torch.isnan(u.real)]
        out["uscat_abs_max"] = float(torch.max(torch.abs(fin)))
    return out, dens.reshape(B, H)


if __name__ == "__main__":
    half = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    n_end = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    grid = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
    o, _ = run(half, n_end, 1.0, grid)
    print(json.dumps(o))
