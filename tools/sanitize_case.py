"""Small end-to-end pass over every kernel family for `compute-sanitizer --tool memcheck` (sizes kept tiny)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import biem_helmholtz_sphere_b200 as bhs

rng = np.random.default_rng(0)
for bt, n_end, k in (("a", 9, 1.3), ("ba", 6, 1.0), ("ba", 5, 0.8 + 0.3j), ("bba", 4, 1.1), ("bbba", 3, 0.9), ("a", 5, 1.0 + 0.2j)):
    c = bhs.create_from_branching_types(bt)
    d = c.c_ndim
    cen = np.zeros((3, d)); cen[0, 1], cen[1, 1], cen[2, 0] = 2.0, -2.2, 3.5
    rad = np.array([1.0, 0.8, 1.1])
    dirn = np.zeros(d); dirn[0] = 1.0
    uin, ug = bhs.plane_wave(k=np.asarray(k), direction=dirn)
    calc = bhs.biem(c, uin=uin, uin_grad=ug, k=np.asarray(k), n_end=n_end, eta=np.asarray(0.9), centers=cen, radii=rad, alpha=1.0, beta=0.2)
    x = rng.uniform(-6, 6, size=(d, 70))
    u = calc.uscat(x); up = calc.uscat(x, per_ball=True)
    xh = x / np.linalg.norm(x, axis=0, keepdims=True)
    uf = calc.uscat(xh, far_field=True)
    xpl = x.copy(); xpl[2:] = 0.0
    upl = calc.uscat(xpl)
    print(bt, n_end, k, complex(u[0]), int(np.isnan(upl).sum()))
# batched sweep (grouped LU, graphs off and on: second call captures)
c = bhs.create_from_branching_types("ba")
cen = np.array([[0.0, 2.0, 0.0], [0.0, -2.0, 0.0]]); ks = np.linspace(0.7, 2.0, 5)
for _ in range(3):
    uin = bhs.plane_wave(k=ks, direction=np.array([[1.0], [0.0], [0.0]]))[0]
    calc = bhs.biem(c, uin=uin, k=ks, n_end=6, eta=np.ones(5), centers=cen[None], radii=np.ones((1, 2)), keep_matrix=False)
    u = calc.uscat(np.zeros((3, 2)))
print("sweep", u.shape, complex(u[0, 0]))
