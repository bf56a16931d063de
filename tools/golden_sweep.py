"""Replay the reference's golden CSVs (tests/golden/*.csv = /root/reference/accuracy/*.csv and jascome/jascome_output.csv)
through the B200 public API, exactly as the reference CLI produced them (cli.py:56-116, 189-271: plane wave built
with k = 1 -- quirk A.7-1 --, eta = 1, unit radii, `_center` geometries, uscat at the origin), and report the
relative deviation of every row.

    python tools/golden_sweep.py [--stride S] [--max-n-end-2d M] [--out summary.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# forced-`triplet` rows of jascome_output.csv carry the reference's own quadrature noise
TRIPLET_TOL = {6: 5e-10, 7: 1e-9, 8: 2e-9, 9: 1e-7}


def run_row(bhs, btype, n_end, k, half):
    from biem_helmholtz_sphere_b200.geometry import grid_centers

    c = bhs.create_from_branching_types(btype)
    d = c.c_ndim
    if "p" in btype:  # the reference CLI swaps the cartesian leaves 0 and d - 1 of the `p` trees (cli.py:63-69)
        c = c.relabel({0: d - 1, d - 1: 0})
    uin = bhs.plane_wave(k=np.asarray(1.0), direction=np.asarray((1.0,) + (0.0,) * (d - 1)))[0]
    cen = grid_centers(half, d)
    calc = bhs.biem(c, uin=uin, k=np.asarray(float(k)), n_end=n_end, eta=np.asarray(1.0), centers=cen,
                    radii=np.ones(len(cen)), kind="outer", keep_matrix=False)
    return complex(calc.uscat(np.zeros(d)))


def sweep(stride=1, max_n_end_2d=512, max_n_end_3d=39, verbose=False):
    import biem_helmholtz_sphere_b200 as bhs
    from golden_util import load

    files = {
        "accuracy_k_ba.csv[ba]": [r for r in load("accuracy_k_ba.csv") if r["branching_types"] == "ba"],
        "accuracy_k_a.csv": load("accuracy_k_a.csv"),
        "accuracy_n_balls_a.csv": load("accuracy_n_balls_a.csv"),
        "jascome_output.csv": load("jascome_output.csv"),  # all six trees of cli.py:41: a, ba, bpa, bba, bpbpa, caa
    }
    out = {}
    for name, rows in files.items():
        t0 = time.perf_counter()
        worst, worst_row, n_run, n_skip, n_bad = 0.0, None, 0, 0, 0
        for i, r in enumerate(rows):
            if i % stride:
                continue
            bt = r.get("branching_types", "a")
            d = len(bt.replace("bp", "b")) + 1
            lim = max_n_end_2d if d == 2 else max_n_end_3d
            if r["n_end"] > lim:
                n_skip += 1
                continue
            nb = r.get("n_balls", 2)
            half = 0 if nb == 2 else int(round(nb ** 0.5)) // 2
            try:
                v = run_row(bhs, bt, r["n_end"], r.get("k", 1.0), half)
            except (NotImplementedError, MemoryError):
                n_skip += 1
                continue
            err = abs(v - r["uscat"]) / max(abs(r["uscat"]), 1e-300)
            tol = TRIPLET_TOL.get(r["n_end"], 1e-10) if name.startswith("jascome") else 1e-10
            n_run += 1
            if not (err <= tol):
                n_bad += 1
                if verbose:
                    print("MISS", name, bt, r["n_end"], r.get("k"), nb, err, flush=True)
            if err / tol > worst:
                worst, worst_row = err / tol, (bt, r["n_end"], r.get("k", 1.0), nb, err)
        out[name] = {"rows": len(rows), "run": n_run, "skipped": n_skip, "outside_tolerance": n_bad,
                     "worst_err_over_tol": worst, "worst_row": worst_row, "seconds": time.perf_counter() - t0}
        if verbose:
            print(name, out[name], flush=True)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--stride", type=int, default=1)
    ap.add_argument("--max-n-end-2d", type=int, default=512)
    ap.add_argument("--max-n-end-3d", type=int, default=39)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    res = sweep(a.stride, a.max_n_end_2d, a.max_n_end_3d, verbose=True)
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)
    print(json.dumps(res))
