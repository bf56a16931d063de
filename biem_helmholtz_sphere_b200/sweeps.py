"""Sweep drivers that write the reference's CSV formats with the B200 path.

Restates the two batch commands of the reference CLI that produced its golden data -- ``jascome`` (cli.py:36-116)
and ``accuracy`` (cli.py:188-271) -- minus the typer/tqdm/matplotlib front-end: same loops, same geometries
(``_center``, cli.py:170-185), same call into ``biem`` (plane wave built with k = 1 whatever the solve wavenumber,
cli.py:239 vs :244), same NaN checks, same CSV columns, so that the output can be diffed row by row against
``accuracy/*.csv`` / ``jascome/jascome_output.csv``.  The default tree list is the reference's (cli.py:41), including the
0 <-> d-1 leaf relabelling of the ``p`` trees (cli.py:63-69).

    python -m biem_helmholtz_sphere_b200.sweeps jascome  [--out jascome_output.csv] [--branching-types a,ba,bpa,bba,bpbpa,caa]
    python -m biem_helmholtz_sphere_b200.sweeps accuracy [--out accuracy.csv] [--branching-types a] [--max-n-end 512]
"""

from __future__ import annotations

import argparse
import logging

import numpy as np

from . import _biem
from ._coords import create_from_branching_types
from .geometry import grid_centers

LOG = logging.getLogger(__name__)


def _row_value(btype: str, n_end: int, k: float, half: int):
    c = create_from_branching_types(btype)
    d = c.c_ndim
    if "p" in btype:  # swap 0 and -1, as the reference CLI does with nx.relabel_nodes (cli.py:63-69)
        c = c.relabel({0: d - 1, d - 1: 0})
    centers = grid_centers(half, d)
    n_balls = len(centers)
    calc = _biem.biem(
        c,
        uin=_biem.plane_wave(k=np.asarray(1.0), direction=np.asarray((1.0,) + (0.0,) * (d - 1)))[0],
        k=np.asarray(float(k)), n_end=int(n_end), eta=np.asarray(1.0), centers=centers, radii=np.ones(n_balls),
        kind="outer", keep_matrix=False,
    )
    if np.any(np.isnan(calc.density)):
        raise ValueError("Density contains NaN")
    uscat = calc.uscat(np.zeros(d))
    if np.isnan(uscat):
        raise ValueError("uscat is NaN")
    return n_balls, calc, complex(uscat)


JASCOME_TREES = "a,ba,bpa,bba,bpbpa,caa"  # cli.py:41


def jascome(out: str = "jascome_output.csv", branching_types: str = JASCOME_TREES, max_n_end_4d: int | None = None) -> str:
    """Two unit spheres at (0, +-2, 0, ...), k = 1, n_end = 1..9 per tree (cli.py:36-116).  ``max_n_end_4d`` caps the 4-D
    trees (the reference's own 4-D runs ended at n_end = 5 / 6 when the `triplet` tensor no longer fitted its host)."""
    with open(out, "w") as f:
        f.write("branching_types,n_end,uscat,device,dtype,density_dtype,density_device,uscat_dtype,uscat_device\n")
    for btype in reversed(branching_types.split(",")):
        try:
            for n_end in range(1, 10):
                if max_n_end_4d is not None and create_from_branching_types(btype).c_ndim >= 4 and n_end > max_n_end_4d:
                    break
                _, calc, uscat = _row_value(btype, n_end, 1.0, 0)
                with open(out, "a") as f:
                    f.write(f"{btype},{n_end},{uscat},cuda,<class 'numpy.float64'>,{calc.density.dtype},cuda,"
                            f"complex128,cuda\n")
        except Exception as e:  # same policy as the reference: log and go on with the next tree (cli.py:113-115)
            LOG.error(e)
            continue
    return out


def accuracy(out: str = "accuracy.csv", branching_types: str = "a", max_n_end: int | None = None,
             k_sweep: bool = True, grids: bool = True) -> str:
    """The accuracy sweeps (cli.py:188-271): for the two-sphere case k = 2^{0, .5, ..., 14.5}; for the 2x2 ... 64x64
    grids k = 1; n_end over unique(int(2^{0, .25, ..., 14.75})); an exception ends the current n_end loop."""
    with open(out, "w") as f:
        f.write("branching_types,n_end,k,n_balls,uscat,device,dtype,density_dtype,density_device,uscat_dtype,uscat_device\n")
    n_ends = np.unique((2 ** np.arange(0, 15, 0.25)).astype(int))
    if max_n_end is not None:
        n_ends = n_ends[n_ends <= max_n_end]
    for btype in reversed(branching_types.split(",")):
        for log2div2 in range(0, 7):
            if (log2div2 == 0 and not k_sweep) or (log2div2 > 0 and not grids):
                continue
            ks = (2 ** np.arange(0, 15, 0.5)) if log2div2 == 0 else (1,)
            half = 0 if log2div2 == 0 else 2 ** (log2div2 - 1)
            for k in ks:
                try:
                    for n_end in n_ends:
                        n_balls, calc, uscat = _row_value(btype, int(n_end), float(k), half)
                        with open(out, "a") as f:
                            f.write(f"{btype},{n_end},{k},{n_balls},{uscat},cuda,<class 'numpy.float64'>,"
                                    f"{calc.density.dtype},cuda,complex128,cuda\n")
                except Exception as e:
                    LOG.error(e)
    return out


def main(argv=None) -> None:
    ap = argparse.ArgumentParser(prog="biem_helmholtz_sphere_b200.sweeps")
    sub = ap.add_subparsers(dest="cmd", required=True)
    j = sub.add_parser("jascome")
    j.add_argument("--out", default="jascome_output.csv")
    j.add_argument("--branching-types", default=JASCOME_TREES)
    a = sub.add_parser("accuracy")
    a.add_argument("--out", default="accuracy.csv")
    a.add_argument("--branching-types", default="a")
    a.add_argument("--max-n-end", type=int, default=None)
    a.add_argument("--no-k-sweep", action="store_true")
    a.add_argument("--no-grids", action="store_true")
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO)
    if args.cmd == "jascome":
        print(jascome(args.out, args.branching_types))
    else:
        print(accuracy(args.out, args.branching_types, args.max_n_end, not args.no_k_sweep, not args.no_grids))


if __name__ == "__main__":
    main()
