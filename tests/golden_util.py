"""Loader for the reference's golden CSVs (copied verbatim from /root/reference/accuracy and
/root/reference/jascome into tests/golden/ -- they are data fixtures, not source)."""
import csv
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    rows = []
    with open(os.path.join(GOLDEN, name)) as f:
        for r in csv.DictReader(f):
            r = dict(r)
            r["n_end"] = int(r["n_end"])
            r["uscat"] = complex(r["uscat"])
            if "k" in r:
                r["k"] = float(r["k"])
            if "n_balls" in r:
                r["n_balls"] = int(r["n_balls"])
            rows.append(r)
    return rows


def find(rows, **kw):
    out = []
    for r in rows:
        ok = True
        for key, v in kw.items():
            if isinstance(v, float):
                ok &= abs(r[key] - v) < 1e-9 * max(1.0, abs(v))
            else:
                ok &= r[key] == v
        if ok:
            out.append(r)
    return out
