"""CPU oracle for the biem_helmholtz_sphere hot path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy/SciPy float64 restatement of the reference algorithm
(`/root/reference/src/biem_helmholtz_sphere/_biem.py:453-977` plus what the
un-vendored `ultrasphere`, `ultrasphere-harmonics` and `batch-tensorsolve`
packages compute for it).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it.  The product package
(`biem_helmholtz_sphere_b200`) never imports anything from here.

Parity status: PINNED against the reference's own golden vectors
(`accuracy/*.csv`, `jascome/jascome_output.csv`, README known answer), copied as
fixtures under `tests/golden/` -- see `tests/test_oracle_golden.py`.
"""

from .biem_oracle import (  # noqa: F401
    OracleCoordinates,
    OracleResult,
    biem,
    biem_u,
    create_from_branching_types,
    expand,
    harm_count,
    harmonics,
    index_tables,
    plane_wave,
    point_source,
    quadrature,
    radial,
    translation_coef,
)
