"""biem_helmholtz_sphere_b200 -- B200-native (sm_100a) implementation of the biem_helmholtz_sphere hot path.

Re-exports the reference's public names (src/biem_helmholtz_sphere/__init__.py:2-24) plus the
``create_from_branching_types`` stand-in for the (un-installable) ``ultrasphere`` coordinate factory.
"""

__version__ = "1.2.0+b200.1"

import os as _os

# k-sweeps keep up to 32 independent systems in flight on their own streams; give the driver enough hardware
# queues (default 8) so those streams do not alias.  Only effective before the CUDA context exists.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from ._biem import (
    BIEMKwargs,
    BIEMResultCalculator,
    BIEMResultCalculatorProtocol,
    UinCallable,
    biem,
    biem_u,
    max_memory,
    max_n_end,
    plane_wave,
    point_source,
)
from ._coords import SphericalCoordinates, create_from_branching_types
from .fields import far_field_pattern, heatmap_field

__all__ = [
    "BIEMKwargs",
    "BIEMResultCalculator",
    "BIEMResultCalculatorProtocol",
    "UinCallable",
    "biem",
    "biem_u",
    "max_memory",
    "max_n_end",
    "plane_wave",
    "point_source",
    "SphericalCoordinates",
    "create_from_branching_types",
    "heatmap_field",
    "far_field_pattern",
]
