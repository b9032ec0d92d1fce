"""sqfa_b200 -- B200-native (sm_100a) implementation of the SQFA hot paths.

Drop-in for the `sqfa` package namespace (`statistics`, `linalg`, `distances`, `constraints`,
`_optim`, `model`; the matplotlib `plot` helpers of the reference are out of scope). All compute
runs in hand-written CUDA kernels behind the C ABI of include/sqfa_b200.h; there is no CPU
fallback -- importing works anywhere, running a kernel needs the built library and a CUDA device.
"""

from . import _optim as _optim
from . import constraints as constraints
from . import distances as distances
from . import linalg as linalg
from . import model as model
from . import statistics as statistics

__version__ = "0.1.0"
