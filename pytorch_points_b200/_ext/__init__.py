"""Stand-ins for the reference's pybind11 modules `pytorch_points._ext.losses` and
`pytorch_points._ext.sampling`: same function names, argument order and ownership rules,
implemented as thin ctypes calls into libpp_b200.so."""
from . import losses, sampling  # noqa: F401
