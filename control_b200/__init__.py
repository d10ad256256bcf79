"""control_b200: B200-native (sm_100a) solver for the all-at-once KKT systems of
``sleveque/control``'s ``Control.Instationary`` -- operator apply, Krylov solve and the
in-built block preconditioner -- behind the reference's own ``P=`` / ``solver_parameters``
/ shell-matrix hooks.  All numerics run in libctl_b200.so (include/ctl_b200.h)."""
from . import _lib
from ._lib import CtlError
from . import partition
from .control import Control
from .system import KSPInfo, MultiBlockSystem, csr_arrays

__all__ = ["Control", "MultiBlockSystem", "KSPInfo", "CtlError", "csr_arrays", "partition", "_lib"]
