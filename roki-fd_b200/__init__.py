"""rokifd_b200 - B200-native batched forward-dynamics engine behind RoKi-FD's rkfd_sim step API.

Host-side Python mirror of the reference interface (rkFDCreate / rkFDChainReg / rkFDUpdateInit /
rkFDUpdate / rkFDUpdateDestroy / rkFDDestroy, reference include/roki_fd/rkfd_sim.h:58-102) over the
C-ABI shared library `librokifd_b200.so` (CUDA sm_100a).  There is no CPU fallback: importing
`rokifd_b200.capi` fails loudly when the library is missing.
"""
from . import chains  # noqa: F401

__all__ = ["chains", "capi"]


def __getattr__(name):
    if name == "capi":
        import importlib
        return importlib.import_module(".capi", __name__)
    raise AttributeError(name)
