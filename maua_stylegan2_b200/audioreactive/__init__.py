"""Device-side drop-in for the reference's `audioreactive` package (signal + latent halves)."""
from .bend import *  # noqa: F401,F403
from .latent import *  # noqa: F401,F403
from .signal import *  # noqa: F401,F403
from .segmentation import laplacian_segmentation  # noqa: F401
from . import signal as _signal

del SMF  # noqa: F821  — served live by __getattr__ so that `ar.SMF` follows set_SMF()


def __getattr__(name):
    if name == "SMF":
        return _signal.SMF
    raise AttributeError(name)
