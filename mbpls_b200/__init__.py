"""mbpls_b200 -- B200-native (sm_100a) implementation of the MB-PLS latent-variable fitting path.

``from mbpls_b200 import MBPLS`` is a drop-in for ``from mbpls.mbpls import MBPLS``.
"""
from .mbpls import MBPLS  # noqa: F401

__all__ = ["MBPLS"]
__version__ = "0.1.0"
