"""lpfun_b200 -- B200-native Fast Newton Transform path of phil-hofmann/lpFun.

    from lpfun_b200 import Transform
    from lpfun_b200.utils import leja_nodes

mirrors ``from lpfun import Transform`` / ``from lpfun.utils import leja_nodes``
(reference: src/lpfun/__init__.py).  The compute path is liblpfnt.so (CUDA, sm_100a);
build it with ``python -m lpfun_b200._build``.
"""
__version__ = "0.1.0"

from . import core, utils  # noqa: F401
from .transform import Transform  # noqa: F401

__all__ = ["Transform", "utils", "core", "__version__"]
