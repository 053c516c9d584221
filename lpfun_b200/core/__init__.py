"""Index-set layer (reference: src/lpfun/core/). Host-side, backed by liblpfnt.so's C++ builders."""
from . import set, utils  # noqa: F401
