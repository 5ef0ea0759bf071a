"""``isplib`` -- alias of :mod:`isplib_b200` so that the reference's documented usage

    from isplib import *
    iSpLibPlugin.patch_pyg()

(/root/reference/README.md:66-69) works unchanged on top of the B200 kernels.
"""
from isplib_b200 import *  # noqa: F401,F403
from isplib_b200 import (iSpLibPlugin, isplib_autotune, SparseTensor, matmul, torch, torch_sparse,  # noqa: F401
                         __version__, __all__)
