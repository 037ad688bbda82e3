"""``import MinkowskiEngine as ME`` shim: the MinkowskiEngine names the GCDLSS reference uses
(SURVEY 8(b)), implemented by gcdlss_b200 on hand-written sm_100a kernels.  With this package on
``sys.path`` the reference's ``models/minkunet.py``, ``models/resnet.py`` and
``models/multiheadminkunet.py`` import and run unmodified."""
from gcdlss_b200.nn import (MinkowskiBatchNorm, MinkowskiConvolution, MinkowskiConvolutionTranspose, MinkowskiDropout, MinkowskiGELU,
                            MinkowskiGlobalMaxPooling, MinkowskiInstanceNorm, MinkowskiLinear, MinkowskiMaxPooling, MinkowskiReLU, cat)
from gcdlss_b200.sparse_tensor import CoordinateMapKey, SparseTensor
from gcdlss_b200.coords import CoordinateManager

from . import modules, utils

__version__ = "0.5.4+gcdlss_b200"
