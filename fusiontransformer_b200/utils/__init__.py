from .kernel_region import KernelRegion  # noqa: F401
from .quantize import sparse_quantize, sparse_quantize_batch  # noqa: F401
from . import helpers, kernel_region  # noqa: F401
