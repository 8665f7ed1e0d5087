"""KernelRegion -- torchsparse.utils.kernel_region v1.1.0.

Imported with ``*`` by FusionTransformer/models/utils.py:6 and spvcnn.py:12, which also rely on the
wildcard to bring ``torch`` into scope (models/utils.py never imports torch itself), so this module
deliberately has no ``__all__``.
"""
import numpy as np
import torch


class KernelRegion:
    def __init__(self, kernel_size: int = 3, tensor_stride: int = 1, dilation: int = 1, dim=(0, 1, 2)):
        self.kernel_size = kernel_size
        self.tensor_stride = tensor_stride
        self.dilation = dilation
        ks = kernel_size
        axis = [v * tensor_stride * dilation for v in range(-ks // 2 + 1, ks // 2 + 1)]
        if ks % 2 == 1:      # odd: z outermost, x innermost  -> k = (dz+1)*9 + (dy+1)*3 + (dx+1)
            offs = [[x, y, z] for z in axis for y in axis for x in axis]
        else:                # even: x outermost, z innermost -> k = dx*4 + dy*2 + dz
            offs = [[x, y, z] for x in axis for y in axis for z in axis]
        self.kernel_offset = np.array(offs, dtype=np.int32).reshape(-1, 3)

    def get_kernel_offset(self):
        return torch.from_numpy(self.kernel_offset.copy())
