"""CPU oracle for the FusionTransformer 3D-branch hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors for
this path (FusionTransformer/tests/test_dataset.py is inert), and its arithmetic
lives in torchsparse v1.1.0 (pinned only by docker/Dockerfile:33), which is not
vendored and not installable here.  This package restates the published
torchsparse v1.1.0 operator semantics plus the
reference's own glue (FusionTransformer/models/utils.py, models/spvcnn.py,
models/middle_fusion.py, models/image_models_billinear.py:88-126,
data/semantic_kitti/semantic_kitti_dataloader.py:216-238, data/collate.py).
It is validated against dense torch.nn.functional.conv3d / conv_transpose3d,
hand-computed FNV known answers and algebraic properties (tests/test_oracle_*.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference leg may import this package.  The product package
(fusiontransformer_b200) never does.
"""
from . import ts_ops, ft_glue  # noqa: F401
