"""pytorch_points_b200 -- B200-native (sm_100a) drop-in for the data-parallel hot path of
yifita/pytorch_points: Chamfer/nndistance, farthest point sampling + gather, ball_query,
group_knn.  Python/torch.autograd.Function signatures follow the reference
(`pytorch_points.network.*`); the kernels live in a C-ABI CUDA library (include/pp_b200.h).
"""
from . import _C  # noqa: F401  (loads libpp_b200.so; raises if it is missing)

__version__ = "0.1.0"
