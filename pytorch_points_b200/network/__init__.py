"""Host-side mirror of the reference's `pytorch_points.network` entry points that sit on the
hot path (same names, argument order and defaults)."""
from .model_loss import (NmDistanceFunction, LabeledNmdistanceFunction, nndistance, labeled_nndistance,  # noqa: F401
                         ChamferSumsFunction, chamfer_sums, ChamferWeightedLossFunction, chamfer_weighted_loss,
                         chamfer_mean_loss, PointLaplacianLoss,
                         PointEdgeLengthLoss, PointStretchLoss, SimplePointRepulsionLoss, NormalLoss)
from .geo_operations import (FurthestPointSampling, FurthestPointSampleGather, furthest_point_sample,  # noqa: F401
                             pointUniformLaplacian, batch_normals)
from .operations import (stage_features, GatherFunction, gather_points, BallQuery, ball_query, GroupingOperation,  # noqa: F401
                         grouping_operation, QueryAndGroup, QueryAndGroupFunction, query_and_group, group_knn,
                         knn_points)
from .pointnet2_utils import ThreeNN, three_nn, ThreeInterpolate, three_interpolate, propagate_features, GroupAll  # noqa: F401
from .layers import Conv2d, SharedMLP, DenseEdgeConv, SampledDenseEdgeConv  # noqa: F401
from .pointnet2_modules import PointnetSAModule, PointnetSAModuleMSG, PointnetFPModule  # noqa: F401
