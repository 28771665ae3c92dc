"""kNN grouping helpers of the small-range model (mirror of src/modules/point_utils.py: index_points_group :5,
grouping :63, square_distance :109, query_ball_point :131, knn_point :164).

Same names, arguments and return tuples.  Underneath: the warp-per-query kNN kernel (csrc/query.cu, never
builds the (B,S,N) distance matrix the reference feeds to torch.topk) and the channels-last row gather with its
vector red.add backward (csrc/gather.cu), instead of transpose -> grouping_operation -> permute.
Neighbour order inside a group is (distance, index); the reference's is unspecified (topk sorted=False).
"""
import torch

from ..pointnet2 import pointnet2_utils
from ..projectPN.utils import gather_rows, knn_point, square_distance  # noqa: F401  (re-exported names)


def index_points_group(points, knn_idx):
    """points (B,N,C), knn_idx (B,S,K) -> (B,S,K,C)"""
    return gather_rows(points.contiguous(), knn_idx.to(torch.int32))


def grouping(feature, K, src_xyz, q_xyz, use_xyz=False, raw_feat_point=False, raw_xyz1=None, raw_xyz2=None):
    """The K nearest src_xyz (B,N,3) of every q_xyz (B,S,3) and their features (B,N,c).
    -> grouped_xyz (B,S,K,3), xyz_diff (B,S,K,3), new_points (B,S,K,c[+3]), point_indices (B,S,K) int64,
       grouped_raw_xyz (B,S,K,3) or None.  With raw_feat_point the offsets are taken between the raw (un-warped)
       coordinates raw_xyz1 (of the sources) and raw_xyz2 (of the queries), as point_utils.py:85-98."""
    q_xyz, src_xyz = q_xyz.contiguous(), src_xyz.contiguous()
    point_indices = knn_point(K, src_xyz, q_xyz)
    idx32 = point_indices.to(torch.int32)
    grouped_xyz = gather_rows(src_xyz, idx32)
    grouped_raw_xyz = None
    if raw_feat_point:
        grouped_raw_xyz = gather_rows(raw_xyz1.contiguous(), idx32)
        xyz_diff = grouped_raw_xyz - raw_xyz2.unsqueeze(2)
    else:
        xyz_diff = grouped_xyz - q_xyz.unsqueeze(2)
    grouped_feature = gather_rows(feature.contiguous(), idx32)
    new_points = torch.cat([xyz_diff, grouped_feature], dim=-1) if use_xyz else grouped_feature
    return grouped_xyz, xyz_diff, new_points, point_indices, grouped_raw_xyz


def query_ball_point(radius, nsample, xyz, new_xyz):
    """First nsample points of xyz (B,N,3), in index order, within `radius` of each new_xyz (B,S,3); groups with
    fewer are padded with their first member -> (B,S,nsample) int64 (point_utils.py:131-161).  The ball-query
    kernel has exactly these semantics up to the boundary (strict `<` there, `<=` here in the reference's
    naive form); a query with no point in range yields index 0 where the naive form yields N."""
    idx = pointnet2_utils.ball_query(radius, nsample, xyz.contiguous(), new_xyz.contiguous())
    return idx.long()
