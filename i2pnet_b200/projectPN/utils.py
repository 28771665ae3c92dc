"""Projection-aware index / gather operators (mirror of the reference's src/projectPN/utils.py).

Two layers live here:

* the reference's public helpers with their exact names, argument order and return values
  (`get_neighbor_copy` :63, `get_neighbor_att` :253, `gather_torch` :36, `project_seq` :111,
  `project` :189, `grouping` :313, `knn_point` :369, `index_points_group` :382, `check_valid`
  :106, `get_idx_cuda` :8, `get_sample_idx` :18, `get_stride_idx_cuda` :28), each a single
  kernel launch through the C ABI instead of the reference's allocate-six-buffers-then-launch
  / matmul+topk / Python-loop-of-index_put formulations;
* the compact forms the modules in PPBackbone_center.py use internally (`select_flat`,
  `gather_rows`, `StrideGrid`): int32 flat indices instead of three int64 tensors, no dead
  `valid_idx` buffers, centres of a regular grid generated inside the kernel.
"""
import numpy as np
import torch
from torch.autograd import Function

from .. import _cabi

FLAG_COPY = 1   # src/projectPN/fused_conv_select/fused_conv_select_k.py:6
FLAG_SHIFT = 2  # :7


# ---------------------------------------------------------------------------------------------
# index grids
# ---------------------------------------------------------------------------------------------
def get_idx_cuda(B, H, W, device):
    """(B, H*W, 2) int32 (h, w) of every pixel."""
    return get_stride_idx_cuda(B, H, W, 1, 1, device)


def get_stride_idx_cuda(B, out_h, out_w, stride_h, stride_w, device):
    """(B, out_h*out_w, 2) int32 (h*stride_h, w*stride_w)."""
    h = torch.arange(0, out_h * stride_h, stride_h, device=device, dtype=torch.int32)
    w = torch.arange(0, out_w * stride_w, stride_w, device=device, dtype=torch.int32)
    grid = torch.stack(torch.meshgrid(h, w, indexing="ij"), dim=-1).reshape(1, -1, 2)
    return grid.expand(B, -1, -1).contiguous()


def get_sample_idx(batch, out_h, out_w, stride_H, stride_W, device):
    """Three (batch, out_h, out_w) int64 tensors (b, h*stride_H, w*stride_W)."""
    h = torch.arange(0, out_h * stride_H, stride_H, device=device, dtype=torch.int64)
    w = torch.arange(0, out_w * stride_W, stride_W, device=device, dtype=torch.int64)
    b = torch.arange(batch, device=device, dtype=torch.int64)
    shape = (batch, out_h, out_w)
    return (b.view(-1, 1, 1).expand(shape).contiguous(), h.view(1, -1, 1).expand(shape).contiguous(),
            w.view(1, 1, -1).expand(shape).contiguous())


class StrideGrid:
    """The regular centre grid a set-abstraction level samples: out_h x out_w centres at
    (h*stride_h, w*stride_w).  Unpacks (`*grid`) to the reference's (b, h, w) index triple
    lazily, so it can be handed to anything that expects `sample_idx`."""

    def __init__(self, batch, out_h, out_w, stride_h, stride_w, device):
        self.batch, self.out_h, self.out_w = batch, out_h, out_w
        self.stride_h, self.stride_w, self.device = stride_h, stride_w, device

    def as_tuple(self):
        return (self.out_h, self.out_w, self.stride_h, self.stride_w)

    def __iter__(self):
        return iter(get_sample_idx(self.batch, self.out_h, self.out_w, self.stride_h, self.stride_w, self.device))

    def take(self, image):
        """image (B,H,W,C) -> the grid's pixels (B,out_h,out_w,C); a strided view, no gather."""
        return image[:, ::self.stride_h, ::self.stride_w][:, :self.out_h, :self.out_w]


# ---------------------------------------------------------------------------------------------
# window select
# ---------------------------------------------------------------------------------------------
def _select_full(xyz1_proj, xyz2_proj, idx_n2, kernel_shape, knn_points, stride_h, stride_w, distance, flag):
    xyz1_proj, xyz2_proj = xyz1_proj.contiguous(), xyz2_proj.contiguous()
    batch, height, width, _ = xyz1_proj.shape
    small_h, small_w = xyz2_proj.shape[1:3]
    n_points = idx_n2.shape[1]
    dev = xyz1_proj.device
    random_hw = torch.arange(0, kernel_shape[0] * kernel_shape[1], device=dev, dtype=torch.int32)
    sel = torch.zeros(3, batch, n_points, knn_points, device=dev, dtype=torch.int64)
    mask = torch.zeros(batch, n_points, knn_points, 1, device=dev, dtype=torch.float32)
    _cabi.fused_conv_select_k(xyz1_proj, xyz2_proj, idx_n2.contiguous(), random_hw, height, width, n_points,
                              kernel_shape[0], kernel_shape[1], knn_points, flag, distance, stride_h, stride_w,
                              sel[0], sel[1], sel[2], mask, small_h, small_w)
    return sel[0], sel[1], sel[2], mask


def get_neighbor_copy(xyz1_proj, xyz2_proj, idx_n2, kernel_shape, knn_points, stride_h=1, stride_w=1, distance=10):
    """For each centre idx_n2 (B,N,2) of xyz1_proj (B,H,W,3): the knn_points nearest valid cells of
    xyz2_proj inside the kernel_shape window, padded with the nearest one (FLAG_COPY), width
    wrapping (FLAG_SHIFT).  -> (b, h, w) int64 (B,N,K) and mask f32 (B,N,K,1)."""
    with torch.no_grad():
        return _select_full(xyz1_proj, xyz2_proj, idx_n2, kernel_shape, knn_points, stride_h, stride_w, distance,
                            FLAG_SHIFT | FLAG_COPY)


def get_neighbor_att(xyz1_proj, xyz2_proj, idx_n2, kernel_shape, knn_points, stride_h=1, stride_w=1, distance=10):
    """Same without padding: empty slots keep index 0 and mask 0."""
    with torch.no_grad():
        return _select_full(xyz1_proj, xyz2_proj, idx_n2, kernel_shape, knn_points, stride_h, stride_w, distance,
                            FLAG_SHIFT)


def select_flat(xyz1_proj, xyz2_proj, centres, kernel_shape, knn_points, flag, distance, stride_h=1, stride_w=1):
    """Compact select.  centres: a StrideGrid or an int32 (B,N,2) tensor.
    -> flat index h*W2 + w int32 (B,N,K), mask f32 (B,N,K,1)."""
    with torch.no_grad():
        xyz1_proj, xyz2_proj = xyz1_proj.contiguous(), xyz2_proj.contiguous()
        B = xyz1_proj.shape[0]
        if isinstance(centres, StrideGrid):
            n, idx_n2, grid = centres.out_h * centres.out_w, None, centres.as_tuple()
        else:
            n, idx_n2, grid = centres.shape[1], centres.contiguous(), None
        flat = torch.empty(B, n, knn_points, device=xyz1_proj.device, dtype=torch.int32)
        mask = torch.empty(B, n, knn_points, 1, device=xyz1_proj.device, dtype=torch.float32)
        _cabi.select_k_flat(xyz1_proj, xyz2_proj, idx_n2, grid, kernel_shape, knn_points, flag, distance,
                            (stride_h, stride_w), flat, mask)
    return flat, mask


# ---------------------------------------------------------------------------------------------
# gathers
# ---------------------------------------------------------------------------------------------
class GatherRows(Function):
    """feature (B,HW,C) channels-last, flat_idx (B,M) int32 -> (B,M,C); backward scatter-adds."""

    @staticmethod
    def forward(ctx, feature, flat_idx):
        feature = feature.contiguous()
        B, HW, C = feature.shape
        M = flat_idx.shape[1]
        out = torch.empty(B, M, C, device=feature.device, dtype=torch.float32)
        _cabi.gather_rows(B, HW, C, M, feature, flat_idx, out)
        ctx.save_for_backward(flat_idx)
        ctx.hw = HW
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (flat_idx,) = ctx.saved_tensors
        B, M, C = grad_out.shape
        from .. import scratch
        grad = scratch.zeros((B, ctx.hw, C), torch.float32, grad_out.device)
        _cabi.gather_rows_grad(B, ctx.hw, C, M, grad_out.contiguous(), flat_idx, grad)
        return grad, None


def gather_rows(feature, flat_idx):
    """feature (B,...,C) with any number of middle dims, flat_idx int32 (B, *S) -> (B, *S, C)."""
    B, C = feature.shape[0], feature.shape[-1]
    out = GatherRows.apply(feature.reshape(B, -1, C), flat_idx.reshape(B, -1))
    return out.view(*flat_idx.shape, C)


def gather_torch(feature, neigh_b_idx, neigh_h_idx, neigh_w_idx, batch, height, width):
    """feature (B,H,W,C); neigh_*_idx (B,H',W') -> (B,H',W',C) (src/projectPN/utils.py:36-60).
    neigh_b_idx is ignored, as in the reference."""
    flat = (neigh_h_idx * width + neigh_w_idx).to(torch.int32)
    return gather_rows(feature.reshape(batch, height * width, -1), flat)


def check_valid(xyz):
    return torch.any(torch.ne(xyz, 0), dim=-1, keepdim=True).float()


# ---------------------------------------------------------------------------------------------
# spherical projection
# ---------------------------------------------------------------------------------------------
def project_seq(xyz, features, H, W, use_rank=True, fup=2.0, fdown=-24.8):
    """xyz (B,N,3), features list of (B,N,D) -> xyz_proj (B,H,W,3), [feature_proj (B,H,W,D)].
    use_rank: the closest point of a cell wins (the reference sorts by descending range and lets
    the last writer win); otherwise the highest point index wins (src/projectPN/utils.py:111-187)."""
    xyz = xyz.float().contiguous()
    features = [f.float().contiguous() for f in features]
    B = xyz.shape[0]
    if use_rank:
        with torch.no_grad():
            rank = torch.argsort(torch.norm(xyz, p=2, dim=2), dim=1, descending=True)
        xyz = torch.gather(xyz, 1, rank[:, :, None].expand(-1, -1, 3)).contiguous()
        features = [torch.gather(f, 1, rank[:, :, None].expand(-1, -1, f.shape[-1])).contiguous() for f in features]
    dev = xyz.device
    xyz_proj = torch.empty(B, H, W, 3, dtype=torch.float32, device=dev)
    feat_projs = [torch.empty(B, H, W, f.shape[-1], dtype=torch.float32, device=dev) for f in features]
    owner = torch.empty(B, H, W, dtype=torch.int32, device=dev)
    with torch.no_grad():
        _cabi.project_seq(xyz, features, H, W, fup, fdown, xyz_proj, feat_projs, owner)
    return xyz_proj, feat_projs


def project(xyz, feature, H, W, fup=2.0, fdown=-24.8):
    """Single-feature, always-ranked form (src/projectPN/utils.py:189-251)."""
    xyz_proj, (feature_proj,) = project_seq(xyz, [feature], H, W, True, fup, fdown)
    return xyz_proj, feature_proj


def get_pixel_posinfo(rf, K):
    B, H, W, _ = rf.shape
    idx = get_idx_cuda(B, H, W, rf.device).float()
    idx = torch.cat([idx, torch.ones(B, H * W, 1, dtype=torch.float32, device=rf.device)], dim=-1)
    return torch.bmm(inverse3x3(K), idx.permute(0, 2, 1)).permute(0, 2, 1).reshape(B, H, W, 3)


def inverse3x3(m):
    """Batched 3x3 inverse by the adjugate, on device.  The reference round-trips through the
    host for this (torch.inverse(K.cpu()).to(device), src/modellearn_proj_center.py:282), a
    synchronisation that would also forbid capturing the step into a CUDA graph."""
    a, b, c = m[:, 0, 0], m[:, 0, 1], m[:, 0, 2]
    d, e, f = m[:, 1, 0], m[:, 1, 1], m[:, 1, 2]
    g, h, i = m[:, 2, 0], m[:, 2, 1], m[:, 2, 2]
    A, Bc, C = e * i - f * h, -(d * i - f * g), d * h - e * g
    det = a * A + b * Bc + c * C
    adj = torch.stack([A, -(b * i - c * h), b * f - c * e,
                       Bc, a * i - c * g, -(a * f - c * d),
                       C, -(a * h - b * g), a * e - b * d], dim=-1).view(-1, 3, 3)
    return adj / det.view(-1, 1, 1)


def pixel_rays(intrinsic, h, w, img_h, img_w):
    """Pixel centres of an (h, w) feature map of an (img_h, img_w) image on the normalised camera plane, (B, h*w, 3):
    K'^-1 [u, v, 1] with K' the intrinsic (B,3,3) rescaled to the map (src/modellearn_proj_center.py:275-287:
    change_intrinsic -> torch.inverse on the host -> set_id_grid -> bmm).  One kernel on the device; no gradient (the
    intrinsic is data)."""
    sx, sy = w / img_w, h / img_h
    B = intrinsic.shape[0]
    if intrinsic.is_cuda:
        from .. import _cabi
        K = intrinsic.detach().float().contiguous()
        rays = torch.empty(B, h * w, 3, dtype=torch.float32, device=K.device)
        _cabi.call("i2p_pixel_rays", K.device, B, h, w, float(sx), float(sy), _cabi._ptr(K, torch.float32, "intrinsic"), rays.data_ptr())
        return rays
    K = torch.cat([intrinsic[:, 0:1] * sx, intrinsic[:, 1:2] * sy, intrinsic[:, 2:3]], dim=1).float()
    v, u = torch.meshgrid(torch.arange(h, dtype=K.dtype), torch.arange(w, dtype=K.dtype), indexing="ij")
    grid = torch.stack([u, v, torch.ones_like(u)], dim=-1).reshape(1, h * w, 3)
    return (grid.unsqueeze(2) * inverse3x3(K).unsqueeze(1)).sum(-1)


# ---------------------------------------------------------------------------------------------
# brute-force kNN grouping
# ---------------------------------------------------------------------------------------------
def square_distance(src, dst):
    """(B,N,C),(B,M,C) -> (B,N,M) as -2 s.d + |s|^2 + |d|^2.  Kept for API parity; knn_point no
    longer materialises this matrix."""
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    dist += torch.sum(src ** 2, -1).unsqueeze(-1)
    dist += torch.sum(dst ** 2, -1).unsqueeze(1)
    return dist


def knn_point(nsample, xyz, new_xyz):
    """The nsample nearest of xyz (B,N,3) for each new_xyz (B,S,3) -> int64 (B,S,nsample), sorted
    by (distance, index).  The reference's order is unspecified (topk sorted=False, :379)."""
    with torch.no_grad():
        xyz, new_xyz = xyz.float().contiguous(), new_xyz.float().contiguous()
        B, N, _ = xyz.shape
        S = new_xyz.shape[1]
        idx = torch.empty(B, S, nsample, dtype=torch.int64, device=xyz.device)
        _cabi.knn_point(B, N, S, nsample, xyz, new_xyz, idx)
    return idx


def index_points_group(points, knn_idx):
    """points (B,N,C), knn_idx (B,S,K) -> (B,S,K,C).  A channels-last row gather; the reference
    transposes to (B,C,N), runs grouping_operation and permutes back (:382-393)."""
    return gather_rows(points, knn_idx.to(torch.int32))


def grouping(feature, K, src_xyz, q_xyz, use_xyz=False):
    """-> grouped_xyz (B,S,K,3), xyz_diff, new_points (B,S,K,c[+3]), point_indices (B,S,K)."""
    q_xyz, src_xyz = q_xyz.contiguous(), src_xyz.contiguous()
    point_indices = knn_point(K, src_xyz, q_xyz)
    idx32 = point_indices.to(torch.int32)
    grouped_xyz = gather_rows(src_xyz, idx32)
    xyz_diff = grouped_xyz - q_xyz.unsqueeze(2)
    grouped_feature = gather_rows(feature, idx32)
    new_points = torch.cat([xyz_diff, grouped_feature], dim=-1) if use_xyz else grouped_feature
    return grouped_xyz, xyz_diff, new_points, point_indices
