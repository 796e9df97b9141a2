"""TPS + rotation warps with the reference's function names (keypoints/tps.py:122-131,154-166), executed by
csrc/kp_warp.cu: the sampling grid is evaluated per pixel inside the sampler kernel."""
import torch

from . import lib as L


def tps_sample_params(batch_size, num_control_points, var=0.05):
    """Same draws, same order, from torch's global CPU generator as tps.py:122-125."""
    theta = torch.randn(batch_size, num_control_points + 3, 2) * var
    cnt_points = torch.rand(batch_size, num_control_points, 2)
    return theta, cnt_points


def _f32(t, device):
    return t.to(device=device, dtype=torch.float32).contiguous()


def tps_transform(x, theta, cnt_points):
    if not x.is_cuda:
        raise RuntimeError('keypoints_b200.tps runs on CUDA only')
    xin = x.float().contiguous()
    n, c, h, w = xin.shape
    if cnt_points.dim() == 2:
        cnt_points = cnt_points.expand(n, *cnt_points.shape)
    theta, cnt_points = _f32(theta, x.device), _f32(cnt_points, x.device)
    T = cnt_points.shape[1]
    rows = theta.shape[1]
    if rows not in (T + 3, T + 2):
        raise ValueError('theta must be N x (T+3) x 2 or N x (T+2) x 2')
    out = torch.empty_like(xin)
    L.call('kp_tps_warp', L.stream(), L.ptr(xin), L.ptr(out), L.ptr(theta), L.ptr(cnt_points), n, c, h, w, T,
           1 if rows == T + 2 else 0)
    return out.to(x.dtype)


def rotate_affine_grid_multi(x, theta):
    if not x.is_cuda:
        raise RuntimeError('keypoints_b200.tps runs on CUDA only')
    xin = x.float().contiguous()
    n, c, h, w = xin.shape
    out = torch.empty_like(xin)
    L.call('kp_rotate_warp', L.stream(), L.ptr(xin), L.ptr(out), L.ptr(_f32(theta, x.device)), n, c, h, w)
    return out.to(x.dtype)
