"""CPU oracle for the unsupervised-keypoint training hot path.

TEST INFRASTRUCTURE ONLY.  This module is a CPU restatement (functional torch
ops on CPU tensors, fp32 or fp64) of the reference algorithm.  It is the
checker for the CUDA path, never the thing measured or shipped: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``keypoints_b200/``
imports it.  The functions follow their inputs' device (linspace rulers are computed on
the CPU exactly as the reference does, then moved), so ``bench.py``'s
``gpu_eager_baseline`` leg can also run this same graph in torch eager on cuda:0 —
the reference's own GPU execution, the "kernel to beat" (SURVEY 8d).

Parity pinning: the reference's own test-suite holds no golden vectors for this
path, so the oracle is pinned against fixtures generated
by importing the *reference itself* in the build container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``) and against the
known answers derived from the literals in the reference tests
(``tests/tests.py:11-36,239-240``; ``keypoints/tps.py:197-212``).
``tests/test_oracle_golden.py`` enforces both.

Every function cites the reference file:line it restates (paths relative to
the reference checkout).  The arithmetic below is all third-party ATen
(torch 2.11, unpinned in the reference's requirements.txt:2); version-sensitive
semantics inherited from this torch are spelled out where they matter
(grid_sample/affine_grid align_corners=False, UpsamplingBilinear2d
align_corners=True).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

# ---------------------------------------------------------------------------
# layer tables (data restated from keypoints/models/vgg.py:48-70)
# ---------------------------------------------------------------------------
ENCODER_CFG = {
    'A': [64, 'M', 128, 'M', 256, 256, 'M', 512, 512, 'M', 512, 512, 'M'],
    'B': [64, 64, 'M', 128, 128, 'M', 256, 256, 'M', 512, 512, 'M', 512, 512, 'M'],
    'D': [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512, 'M'],
    'E': [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 256, 'M', 512, 512, 512, 512, 'M',
          512, 512, 512, 512, 'M'],
    'F': [64, 128, 'M', 256, 256, 'M', 512, 512, 'M', 512, 512],
    'VGG_PONG': [16, 'M', 16, 'M', 32],
    'VGG_PONG_TRIVIAL': [16, 16],
    'VGG_PONG_LAYERNECK': [16, 32],
    'VGG_PACMAN': [16, 32, 32, 16],
    'VGG_PACMAN_2': [16, 32, 32, 'M', 64],
    'MAPPER': [8, 8],
}
DECODER_CFG = {
    'A': [512, 512, 'U', 256, 256, 'U', 256, 256, 'U', 128, 'U', 64, 'U'],
    'F': [512, 512, 'U', 256, 256, 'U', 256, 256, 'U', 128, 64],
    'VGG_PONG': [32, 'U', 16, 'U', 16],
    'VGG_PONG_TRIVIAL': [16, 16],
    'VGG_PONG_LAYERNECK': [32, 32, 16, 16],
    'VGG_PACMAN': [16, 32, 32, 16],
    'VGG_PACMAN_2': [64, 'U', 32, 32, 16],
}

BN_EPS = 1e-5        # nn.BatchNorm2d default, vgg.py:35 / knn.py:117
BN_MOMENTUM = 0.1
LEAKY_SLOPE = 0.01   # nn.LeakyReLU default, knn.py:118,124 / transporter.py:122


def unit_ops(cfg: Sequence, in_channels: int, out_channels: int, core_act: str) -> List[dict]:
    """Flatten one ``knn.Unit`` (knn.py:110-130) into a list of ops.

    in_block  = RepPad1 + Conv3x3(in->cfg[0]) + BN + LeakyReLU       (knn.py:115-119)
    core      = vgg.make_layers(cfg): cfg[0] is the *input* width     (vgg.py:21-22)
                'M' -> MaxPool2x2, 'U' -> bilinear x2, int v -> RepPad1+Conv3x3+BN+act (vgg.py:23-37)
    out_block = Conv1x1(core_out->out) + LeakyReLU, no BN             (knn.py:123-125)

    ``key`` fields reproduce the Sequential indices the reference's state_dict uses.
    """
    ops = [dict(op='conv', k=3, cin=in_channels, cout=cfg[0], conv='in_block.1', bn='in_block.2', act='leaky')]
    idx, c = 0, cfg[0]
    for v in cfg[1:]:
        if v == 'M':
            ops.append(dict(op='pool')); idx += 1
        elif v == 'U':
            ops.append(dict(op='up')); idx += 1
        elif v == 'L':
            idx += 1                      # ActivationMap is identity (knn.py:79-84)
        else:
            ops.append(dict(op='conv', k=3, cin=c, cout=v, conv=f'core.{idx + 1}', bn=f'core.{idx + 2}', act=core_act))
            idx += 4; c = v
    ops.append(dict(op='conv', k=1, cin=c, cout=out_channels, conv='out_block.0', bn=None, act='leaky'))
    return ops


def unit_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str, ops: List[dict],
                 training: bool = True) -> torch.Tensor:
    """``knn.Unit.forward`` (knn.py:127-130) on a flat parameter dict.

    Train-mode BatchNorm uses batch statistics and updates ``running_mean`` /
    ``running_var`` / ``num_batches_tracked`` in ``sd`` in place (the reference
    never calls ``.eval()``).
    """
    h = x
    for o in ops:
        if o['op'] == 'pool':
            h = F.max_pool2d(h, kernel_size=2, stride=2)                    # vgg.py:24
        elif o['op'] == 'up':
            h = F.interpolate(h, scale_factor=2, mode='bilinear', align_corners=True)   # vgg.py:26
        else:
            w, b = sd[f'{prefix}{o["conv"]}.weight'], sd[f'{prefix}{o["conv"]}.bias']
            if o['k'] == 3:
                h = F.pad(h, (1, 1, 1, 1), mode='replicate')               # vgg.py:30, knn.py:115
            h = F.conv2d(h, w, b)
            if o['bn'] is not None:
                p = f'{prefix}{o["bn"]}'
                if training:
                    sd[p + '.num_batches_tracked'] += 1
                h = F.batch_norm(h, sd[p + '.running_mean'], sd[p + '.running_var'], sd[p + '.weight'],
                                 sd[p + '.bias'], training=training, momentum=BN_MOMENTUM, eps=BN_EPS)
            h = F.leaky_relu(h, LEAKY_SLOPE) if o['act'] == 'leaky' else F.relu(h)
    return h


# ---------------------------------------------------------------------------
# bottleneck math (keypoints/models/functional.py)
# ---------------------------------------------------------------------------
def spatial_logsoftmax(heat: torch.Tensor) -> Tuple[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
    """``spacial_logsoftmax(heatmap, probs=True)`` (functional.py:37-44).

    Two independent 1-D log-softmaxes of the row means and column means
    (functional.py:11-14), expectation against linspace(0,1) taken in log space
    (functional.py:22-24).  Returns k (N,K,2) ordered (y,x) and (p_h, p_w).
    """
    n, k, h, w = heat.shape
    logp_h = F.log_softmax(heat.mean(dim=3), dim=2)
    logp_w = F.log_softmax(heat.mean(dim=2), dim=2)
    ruler_h = torch.log(torch.linspace(0, 1, h)).to(heat)
    ruler_w = torch.log(torch.linspace(0, 1, w)).to(heat)
    ky = torch.exp(logp_h + ruler_h).sum(dim=2)
    kx = torch.exp(logp_w + ruler_w).sum(dim=2)
    return torch.stack((ky, kx), dim=2), (torch.exp(logp_h), torch.exp(logp_w))


def spatial_softmax(heat: torch.Tensor) -> Tuple[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
    """``spacial_softmax(heatmap, probs=True)`` (functional.py:27-34)."""
    n, k, h, w = heat.shape
    p_h = F.softmax(heat.mean(dim=3), dim=2)
    p_w = F.softmax(heat.mean(dim=2), dim=2)
    ky = (p_h * torch.linspace(0, 1, h).to(heat)).sum(dim=2)
    kx = (p_w * torch.linspace(0, 1, w).to(heat)).sum(dim=2)
    return torch.stack((ky, kx), dim=2), (p_h, p_w)


def gaussian_like(kp: torch.Tensor, height: int, width: int, sigma: float = 0.1, eps: float = 1e-6) -> torch.Tensor:
    """``gaussian_like_function`` (functional.py:56-63): exp(-sqrt(dy^2+dx^2+eps)/(2 sigma^2))."""
    ys = torch.linspace(0, 1, height).to(kp).view(1, 1, height, 1)
    xs = torch.linspace(0, 1, width).to(kp).view(1, 1, 1, width)
    dy2 = (ys - kp[:, :, 0, None, None]) ** 2
    dx2 = (xs - kp[:, :, 1, None, None]) ** 2
    return torch.exp(-torch.sqrt(dy2 + dx2 + eps) / (2 * sigma ** 2))


def transport(phi_s, m_s, phi_t, m_t, mode: str = 'max'):
    """Feature transport (models/transporter.py:41-60).  Returns (phi, mask_s, mask_t)."""
    if mode == 'max':
        mask_s = m_s.max(dim=1, keepdim=True)[0]
        mask_t = m_t.max(dim=1, keepdim=True)[0]
    elif mode == 'sum_and_clamp':
        mask_s = m_s.sum(dim=1, keepdim=True).clamp(0.0, 1.0)
        mask_t = m_t.sum(dim=1, keepdim=True).clamp(0.0, 1.0)
    elif mode == 'loop':
        phi = phi_s
        for i in range(m_t.shape[1]):
            mask_s, mask_t = m_s[:, i:i + 1], m_t[:, i:i + 1]
            phi = phi * (1 - mask_s) * (1 - mask_t) + phi_t * mask_t
        return phi, mask_s, mask_t
    else:
        raise ValueError(mode)
    return phi_s * (1 - mask_s) * (1 - mask_t) + phi_t * mask_t, mask_s, mask_t


def l2_reconstruction_loss(x, x_, loss_mask=None):
    """transporter.py:56-60 == keypoints.py:54-58."""
    loss = (x - x_) ** 2
    if loss_mask is not None:
        loss = loss * loss_mask
    return loss.mean()


# ---------------------------------------------------------------------------
# models
# ---------------------------------------------------------------------------
def transporter_ops(model_type: str, cin: int, z: int, K: int):
    """Layer lists of ``transporter.make`` (models/transporter.py:112-128): encoder and
    keypoint cores use LeakyReLU (:122-123,127), the decoder core the default ReLU (:125, vgg.py:19)."""
    return dict(feature=unit_ops(ENCODER_CFG[model_type], cin, z, 'leaky'),
                keypoint=unit_ops(ENCODER_CFG[model_type], cin, K, 'leaky'),
                decoder=unit_ops(DECODER_CFG[model_type], z, cin, 'relu'))


def keynet_ops(model_type: str, cin: int, z: int, K: int):
    """Layer lists of ``keynet.make`` (models/keynet.py:50-59); decoder input width z+K (:55)."""
    return dict(encoder=unit_ops(ENCODER_CFG[model_type], cin, z, 'leaky'),
                keypoint=unit_ops(ENCODER_CFG[model_type], cin, K, 'leaky'),
                decoder=unit_ops(DECODER_CFG[model_type], z + K, cin, 'relu'))


def autoencoder_ops(model_type: str, cin: int, z: int):
    """Layer lists of the pre-training auto-encoder (autoencode.py:59-66): encoder AND decoder cores use LeakyReLU."""
    return dict(encoder=unit_ops(ENCODER_CFG[model_type], cin, z, 'leaky'),
                decoder=unit_ops(DECODER_CFG[model_type], z, cin, 'leaky'))


def autoencoder_forward(x, sd, ops, training: bool = True):
    """``AutoEncoder.forward`` (models/autoencoder.py:13-16).  Returns (z, x_hat)."""
    z = unit_forward(x, sd, 'encoder.', ops['encoder'], training)
    return z, unit_forward(z, sd, 'decoder.', ops['decoder'], training)


def transporter_forward(xs, xt, sd, ops, mode: str = 'max', sigma: float = 0.1, training: bool = True):
    """``TransporterNet.forward`` (models/transporter.py:34-64).

    Source branch under no_grad (:36-37) - its BatchNorm running stats still update.
    Returns (x_t, phi, k_xt, m_xt, (p_h, p_w), heatmap_xt, mask_xs, mask_xt).
    """
    def extract(x):                                           # models/transporter.py:27-32
        phi = unit_forward(x, sd, 'feature.', ops['feature'], training)
        heat = unit_forward(x, sd, 'keypoint.', ops['keypoint'], training)
        k, p = spatial_logsoftmax(heat)
        m = gaussian_like(k, phi.shape[2], phi.shape[3], sigma)
        return phi, heat, k, p, m

    with torch.no_grad():
        phi_s, _, _, _, m_s = extract(xs)
    phi_t, heat_t, k_t, p_t, m_t = extract(xt)
    phi, mask_s, mask_t = transport(phi_s, m_s, phi_t, m_t, mode)
    x_hat = unit_forward(phi, sd, 'decoder.', ops['decoder'], training)
    return x_hat, phi, k_t, m_t, p_t, heat_t, mask_s, mask_t


def keynet_forward(x, x_t, sd, ops, sigma: float = 0.1, training: bool = True):
    """``KeyNet.forward`` (models/keynet.py:21-31).  Returns (x_hat, z, k, m, (p_h,p_w), heatmap)."""
    z = unit_forward(x, sd, 'encoder.', ops['encoder'], training)
    heat = unit_forward(x_t, sd, 'keypoint.', ops['keypoint'], training)
    k, p = spatial_logsoftmax(heat)
    m = gaussian_like(k, z.shape[2], z.shape[3], sigma)
    x_hat = unit_forward(torch.cat((z, m), dim=1), sd, 'decoder.', ops['decoder'], training)
    return x_hat, z, k, m, p, heat


def init_state_dict(ops_by_unit: Dict[str, List[dict]], seed: int, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic parameters with the reference's init *distribution*
    (knn.py:12-23: conv kaiming-normal fan_out/relu, bias 0; BN gamma 1, beta 0) drawn from a
    numpy PCG64 stream so fixtures do not depend on torch's RNG implementation.
    Keys/shapes equal the reference's state_dict."""
    import numpy as np
    rng = np.random.default_rng(seed)
    sd: Dict[str, torch.Tensor] = {}
    for unit, ops in ops_by_unit.items():
        for o in ops:
            if o['op'] != 'conv':
                continue
            k, cin, cout = o['k'], o['cin'], o['cout']
            std = math.sqrt(2.0 / (cout * k * k))
            sd[f'{unit}.{o["conv"]}.weight'] = torch.from_numpy(
                (rng.standard_normal((cout, cin, k, k)) * std).astype('float32')).to(dtype)
            # reference init is bias 0 / gamma 1 / beta 0; perturb so tests exercise those terms
            sd[f'{unit}.{o["conv"]}.bias'] = torch.from_numpy(
                (rng.standard_normal(cout) * 0.05).astype('float32')).to(dtype)
            if o['bn'] is not None:
                sd[f'{unit}.{o["bn"]}.weight'] = torch.from_numpy(
                    (1.0 + 0.1 * rng.standard_normal(cout)).astype('float32')).to(dtype)
                sd[f'{unit}.{o["bn"]}.bias'] = torch.from_numpy(
                    (0.1 * rng.standard_normal(cout)).astype('float32')).to(dtype)
                sd[f'{unit}.{o["bn"]}.running_mean'] = torch.zeros(cout, dtype=dtype)
                sd[f'{unit}.{o["bn"]}.running_var'] = torch.ones(cout, dtype=dtype)
                sd[f'{unit}.{o["bn"]}.num_batches_tracked'] = torch.zeros((), dtype=torch.long)
    return sd


def trainable_keys(sd: Dict[str, torch.Tensor]) -> List[str]:
    return [k for k in sd if k.endswith('.weight') or k.endswith('.bias')]


# ---------------------------------------------------------------------------
# TPS + rotate augmentation (keypoints/tps.py, keypoints/data_augments.py)
# ---------------------------------------------------------------------------
def tps_grid(theta: torch.Tensor, ctrl: torch.Tensor, size: Tuple[int, int, int, int]) -> torch.Tensor:
    """``tps_grid`` (tps.py:60-87) + ``tps`` (tps.py:10-57), full (T+3) and reduced (T+2) forms.

    For every output pixel (x,y) in [0,1]^2: z = sum_t w_t U(|p-c_t|) + a0 + a1 x + a2 y with
    U(d) = d^2 log(d + 1e-6); returns ((x,y) + z) * 2 - 1 as an (N,H,W,2) sampling grid.
    """
    N, _, H, W = size
    xs = torch.linspace(0, 1, W).to(theta)
    ys = torch.linspace(0, 1, H).to(theta)
    gx = xs.view(1, 1, W).expand(N, H, W)
    gy = ys.view(1, H, 1).expand(N, H, W)
    if ctrl.dim() == 2:
        ctrl = ctrl.expand(N, *ctrl.shape)
    T = ctrl.shape[1]
    w, a = theta[:, :-3, :], theta[:, -3:, :]
    if theta.shape[1] == T + 2:                                   # reduced form, tps.py:47-50
        w = torch.cat((-w.sum(dim=1, keepdim=True), w), dim=1)
    dx = gx.unsqueeze(-1) - ctrl[:, None, None, :, 0]
    dy = gy.unsqueeze(-1) - ctrl[:, None, None, :, 1]
    D = torch.sqrt(dx * dx + dy * dy)
    U = D * D * torch.log(D + 1e-6)                               # tps.py:42-43
    b = torch.einsum('nhwt,ntc->nhwc', U, w)
    lin = a[:, None, None, 0, :] + gx.unsqueeze(-1) * a[:, None, None, 1, :] + gy.unsqueeze(-1) * a[:, None, None, 2, :]
    xy = torch.stack((gx, gy), dim=-1)
    return (xy + lin + b) * 2 - 1


def tps_transform(x, theta, ctrl):
    """``tps_transform`` (tps.py:128-131): bilinear grid_sample, zeros padding, align_corners=False
    (torch >= 1.3 default)."""
    grid = tps_grid(theta, ctrl, tuple(x.shape)).to(x.dtype)
    return F.grid_sample(x, grid, mode='bilinear', padding_mode='zeros', align_corners=False)


def rotate_affine_grid_multi(x, theta):
    """``rotate_affine_grid_multi`` (tps.py:154-166)."""
    c, s = torch.cos(theta), torch.sin(theta)
    A = torch.zeros(x.shape[0], 2, 3, dtype=x.dtype, device=x.device)
    A[:, 0, 0], A[:, 0, 1], A[:, 1, 0], A[:, 1, 1] = c, s, -s, c
    grid = F.affine_grid(A, list(x.shape), align_corners=False)
    return F.grid_sample(x, grid, mode='bilinear', padding_mode='zeros', align_corners=False)


def perturb(x, theta_tps, ctrl, theta_rot):
    """``peturb`` (data_augments.py:13-16)."""
    return rotate_affine_grid_multi(tps_transform(x, theta_tps, ctrl), theta_rot)


def tps_and_rotate(x, params1, params2):
    """``TpsAndRotate.__call__`` (data_augments.py:27-38) with the two random draws passed in
    explicitly (each = (theta_tps, ctrl, theta_rot)).  Returns (x1, x2, loss_mask); note x2 is
    warped twice and the mask follows both warps."""
    mask = torch.ones_like(x)
    x1 = perturb(x, *params1)
    mask = perturb(mask, *params1)
    x2 = perturb(x1, *params2)
    mask = perturb(mask, *params2)
    return x1, x2, mask


def sample_perturb_params(n: int, ctrl_pts: int, variance: float, max_rotate: float, generator=None):
    """``rand_peturb_params`` (data_augments.py:6-10) + ``tps_sample_params`` (tps.py:122-125)."""
    theta = torch.randn(n, ctrl_pts + 3, 2, generator=generator) * variance
    ctrl = torch.rand(n, ctrl_pts, 2, generator=generator)
    rot = (torch.rand(n, generator=generator) * 2 - 1) * max_rotate
    return theta, ctrl, rot


# ---------------------------------------------------------------------------
# optimiser (torch.optim.Adam defaults, transporter.py:47)
# ---------------------------------------------------------------------------
def adam_step(p, g, m, v, step: int, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8):
    """One Adam update, torch.optim.Adam semantics (no weight decay, no amsgrad); in place."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


# ---------------------------------------------------------------------------
# one whole training step (transporter.py:75-89 / keypoints.py:70-84), used as the CPU baseline
# ---------------------------------------------------------------------------
class OracleTrainer:
    """Functional train step on a flat parameter dict: forward, L2 loss, autograd backward, Adam."""

    def __init__(self, kind: str, model_type: str, cin: int, z: int, K: int, sd: Dict[str, torch.Tensor],
                 mode: str = 'max'):
        self.kind = kind
        self.mode = mode
        if kind == 'autoencoder':
            self.ops = autoencoder_ops(model_type, cin, z)
        else:
            self.ops = transporter_ops(model_type, cin, z, K) if kind == 'transporter' else keynet_ops(model_type, cin, z, K)
        self.sd = sd
        self.keys = trainable_keys(sd)
        for k in self.keys:
            sd[k].requires_grad_(True)
        self.m = {k: torch.zeros_like(sd[k]) for k in self.keys}
        self.v = {k: torch.zeros_like(sd[k]) for k in self.keys}
        self.t = 0

    def forward(self, a, b):
        if self.kind == 'transporter':
            return transporter_forward(a, b, self.sd, self.ops, mode=self.mode)
        if self.kind == 'autoencoder':          # autoencode.py:87-88: z, x_ = net(x); MSELoss(x_, x); (x_hat first here)
            z, xh = autoencoder_forward(a, self.sd, self.ops)
            return xh, z
        return keynet_forward(a, b, self.sd, self.ops)

    def step(self, a, b, mask=None, lr=1e-4):
        for k in self.keys:
            self.sd[k].grad = None
        out = self.forward(a, b)
        loss = l2_reconstruction_loss(out[0], b, mask)
        loss.backward()
        self.t += 1
        with torch.no_grad():
            for k in self.keys:
                g = self.sd[k].grad
                if g is None:
                    continue
                adam_step(self.sd[k], g, self.m[k], self.v[k], self.t, lr=lr)
        return loss.detach(), out
