"""Bottleneck math with the reference's names and signatures (keypoints/models/functional.py), executed by
hand-written CUDA kernels (csrc/kp_bottleneck.cu) through autograd Functions with closed-form backward.

Note the reference's definition (functional.py:5-24): the "spatial soft-max" is two independent 1-D
soft-maxes of the row means and the column means of the heat-map, and keypoints are ordered (y, x).
"""
import torch

from .. import lib as L


def _planes(t):
    t = t.contiguous()
    if t.dtype != torch.float32:
        t = t.float()
    return t


class _SpatialSoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, heat):
        heat = _planes(heat)
        n, k, h, w = heat.shape
        kp = torch.empty((n, k, 2), dtype=torch.float32, device=heat.device)
        ph = torch.empty((n, k, h), dtype=torch.float32, device=heat.device)
        pw = torch.empty((n, k, w), dtype=torch.float32, device=heat.device)
        L.call('kp_spatial_softmax_fwd', L.stream(), L.ptr(heat), n * k, h, w, L.ptr(kp), L.ptr(ph), L.ptr(pw))
        ctx.save_for_backward(kp, ph, pw)
        ctx.hw = (h, w)
        ctx.mark_non_differentiable(ph, pw)
        return kp, ph, pw

    @staticmethod
    def backward(ctx, dk, _dph, _dpw):
        kp, ph, pw = ctx.saved_tensors
        h, w = ctx.hw
        n, k, _ = kp.shape
        dheat = torch.empty((n, k, h, w), dtype=torch.float32, device=kp.device)
        L.call('kp_spatial_softmax_bwd', L.stream(), L.ptr(_planes(dk)), L.ptr(kp), L.ptr(ph), L.ptr(pw), n * k, h, w,
               L.ptr(dheat))
        return dheat


class _GaussianFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kp, height, width, sigma, eps):
        kp = _planes(kp)
        n, k, _ = kp.shape
        m = torch.empty((n, k, height, width), dtype=torch.float32, device=kp.device)
        L.call('kp_gaussian_fwd', L.stream(), L.ptr(kp), n * k, height, width, float(sigma), float(eps), L.ptr(m))
        ctx.save_for_backward(kp)
        ctx.cfg = (height, width, float(sigma), float(eps))
        return m

    @staticmethod
    def backward(ctx, dm):
        (kp,) = ctx.saved_tensors
        h, w, sigma, eps = ctx.cfg
        n, k, _ = kp.shape
        dm = _planes(dm)
        dk = torch.empty_like(kp)
        L.call('kp_gaussian_bwd', L.stream(), L.nchw(dm), 0, L.ptr(kp), None, n, k, h, w, sigma, eps, L.ptr(dk))
        return dk, None, None, None, None


class _TransportMaxFn(torch.autograd.Function):
    """phi_s (1-M_s)(1-M_t) + phi_t M_t with M = max_k gaussian(k) (models/transporter.py:57-60)."""

    @staticmethod
    def forward(ctx, phi_s, phi_t, k_s, k_t, sigma, eps):
        phi_s, phi_t, k_s, k_t = _planes(phi_s), _planes(phi_t), _planes(k_s), _planes(k_t)
        n, c, h, w = phi_t.shape
        K = k_t.shape[1]
        dev = phi_t.device
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=dev)
        mask_s = torch.empty((n, 1, h, w), dtype=torch.float32, device=dev)
        mask_t = torch.empty((n, 1, h, w), dtype=torch.float32, device=dev)
        amax = torch.empty((n, h, w), dtype=torch.int32, device=dev)
        L.call('kp_transport_fwd', L.stream(), L.nchw(phi_s), L.nchw(phi_t), L.ptr(k_s), L.ptr(k_t), L.nchw(out), 0,
               L.ptr(mask_s), L.ptr(mask_t), L.ptr(amax), n, h, w, c, K, float(sigma), float(eps))
        ctx.save_for_backward(phi_s, phi_t, k_t, mask_s, mask_t, amax)
        ctx.cfg = (float(sigma), float(eps))
        ctx.mark_non_differentiable(mask_s)
        return out, mask_s, mask_t

    @staticmethod
    def backward(ctx, dout, _dms, _dmt):
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[2]:
            raise NotImplementedError('the source branch of the Transporter is a constant (reference runs it under no_grad)')
        phi_s, phi_t, k_t, mask_s, mask_t, amax = ctx.saved_tensors
        sigma, eps = ctx.cfg
        n, c, h, w = phi_t.shape
        K = k_t.shape[1]
        dout = _planes(dout)
        dphi_t = torch.empty_like(phi_t)
        dmask = torch.empty((n, h, w, 1), dtype=torch.float32, device=phi_t.device)
        L.call('kp_transport_bwd', L.stream(), L.nchw(dout), 0, L.nchw(phi_s), L.nchw(phi_t), L.ptr(mask_s),
               L.ptr(mask_t), L.nchw(dphi_t), L.ptr(dmask), n, h, w, c)
        dk = torch.empty_like(k_t)
        L.call('kp_gaussian_bwd', L.stream(), L.view(dmask), 0, L.ptr(k_t), L.ptr(amax), n, K, h, w, sigma, eps,
               L.ptr(dk))
        return None, dphi_t, None, dk, None, None


class _TransportModeFn(torch.autograd.Function):
    """The 'sum_and_clamp' and 'loop' combine modes (models/transporter.py:41-50); 'max' is accepted as well."""

    @staticmethod
    def forward(ctx, phi_s, phi_t, k_s, k_t, mode, sigma, eps):
        phi_s, phi_t, k_s, k_t = _planes(phi_s), _planes(phi_t), _planes(k_s), _planes(k_t)
        n, c, h, w = phi_t.shape
        K = k_t.shape[1]
        dev = phi_t.device
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=dev)
        mask_s = torch.empty((n, 1, h, w), dtype=torch.float32, device=dev)
        mask_t = torch.empty((n, 1, h, w), dtype=torch.float32, device=dev)
        aux = torch.zeros((n, h, w), dtype=torch.int32, device=dev)
        coef = torch.empty((n, h, w, 2), dtype=torch.float32, device=dev)
        L.call('kp_transport_mode_fwd', L.stream(), int(mode), L.nchw(phi_s), L.nchw(phi_t), L.ptr(k_s), L.ptr(k_t),
               L.nchw(out), 0, L.ptr(mask_s), L.ptr(mask_t), L.ptr(aux), L.ptr(coef), n, h, w, c, K, float(sigma), float(eps))
        ctx.save_for_backward(phi_s, phi_t, k_s, k_t, mask_s, mask_t, aux, coef)
        ctx.cfg = (int(mode), float(sigma), float(eps))
        ctx.mark_non_differentiable(mask_s)
        return out, mask_s, mask_t

    @staticmethod
    def backward(ctx, dout, _dms, _dmt):
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[2]:
            raise NotImplementedError('the source branch of the Transporter is a constant (reference runs it under no_grad)')
        phi_s, phi_t, k_s, k_t, mask_s, mask_t, aux, coef = ctx.saved_tensors
        mode, sigma, eps = ctx.cfg
        n, c, h, w = phi_t.shape
        K = k_t.shape[1]
        dout = _planes(dout)
        dphi_t = torch.empty_like(phi_t)
        dm_t = torch.empty((n, K, h, w), dtype=torch.float32, device=phi_t.device)
        L.call('kp_transport_mode_bwd', L.stream(), mode, L.nchw(dout), 0, L.nchw(phi_s), L.nchw(phi_t), L.ptr(k_s),
               L.ptr(k_t), L.ptr(mask_s), L.ptr(mask_t), L.ptr(aux), L.ptr(coef), L.nchw(dphi_t), L.ptr(dm_t), n, h, w, c, K,
               sigma, eps)
        dk = torch.empty_like(k_t)
        L.call('kp_gaussian_bwd', L.stream(), L.nchw(dm_t), 0, L.ptr(k_t), None, n, K, h, w, sigma, eps, L.ptr(dk))
        return None, dphi_t, None, dk, None, None, None


def spacial_softmax(heatmap, probs=False):
    """functional.py:27-34."""
    k, ph, pw = _SpatialSoftmaxFn.apply(heatmap)
    return (k, (ph, pw)) if probs else k


def spacial_logsoftmax(heatmap, probs=False):
    """functional.py:37-44 — same values as spacial_softmax (the log-space detour cancels)."""
    k, ph, pw = _SpatialSoftmaxFn.apply(heatmap)
    return (k, (ph, pw)) if probs else k


def gaussian_like_function(kp, height, width, sigma=0.1, eps=1e-6):
    """functional.py:56-63: exp(-sqrt((y-ky)^2 + (x-kx)^2 + eps) / (2 sigma^2))."""
    return _GaussianFn.apply(kp, int(height), int(width), sigma, eps)


def transport_max(phi_s, phi_t, k_s, k_t, sigma=0.1, eps=1e-6):
    """Fused render + max-over-keypoints + transport; returns (phi, mask_s, mask_t)."""
    return _TransportMaxFn.apply(phi_s, phi_t, k_s, k_t, sigma, eps)


def transport(phi_s, phi_t, k_s, k_t, mode='max', sigma=0.1, eps=1e-6):
    """Render + combine + transport for any combine mode of TransporterNet.forward ('max' | 'sum_and_clamp' | 'loop',
    models/transporter.py:41-60); returns (phi, mask_s, mask_t) exactly as the reference's forward leaves them."""
    if mode == 'max':
        return transport_max(phi_s, phi_t, k_s, k_t, sigma, eps)
    if mode not in L.COMBINE:
        raise NotImplementedError(f"combine mode {mode!r}: 'pretrained_network' needs a MaskMaker network that the "
                                  "reference never builds in make() (models/transporter.py:112-128)")
    return _TransportModeFn.apply(phi_s, phi_t, k_s, k_t, L.COMBINE[mode], sigma, eps)


# -- small helpers kept for API completeness (not on the training path; plain tensor algebra) ------------
def marginal_softmax(heatmap, dim):
    return torch.softmax(heatmap.mean(dim=dim), dim=2)


def marginal_logsoftmax(heatmap, dim):
    return torch.log_softmax(heatmap.mean(dim=dim), dim=2)


def prob_to_keypoints(prob, length):
    return (prob * torch.linspace(0, 1, length, device=prob.device, dtype=prob.dtype)).sum(dim=2)


def logprob_to_keypoints(prob, length):
    return prob.exp().mul(torch.linspace(0, 1, length, device=prob.device, dtype=prob.dtype)).sum(dim=2)


def squared_diff(h, height):
    ruler = torch.linspace(0, 1, height, device=h.device, dtype=h.dtype)
    return (ruler.view(1, 1, -1) - h.unsqueeze(-1)) ** 2


def point_map(kp, h, w):
    return -(squared_diff(kp[:, :, 0], h).unsqueeze(-1) + squared_diff(kp[:, :, 1], w).unsqueeze(-2))
