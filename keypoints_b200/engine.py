"""Layer engine: runs one ``knn.Unit`` (reference keypoints/models/knn.py:110-130) forward and backward
as a sequence of C-ABI kernel launches on padded NHWC buffers.

Data layout (DESIGN.md "Data layout in HBM").  With T = fp32 ('fp32' precision, parity grade) or bf16
('bf16' precision, tensor cores):
  x[i]   : [N][H+2][W+2][Cp] T   replicate-padded input of conv i (Cp = channel pitch >= Cin)
  y[i]   : [N][H+2][W+2][Cout] T raw conv output (+bias), valid region top-left aligned
  dy[i]  : [N][H+2][W+2][Cout] T gradient w.r.t. y[i], interior aligned, ZERO border
  dx[i]  : [N][H+2][W+2][Cp] T   gradient w.r.t. the padded x[i] (border folded by the consumer)
Every buffer is also the flat matrix [Q = N (H+2) (W+2)][C] the tensor-core kernels index.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional

import torch

from . import lib as L

import os

BN_EPS = 1e-5
REGATHER = os.environ.get('KP_BN_REGATHER', '1') != '0'     # BN backward pass 2 recomputes dz instead of reading it back
REGATHER_POSTS = tuple(os.environ.get('KP_BN_REGATHER_POSTS', 'none,pool').split(','))


def regather_ok(s, h: int, w: int, dout: torch.Tensor, precision: str) -> bool:
    """True when both BatchNorm backward passes of this layer run as bulk-async streaming kernels (kp_bn_pipe.cuh:
    bf16, dense rows, whole 512-item chunks); only then is repeating the gather cheaper than the dz round trip."""
    if not (REGATHER and precision == 'bf16' and s.post in REGATHER_POSTS and s.cout % 8 == 0):
        return False
    ncg = s.cout // 8
    if ncg & (ncg - 1) or ncg > 64 or dout.dtype != torch.bfloat16 or dout.stride(3) != 1 or dout.stride(2) != s.cout:
        return False
    if s.post == 'none':
        return (w * ncg) % 512 == 0
    return h % 2 == 0 and w % 2 == 0 and ((w // 2) * ncg) % 256 == 0
BN_MOMENTUM = 0.1
WGRAD_IMG = os.environ.get('KP_WGRAD_IMG', '1') != '0'


@dataclass
class ConvSpec:
    """One conv of a Unit together with what follows it up to the next conv."""
    k: int
    cin: int
    cout: int
    bn: bool
    act: str                 # 'leaky' | 'relu' | 'none'
    post: str = 'none'       # 'none' | 'pool' | 'up'
    conv_key: str = ''       # state_dict prefix, e.g. 'core.1'
    bn_key: str = ''


@dataclass
class LayerParams:
    w: torch.Tensor
    b: Optional[torch.Tensor]
    gamma: Optional[torch.Tensor] = None
    beta: Optional[torch.Tensor] = None
    rmean: Optional[torch.Tensor] = None
    rvar: Optional[torch.Tensor] = None
    nbt: Optional[torch.Tensor] = None


@dataclass
class LayerGrads:
    dw: torch.Tensor
    db: Optional[torch.Tensor]
    dgamma: Optional[torch.Tensor] = None
    dbeta: Optional[torch.Tensor] = None


@dataclass
class LayerCtx:
    x: torch.Tensor
    y: torch.Tensor
    h: int
    w: int
    tc: bool
    scale: Optional[torch.Tensor] = None
    shift: Optional[torch.Tensor] = None
    mean: Optional[torch.Tensor] = None
    invstd: Optional[torch.Tensor] = None
    pack: dict = field(default_factory=dict)


def act_dtype(precision: str):
    if precision == 'fp32':
        return torch.float32
    if precision == 'bf16':
        return torch.bfloat16
    raise ValueError(f"precision must be 'fp32' or 'bf16', got {precision!r}")


def pitch(c: int, precision: str) -> int:
    """Channel pitch of an activation buffer: the tensor-core path wants multiples of 64."""
    if precision == 'bf16' and c > 64 and c % 64:
        return (c + 63) // 64 * 64
    return c


def post_dims(post: str, h: int, w: int):
    if post == 'pool':
        return h // 2, w // 2
    if post == 'up':
        return 2 * h, 2 * w
    return h, w


def default_alloc(name, shape, dtype, device, zero=False):
    return (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=device)


class CachedAlloc:
    """Static buffers for the fused trainer: allocated (and border-zeroed) once, reused every step."""

    def __init__(self, prefix=''):
        self.bufs = {}
        self.prefix = prefix

    def __call__(self, name, shape, dtype, device, zero=False):
        key = (self.prefix, name)
        t = self.bufs.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=device)
            self.bufs[key] = t
        return t

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.bufs.values())


def uses_tc(spec: ConvSpec, cinp: int, precision: str) -> bool:
    return (precision == 'bf16' and cinp % 64 == 0 and spec.cout % 64 == 0
            and (spec.cout % 128 == 0 or cinp % 128 == 0))


def pack_layer(spec: ConvSpec, p: LayerParams, cinp: int, precision: str, alloc, tag, launch: bool = True) -> dict:
    """OIHW fp32 master weights -> the layouts the kernels read (kp_pack_weights).  With launch=False only the
    buffers are created and ``out['desc']`` holds the record for kp_pack_weights_multi."""
    dev = p.w.device
    T = spec.k * spec.k
    out = {}
    if uses_tc(spec, cinp, precision):
        out['tc_f'] = alloc(f'{tag}.tc_f', (T, spec.cout, cinp), torch.bfloat16, dev)
        out['tc_d'] = alloc(f'{tag}.tc_d', (T, cinp, spec.cout), torch.bfloat16, dev)
        out['desc'] = [p.w.data_ptr(), 0, 0, out['tc_f'].data_ptr(), out['tc_d'].data_ptr(), spec.cout, spec.cin, spec.k, cinp]
        if launch:
            L.call('kp_pack_weights', L.stream(), L.ptr(p.w), spec.cout, spec.cin, spec.k, cinp, None, None,
                   L.ptr(out['tc_f']), L.ptr(out['tc_d']))
    else:
        out['simt_f'] = alloc(f'{tag}.simt_f', (T, spec.cin, spec.cout), torch.float32, dev)
        out['simt_d'] = alloc(f'{tag}.simt_d', (T, spec.cout, spec.cin), torch.float32, dev)
        out['desc'] = [p.w.data_ptr(), out['simt_f'].data_ptr(), out['simt_d'].data_ptr(), 0, 0, spec.cout, spec.cin, spec.k,
                       spec.cin]
        if launch:
            L.call('kp_pack_weights', L.stream(), L.ptr(p.w), spec.cout, spec.cin, spec.k, spec.cin,
                   L.ptr(out['simt_f']), L.ptr(out['simt_d']), None, None)
    return out


def fshifts(k, PW):
    return [PW + 1] if k == 1 else [ty * PW + tx for ty in range(3) for tx in range(3)]


def bshifts(k, PW):
    return [0] if k == 1 else [(ty - 1) * PW + (tx - 1) for ty in range(3) for tx in range(3)]


def to_padded(x_nchw: torch.Tensor, precision: str, alloc=default_alloc, name='x0', cp: Optional[int] = None):
    """NCHW fp32 API tensor -> replicate-padded NHWC buffer (kp_bn_act_fwd as layout converter)."""
    n, c, h, w = x_nchw.shape
    cp = cp or pitch(c, precision)
    buf = alloc(name, (n, h + 2, w + 2, cp), act_dtype(precision), x_nchw.device, zero=cp != c)
    L.call('kp_bn_act_fwd', L.stream(), L.nchw(x_nchw), L.view(buf[..., :c]), None, None, L.ACT_NONE, L.POST_NONE, 1,
           n, h, w, c)
    return buf


def unit_forward(specs: List[ConvSpec], params: List[LayerParams], x_pad: torch.Tensor, H: int, W: int,
                 precision: str, out: torch.Tensor, out_pad: int, alloc: Callable = default_alloc,
                 training: bool = True, packs: Optional[List[dict]] = None, tag: str = 'u',
                 coresident: bool = False) -> List[LayerCtx]:
    """Forward of one Unit.  ``x_pad``: [N][H+2][W+2][Cp].  ``out``: tensor indexed [n,y,x,c] receiving the
    activation of the last conv (ptr at the padded origin when out_pad=1).  ``coresident``: another Unit's convolutions
    run concurrently on another stream (the BatchNorm passes then use their small-footprint variants)."""
    co = L.POST_CORESIDENT if coresident else 0
    dev = x_pad.device
    N = x_pad.shape[0]
    T = act_dtype(precision)
    st = L.stream()
    nstat = sum(2 * s.cout for s in specs if s.bn)
    stats_all = alloc(f'{tag}.stats', (max(nstat, 1),), torch.float64, dev)
    if nstat and training:
        L.zero(stats_all)
    soff = 0
    cur, h, w = x_pad, H, W
    ctxs: List[LayerCtx] = []
    for i, (s, p) in enumerate(zip(specs, params)):
        last = i == len(specs) - 1
        PH, PW = h + 2, w + 2
        cinp = cur.shape[3]
        tc = uses_tc(s, cinp, precision)
        pk = packs[i] if packs is not None else pack_layer(s, p, cinp, precision, alloc, f'{tag}.{i}')
        y = alloc(f'{tag}.y{i}', (N, PH, PW, s.cout), T, dev)
        stats = None
        if s.bn and training:
            stats = stats_all[soff:soff + 2 * s.cout]
            soff += 2 * s.cout
        if tc:
            sh = fshifts(s.k, PW)
            L.call('kp_conv_tc', st, L.ptr(cur), N * PH * PW, cinp, L.ptr(pk['tc_f']), len(sh), L.shifts_array(sh),
                   L.ptr(p.b), L.ptr(y), s.cout, L.ptr(stats), PH, PW, h, w,
                   flops=2.0 * N * h * w * s.cin * s.cout * s.k * s.k, tag=f'{tag}{i} {s.cin}->{s.cout}@{h}x{w} k{s.k}')
        else:
            src = cur if s.k == 3 else cur[:, 1:, 1:, :]
            L.call('kp_conv_simt', st, L.view(src), L.ptr(pk['simt_f']), L.ptr(p.b), L.view(y[:, :h, :w, :]),
                   L.ptr(stats), N, h, w, PH if s.k == 3 else h, PW if s.k == 3 else w, s.cin, s.cout, s.k, 0,
                   tag=f'{tag}{i} {s.cin}->{s.cout}@{h}x{w} k{s.k}')
        c = LayerCtx(x=cur, y=y, h=h, w=w, tc=tc, pack=pk)
        oh, ow = post_dims(s.post, h, w)
        if last:
            dst, pad = out, out_pad
        else:
            cp = pitch(s.cout, precision)
            nxt = alloc(f'{tag}.x{i + 1}', (N, oh + 2, ow + 2, cp), T, dev, zero=cp != s.cout)
            dst, pad = nxt[..., :s.cout], 1
        ltag = f'{tag}{i} {s.cin}->{s.cout}@{h}x{w} {s.post}'
        if s.bn:
            c.scale = alloc(f'{tag}.scale{i}', (s.cout,), torch.float32, dev)
            c.shift = alloc(f'{tag}.shift{i}', (s.cout,), torch.float32, dev)
            if training:
                c.mean = alloc(f'{tag}.mean{i}', (s.cout,), torch.float32, dev)
                c.invstd = alloc(f'{tag}.invstd{i}', (s.cout,), torch.float32, dev)
                L.call('kp_bn_finalize_act_fwd', st, L.ptr(stats), float(N * h * w), L.ptr(p.gamma), L.ptr(p.beta),
                       BN_EPS, BN_MOMENTUM, L.ptr(p.rmean), L.ptr(p.rvar), L.ptr(p.nbt), L.ptr(c.scale), L.ptr(c.shift),
                       L.ptr(c.mean), L.ptr(c.invstd), L.view(y[:, :h, :w, :]), L.view(dst), L.ACTS[s.act], L.POSTS[s.post] | co,
                       pad, N, h, w, s.cout, tag=ltag)
            else:   # eval: running statistics (host-side plumbing on [C] vectors)
                inv = torch.rsqrt(p.rvar + BN_EPS)
                c.scale.copy_(p.gamma * inv)
                c.shift.copy_(p.beta - p.rmean * p.gamma * inv)
        if not (s.bn and training):
            L.call('kp_bn_act_fwd', st, L.view(y[:, :h, :w, :]), L.view(dst), L.ptr(c.scale), L.ptr(c.shift), L.ACTS[s.act],
                   L.POSTS[s.post], pad, N, h, w, s.cout, tag=ltag)
        ctxs.append(c)
        if not last:
            cur, h, w = nxt, oh, ow
    return ctxs


def unit_backward(specs: List[ConvSpec], params: List[LayerParams], grads: List[LayerGrads], ctxs: List[LayerCtx],
                  dout: torch.Tensor, dout_pad: int, precision: str, need_dx: bool,
                  alloc: Callable = default_alloc, tag: str = 'u', defer_stg: Optional[List] = None,
                  after_wgrad: Optional[Callable] = None, wgrad_stream=None) -> Optional[torch.Tensor]:
    """Backward of one Unit.  ``dout``: gradient w.r.t. the last activation, indexed [n,y,x,c] (ptr at the
    padded origin when dout_pad=1).  Weight gradients are ACCUMULATED into ``grads`` (zero them per step);
    BatchNorm / bias gradients are overwritten.  ``after_wgrad(i)`` (optional) is called once every gradient of layer i
    (weight, bias, BatchNorm affine) has been issued on the current stream — the fused trainer starts the bucket-wise
    gradient all-reduce from it.  ``wgrad_stream`` (optional, deferred-staging mode only): the weight-gradient kernels are
    issued on that stream, off the dgrad -> BatchNorm-backward chain that the next layer waits for; joined before returning.
    Returns d(padded input) [N][H+2][W+2][Cp] or None."""
    dev = ctxs[0].x.device
    N = ctxs[0].x.shape[0]
    T = act_dtype(precision)
    st = L.stream()
    nsum = sum(2 * s.cout for s in specs)
    sums_all = alloc(f'{tag}.sums', (nsum,), torch.float64, dev)
    L.zero(sums_all)
    soff = 0
    dx = None
    if defer_stg is None or after_wgrad is not None:
        wgrad_stream = None

    def wgrad_call(name, *args, **kw):
        """Launch a weight-gradient kernel (args after the stream) on the wgrad stream if there is one."""
        if wgrad_stream is None:
            return L.call(name, L.stream(), *args, **kw)
        wgrad_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(wgrad_stream):
            L.call(name, L.stream(), *args, **kw)
    for i in range(len(specs) - 1, -1, -1):
        s, p, g, c = specs[i], params[i], grads[i], ctxs[i]
        h, w = c.h, c.w
        PH, PW = h + 2, w + 2
        cinp = c.x.shape[3]
        Q = N * PH * PW
        dyp = alloc(f'{tag}.dy{i}', (N, PH, PW, s.cout), T, dev, zero=True)
        dy_int = dyp[:, 1:h + 1, 1:w + 1, :]
        yv = c.y[:, :h, :w, :]
        sums = sums_all[soff:soff + 2 * s.cout]
        soff += 2 * s.cout
        a, po = L.ACTS[s.act], L.POSTS[s.post]
        if s.bn:
            if c.mean is None:
                raise NotImplementedError('backward through eval-mode BatchNorm is not supported')
            ltag = f'{tag}{i} {s.cin}->{s.cout}@{h}x{w} {s.post}'
            regather = regather_ok(s, h, w, dout, precision)
            L.call('kp_bn_act_bwd_reduce', st, L.view(dout), L.view(yv), None if regather else L.view(dy_int), L.ptr(c.scale),
                   L.ptr(c.shift), L.ptr(c.mean), L.ptr(c.invstd), L.ptr(sums), a, po, dout_pad, N, h, w, s.cout, tag=ltag)
            if regather:     # pass 2 repeats the (cheap) gather instead of reading dz back
                L.call('kp_bn_act_bwd_apply_gather', st, L.view(dout), L.view(yv), L.view(dy_int), L.ptr(c.scale),
                       L.ptr(c.shift), L.ptr(c.mean), L.ptr(c.invstd), L.ptr(sums), float(N * h * w), a, po, dout_pad,
                       N, h, w, s.cout, L.ptr(g.dgamma), L.ptr(g.dbeta), tag=ltag)
            else:
                L.call('kp_bn_act_bwd_apply', st, L.view(yv), L.view(dy_int), L.ptr(c.scale), L.ptr(c.mean), L.ptr(c.invstd),
                       L.ptr(sums), float(N * h * w), N, h, w, s.cout, L.ptr(g.dgamma), L.ptr(g.dbeta), tag=ltag)
            # the bias of a conv feeding train-mode BatchNorm has an exactly zero gradient
        else:
            L.call('kp_bn_act_bwd_reduce', st, L.view(dout), L.view(yv), L.view(dy_int), None, None, None, None,
                   L.ptr(sums), a, po, dout_pad, N, h, w, s.cout)
            if g.db is not None:
                L.call('kp_bn_grad_finalize', st, L.ptr(sums), s.cout, None, L.ptr(g.db))
        want_dx = i > 0 or need_dx
        if c.tc:
            sh = bshifts(s.k, PW)
            if defer_stg is not None:       # fused trainer: per-layer slice of one arena, zeroed and folded once per step
                stg, dw_ptr = defer_stg[i], None
            else:
                stg = alloc(f'{tag}.stg', (max(sp.k * sp.k * sp.cout * cx.x.shape[3] for sp, cx in zip(specs, ctxs) if cx.tc),),
                            torch.float32, dev)
                dw_ptr = L.ptr(g.dw)
            # measured (gpurun_out/calls_s10/s11): the 4-D loads cost more per k-block than the flat 2-D ones, so the image form
            # wins where the flat form wastes many pixels (W <= 32: 13-27 %) or the tiles are wide (>= 256 channels both sides)
            if (WGRAD_IMG and w >= 16 and w & (w - 1) == 0 and (w >= 64 or h % (64 // w) == 0)
                    and (w <= 32 or min(cinp, s.cout) >= 256)):
                # contraction over the valid pixels only (no MMA work on the pad / zero border)
                wgrad_call('kp_conv_wgrad_tc_img', L.ptr(c.x), L.ptr(dyp), N, h, w, s.cin, cinp, s.cout, s.k, L.ptr(stg),
                       dw_ptr, flops=2.0 * N * h * w * s.cin * s.cout * s.k * s.k, tag=f'{tag}{i} {s.cin}->{s.cout}@{h}x{w} k{s.k}')
            else:
                wgrad_call('kp_conv_wgrad_tc', L.ptr(c.x), L.ptr(dyp), Q, s.cin, cinp, s.cout, len(sh), L.shifts_array(sh),
                       L.ptr(stg), dw_ptr, flops=2.0 * N * h * w * s.cin * s.cout * s.k * s.k, tag=f'{tag}{i} {s.cin}->{s.cout}@{h}x{w} k{s.k}')
            if after_wgrad is not None:
                after_wgrad(i)
            if want_dx:
                dx = alloc(f'{tag}.dx{i}', (N, PH, PW, cinp), T, dev)
                L.call('kp_conv_tc', st, L.ptr(dyp), Q, s.cout, L.ptr(c.pack['tc_d']), len(sh), L.shifts_array(sh), None,
                       L.ptr(dx), cinp, None, PH, PW, PH, PW, flops=2.0 * N * h * w * s.cin * s.cout * s.k * s.k, tag=f'{tag}{i} {s.cin}->{s.cout}@{h}x{w} k{s.k}')
        else:
            src = c.x if s.k == 3 else c.x[:, 1:, 1:, :]
            wgrad_call('kp_conv_wgrad_simt', L.view(src), L.view(dy_int), L.ptr(g.dw), N, h, w, s.cin, s.cout, s.k,
                   tag=f'{tag}{i} {s.cin}->{s.cout}@{h}x{w} k{s.k}')
            if after_wgrad is not None:
                after_wgrad(i)
            if want_dx:
                dx = alloc(f'{tag}.dx{i}', (N, PH, PW, cinp), T, dev, zero=s.k == 1)
                if s.k == 3:
                    L.call('kp_conv_simt', st, L.view(dy_int), L.ptr(c.pack['simt_d']), None, L.view(dx[..., :s.cin]),
                           None, N, PH, PW, h, w, s.cout, s.cin, 3, -2, tag=f'{tag}{i} dgrad {s.cin}->{s.cout}@{h}x{w} k{s.k}')
                else:
                    L.call('kp_conv_simt', st, L.view(dy_int), L.ptr(c.pack['simt_d']), None,
                           L.view(dx[:, 1:h + 1, 1:w + 1, :s.cin]), None, N, h, w, h, w, s.cout, s.cin, 1, 0,
                           tag=f'{tag}{i} dgrad {s.cin}->{s.cout}@{h}x{w} k{s.k}')
        if want_dx and i > 0:
            dout, dout_pad = dx[..., :specs[i - 1].cout], 1
    if wgrad_stream is not None:
        torch.cuda.current_stream().wait_stream(wgrad_stream)
    return dx if need_dx else None


def fold_to_nchw(dx_pad: torch.Tensor, C: int, H: int, W: int) -> torch.Tensor:
    """d(padded NHWC) -> d(NCHW fp32): replication_pad2d backward + layout change (API boundary only)."""
    N = dx_pad.shape[0]
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=dx_pad.device)
    sums = torch.zeros(2 * C, dtype=torch.float64, device=dx_pad.device)
    L.call('kp_bn_act_bwd_reduce', L.stream(), L.view(dx_pad[..., :C]), L.view(dx_pad[:, 1:H + 1, 1:W + 1, :C]),
           L.nchw(out), None, None, None, None, L.ptr(sums), L.ACT_NONE, L.POST_NONE, 1, N, H, W, C)
    return out
