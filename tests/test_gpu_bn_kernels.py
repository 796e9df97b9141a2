"""GPU tests of the streaming BatchNorm / activation / pool / upsample kernels (bf16 throughput mode) in isolation.

One conv+BN+act(+post) layer is run forward and backward through the engine at shapes that select the bulk-async
("pipe") kernels (kp_bn_pipe.cuh) and at shapes that select the register-staged ("lean" / generic) ones; the BN passes are
then recomputed in fp32 with torch from the SAME stored raw conv output (bf16) and the SAME per-channel vectors, so the
comparison isolates the kernels under test from the convolution.  Reference semantics: vgg.py:16-39 of the reference
(BatchNorm2d(train) -> LeakyReLU/ReLU -> MaxPool2d(2,2) | UpsamplingBilinear2d(2), then ReplicationPad2d(1) of the
next block), backward by torch.autograd.  Tolerance: bf16 output rounding (2^-8 relative) -> 1e-2 of max |ref|."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from keypoints_b200 import lib
    lib.device_info()
    return torch.device('cuda:0')


def _post(a, post):
    if post == 'pool':
        return F.max_pool2d(a, 2, 2)
    if post == 'up':
        return F.interpolate(a, scale_factor=2, mode='bilinear', align_corners=True)
    return a


def _act(z, act):
    if act == 'leaky':
        return F.leaky_relu(z, 0.01)
    if act == 'relu':
        return F.relu(z)
    return z


CASES = [
    # (cin, cout, h, post, act, n)    pipe-eligible: W*C/8 % 512 == 0 (none), OW*C/8 % 256 == 0 (pool), W*C/8 % 256 == 0 (up)
    (128, 64, 64, 'none', 'leaky', 3),
    (128, 128, 32, 'none', 'relu', 2),
    (128, 64, 64, 'pool', 'leaky', 3),
    (128, 256, 16, 'pool', 'relu', 2),
    (128, 64, 32, 'up', 'relu', 3),
    (128, 256, 16, 'up', 'leaky', 2),
    (128, 512, 16, 'up', 'relu', 2),
    (128, 64, 48, 'up', 'relu', 2),          # 3 row bands of 16
    (128, 64, (24, 32), 'up', 'leaky', 2),   # bands of 16 + 8 rows, H != W
    (128, 64, (20, 64), 'pool', 'relu', 2),  # H != W
    (128, 128, (10, 32), 'none', 'relu', 3),
    # not pipe-eligible (row length): register-staged kernels
    (128, 64, 24, 'none', 'leaky', 2),
    (128, 64, 24, 'pool', 'relu', 2),
    (128, 64, 12, 'up', 'relu', 2),
]


@pytest.mark.parametrize('cin,cout,h,post,act,n', CASES)
def test_bn_act_post_kernels_vs_torch(dev, cin, cout, h, post, act, n):
    from keypoints_b200 import engine
    from keypoints_b200.engine import ConvSpec, LayerGrads, LayerParams
    h, w = h if isinstance(h, tuple) else (h, h)
    torch.manual_seed(h * 1000 + cout)
    spec = ConvSpec(k=3, cin=cin, cout=cout, bn=True, act=act, post=post)
    x = torch.randn(n, cin, h, w, device=dev)
    p = LayerParams(w=torch.randn(cout, cin, 3, 3, device=dev) / (3 * cin ** 0.5), b=torch.randn(cout, device=dev) * 0.1,
                    gamma=torch.rand(cout, device=dev) + 0.5, beta=torch.randn(cout, device=dev) * 0.3,
                    rmean=torch.zeros(cout, device=dev), rvar=torch.ones(cout, device=dev),
                    nbt=torch.zeros((), dtype=torch.int64, device=dev))
    g = LayerGrads(dw=torch.zeros_like(p.w), db=torch.zeros(cout, device=dev), dgamma=torch.zeros(cout, device=dev),
                   dbeta=torch.zeros(cout, device=dev))
    alloc = engine.CachedAlloc('t')
    xp = engine.to_padded(x, 'bf16', alloc, 'x0', cin)
    oh, ow = engine.post_dims(post, h, w)
    out = torch.full((n, oh + 2, ow + 2, cout), float('nan'), device=dev, dtype=torch.bfloat16)
    dout = (torch.randn(n, oh + 2, ow + 2, cout, device=dev) * 0.1).to(torch.bfloat16)
    ctxs = engine.unit_forward([spec], [p], xp, h, w, 'bf16', out, 1, alloc=alloc, tag='f')
    engine.unit_backward([spec], [p], [g], ctxs, dout, 1, 'bf16', True, alloc=alloc, tag='b')
    torch.cuda.synchronize()
    c = ctxs[0]
    # ---- forward reference from the stored raw conv output and the published affine ----
    yraw = c.y[:, :h, :w, :].permute(0, 3, 1, 2).float().requires_grad_(True)           # NCHW fp32 copy of the bf16 y
    sc, sh = c.scale.view(1, -1, 1, 1), c.shift.view(1, -1, 1, 1)
    a = _post(_act(yraw * sc + sh, act), post)
    ref_out = F.pad(a, (1, 1, 1, 1), mode='replicate')
    got = out.permute(0, 3, 1, 2).float()
    assert torch.isfinite(got).all(), 'forward left unwritten (NaN) output pixels'
    e_f = float((got - ref_out).abs().max() / ref_out.abs().max())
    # the affine itself: batch statistics of the raw output
    m = yraw.detach().mean((0, 2, 3))
    v = yraw.detach().var((0, 2, 3), unbiased=False)
    e_s = float((c.scale - p.gamma / torch.sqrt(v + 1e-5)).abs().max() / c.scale.abs().max())
    # ---- backward reference ----
    ref_out.backward(dout.permute(0, 3, 1, 2).float())
    dz = yraw.grad * 1.0 / sc                       # d/d(bn output) = d/dy / scale
    xhat = (yraw.detach() - c.mean.view(1, -1, 1, 1)) * c.invstd.view(1, -1, 1, 1)
    cnt = n * h * w
    s1 = dz.sum((0, 2, 3))
    s2 = (dz * xhat).sum((0, 2, 3))
    dy_ref = sc * (dz - s1.view(1, -1, 1, 1) / cnt - xhat * s2.view(1, -1, 1, 1) / cnt)
    dyp = alloc.bufs[('t', 'b.dy0')]
    dy = dyp[:, 1:h + 1, 1:w + 1, :].permute(0, 3, 1, 2).float()
    e_b = float((dy - dy_ref).abs().max() / dy_ref.abs().max())
    e_beta = float((g.dbeta - s1).abs().max() / s1.abs().max())
    e_gamma = float((g.dgamma - s2).abs().max() / s2.abs().max())
    border = float(dyp[:, 0].abs().max() + dyp[:, -1].abs().max() + dyp[:, :, 0].abs().max() + dyp[:, :, -1].abs().max())
    print(f'{post}/{act} C={cout} {h}x{w}: fwd {e_f:.2e} scale {e_s:.2e} dy {e_b:.2e} dbeta {e_beta:.2e} dgamma {e_gamma:.2e}')
    assert e_f < 1e-2 and e_s < 2e-2 and e_b < 1.5e-2 and e_beta < 5e-3 and e_gamma < 5e-3
    assert border == 0.0, 'dy border must stay zero (wgrad / dgrad sum over all flat pixels)'
