"""Per-layer micro-benchmark: one conv+BN+act(+pool/up) layer forward and backward at the bench shapes, every C-ABI call
timed with CUDA events (min over repeats), plus a quick numerical check against torch (fp32 on bf16-rounded operands).

  python scripts/bench_layer.py [cin:cout:hw:post:act ...]      e.g. 3:64:128:none:leaky 256:256:64:up:relu
Environment switches of the kernels (KP_*) apply; N via KP_N (default 64).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoints_b200 import engine, lib as L                      # noqa: E402
from keypoints_b200.engine import ConvSpec, LayerGrads, LayerParams   # noqa: E402

dev = torch.device('cuda:0')
N = int(os.environ.get('KP_N', 64))
REPS = int(os.environ.get('KP_REPS', 5))
CHECK = os.environ.get('KP_CHECK', '1') != '0'
DEFAULT = ['3:64:128:none:leaky', '64:128:128:pool:leaky', '128:256:64:none:leaky', '256:256:64:pool:leaky',
           '256:256:64:up:relu', '256:128:128:none:relu', '128:64:128:none:relu', '512:512:16:up:relu', '512:512:16:none:leaky']


def run(spec_s):
    cin, cout, hw, post, act = spec_s.split(':')
    cin, cout, h = int(cin), int(cout), int(hw)
    w = h
    spec = ConvSpec(k=3, cin=cin, cout=cout, bn=True, act=act, post=post)
    torch.manual_seed(0)
    x = torch.randn(N, cin, h, w, device=dev)
    p = LayerParams(w=torch.randn(cout, cin, 3, 3, device=dev) / (3 * cin ** 0.5), b=torch.zeros(cout, device=dev),
                    gamma=torch.rand(cout, device=dev) + 0.5, beta=torch.randn(cout, device=dev) * 0.1,
                    rmean=torch.zeros(cout, device=dev), rvar=torch.ones(cout, device=dev),
                    nbt=torch.zeros((), dtype=torch.int64, device=dev))
    g = LayerGrads(dw=torch.zeros_like(p.w), db=torch.zeros(cout, device=dev), dgamma=torch.zeros(cout, device=dev),
                   dbeta=torch.zeros(cout, device=dev))
    alloc = engine.CachedAlloc('b')
    cp = engine.pitch(cin, 'bf16')
    xp = engine.to_padded(x, 'bf16', alloc, 'x0', cp)
    oh, ow = engine.post_dims(post, h, w)
    out = torch.empty(N, oh + 2, ow + 2, cout, device=dev, dtype=torch.bfloat16)
    dout = (torch.randn(N, oh + 2, ow + 2, cout, device=dev) * 0.1).to(torch.bfloat16)
    pk = [engine.pack_layer(spec, p, cp, 'bf16', alloc, 'pk')]

    def step():
        ctxs = engine.unit_forward([spec], [p], xp, h, w, 'bf16', out, 1, alloc=alloc, packs=pk, tag='f')
        g.dw.zero_()
        dx = engine.unit_backward([spec], [p], [g], ctxs, dout, 1, 'bf16', cin > 3, alloc=alloc, tag='b')
        return ctxs, dx

    for _ in range(2):
        ctxs, dx = step()
    torch.cuda.synchronize()
    best = {}
    for _ in range(REPS):
        L.timing = []
        step()
        torch.cuda.synchronize()
        seen = {}
        for name, fl, a, b, tg in L.timing:
            k = seen.get(name, 0)
            seen[name] = k + 1
            key = f'{name}#{k}' if k else name
            t = a.elapsed_time(b)
            if key not in best or t < best[key][0]:
                best[key] = (t, fl)
        L.timing = None
    elems = N * h * w * cout
    print(f'--- {spec_s}  N={N}  ({elems / 1e6:.0f} M y-elements)')
    for key, (t, fl) in best.items():
        extra = f'{fl / t / 1e9:7.0f} TFLOP/s' if fl else f'{elems * 2 / t / 1e6:7.0f} GB/s per 2B/elt'
        print(f'  {t:8.4f} ms  {key:32s} {extra}')
    if CHECK:
        xb = x.to(torch.bfloat16).float()
        wb = p.w.to(torch.bfloat16).float()
        ref = torch.nn.functional.conv2d(torch.nn.functional.pad(xb, (1, 1, 1, 1), mode='replicate'), wb)
        y = ctxs[0].y[:, :h, :w, :].permute(0, 3, 1, 2).float()
        e = float((y - ref).abs().max() / ref.abs().max())
        # weight gradient: dy (bf16, as stored) x bf16 input
        dyp = alloc.bufs[('b', 'b.dy0')]
        dy = dyp[:, 1:h + 1, 1:w + 1, :].permute(0, 3, 1, 2).float()
        xpad = torch.nn.functional.pad(xb, (1, 1, 1, 1), mode='replicate')
        gw = torch.nn.grad.conv2d_weight(xpad, p.w.shape, dy)
        ew = float((g.dw - gw).abs().max() / gw.abs().max())
        print(f'  check: fprop rel err {e:.2e}   wgrad rel err {ew:.2e}')


if __name__ == '__main__':
    L.device_info()
    for s in (sys.argv[1:] or DEFAULT):
        run(s)
