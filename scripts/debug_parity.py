"""Debug aid: per-tensor error of the CUDA fp32 path and of the fp32 oracle against the fp64 oracle."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import keypoints_oracle as O
import keypoints_b200
from keypoints_b200.models import keynet, transporter

dev = torch.device('cuda:0')
def rel(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

for name, kind, mt in [('keynet_F', 'keynet', 'F'), ('transporter_F', 'transporter', 'F')]:
    g = dict(np.load(f'tests/golden/{name}.npz'))
    cin, z, K, n, h, w, seed = (int(v) for v in g['meta'])
    ops = O.transporter_ops(mt, cin, z, K) if kind == 'transporter' else O.keynet_ops(mt, cin, z, K)
    ref = {}
    for dt in (torch.float32, torch.float64):
        sd = O.init_state_dict(ops, seed, dtype=dt)
        tr = O.OracleTrainer(kind, mt, cin, z, K, sd)
        a, b = torch.from_numpy(g['a']).to(dt), torch.from_numpy(g['b']).to(dt)
        mask = torch.from_numpy(g['mask']).to(dt) if 'mask' in g else None
        out = tr.forward(a, b)
        for o in out:
            if isinstance(o, torch.Tensor) and o.requires_grad:
                o.retain_grad()
        loss = O.l2_reconstruction_loss(out[0], b, mask); loss.backward()
        ref[dt] = (out, {k: sd[k].grad for k in tr.keys}, loss)
    net = (transporter.make(mt, cin, z, K) if kind == 'transporter' else keynet.build(mt, cin, z, K))
    net.load_state_dict(O.init_state_dict(ops, seed)); net = net.to(dev)
    a, b = torch.from_numpy(g['a']).to(dev), torch.from_numpy(g['b']).to(dev)
    mask = torch.from_numpy(g['mask']).to(dev) if 'mask' in g else None
    res = net(a, b)
    for o in res:
        if isinstance(o, torch.Tensor) and o.requires_grad:
            o.retain_grad()
    loss = ((res[0] - b) ** 2 * mask).mean() if mask is not None else ((res[0] - b) ** 2).mean()
    loss.backward()
    o32, g32, l32 = ref[torch.float32]; o64, g64, l64 = ref[torch.float64]
    print(f'== {name}: loss ours {float(loss):.8f} ref32 {float(l32):.8f} ref64 {float(l64):.8f}')
    names = ['x_hat', 'phi', 'k', 'm', 'p', 'heat', 'mask_s', 'mask_t'] if kind == 'transporter' else ['x_hat', 'z', 'k', 'm', 'p', 'heat']
    for i, nm in enumerate(names):
        if nm == 'p': continue
        line = f'  out {nm:8s} ours-vs-64 {rel(res[i], o64[i]):.2e}   ref32-vs-64 {rel(o32[i], o64[i]):.2e}'
        if res[i].grad is not None and o64[i].grad is not None:
            line += f'   | d/d{nm}: ours {rel(res[i].grad, o64[i].grad):.2e} ref32 {rel(o32[i].grad, o64[i].grad):.2e}'
        print(line)
    params = dict(net.named_parameters())
    for k in g64:
        if g64[k] is None or (k.endswith('.bias') and float(g64[k].abs().max()) < 1e-12): continue
        print(f'  grad {k:32s} ours-vs-64 {rel(params[k].grad, g64[k]):.2e}   ref32-vs-64 {rel(g32[k], g64[k]):.2e}   |g|max {float(g64[k].abs().max()):.2e}')
