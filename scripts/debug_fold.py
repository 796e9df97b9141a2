import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoints_b200 import trainer as T
from keypoints_b200.models import keynet
dev = torch.device('cuda:0')
torch.manual_seed(3)
x = torch.rand(4, 3, 64, 64, device=dev)
xb = x.flip(0).contiguous()
res = {}
for defer in (False, True):
    T.DEFER_FOLD = defer
    torch.manual_seed(5)
    net = keynet.build('F', 3, 64, 10)
    tr = T.Trainer(net, precision='bf16', use_graph=False)
    tr.step(x, xb)
    torch.cuda.synchronize()
    res[defer] = {(u.name, i): u.grads[i].dw.clone() for u in tr.units.values() for i in range(len(u.specs))}
    specs = {(u.name, i): (u.specs[i].cin, u.specs[i].cout, u.specs[i].k) for u in tr.units.values() for i in range(len(u.specs))}
for k in res[False]:
    a, b = res[False][k], res[True][k]
    e = float((a - b).abs().max() / a.abs().max().clamp_min(1e-20))
    print(k, specs[k], f'rel diff {e:.3e}', f'norm ratio {float(b.norm() / a.norm()):.3f}')
