"""Experiment: time the 64->128 @128x128 fprop (resident-weights tcgen05 kernel) in isolation."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoints_b200 import engine, lib as L
from keypoints_b200.engine import ConvSpec, LayerParams
dev = torch.device('cuda:0')
n, cin, cout, h, w = 64, int(os.environ.get('CIN', 64)), int(os.environ.get('COUT', 128)), 128, 128
spec = ConvSpec(k=3, cin=cin, cout=cout, bn=False, act='none')
x = torch.randn(n, cin, h, w, device=dev)
p = LayerParams(w=torch.randn(cout, cin, 3, 3, device=dev) / 24, b=None)
xp = engine.to_padded(x, 'bf16')
alloc = engine.CachedAlloc('e')
out = torch.empty(n, h, w, cout, device=dev, dtype=torch.bfloat16)
pk = [engine.pack_layer(spec, p, cin, 'bf16', alloc, 'pk')]
for _ in range(3):
    engine.unit_forward([spec], [p], xp, h, w, 'bf16', out, 0, alloc=alloc, packs=pk)
torch.cuda.synchronize()
L.timing = []
for _ in range(5):
    engine.unit_forward([spec], [p], xp, h, w, 'bf16', out, 0, alloc=alloc, packs=pk)
torch.cuda.synchronize()
ts = [a.elapsed_time(b) for name, fl, a, b, tg in L.timing if name == 'kp_conv_tc']
fl = 2.0 * n * h * w * cin * cout * 9
print(f'dbg={os.environ.get("KP_TC_DBG","0")} resident={os.environ.get("KP_TC_RESIDENT","1")} pair={os.environ.get("KP_TC_PAIR","1")} conv_tc ms: {min(ts):.4f}  -> {fl / min(ts) / 1e9:.0f} TFLOP/s')
