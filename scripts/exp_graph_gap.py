"""How long does one dependent kernel node of a CUDA graph take on this GPU when the kernel itself is (almost) empty?
Upper bound of what programmatic dependent launch could recover per kernel boundary of the training step."""
import torch
x = torch.zeros(32, device='cuda')
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3):
        x.add_(1)
torch.cuda.synchronize()
for n in (200, 1000):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            x.add_(1)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f'{n} chained tiny kernels per graph: {e0.elapsed_time(e1) / 10 / n * 1e3:.2f} us per kernel node')
