"""Synthetic inputs shared by make_golden.py and the tests: smooth noise + bright rectangles from a numpy
PCG64 stream.  The full-size fixtures (configs 4 / 5) do not store their inputs — the tests regenerate them here and
check the stored sums."""
import numpy as np
import torch


def synth_images(rng, n, c, h, w, lo=0.0, hi=1.0):
    base = torch.from_numpy(rng.random((n, c, max(h // 8, 2), max(w // 8, 2))).astype('float32'))
    x = torch.nn.functional.interpolate(base, size=(h, w), mode='bilinear', align_corners=False)
    for i in range(n):
        for _ in range(3):
            y0, x0 = int(rng.integers(0, h - 4)), int(rng.integers(0, w - 4))
            hh, ww = int(rng.integers(2, max(h // 4, 3))), int(rng.integers(2, max(w // 4, 3)))
            x[i, :, y0:y0 + hh, x0:x0 + ww] = torch.from_numpy(rng.random(c).astype('float32')).view(c, 1, 1) * 0.5 + 0.5
    return (x * (hi - lo) + lo).contiguous()


def sample(a, limit=30000):
    """Strided sample of a large array (flat, prime stride so it does not alias rows / channels); small arrays in full."""
    a = np.asarray(a)
    flat = a.reshape(-1)
    if flat.size <= limit:
        return flat.copy()
    stride = [s for s in (7, 13, 29, 61, 127, 251, 509, 997) if flat.size // s <= limit]
    return flat[::(stride[0] if stride else 997)].copy()


def sample_like(a, ref_len):
    """The same sample of `a` as the fixture holds (length tells the stride)."""
    flat = np.asarray(a).reshape(-1)
    if flat.size == ref_len:
        return flat
    for s in (7, 13, 29, 61, 127, 251, 509, 997):
        if len(flat[::s]) == ref_len:
            return flat[::s]
    raise ValueError((flat.size, ref_len))


def full_size_inputs(meta):
    cin, z, K, n, h, w, seed = (int(v) for v in meta)
    rng = np.random.default_rng(seed)
    a = synth_images(rng, n, cin, h, w, 0.0, 1.0)
    b = synth_images(rng, n, cin, h, w, 0.0, 1.0)
    mask = (torch.from_numpy(rng.random((n, cin, h, w)).astype('float32')) > 0.2).float()
    return a, b, mask
