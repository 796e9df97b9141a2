"""Ragged image sizes through the CUDA path (sizes that are not powers of two, not multiples of 16, odd after pooling):
one train step of the fused trainer in fp32 parity mode against the CPU oracle (1e-3), and the same step on the bf16
tensor-core path (finite, loss within 5 % of the oracle's).  The benchmark shapes only exercise 84, 128 and 256; these
cover the dispatch edges: flat-only weight gradients (W not a power of two), the mma.sync first layer's W % 16 != 0 fallback,
MaxPool floor on odd sizes, BatchNorm row kernels on rows that are no multiple of their chunk."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from keypoints_b200 import lib
    lib.device_info()
    return torch.device('cuda:0')


def _inputs(n, c, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(n, c, max(h // 8, 2), max(w // 8, 2), generator=g)
    x = torch.nn.functional.interpolate(low, size=(h, w), mode='bilinear', align_corners=False)
    x = (x + 0.1 * torch.rand(n, c, h, w, generator=g)).clamp(0, 1)
    return x, x.flip(0).roll(3, 3).contiguous()


# F pools / upsamples three times: the reference itself needs H, W multiples of 8 there (else x_hat and x_ differ in size)
CASES = [('keynet', 'F', 3, 64, 10, 104, 88), ('keynet', 'F', 3, 64, 10, 72, 120), ('transporter', 'F', 3, 64, 6, 40, 56),
         ('transporter', 'VGG_PONG_LAYERNECK', 1, 16, 4, 50, 70), ('transporter', 'VGG_PONG', 1, 16, 4, 84, 84)]


@pytest.mark.parametrize('kind,mt,cin,z,K,H,W', CASES)
def test_ragged_sizes_fp32_vs_oracle_and_bf16_sane(dev, kind, mt, cin, z, K, H, W):
    from oracle import keypoints_oracle as O
    from keypoints_b200.models import keynet, transporter
    from keypoints_b200.trainer import Trainer
    n = 2
    a, b = _inputs(n, cin, H, W, 11 + H)
    ops = O.transporter_ops(mt, cin, z, K) if kind == 'transporter' else O.keynet_ops(mt, cin, z, K)
    sd = O.init_state_dict(ops, 3)
    ref = O.OracleTrainer(kind, mt, cin, z, K, {k: v.clone() for k, v in sd.items()})
    ref_loss, ref_out = ref.step(a, b, None)
    ref_x, ref_k = ref_out[0].detach(), ref_out[2].detach()
    for precision in ('fp32', 'bf16'):
        net = transporter.make(mt, cin, z, K) if kind == 'transporter' else keynet.build(mt, cin, z, K)
        net.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
        tr = Trainer(net, precision=precision, use_graph=False)
        tr.step(a.to(dev), b.to(dev), None)
        k_t, xhat = tr.outputs()
        loss = tr.loss()
        assert torch.isfinite(tr.flat_p).all() and torch.isfinite(tr.flat_g).all()
        ek = float((k_t.cpu() - ref_k).abs().max())
        ex = float((xhat.cpu() - ref_x).abs().max() / ref_x.abs().max())
        el = abs(loss - float(ref_loss)) / abs(float(ref_loss))
        print(f'{kind} {mt} {H}x{W} {precision}: k {ek:.2e}  x_hat {ex:.2e}  loss {el:.2e}')
        if precision == 'fp32':
            assert ek <= 1e-3 and ex <= 1e-3 and el <= 1e-3, (ek, ex, el)
        else:
            assert ek <= 0.1 and el <= 0.05, (ek, ex, el)
        del tr, net
        torch.cuda.empty_cache()
