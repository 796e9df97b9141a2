"""GPU parity of the SURVEY 8f rows built so far, through the reference's module API in fp32 (1e-3 bar):
the 'loop' and 'sum_and_clamp' combine modes of TransporterNet.forward (models/transporter.py:41-50) and the
auto-encoder pre-training step (models/autoencoder.py, autoencode.py:84-96), against fixtures produced by the
unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from test_gpu_parity import Report, TOL, bn_sibling, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from keypoints_b200 import lib
    lib.device_info()
    return torch.device('cuda:0')


def _compare(R, g, net, res, names, loss, gtol=2e-3):
    for nm, r in zip(names, res):
        if nm == 'p':
            R.close(r[0], g['out/p_h'], TOL, 'p_h'); R.close(r[1], g['out/p_w'], TOL, 'p_w')
        else:
            R.close(r, g[f'out/{nm}'], TOL, nm)
    R.close(loss, g['loss'], TOL, 'loss')
    grads = dict(net.named_parameters())
    for key in g:
        if key.startswith('grad/'):
            sib = bn_sibling(key, g)
            if sib is not None:
                R.check(float(grads[key[5:]].grad.abs().max()) <= 1e-3 * np.abs(g[sib]).max(), key)
            else:
                R.close(grads[key[5:]].grad, g[key], gtol, key)
                R.rows.append((key + ' L2', rel_l2(grads[key[5:]].grad, g[key]), 1e-3))
        elif key.startswith('stat/'):
            R.close(net.state_dict()[key[5:]], g[key], TOL, key)


@pytest.mark.parametrize('name,mode', [('transporter_pong_loop', 'loop'), ('transporter_pong_sum', 'sum_and_clamp')])
def test_combine_modes_vs_reference_golden(dev, golden, name, mode):
    import keypoints_b200
    from oracle import keypoints_oracle as O
    from keypoints_b200.models import transporter
    keypoints_b200.set_precision('fp32')
    g = golden(name)
    cin, z, K, n, h, w, seed = (int(v) for v in g['meta'])
    net = transporter.make('VGG_PONG_LAYERNECK', cin, z, K, combine_mode=mode)
    net.load_state_dict(O.init_state_dict(O.transporter_ops('VGG_PONG_LAYERNECK', cin, z, K), seed), strict=True)
    net = net.to(dev)
    a, b = torch.from_numpy(g['a']).to(dev), torch.from_numpy(g['b']).to(dev)
    res = net(a, b)
    loss = ((res[0] - b) ** 2).mean()
    loss.backward()
    R = Report()
    _compare(R, g, net, res, ['x_hat', 'phi', 'k', 'm', 'p', 'heat', 'mask_s', 'mask_t'], loss)
    R.finish()


def test_autoencoder_vs_reference_golden(dev, golden):
    import keypoints_b200
    from oracle import keypoints_oracle as O
    from keypoints_b200.models import autoencoder
    keypoints_b200.set_precision('fp32')
    g = golden('autoencoder_pong')
    cin, z, _, n, h, w, seed = (int(v) for v in g['meta'])
    net = autoencoder.make('VGG_PONG', cin, z)
    net.load_state_dict(O.init_state_dict(O.autoencoder_ops('VGG_PONG', cin, z), seed), strict=True)
    net = net.to(dev)
    x = torch.from_numpy(g['a']).to(dev)
    optim = torch.optim.Adam(net.parameters(), lr=1e-4)
    optim.zero_grad()
    zz, xh = net(x)
    loss = torch.nn.MSELoss()(xh, x)
    loss.backward()
    R = Report()
    _compare(R, g, net, (xh, zz), ['x_hat', 'z'], loss)
    optim.step()
    sd = net.state_dict()
    for key in g:
        if key.startswith('adam/') and bn_sibling('grad/' + key[5:], g) is None:
            e = np.abs(sd[key[5:]].cpu().numpy().astype(np.float64) - g[key]).max()
            R.rows.append((key, e, 0.05 * 1e-4))
    R.finish()


def test_trainer_resume_matches_uninterrupted_run(dev, tmp_path):
    """`Trainer.save` writes the reference's 9 .mdl files plus the Adam state; a fresh trainer that loads
    them continues exactly where the first one stopped (3 steps == 2 steps + save/load + 1 step, up to the fp32-atomic
    noise of one step)."""
    import os
    from keypoints_b200.models import transporter
    from keypoints_b200.trainer import Trainer
    torch.manual_seed(2)
    xa = torch.rand(4, 1, 32, 32, device=dev) * 2 - 1
    xb = torch.rand(4, 1, 32, 32, device=dev) * 2 - 1

    def fresh():
        torch.manual_seed(9)
        return Trainer(transporter.make('VGG_PONG', 1, 8, 3), precision='fp32', use_graph=False)

    ref = fresh()
    for _ in range(3):
        ref.step(xa, xb)
    a = fresh()
    for _ in range(2):
        a.step(xa, xb)
    a.save(str(tmp_path / 'ck'))
    files = sorted(os.listdir(tmp_path / 'ck'))
    assert files == ['decoder', 'encoder', 'keypoint', 'trainer.pt']
    b = Trainer(transporter.make('VGG_PONG', 1, 8, 3), precision='fp32', use_graph=False)
    b.load(str(tmp_path / 'ck'))
    assert int(b.step_dev.item()) == 2
    assert torch.equal(b.flat_p, a.flat_p) and torch.equal(b.flat_m, a.flat_m) and torch.equal(b.flat_v, a.flat_v)
    b.step(xa, xb)
    d = float((b.flat_p - ref.flat_p).abs().max())
    assert d <= 2e-5, d                                      # a fifth of one lr-sized (1e-4) Adam step
    # running statistics travel with the module files
    for (n1, t1), (n2, t2) in zip(ref.net.named_buffers(), b.net.named_buffers()):
        assert n1 == n2 and torch.allclose(t1.float(), t2.float(), rtol=1e-4, atol=1e-5), n1


def test_eval_mode_inference_vs_reference_golden(dev, golden):
    """The demo path (atari_demo.py:35-36): `net.eval()`, `net.keypoint(s_t)` + spatial soft-max at batch 1 and the full
    eval forward, BatchNorm on the running statistics the reference accumulated; fp32, 1e-3 bar."""
    import keypoints_b200
    from keypoints_b200.models import functional as MF, transporter
    keypoints_b200.set_precision('fp32')
    g = golden('transporter_pong_eval')
    cin, z, K, n, h, w, seed = (int(v) for v in g['meta'])
    net = transporter.make('VGG_PONG', cin, z, K)
    net.load_state_dict({k[6:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('state/')}, strict=True)
    net = net.to(dev).eval()
    s_t, s_u = torch.from_numpy(g['s_t']).to(dev), torch.from_numpy(g['s_u']).to(dev)
    R = Report()
    with torch.no_grad():
        heat = net.keypoint(s_t)
        k = MF.spacial_logsoftmax(heat)
        res = net(s_t, s_u)
    R.close(heat, g['eval/heat'], TOL, 'heat'); R.close(k, g['eval/k'], TOL, 'k')
    R.close(res[0], g['eval/x_hat'], TOL, 'x_hat'); R.close(res[2], g['eval/k_full'], TOL, 'k_full')
    R.rows.append(('k max-abs', float((k.cpu() - torch.from_numpy(g['eval/k'])).abs().max()), TOL))
    # running statistics must not move in eval mode
    for key, v in g.items():
        if key.startswith('state/') and 'running' in key:
            R.close(net.state_dict()[key[6:]], v, 1e-7, key)
    R.finish()
