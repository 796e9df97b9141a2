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


# ---- round 2: the fused Trainer covers the auto-encoder and every combine mode; run services -----------------------
@pytest.mark.parametrize('use_graph', [False, True])
def test_fused_trainer_autoencoder_vs_reference_golden(dev, golden, use_graph):
    """autoencode.py:84-96 as one fused step (fp32 parity mode): loss, reconstruction, gradients and the Adam update
    against the fixture written by the unmodified reference."""
    from oracle import keypoints_oracle as O
    from keypoints_b200.models import autoencoder
    from keypoints_b200.trainer import Trainer
    g = golden('autoencoder_pong')
    cin, z, _, n, h, w, seed = (int(v) for v in g['meta'])
    net = autoencoder.make('VGG_PONG', cin, z)
    net.load_state_dict(O.init_state_dict(O.autoencoder_ops('VGG_PONG', cin, z), seed), strict=True)
    tr = Trainer(net, precision='fp32', use_graph=use_graph)
    tr.step(torch.from_numpy(g['a']).to(dev))
    R = Report()
    R.close(torch.tensor(tr.loss()), g['loss'], TOL, 'loss')
    R.close(tr.outputs()[1], g['out/x_hat'], TOL, 'x_hat')
    grads = tr.named_grads()
    for key in g:
        if key.startswith('grad/') and bn_sibling(key, g) is None:
            R.close(grads[key[5:]], g[key], 2e-3, key)
    sd = net.state_dict()
    for key in g:
        if key.startswith('adam/') and bn_sibling('grad/' + key[5:], g) is None:
            d = np.abs(sd[key[5:]].cpu().numpy().astype(np.float64) - g[key])
            gk = 'grad/' + key[5:]
            if gk in g:
                d = d[np.abs(g[gk]) >= 1e-6]
            R.rows.append((key, d.max() if d.size else 0.0, 0.05 * 1e-4))
        elif key.startswith('stat/'):
            R.close(sd[key[5:]], g[key], TOL, key)
    R.finish()


@pytest.mark.parametrize('name,mode', [('transporter_pong_loop', 'loop'), ('transporter_pong_sum', 'sum_and_clamp')])
def test_fused_trainer_combine_modes_vs_reference_golden(dev, golden, name, mode):
    """models/transporter.py:41-50 inside the fused step (the reference constructor's default is 'loop')."""
    from oracle import keypoints_oracle as O
    from keypoints_b200.models import transporter
    from keypoints_b200.trainer import Trainer
    g = golden(name)
    cin, z, K, n, h, w, seed = (int(v) for v in g['meta'])
    net = transporter.make('VGG_PONG_LAYERNECK', cin, z, K, combine_mode=mode)
    net.load_state_dict(O.init_state_dict(O.transporter_ops('VGG_PONG_LAYERNECK', cin, z, K), seed), strict=True)
    tr = Trainer(net, precision='fp32', use_graph=False)
    tr.step(torch.from_numpy(g['a']).to(dev), torch.from_numpy(g['b']).to(dev))
    k_t, xhat = tr.outputs()
    R = Report()
    R.close(torch.tensor(tr.loss()), g['loss'], TOL, 'loss')
    R.close(k_t, g['out/k'], TOL, 'k')
    R.close(xhat, g['out/x_hat'], TOL, 'x_hat')
    R.close(tr.misc.bufs[('misc', 'mask_s')], g['out/mask_s'], TOL, 'mask_s')
    R.close(tr.misc.bufs[('misc', 'mask_t')], g['out/mask_t'], TOL, 'mask_t')
    grads = tr.named_grads()
    for key in g:
        if key.startswith('grad/') and bn_sibling(key, g) is None:
            R.close(grads[key[5:]], g[key], 2e-3, key)
    R.finish()


def test_checkpoint_files_hold_only_their_own_block(dev, tmp_path):
    """parameters alias the trainer's flat bucket; every .mdl must still be about its block's own size."""
    import os
    from keypoints_b200.models import transporter
    from keypoints_b200.trainer import Trainer
    net = transporter.make('VGG_PONG', 1, 8, 3)
    tr = Trainer(net, precision='fp32', use_graph=False)
    tr.save(str(tmp_path / 'ck'))
    for unit_dir, unit in (('encoder', net.feature), ('keypoint', net.keypoint), ('decoder', net.decoder)):
        for block_name, block in unit._blocks().items():
            own = sum(v.numel() * v.element_size() for v in block.state_dict().values())
            size = os.path.getsize(tmp_path / 'ck' / unit_dir / f'{block_name}.mdl')
            assert size <= own + 16384, (unit_dir, block_name, size, own)
    assert tr.n_params * 4 > 4 * 16384      # the bucket is much larger than any slack allowed above


def test_aug_draw_distribution_and_replay(dev):
    """kp_aug_draw: moments of the drawn TpsAndRotate parameters (tps.py:122-125, data_augments.py:6-10), different
    values for different steps / seeds, identical values for the same (seed, step)."""
    from keypoints_b200 import lib as L
    n, T = 4096, 4
    theta = torch.empty(2, n, T + 3, 2, device=dev)
    ctrl = torch.empty(2, n, T, 2, device=dev)
    rot = torch.empty(2, n, device=dev)
    step = torch.zeros(1, dtype=torch.int32, device=dev)

    def draw(seed):
        L.call('kp_aug_draw', L.stream(), seed, L.ptr(step), 2, n, T, 0.05, 0.1, L.ptr(theta), L.ptr(ctrl), L.ptr(rot))
        torch.cuda.synchronize()
        return theta.clone(), ctrl.clone(), rot.clone()

    t0, c0, r0 = draw(11)
    assert abs(float(t0.mean())) < 2e-3 and abs(float(t0.std()) - 0.05) < 1.5e-3
    assert abs(float((t0 / 0.05).pow(4).mean()) - 3.0) < 0.2                      # Gaussian kurtosis
    assert 0.0 <= float(c0.min()) and float(c0.max()) < 1.0 and abs(float(c0.mean()) - 0.5) < 0.01
    assert abs(float(c0.var()) - 1 / 12) < 3e-3
    assert float(r0.abs().max()) <= 0.1 and abs(float(r0.mean())) < 5e-3 and abs(float(r0.var()) - 0.01 / 3) < 3e-4
    assert not torch.equal(t0[0], t0[1])                                          # the two perturbations of a step differ
    t1, c1, r1 = draw(11)
    assert torch.equal(t0, t1) and torch.equal(c0, c1) and torch.equal(r0, r1)
    step.fill_(1)
    t2, _, _ = draw(11)
    assert not torch.equal(t0, t2)
    step.fill_(0)
    t3, _, _ = draw(12)
    assert not torch.equal(t0, t3)
    # constant-one source == warping a tensor of ones
    x1 = torch.ones(3, 2, 20, 24, device=dev)
    a, b = torch.empty_like(x1), torch.empty_like(x1)
    L.call('kp_tps_warp', L.stream(), L.ptr(x1), L.ptr(a), L.ptr(t0[0]), L.ptr(c0[0]), 3, 2, 20, 24, T, 0)
    L.call('kp_tps_warp', L.stream(), None, L.ptr(b), L.ptr(t0[0]), L.ptr(c0[0]), 3, 2, 20, 24, T, 0)
    assert torch.allclose(a, b, atol=1e-6)


def test_loss_ring_and_plateau_scheduler(dev):
    """the non-blocking logger.  Losses read from the device ring equal the per-step `loss()` reads; the
    plateau scheduler lowers lr like torch's ReduceLROnPlateau fed the same values."""
    from keypoints_b200 import runlog
    from keypoints_b200.models import transporter
    from keypoints_b200.trainer import Trainer
    torch.manual_seed(4)
    xa = torch.rand(4, 1, 32, 32, device=dev) * 2 - 1
    xb = torch.rand(4, 1, 32, 32, device=dev) * 2 - 1
    tr = Trainer(transporter.make('VGG_PONG', 1, 8, 3), precision='fp32', use_graph=True)
    log = runlog.RunLog(tr, slots=16, every=4)
    direct = []
    for _ in range(21):
        tr.step(xa, xb)
        direct.append(tr.loss())                 # the blocking read the ring replaces (here only to compare)
        log.after_step()
    log.flush()
    steps = [s for s, _ in log.history]
    assert steps == list(range(21)), steps
    assert np.allclose([v for _, v in log.history], direct, rtol=1e-12)
    # scheduler parity with torch on a synthetic plateau
    class _T:
        lr = 1e-3
    t = _T()
    mine = runlog.PlateauLR(t, factor=0.5, patience=2)
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=1e-3)
    ref = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, factor=0.5, patience=2)
    for v in [1.0, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.8, 0.8, 0.8, 0.8, 0.8]:
        mine.step(v); ref.step(v)
        assert abs(t.lr - opt.param_groups[0]['lr']) < 1e-12, (v, t.lr, opt.param_groups[0]['lr'])
    # lr assignment drops the graph and the next step honours it
    tr.lr = 5e-5
    assert tr.graph is None
    tr.step(xa, xb)
    assert tr.graph is not None


def test_ctx_tensor_map_cache_and_sm_limit(dev):
    """kp_ctx: with a current context the tensor-core launchers take their TMA tensor maps from the cache on the
    second call (same buffers) and produce bit-identical results; the SM limit shrinks the persistent grids without
    changing the result beyond the order of the split-K atomics."""
    from keypoints_b200 import engine, lib as L
    from keypoints_b200.engine import ConvSpec, LayerParams
    torch.manual_seed(0)
    n, cin, cout, h = 2, 64, 128, 32
    spec = ConvSpec(k=3, cin=cin, cout=cout, bn=False, act='none')
    p = LayerParams(w=(torch.randn(cout, cin, 3, 3, device=dev) / 24), b=torch.zeros(cout, device=dev))
    x = torch.randn(n, cin, h, h, device=dev)
    alloc = engine.CachedAlloc('ctxtest')
    ctx = L.Context()
    try:
        L.load().kp_ctx_set_current(None)
        xp = engine.to_padded(x, 'bf16', alloc, 'x')
        out0 = torch.empty(n, h, h, cout, device=dev)
        engine.unit_forward([spec], [p], xp, h, h, 'bf16', out0, 0, alloc=alloc)          # no context: encode per call
        ctx.use()
        outs = []
        for _ in range(2):
            o = torch.empty(n, h, h, cout, device=dev)
            engine.unit_forward([spec], [p], xp, h, h, 'bf16', o, 0, alloc=alloc)
            outs.append(o)
        info = ctx.info()
        assert info['map_misses'] > 0 and info['map_hits'] >= info['map_misses'], info
        assert torch.equal(outs[0], out0) and torch.equal(outs[1], out0)
        ctx.set_sm_limit(64)
        assert ctx.info()['sm_limit'] == 64
        o = torch.empty(n, h, h, cout, device=dev)
        engine.unit_forward([spec], [p], xp, h, h, 'bf16', o, 0, alloc=alloc)
        assert torch.equal(o, out0)                                                        # fprop has no atomics: identical
        ctx.set_sm_limit(0)
    finally:
        ctx.close()
