"""Pins the CPU oracle (oracle/keypoints_oracle.py) to outputs of the reference itself
(tests/golden/*.npz, produced by tests/golden/make_golden.py) and to the known answers derived from
the reference tests' literals.  fp32 tolerance: 1e-5 relative-to-max
(oracle and reference issue the same ATen ops, so most entries are bit-identical)."""
import numpy as np
import pytest
import torch

from oracle import keypoints_oracle as O

torch.set_num_threads(1)      # the fixtures were made single-threaded; same reduction order -> same bits


def bn_sibling(key, keys):
    """For 'grad/<unit>.<block>.<i>.bias' return the key of the BatchNorm bias that follows the conv
    (index i+1) if there is one.  The bias of a conv feeding train-mode BN has a mathematically zero
    gradient (BN subtracts the batch mean); what the reference stores there is rounding noise, so it
    is compared on the scale of the BN bias gradient instead of its own."""
    if not key.endswith('.bias'):
        return None
    head, idx, _ = key.rsplit('.', 2)
    sib = f'{head}.{int(idx) + 1}.bias'
    return sib if sib in keys else None


def close(a, b, tol=1e-5, name='', scale=None):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    scale = max(np.abs(b).max(), 1e-12) if scale is None else scale
    err = np.abs(a - b).max() / scale
    assert err <= tol, f'{name}: rel-to-max err {err:.3e} > {tol}'


def test_known_answers(golden):
    g = golden('known_answers')
    for v in (1, 5, 100):
        hm = torch.zeros(1, 1, 5, 5); hm[0, 0, 2, 2] = float(v)
        close(O.spatial_logsoftmax(hm)[0], g[f'peak5_center_{v}/logsoft'])
        close(O.spatial_softmax(hm)[0], g[f'peak5_center_{v}/soft'])
        close(O.spatial_logsoftmax(hm)[0], np.full((1, 1, 2), 0.5))
    hm = torch.zeros(1, 1, 5, 5); hm[0, 0, 4, 4] = 5.0
    close(O.spatial_logsoftmax(hm)[0], g['peak5_corner/logsoft'])
    close(O.spatial_logsoftmax(hm)[0], np.full((1, 1, 2), 0.627881), tol=1e-5)
    hm = torch.zeros(1, 1, 16, 16); hm[0, 0, 0, 15] = 20.0
    k = O.spatial_logsoftmax(hm)[0]
    close(k, g['coords16/k'])
    close(k, np.array([[[0.432658, 0.567342]]]), tol=1e-5)
    gm = O.gaussian_like(k, 16, 16)
    close(gm, g['coords16/g'])
    assert abs(float(gm.max()) - 0.099278) < 1e-5 and int(gm.argmax()) == 6 * 16 + 9
    c = torch.tensor([[0., 0], [1., 0], [1., 1], [0, 1]]).unsqueeze(0)
    grid = O.tps_grid(torch.zeros(1, 7, 2), c, (1, 1, 6, 3))
    close(grid, g['tps_identity/grid'])
    close(grid[0, :, :, 0], np.tile(np.array([-1., 0., 1.]), (6, 1)), tol=1e-6)
    close(grid[0, :, :, 1], np.tile(np.linspace(-1, 1, 6)[:, None], (1, 3)), tol=1e-6)


def test_functional(golden):
    g = golden('functional')
    heat = torch.from_numpy(g['ssm/heat']).requires_grad_(True)
    k, (ph, pw) = O.spatial_logsoftmax(heat)
    close(k, g['ssm/k']); close(ph, g['ssm/ph']); close(pw, g['ssm/pw'])
    (k * torch.from_numpy(g['ssm/gk'])).sum().backward()
    close(heat.grad, g['ssm/dheat'])
    k2, (ph2, _) = O.spatial_softmax(heat.detach())
    close(k2, g['ssm/k_soft']); close(ph2, g['ssm/ph_soft'])
    kp = torch.from_numpy(g['gauss/kp']).requires_grad_(True)
    m = O.gaussian_like(kp, 6, 11)
    close(m, g['gauss/m'])
    (m * torch.from_numpy(g['gauss/gm'])).sum().backward()
    close(kp.grad, g['gauss/dkp'])


def test_tps(golden):
    g = golden('tps')
    x, theta, ctrl, rot = (torch.from_numpy(g[k]) for k in ('x', 'theta', 'ctrl', 'rot'))
    close(O.tps_grid(theta, ctrl, tuple(x.shape)), g['grid'], tol=2e-6)
    close(O.tps_grid(torch.from_numpy(g['theta_reduced']), ctrl, tuple(x.shape)), g['grid_reduced'], tol=2e-6)
    close(O.tps_transform(x, theta, ctrl), g['tps'], tol=2e-5)
    close(O.rotate_affine_grid_multi(x, rot), g['rot_out'])
    torch.manual_seed(77)                       # same RNG consumption order as data_augments.py:31,35
    p1 = O.sample_perturb_params(3, 4, 0.05, 0.1)
    p2 = O.sample_perturb_params(3, 4, 0.05, 0.1)
    x1, x2, mask = O.tps_and_rotate(x, p1, p2)
    close(x1, g['aug/x1'], tol=5e-5); close(x2, g['aug/x2'], tol=5e-5); close(mask, g['aug/mask'], tol=5e-5)


MODELS = [('transporter_pong', 'transporter', 'VGG_PONG_LAYERNECK'), ('keynet_F', 'keynet', 'F'),
          ('transporter_F', 'transporter', 'F'), ('keynet_pong_mu', 'keynet', 'VGG_PONG')]


@pytest.mark.parametrize('name,kind,model_type', MODELS)
def test_model_step(golden, name, kind, model_type):
    g = golden(name)
    cin, z, K, n, h, w, seed = (int(v) for v in g['meta'])
    ops = O.transporter_ops(model_type, cin, z, K) if kind == 'transporter' else O.keynet_ops(model_type, cin, z, K)
    sd = O.init_state_dict(ops, seed)
    tr = O.OracleTrainer(kind, model_type, cin, z, K, sd)
    a, b = torch.from_numpy(g['a']), torch.from_numpy(g['b'])
    mask = torch.from_numpy(g['mask']) if 'mask' in g else None
    loss, out = tr.step(a, b, mask)
    names = ['x_hat', 'phi', 'k', 'm', 'p', 'heat', 'mask_s', 'mask_t'] if kind == 'transporter' else \
            ['x_hat', 'z', 'k', 'm', 'p', 'heat']
    tol = 2e-5
    for nm, r in zip(names, out):
        if nm == 'p':
            close(r[0], g['out/p_h'], tol, 'p_h'); close(r[1], g['out/p_w'], tol, 'p_w')
        else:
            close(r, g[f'out/{nm}'], tol, nm)
    close(loss, g['loss'], tol, 'loss')
    for key in g:
        if key.startswith('grad/'):
            sib = bn_sibling(key, g)
            close(sd[key[5:]].grad, g[key], 2e-4, key, scale=None if sib is None else np.abs(g[sib]).max())
        elif key.startswith('gradsample/'):
            close(sd[key[11:]].grad.reshape(-1)[::997], g[key], 2e-4, key)
        elif key.startswith('stat/'):
            close(sd[key[5:]], g[key], tol, key)
        elif key.startswith('adam/'):
            if bn_sibling('grad/' + key[5:], g) is not None:
                # Adam normalises the rounding-noise gradient of a pre-BN conv bias to a +-lr step:
                # the reference's value is noise-determined, only |step| <= lr is meaningful
                close(sd[key[5:]], g[key], 1.0, key, scale=2.001e-4)
            else:
                close(sd[key[5:]], g[key], 0.02, key, scale=1e-4)      # within 2 % of one lr-sized step


FULL = [('transporter_F_128_K30', 'transporter', 'F'), ('keynet_F_256_K64', 'keynet', 'F')]


@pytest.mark.parametrize('name,kind,model_type', FULL)
def test_model_step_full_size(golden, name, kind, model_type):
    """BASELINE configs 4 / 5 at their real shapes (Transporter F 128x128 K=30, KeyNet F 256x256 K=64; batch 2): the oracle
    against the reference's outputs; inputs and weights regenerate from the fixture's seed."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    from synth import full_size_inputs, sample_like
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    try:
        g = golden(name)
        cin, z, K, n, h, w, seed = (int(v) for v in g['meta'])
        a, b, mask = full_size_inputs(g['meta'])
        close(np.array([float(a.double().sum()), float(b.double().sum()), float(mask.double().sum())]), g['in/check'], 1e-9, 'inputs')
        ops = O.transporter_ops(model_type, cin, z, K) if kind == 'transporter' else O.keynet_ops(model_type, cin, z, K)
        sd = O.init_state_dict(ops, seed)
        tr = O.OracleTrainer(kind, model_type, cin, z, K, sd)
        loss, out = tr.step(a, b, mask)
        names = ['x_hat', 'phi', 'k', 'm', 'p', 'heat', 'mask_s', 'mask_t'] if kind == 'transporter' else \
                ['x_hat', 'z', 'k', 'm', 'p', 'heat']
        tol = 5e-5        # multi-threaded fp32 convolutions may reduce in another order than the fixture's run
        for nm, r in zip(names, out):
            if nm == 'p':
                close(r[0], g['out/p_h'], tol, 'p_h'); close(r[1], g['out/p_w'], tol, 'p_w')
            elif nm == 'k':
                close(r, g['out/k'], tol, 'k')
            else:
                ref = g[f'outsample/{nm}']
                close(sample_like(r.detach().numpy(), len(ref)), ref, tol, nm, scale=float(g[f'outmax/{nm}']))
        close(loss, g['loss'], tol, 'loss')
        for key in g:
            if key.startswith('grad/'):
                sib = bn_sibling(key, g)
                close(sd[key[5:]].grad, g[key], 5e-3, key, scale=float(g['gradmax/' + (sib or key)[5:]]))
            elif key.startswith('gradsample/'):
                close(sd[key[11:]].grad.reshape(-1)[::997], g[key], 5e-3, key, scale=float(g['gradmax/' + key[11:]]))
            elif key.startswith('stat/'):
                close(sd[key[5:]], g[key], tol, key)
    finally:
        torch.set_num_threads(1)


# ---- SURVEY 8f widening: non-default combine modes, auto-encoder pre-training, checkpoint transfer ----------------
def _check_step(g, sd, loss, outs, names, tol=2e-5):
    for nm, r in zip(names, outs):
        if nm == 'p':
            close(r[0], g['out/p_h'], tol, 'p_h'); close(r[1], g['out/p_w'], tol, 'p_w')
        else:
            close(r, g[f'out/{nm}'], tol, nm)
    close(loss, g['loss'], tol, 'loss')
    for key in g:
        if key.startswith('grad/'):
            sib = bn_sibling(key, g)
            close(sd[key[5:]].grad, g[key], 2e-4, key, scale=None if sib is None else np.abs(g[sib]).max())
        elif key.startswith('stat/'):
            close(sd[key[5:]], g[key], tol, key)


@pytest.mark.parametrize('name,mode', [('transporter_pong_loop', 'loop'), ('transporter_pong_sum', 'sum_and_clamp')])
def test_transporter_combine_modes(golden, name, mode):
    """models/transporter.py:41-50 ('loop', 'sum_and_clamp') against the reference's own forward/backward."""
    g = golden(name)
    cin, z, K, n, h, w, seed = (int(v) for v in g['meta'])
    sd = O.init_state_dict(O.transporter_ops('VGG_PONG_LAYERNECK', cin, z, K), seed)
    tr = O.OracleTrainer('transporter', 'VGG_PONG_LAYERNECK', cin, z, K, sd, mode=mode)
    loss, out = tr.step(torch.from_numpy(g['a']), torch.from_numpy(g['b']))
    _check_step(g, sd, loss, out, ['x_hat', 'phi', 'k', 'm', 'p', 'heat', 'mask_s', 'mask_t'])


def test_autoencoder_step(golden):
    """autoencode.py:84-96 (MSE of the reconstruction against the input) against the reference."""
    g = golden('autoencoder_pong')
    cin, z, _, n, h, w, seed = (int(v) for v in g['meta'])
    sd = O.init_state_dict(O.autoencoder_ops('VGG_PONG', cin, z), seed)
    tr = O.OracleTrainer('autoencoder', 'VGG_PONG', cin, z, 0, sd)
    x = torch.from_numpy(g['a'])
    loss, out = tr.step(x, x)
    _check_step(g, sd, loss, out, ['x_hat', 'z'])


def test_reference_checkpoint_interop(golden, tmp_path):
    """The .mdl files written by the reference's AutoEncoder.save (tests/golden/ckpt_autoencoder_pong, made by
    make_golden.py) load into keypoints_b200's modules, transfer into a Transporter exactly as
    models/transporter.py:71-75 does, and a save/load round trip through our modules preserves them."""
    import os
    from conftest import GOLDEN
    from keypoints_b200.models import autoencoder, transporter
    g = golden('autoencoder_pong')
    cin, z, _, n, h, w, seed = (int(v) for v in g['meta'])
    ck = os.path.join(GOLDEN, 'ckpt_autoencoder_pong')
    net = autoencoder.make('VGG_PONG', cin, z, load=ck)
    sd = net.state_dict()
    for key in g:
        if key.startswith('adam/'):                       # the checkpoint was written after the Adam step
            close(sd[key[5:]], g[key], 1e-7, key)
    t = transporter.make('VGG_PONG', cin, z, 3, transfer_load=ck)
    tsd = t.state_dict()
    moved = [k for k in g if k.startswith('transfer/')]
    assert len(moved) > 20
    for key in moved:
        close(tsd[key[9:]].float(), g[key], 1e-7, key)
    net.save(str(tmp_path / 'ours'))
    ref_files = sorted(os.path.relpath(os.path.join(d, f), ck) for d, _, fs in os.walk(ck) for f in fs)
    our_files = sorted(os.path.relpath(os.path.join(d, f), tmp_path / 'ours') for d, _, fs in os.walk(tmp_path / 'ours') for f in fs)
    assert ref_files == our_files
    for f in ref_files:
        a, b = torch.load(os.path.join(ck, f)), torch.load(str(tmp_path / 'ours' / f))
        assert list(a.keys()) == list(b.keys())
        for k in a:
            assert torch.equal(a[k], b[k]), (f, k)


def test_eval_mode_inference_path(golden):
    """atari_demo.py:20-36: eval() forward (BatchNorm on running statistics), key-points of one frame."""
    g = golden('transporter_pong_eval')
    cin, z, K, n, h, w, seed = (int(v) for v in g['meta'])
    ops = O.transporter_ops('VGG_PONG', cin, z, K)
    sd = {k[6:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('state/')}
    s_t, s_u = torch.from_numpy(g['s_t']), torch.from_numpy(g['s_u'])
    with torch.no_grad():
        heat = O.unit_forward(s_t, sd, 'keypoint.', ops['keypoint'], False)
        k, _ = O.spatial_logsoftmax(heat)
        res = O.transporter_forward(s_t, s_u, sd, ops, training=False)
    close(heat, g['eval/heat'], 2e-5, 'heat'); close(k, g['eval/k'], 2e-5, 'k')
    close(res[0], g['eval/x_hat'], 2e-5, 'x_hat'); close(res[2], g['eval/k_full'], 2e-5, 'k_full')
