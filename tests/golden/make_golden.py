"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (it needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference modules are imported read-only from /root/reference (the model path
imports under torch 2.11 with /root/reference and /root/reference/keypoints on sys.path).  Weights
come from ``oracle.init_state_dict`` (numpy PCG64 stream) and are *loaded into the reference
modules* with load_state_dict(strict=True), so a fixture stores only inputs and the reference's
outputs.  Everything is fp32 on CPU, torch deterministic single-thread.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path[:0] = ['/root/reference', '/root/reference/keypoints']

import warnings  # noqa: E402
warnings.filterwarnings('ignore')

from keypoints.models import transporter as ref_transporter  # noqa: E402
from keypoints.models import knn as ref_knn, vgg as ref_vgg, keynet as ref_keynet  # noqa: E402
import keypoints.models as _ref_models  # noqa: E402
# models/autoencoder.py:1 does `from keypoints.models import Container`, a name the package __init__ never exports (the
# reference's autoencode.py cannot start either); supply it from knn, where the class lives, without touching the file
_ref_models.Container = ref_knn.Container
from keypoints.models import autoencoder as ref_autoencoder  # noqa: E402
from keypoints.models import functional as RF  # noqa: E402
import tps as ref_tps  # noqa: E402
import data_augments as ref_aug  # noqa: E402

from oracle import keypoints_oracle as O  # noqa: E402

torch.set_num_threads(1)
torch.use_deterministic_algorithms(True)


def npy(t):
    return t.detach().cpu().numpy()


sys.path.insert(0, HERE)
from synth import synth_images, sample  # noqa: E402  (shared with the tests: the full-size fixtures regenerate their inputs)


def grad_summary(out, name, g):
    """Full tensor when small, otherwise sum / abs-sum / strided sample."""
    g = npy(g).astype('float32')
    if g.size <= 40000:
        out[f'grad/{name}'] = g
    else:
        flat = g.reshape(-1)
        out[f'gradsum/{name}'] = np.array([flat.astype('float64').sum(), np.abs(flat).astype('float64').sum()])
        out[f'gradsample/{name}'] = flat[::997].copy()


def known_answers():
    out = {}
    for v in (1.0, 5.0, 100.0):                                   # tests/tests.py:11-17 style literal
        hm = torch.zeros(1, 1, 5, 5); hm[0, 0, 2, 2] = v
        out[f'peak5_center_{int(v)}/soft'] = npy(RF.spacial_softmax(hm))
        out[f'peak5_center_{int(v)}/logsoft'] = npy(RF.spacial_logsoftmax(hm))
    hm = torch.zeros(1, 1, 5, 5); hm[0, 0, 4, 4] = 5.0             # tests/tests.py:19-25
    out['peak5_corner/soft'] = npy(RF.spacial_softmax(hm))
    out['peak5_corner/logsoft'] = npy(RF.spacial_logsoftmax(hm))
    hm = torch.zeros(1, 1, 16, 16); hm[0, 0, 0, 15] = 20.0         # tests/tests.py:239-240
    k = RF.spacial_logsoftmax(hm)
    out['coords16/k'] = npy(k)
    out['coords16/g'] = npy(RF.gaussian_like_function(k, 16, 16))
    c = torch.tensor([[0., 0], [1., 0], [1., 1], [0, 1]]).unsqueeze(0)   # tps.py:197-212
    out['tps_identity/grid'] = npy(ref_tps.tps_grid(torch.zeros(1, 7, 2), c, (1, 1, 6, 3)))
    np.savez_compressed(os.path.join(HERE, 'known_answers.npz'), **out)


def functional_fixture():
    rng = np.random.default_rng(11)
    out = {}
    heat = torch.from_numpy((rng.standard_normal((2, 3, 7, 9)) * 3).astype('float32')).requires_grad_(True)
    k, (ph, pw) = RF.spacial_logsoftmax(heat, probs=True)
    gk = torch.from_numpy(rng.standard_normal((2, 3, 2)).astype('float32'))
    (k * gk).sum().backward()
    out.update({'ssm/heat': npy(heat), 'ssm/k': npy(k), 'ssm/ph': npy(ph), 'ssm/pw': npy(pw), 'ssm/gk': npy(gk),
                'ssm/dheat': npy(heat.grad)})
    k2, (ph2, pw2) = RF.spacial_softmax(heat.detach(), probs=True)
    out.update({'ssm/k_soft': npy(k2), 'ssm/ph_soft': npy(ph2)})
    kp = torch.from_numpy(rng.random((2, 3, 2)).astype('float32')).requires_grad_(True)
    m = RF.gaussian_like_function(kp, 6, 11)
    gm = torch.from_numpy(rng.standard_normal((2, 3, 6, 11)).astype('float32'))
    (m * gm).sum().backward()
    out.update({'gauss/kp': npy(kp), 'gauss/m': npy(m), 'gauss/gm': npy(gm), 'gauss/dkp': npy(kp.grad)})
    np.savez_compressed(os.path.join(HERE, 'functional.npz'), **out)


def tps_fixture():
    rng = np.random.default_rng(12)
    out = {}
    x = synth_images(rng, 3, 3, 20, 14)
    theta = torch.from_numpy((rng.standard_normal((3, 7, 2)) * 0.05).astype('float32'))
    ctrl = torch.from_numpy(rng.random((3, 4, 2)).astype('float32'))
    rot = torch.from_numpy(((rng.random(3) * 2 - 1) * 0.1).astype('float32'))
    out.update({'x': npy(x), 'theta': npy(theta), 'ctrl': npy(ctrl), 'rot': npy(rot)})
    out['grid'] = npy(ref_tps.tps_grid(theta, ctrl, tuple(x.shape)))
    out['tps'] = npy(ref_tps.tps_transform(x, theta, ctrl))
    out['rot_out'] = npy(ref_tps.rotate_affine_grid_multi(x, rot))
    theta_r = torch.from_numpy((rng.standard_normal((3, 6, 2)) * 0.05).astype('float32'))    # reduced form
    out['theta_reduced'] = npy(theta_r)
    out['grid_reduced'] = npy(ref_tps.tps_grid(theta_r, ctrl, tuple(x.shape)))
    # the full augmentation with the reference's own RNG consumption order
    torch.manual_seed(77)
    x1, x2, mask = ref_aug.TpsAndRotate(4, 0.05, 0.1)(x, x)
    out.update({'aug/x1': npy(x1), 'aug/x2': npy(x2), 'aug/mask': npy(mask)})
    np.savez_compressed(os.path.join(HERE, 'tps.npz'), **out)


def build_ref_keynet(model_type, cin, z, K):
    """keynet.make is broken in the reference; assemble exactly what keynet.py:51-59 intends."""
    nn = torch.nn
    enc = ref_knn.Unit(cin, z, ref_vgg.make_layers(ref_vgg.vgg_cfg[model_type], nonlinearity=nn.LeakyReLU,
                                                   nonlinearity_kwargs={'inplace': True}))
    dec = ref_knn.Unit(z + K, cin, ref_vgg.make_layers(ref_vgg.decoder_cfg[model_type]))
    kp = ref_knn.Unit(cin, K, ref_vgg.make_layers(ref_vgg.vgg_cfg[model_type], nonlinearity=nn.LeakyReLU,
                                                  nonlinearity_kwargs={'inplace': True}))
    return ref_keynet.KeyNet(enc, kp, ref_knn.GaussianLike(sigma=0.1), dec, init_weights=True)


def build_ref_autoencoder(model_type, cin, z):
    """autoencode.py:59-66."""
    nn = torch.nn
    kw = dict(nonlinearity=nn.LeakyReLU, nonlinearity_kwargs={'inplace': True})
    enc = ref_knn.Unit(cin, z, ref_vgg.make_layers(ref_vgg.vgg_cfg[model_type], **kw))
    dec = ref_knn.Unit(z, cin, ref_vgg.make_layers(ref_vgg.decoder_cfg[model_type], **kw))
    return ref_autoencoder.AutoEncoder(enc, dec, init_weights=True)


def autoencoder_fixture(name, model_type, cin, z, n, h, w, seed, lo, hi):
    """One auto-encoder pre-training step (autoencode.py:84-96) + the checkpoint the reference writes, and what
    TransporterNet.load_from_autoencoder (models/transporter.py:71-75) makes of it."""
    import shutil
    rng = np.random.default_rng(seed)
    net = build_ref_autoencoder(model_type, cin, z)
    ops = O.autoencoder_ops(model_type, cin, z)
    sd = O.init_state_dict(ops, seed)
    net.load_state_dict(sd, strict=True)
    x = synth_images(rng, n, cin, h, w, lo, hi)
    out = {'a': npy(x), 'meta': np.array([cin, z, 0, n, h, w, seed])}
    optim = torch.optim.Adam(net.parameters(), lr=1e-4)
    optim.zero_grad()
    zz, xh = net(x)
    loss = torch.nn.MSELoss()(xh, x)
    loss.backward()
    out['out/z'], out['out/x_hat'], out['loss'] = npy(zz), npy(xh), npy(loss)
    for pname, p in net.named_parameters():
        grad_summary(out, pname, p.grad)
    new_sd = net.state_dict()
    for key in new_sd:
        if 'running_' in key or 'num_batches' in key:
            out[f'stat/{key}'] = npy(new_sd[key])
    optim.step()
    for pname, p in net.named_parameters():
        out[f'adam/{pname}'] = npy(p)
    # checkpoint written by the reference itself (9 -> 6 .mdl files) and the transfer into a Transporter
    ck = os.path.join(HERE, f'ckpt_{name}')
    shutil.rmtree(ck, ignore_errors=True)
    net.save(ck)
    K = 3
    tnet = ref_transporter.make(model_type, cin, z, K, transfer_load=ck)
    for key, v in tnet.state_dict().items():
        unit, block = key.split('.')[0], key.split('.')[1]
        transferred = (unit == 'feature' and block != 'out_block') or (unit == 'keypoint' and block != 'out_block') or \
                      (unit == 'decoder' and block != 'in_block')
        if transferred:
            out[f'transfer/{key}'] = npy(v)
    np.savez_compressed(os.path.join(HERE, f'{name}.npz'), **out)
    print(name, 'loss', float(loss), 'bytes', os.path.getsize(os.path.join(HERE, f'{name}.npz')))


def eval_fixture(name, model_type, cin, z, K, h, w, seed):
    """The demo / inference path (atari_demo.py:20-36): a trained-for-one-step Transporter in eval() mode, key-points of a
    single frame from `net.keypoint` + the spatial soft-max, and the full eval forward.  BatchNorm uses running statistics."""
    rng = np.random.default_rng(seed)
    net = ref_transporter.make(model_type, cin, z, K)
    ops = O.transporter_ops(model_type, cin, z, K)
    net.load_state_dict(O.init_state_dict(ops, seed), strict=True)
    a = synth_images(rng, 3, cin, h, w, -1.0, 1.0)
    b = synth_images(rng, 3, cin, h, w, -1.0, 1.0)
    net.train()
    for _ in range(3):                       # move the running statistics away from (0, 1)
        net(a, b)
    net.eval()
    out = {'meta': np.array([cin, z, K, 1, h, w, seed])}
    for key, v in net.state_dict().items():
        out[f'state/{key}'] = npy(v)
    s_t = synth_images(rng, 1, cin, h, w, -1.0, 1.0)
    s_u = synth_images(rng, 1, cin, h, w, -1.0, 1.0)
    with torch.no_grad():
        heat = net.keypoint(s_t)
        k = RF.spacial_logsoftmax(heat)
        res = net(s_t, s_u)
    out.update({'s_t': npy(s_t), 's_u': npy(s_u), 'eval/heat': npy(heat), 'eval/k': npy(k), 'eval/x_hat': npy(res[0]),
                'eval/k_full': npy(res[2])})
    np.savez_compressed(os.path.join(HERE, f'{name}.npz'), **out)
    print(name, 'bytes', os.path.getsize(os.path.join(HERE, f'{name}.npz')))


def model_fixture(name, kind, model_type, cin, z, K, n, h, w, seed, lo, hi, with_mask, adam=False, combine_mode='max'):
    rng = np.random.default_rng(seed)
    if kind == 'transporter':
        net = ref_transporter.make(model_type, cin, z, K, combine_mode=combine_mode)
        ops = O.transporter_ops(model_type, cin, z, K)
    else:
        net = build_ref_keynet(model_type, cin, z, K)
        ops = O.keynet_ops(model_type, cin, z, K)
    sd = O.init_state_dict(ops, seed)
    net.load_state_dict(sd, strict=True)
    a = synth_images(rng, n, cin, h, w, lo, hi)
    b = synth_images(rng, n, cin, h, w, lo, hi)
    mask = None
    if with_mask:
        mask = (torch.from_numpy(rng.random((n, cin, h, w)).astype('float32')) > 0.2).float()
    out = {'a': npy(a), 'b': npy(b), 'meta': np.array([cin, z, K, n, h, w, seed])}
    if mask is not None:
        out['mask'] = npy(mask)
    optim = torch.optim.Adam(net.parameters(), lr=1e-4)
    optim.zero_grad()
    res = net(a, b)
    loss = ((res[0] - b) ** 2 * mask).mean() if mask is not None else ((res[0] - b) ** 2).mean()
    loss.backward()
    names = ['x_hat', 'phi', 'k', 'm', 'p', 'heat', 'mask_s', 'mask_t'] if kind == 'transporter' else \
            ['x_hat', 'z', 'k', 'm', 'p', 'heat']
    for nm, r in zip(names, res):
        if nm == 'p':
            out['out/p_h'], out['out/p_w'] = npy(r[0]), npy(r[1])
        else:
            out[f'out/{nm}'] = npy(r)
    out['loss'] = npy(loss)
    for pname, p in net.named_parameters():
        grad_summary(out, pname, p.grad)
    new_sd = net.state_dict()
    for key in new_sd:
        if 'running_' in key or 'num_batches' in key:
            out[f'stat/{key}'] = npy(new_sd[key])
    if adam:
        optim.step()
        for pname, p in net.named_parameters():
            out[f'adam/{pname}'] = npy(p)
    np.savez_compressed(os.path.join(HERE, f'{name}.npz'), **out)
    print(name, 'loss', float(loss), 'bytes', os.path.getsize(os.path.join(HERE, f'{name}.npz')))


def full_size_fixture(name, kind, model_type, cin, z, K, n, h, w, seed):
    """BASELINE configs 4 / 5 at their real shapes (Transporter F 128x128 K=30; KeyNet F 256x256 K=64), batch 2, fp32.
    Inputs and weights regenerate from the seed (tests/golden/synth.py, oracle.init_state_dict), so the fixture holds
    only the reference's outputs: small tensors in full, large ones as a strided sample (synth.sample)."""
    torch.set_num_threads(8)                      # fp32 CPU convs; thread count does not change ATen's conv arithmetic order
    rng = np.random.default_rng(seed)
    if kind == 'transporter':
        net = ref_transporter.make(model_type, cin, z, K, combine_mode='max')
        ops = O.transporter_ops(model_type, cin, z, K)
    else:
        net = build_ref_keynet(model_type, cin, z, K)
        ops = O.keynet_ops(model_type, cin, z, K)
    net.load_state_dict(O.init_state_dict(ops, seed), strict=True)
    a = synth_images(rng, n, cin, h, w, 0.0, 1.0)
    b = synth_images(rng, n, cin, h, w, 0.0, 1.0)
    mask = (torch.from_numpy(rng.random((n, cin, h, w)).astype('float32')) > 0.2).float()
    out = {'meta': np.array([cin, z, K, n, h, w, seed]),
           'in/check': np.array([float(a.double().sum()), float(b.double().sum()), float(mask.double().sum())])}
    optim = torch.optim.Adam(net.parameters(), lr=1e-4)
    optim.zero_grad()
    res = net(a, b)
    loss = ((res[0] - b) ** 2 * mask).mean()
    loss.backward()
    names = ['x_hat', 'phi', 'k', 'm', 'p', 'heat', 'mask_s', 'mask_t'] if kind == 'transporter' else \
            ['x_hat', 'z', 'k', 'm', 'p', 'heat']
    for nm, r in zip(names, res):
        if nm == 'p':
            out['out/p_h'], out['out/p_w'] = npy(r[0]), npy(r[1])
        elif nm == 'k':
            out['out/k'] = npy(r)
        else:
            out[f'outsample/{nm}'] = sample(npy(r))
            out[f'outmax/{nm}'] = np.array(float(r.detach().abs().max()))
    out['loss'] = npy(loss)
    for pname, p in net.named_parameters():
        grad_summary(out, pname, p.grad)
        out[f'gradmax/{pname}'] = np.array(float(p.grad.abs().max()))
    new_sd = net.state_dict()
    for key in new_sd:
        if 'running_' in key or 'num_batches' in key:
            out[f'stat/{key}'] = npy(new_sd[key])
    np.savez_compressed(os.path.join(HERE, f'{name}.npz'), **out)
    print(name, 'loss', float(loss), 'bytes', os.path.getsize(os.path.join(HERE, f'{name}.npz')))
    torch.set_num_threads(1)


if __name__ == '__main__':
    only = sys.argv[1:]
    if 'full' in only:           # `make_golden.py full`: only the full-size config 4 / 5 fixtures
        full_size_fixture('transporter_F_128_K30', 'transporter', 'F', 3, 64, 30, 2, 128, 128, 109)
        full_size_fixture('keynet_F_256_K64', 'keynet', 'F', 3, 64, 64, 2, 256, 256, 110)
        print('done')
        sys.exit(0)          # e.g. `make_golden.py modes` regenerates only the SURVEY 8f fixtures
    if only:
        known_answers = functional_fixture = tps_fixture = lambda: None
        _mf = model_fixture
        model_fixture = lambda name, *a, **k: _mf(name, *a, **k) if ('loop' in name or 'sum' in name) else None
    known_answers()
    functional_fixture()
    tps_fixture()
    # BASELINE config 1 verbatim: Transporter Pong-grey 84x84 K=4 batch=2 (range [-1,1], datasets.py:292-295)
    model_fixture('transporter_pong', 'transporter', 'VGG_PONG_LAYERNECK', 1, 16, 4, 2, 84, 84, 101, -1.0, 1.0,
                  with_mask=False, adam=True)
    # the F stacks (configs 3-5) at reduced spatial size; weights regenerate from the seed
    model_fixture('keynet_F', 'keynet', 'F', 3, 64, 10, 2, 32, 32, 102, 0.0, 1.0, with_mask=True)
    model_fixture('transporter_F', 'transporter', 'F', 3, 64, 6, 2, 32, 32, 103, 0.0, 1.0, with_mask=True)
    # a pooled + upsampled small net (exercises 'M' right after in_block and 'U' chains)
    model_fixture('keynet_pong_mu', 'keynet', 'VGG_PONG', 1, 8, 3, 3, 24, 16, 104, -1.0, 1.0, with_mask=False,
                  adam=True)
    # the non-default combine modes and the auto-encoder pre-training / transfer path
    if not only or 'modes' in only:
        model_fixture('transporter_pong_loop', 'transporter', 'VGG_PONG_LAYERNECK', 1, 16, 4, 2, 36, 28, 105, -1.0, 1.0,
                      with_mask=False, combine_mode='loop')
        model_fixture('transporter_pong_sum', 'transporter', 'VGG_PONG_LAYERNECK', 1, 16, 4, 2, 36, 28, 106, -1.0, 1.0,
                      with_mask=False, combine_mode='sum_and_clamp')
        autoencoder_fixture('autoencoder_pong', 'VGG_PONG', 1, 8, 3, 24, 16, 107, -1.0, 1.0)
        eval_fixture('transporter_pong_eval', 'VGG_PONG', 1, 8, 3, 32, 24, 108)
    if not only:
        full_size_fixture('transporter_F_128_K30', 'transporter', 'F', 3, 64, 30, 2, 128, 128, 109)
        full_size_fixture('keynet_F_256_K64', 'keynet', 'F', 3, 64, 64, 2, 256, 256, 110)
    print('done')
