"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol declared in
include/keypoints_b200.h (no compute call is made: there is no GPU here and no CPU fallback)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def built():
    subprocess.run(['make', '-C', os.path.join(ROOT, 'keypoints_b200', 'csrc'), '-j', '8'], check=True,
                   stdout=subprocess.DEVNULL)
    from keypoints_b200 import lib
    return lib


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'keypoints_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(kp_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported(built):
    handle = built.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(handle, n), f'{n} declared in include/keypoints_b200.h but not exported'
    assert sorted(built.EXPORTS) == names, 'lib.py binding table and the header disagree'
    assert handle.kp_version() == 100


def test_sass_is_blackwell_native(built):
    """The tensor-core kernels must be tcgen05 / TMA code, not a legacy mma.sync path."""
    sass = subprocess.run(['cuobjdump', '-sass', built.LIB_PATH], capture_output=True, text=True).stdout
    assert 'UTCHMMA' in sass and 'UTMALDG' in sass and 'LDTM' in sass
    # per kernel: warp-level HMMA only in the first-layer (Cin <= 3, K = 27) kernels and the 16/32-channel Pong layers, where
    # a tcgen05 k-block / tile would be 50-95 % zero padding (kp_conv_thin_mma.cu, kp_conv_small_mma.cu); the conv / wgrad families are tcgen05 + TMA, the BatchNorm streaming
    # kernels use the bulk-copy engine (UBLKCP) with mbarrier transaction counts (SYNCS)
    funcs = {}
    name = None
    for line in sass.splitlines():
        if 'Function :' in line:
            name = line.split('Function :')[1].strip()
            funcs[name] = []
        elif name is not None:
            funcs[name].append(line)
    body = {k: '\n'.join(v) for k, v in funcs.items()}
    for k, text in body.items():
        if 'HMMA' in text.replace('UTCHMMA', ''):
            assert 'thin_mma' in k or 'small_mma' in k or 'head_mma' in k, f'legacy mma.sync code in {k}'
    conv = [k for k in body if 'conv_tc_pair' in k or 'wgrad_tc_pair_k' in k]
    assert conv and all('UTCHMMA' in body[k] and 'UTMALDG' in body[k] for k in conv)
    pipe = [k for k in body if '_pipe_k' in k and 'small_mma' not in k]
    small = [k for k in body if '_pipe_k' in k and 'small_mma' in k]       # 16/32-channel layers: cp.async ring + ldmatrix
    assert len(small) >= 8 and all('LDGSTS' in body[k] and 'LDSM' in body[k] and 'HMMA' in body[k] for k in small)
    assert len(pipe) >= 10 and all('UBLKCP' in body[k] and 'SYNCS' in body[k] for k in pipe)
    assert any('FFMA2' in body[k] for k in pipe), 'packed fp32x2 math expected in the streaming kernels'


def test_no_cpu_fallback(built):
    import torch
    from keypoints_b200.models import transporter
    net = transporter.make('VGG_PONG_LAYERNECK', 1, 16, 4)
    with pytest.raises(RuntimeError):
        net(torch.zeros(2, 1, 16, 16), torch.zeros(2, 1, 16, 16))
    with pytest.raises(built.KpError):
        built.ptr(torch.zeros(1))


def test_state_dict_keys_match_reference_layout(built):
    from oracle import keypoints_oracle as O
    from keypoints_b200.models import keynet, transporter
    net = transporter.make('F', 3, 64, 10)
    ref = O.init_state_dict(O.transporter_ops('F', 3, 64, 10), 0)
    assert set(net.state_dict().keys()) == set(ref.keys())
    net.load_state_dict(ref, strict=True)
    assert sum(p.numel() for p in net.parameters()) == 24468813
    kn = keynet.build('VGG_PONG', 1, 8, 3)
    assert set(kn.state_dict().keys()) == set(O.init_state_dict(O.keynet_ops('VGG_PONG', 1, 8, 3), 0).keys())


def test_checkpoint_roundtrip(built, tmp_path):
    import torch
    from keypoints_b200.models import transporter
    a = transporter.make('VGG_PONG_LAYERNECK', 1, 16, 4)
    a.save(str(tmp_path / 'ck'))
    files = sorted(str(p.relative_to(tmp_path / 'ck')) for p in (tmp_path / 'ck').rglob('*.mdl'))
    assert files == sorted(f'{u}/{b}.mdl' for u in ('encoder', 'keypoint', 'decoder') for b in ('in_block', 'core', 'out_block'))
    b = transporter.make('VGG_PONG_LAYERNECK', 1, 16, 4, load=str(tmp_path / 'ck'))
    for (k1, v1), (k2, v2) in zip(a.state_dict().items(), b.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)


def test_dropin_aliases_reference_import_paths(built):
    """after dropin.install() the reference scripts' own import statements resolve to this package."""
    import importlib
    import sys
    saved = {k: sys.modules.get(k) for k in ('keypoints', 'keypoints.models', 'tps', 'data_augments', 'apex', 'apex.amp')}
    try:
        from keypoints_b200 import dropin
        dropin.install()
        from keypoints.models import transporter, keynet, knn, vgg, functional      # noqa: F401  (transporter.py:11-12)
        from tps import tps_transform, rotate_affine_grid_multi, tps_sample_params    # noqa: F401
        from data_augments import TpsAndRotate, nop                                   # noqa: F401  (transporter.py:6)
        from apex import amp
        net = transporter.make('VGG_PONG_LAYERNECK', 1, 16, 4)
        assert hasattr(net, 'feature') and hasattr(net, 'keypoint') and hasattr(net, 'decoder') and hasattr(net, 'ssm')
        m, o = amp.initialize(net, None, opt_level='O0')
        assert m is net
        with amp.scale_loss(1.0, None) as sl:
            assert sl == 1.0
        assert nop(1, 2) == (1, 2, None)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in [k for k in sys.modules if k.startswith('keypoints.models.')]:
            sys.modules.pop(k, None)
        import keypoints_b200
        keypoints_b200.set_precision('fp32')
