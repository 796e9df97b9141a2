"""VGG-style layer factory with the reference's signature and Sequential indices (keypoints/models/vgg.py:16-39),
so state_dict keys such as 'core.1.weight' / 'core.2.running_mean' round-trip with reference checkpoints.

The nn.Sequential returned here is only a parameter container + layer description: ``knn.Unit`` walks it,
derives ``engine.ConvSpec`` records and runs them through the CUDA engine (no ATen convolution is called).
Token grammar of a cfg list (vgg.py:23-37): the first entry is the core's INPUT width; an int is
ReplicationPad2d(1)+Conv3x3+BatchNorm+activation, 'M' MaxPool2x2, 'U' bilinear x2 (align_corners=True),
'L' an identity tap.
"""
import torch.nn as nn

from . import knn


def _parse(table):
    return {name: [t if t in ('M', 'U', 'L') else int(t) for t in spec.split()] for name, spec in table.items()}


def _conv_block(cin, cout, batch_norm, act):
    block = [nn.ReplicationPad2d(1), nn.Conv2d(cin, cout, kernel_size=3)]
    if batch_norm:
        block.append(nn.BatchNorm2d(cout))
    block.append(act)
    return block


def make_layers(cfg, batch_norm=True, extra_in_channels=0, nonlinearity=None, nonlinearity_kwargs=None,
                co_ord_conv=False):
    if co_ord_conv:
        raise NotImplementedError('co_ord_conv is outside the hot path (it crashes in the reference too, SURVEY.md 2)')
    act = nn.ReLU(inplace=True) if nonlinearity is None else nonlinearity(**(nonlinearity_kwargs or {}))
    simple = {'M': lambda: nn.MaxPool2d(kernel_size=2, stride=2),
              'U': lambda: nn.UpsamplingBilinear2d(scale_factor=2),
              'L': lambda: knn.ActivationMap()}
    width = cfg[0] + extra_in_channels
    modules = []
    for token in cfg[1:]:
        if token in simple:
            modules.append(simple[token]())
        else:
            modules.extend(_conv_block(width, token, batch_norm, act))   # one shared activation object, as the reference
            width = token
    return nn.Sequential(*modules)


# the reference's layer tables (data, vgg.py:48-70)
decoder_cfg = _parse({
    'A': '512 512 U 256 256 U 256 256 U 128 U 64 U',
    'F': '512 512 U 256 256 U 256 256 U 128 64',
    'VGG_PONG': '32 U 16 U 16',
    'VGG_PONG_TRIVIAL': '16 16',
    'VGG_PONG_LAYERNECK': '32 32 16 16',
    'VGG_PACMAN': '16 32 32 16',
    'VGG_PACMAN_2': '64 U 32 32 16',
})

vgg_cfg = _parse({
    'A': '64 M 128 M 256 256 M 512 512 M 512 512 M',
    'B': '64 64 M 128 128 M 256 256 M 512 512 M 512 512 M',
    'D': '64 64 M 128 128 M 256 256 256 M 512 512 512 M 512 512 512 M',
    'E': '64 64 M 128 128 M 256 256 256 256 M 512 512 512 512 M 512 512 512 512 M',
    'F': '64 128 M 256 256 M 512 512 M 512 512',
    'VGG_PONG': '16 M 16 M 32',
    'VGG_PONG_TRIVIAL': '16 16',
    'VGG_PONG_LAYERNECK': '16 32',
    'VGG_PACMAN': '16 32 32 16',
    'VGG_PACMAN_2': '16 32 32 M 64',
    'MAPPER': '8 8',
})
