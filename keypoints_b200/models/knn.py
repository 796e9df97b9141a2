"""Building blocks with the reference's names (keypoints/models/knn.py): Container, GaussianLike,
SpatialSoftmax / SpatialLogSoftmax, Unit.  ``Unit`` keeps the reference's module tree (in_block / core /
out_block Sequentials -> identical state_dict keys and .mdl checkpoint files) but its forward and backward
run as hand-written CUDA kernels on padded NHWC buffers through ``keypoints_b200.engine``.
"""
from pathlib import Path

import torch
from torch import nn

from .. import config, engine
from ..engine import ConvSpec, LayerGrads, LayerParams
from . import functional as MF


class Container(nn.Module):
    """Initialisation scheme of the reference (knn.py:12-23): conv kaiming-normal(fan_out, relu) with zero
    bias, BatchNorm gamma=1 beta=0, Linear N(0, 0.01)."""

    def _initialize_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, 0.01)
                nn.init.zeros_(m.bias)

    def forward(self, *input):
        raise NotImplementedError()

    def save(self, directory):
        raise NotImplementedError()

    def load(self, directory):
        raise NotImplementedError()


class GaussianLike(nn.Module):
    def __init__(self, sigma=0.1):
        super().__init__()
        self.sigma = sigma

    def forward(self, kp, height, width):
        return MF.gaussian_like_function(kp, height, width, self.sigma)


class SpatialSoftmax(nn.Module):
    def forward(self, heatmap, probs=False):
        return MF.spacial_softmax(heatmap, probs)


class SpatialLogSoftmax(nn.Module):
    def forward(self, heatmap, probs=False):
        return MF.spacial_logsoftmax(heatmap, probs)


class ActivationMap(nn.Module):
    """Identity tap ('L' token of vgg.make_layers, knn.py:79-84)."""

    def forward(self, x):
        return x


class Identity(nn.Module):
    def forward(self, x):
        return x


def _act_name(m):
    if isinstance(m, nn.LeakyReLU):
        if abs(m.negative_slope - 0.01) > 1e-12:
            raise NotImplementedError('only LeakyReLU(0.01) is implemented')
        return 'leaky'
    if isinstance(m, nn.ReLU):
        return 'relu'
    return None


def specs_of(blocks):
    """Walk [(prefix, nn.Sequential)] and return ([ConvSpec], [(conv_module, bn_module|None)])."""
    specs, mods = [], []
    for prefix, seq in blocks:
        for idx, m in enumerate(seq):
            if isinstance(m, nn.Conv2d):
                k = m.kernel_size[0]
                if m.kernel_size not in ((1, 1), (3, 3)) or m.stride != (1, 1) or m.padding != (0, 0) or m.groups != 1:
                    raise NotImplementedError(f'unsupported conv {m}')
                if k == 3 and (idx == 0 or not isinstance(seq[idx - 1], nn.ReplicationPad2d)):
                    raise NotImplementedError('3x3 convs must follow ReplicationPad2d(1) as in the reference')
                specs.append(ConvSpec(k=k, cin=m.in_channels, cout=m.out_channels, bn=False, act='none',
                                      conv_key=f'{prefix}.{idx}'))
                mods.append([m, None])
            elif isinstance(m, nn.BatchNorm2d):
                specs[-1].bn = True
                specs[-1].bn_key = f'{prefix}.{idx}'
                mods[-1][1] = m
            elif _act_name(m) is not None:
                if specs[-1].post != 'none':
                    raise NotImplementedError('activation after pool/upsample')
                specs[-1].act = _act_name(m)
            elif isinstance(m, (nn.MaxPool2d, nn.UpsamplingBilinear2d)):
                if not specs or specs[-1].post != 'none':
                    raise NotImplementedError('two resampling layers in a row')
                specs[-1].post = 'pool' if isinstance(m, nn.MaxPool2d) else 'up'
            elif isinstance(m, (nn.ReplicationPad2d, ActivationMap, Identity)):
                continue
            else:
                raise NotImplementedError(f'unsupported layer {m}')
    return specs, mods


def layer_params(mods):
    out = []
    for conv, bn in mods:
        p = LayerParams(w=conv.weight.data, b=None if conv.bias is None else conv.bias.data)
        if bn is not None:
            p.gamma, p.beta = bn.weight.data, bn.bias.data
            p.rmean, p.rvar, p.nbt = bn.running_mean, bn.running_var, bn.num_batches_tracked
        out.append(p)
    return out


def trainable(mods):
    """Flat list of the trainable tensors in a fixed order: per layer w, b, (gamma, beta)."""
    flat = []
    for conv, bn in mods:
        flat.append(conv.weight)
        if conv.bias is not None:
            flat.append(conv.bias)
        if bn is not None:
            flat += [bn.weight, bn.bias]
    return flat


class _UnitFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, unit, x, *flat):
        specs, mods = unit.layers()
        prec = config.get_precision()
        x = x.float()
        n, c, h, w = x.shape
        oh, ow = h, w
        for s in specs:
            oh, ow = engine.post_dims(s.post, oh, ow)
        out = torch.empty((n, specs[-1].cout, oh, ow), dtype=torch.float32, device=x.device)
        x_pad = engine.to_padded(x, prec)
        ctxs = engine.unit_forward(specs, layer_params(mods), x_pad, h, w, prec, out.permute(0, 2, 3, 1), 0,
                                   training=unit.training)
        if any(ctx.needs_input_grad):
            ctx.state = (unit, specs, mods, ctxs, prec, (c, h, w))
        return out

    @staticmethod
    def backward(ctx, dout):
        unit, specs, mods, ctxs, prec, (c, h, w) = ctx.state
        dev = dout.device
        grads, flat = [], []
        for conv, bn in mods:
            g = LayerGrads(dw=torch.zeros_like(conv.weight), db=None if conv.bias is None else torch.zeros_like(conv.bias))
            flat.append(g.dw)
            if g.db is not None:
                flat.append(g.db)
            if bn is not None:
                g.dgamma, g.dbeta = torch.zeros_like(bn.weight), torch.zeros_like(bn.bias)
                flat += [g.dgamma, g.dbeta]
            grads.append(g)
        dout = dout.float().contiguous()
        dx_pad = engine.unit_backward(specs, layer_params(mods), grads, ctxs, dout.permute(0, 2, 3, 1), 0, prec,
                                      need_dx=ctx.needs_input_grad[1])
        dx = engine.fold_to_nchw(dx_pad, c, h, w) if ctx.needs_input_grad[1] else None
        ctx.state = None
        return (None, dx, *flat)


class Unit(nn.Module):
    """in_block (RepPad+Conv3x3+BN+LeakyReLU) -> core (vgg.make_layers) -> out_block (Conv1x1+LeakyReLU),
    knn.py:110-130."""

    def __init__(self, in_channels, out_channels, core, batch_norm=True):
        super().__init__()
        core_in, core_out = self._core_channels(core)
        head = [nn.ReplicationPad2d(1), nn.Conv2d(in_channels, core_in, kernel_size=3, stride=1)]
        if batch_norm:
            head.append(nn.BatchNorm2d(core_in))
        head.append(nn.LeakyReLU(inplace=True))
        self.in_block = nn.Sequential(*head)
        self.core = core
        self.out_block = nn.Sequential(nn.Conv2d(core_out, out_channels, kernel_size=1, stride=1),
                                       nn.LeakyReLU(inplace=True))
        self._layers = None

    @staticmethod
    def _core_channels(core):
        convs = [m for m in core.modules() if isinstance(m, nn.Conv2d)]
        return convs[0].in_channels, convs[-1].out_channels

    def layers(self):
        if self._layers is None:
            self._layers = specs_of([('in_block', self.in_block), ('core', self.core), ('out_block', self.out_block)])
        return self._layers

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError('keypoints_b200.Unit runs on CUDA (sm_100a) only; there is no CPU path')
        _, mods = self.layers()
        return _UnitFn.apply(self, x, *trainable(mods))

    # checkpoint files: {dir}/{in_block,core,out_block}.mdl, each a Sequential state_dict (knn.py:143-167)
    def _blocks(self):
        return {'in_block': self.in_block, 'core': self.core, 'out_block': self.out_block}

    def save(self, directory):
        for name, block in self._blocks().items():
            path = Path(f'{directory}/{name}.mdl')
            path.parent.mkdir(parents=True, exist_ok=True)
            # detached copies: a parameter that aliases the fused trainer's flat bucket would otherwise drag the whole
            # bucket's storage into every file (torch.save serialises the underlying storage of a view)
            torch.save({k: v.detach().clone() for k, v in block.state_dict().items()}, str(path))

    def load(self, directory, in_block=True, core=True, out_block=True, map_device=None):
        wanted = {'in_block': in_block, 'core': core, 'out_block': out_block}
        for name, block in self._blocks().items():
            if wanted[name]:
                sd = torch.load(str(Path(f'{directory}/{name}.mdl')), map_location=map_device)
                block.load_state_dict(sd)
