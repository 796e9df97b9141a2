"""KeyNet (Jakab et al. 2018) with the reference's API (keypoints/models/keynet.py:7-66)."""
import torch
from torch import nn

from . import knn, vgg


class KeyNet(knn.Container):
    def __init__(self, encoder, keypoint, key2map, decoder, init_weights=True):
        super().__init__()
        self.encoder = encoder
        self.keypoint = keypoint
        self.ssm = knn.SpatialLogSoftmax()
        self.key2map = key2map
        self.decoder = decoder
        if init_weights:
            self._initialize_weights()

    def forward(self, x, x_t):
        """-> (x_hat, z, k, m, (p_h, p_w), heatmap), keynet.py:21-31."""
        z = self.encoder(x)
        heatmap = self.keypoint(x_t)
        k, p = self.ssm(heatmap, probs=True)
        m = self.key2map(k, height=z.size(2), width=z.size(3))
        x_hat = self.decoder(torch.cat((z, m), dim=1))
        return x_hat, z, k, m, p, heatmap

    def _units(self):
        return {'encoder': self.encoder, 'keypoint': self.keypoint, 'decoder': self.decoder}

    def save(self, directory):
        for name, unit in self._units().items():
            unit.save(f'{directory}/{name}')

    def load(self, directory, map_device=None):
        for name, unit in self._units().items():
            unit.load(f'{directory}/{name}', map_device=map_device)

    def load_from_autoencoder(self, directory):
        """Selective transfer from an autoencoder checkpoint (keynet.py:38-42)."""
        self._initialize_weights()
        self.encoder.load(directory + '/encoder', out_block=False)
        self.keypoint.load(directory + '/encoder', in_block=True, core=True, out_block=False)
        self.decoder.load(directory + '/decoder', in_block=False, core=True, out_block=True)


def build(model_type, in_channels, z_channels, keypoints, sigma=0.1):
    leaky = dict(nonlinearity=nn.LeakyReLU, nonlinearity_kwargs={'inplace': True})
    encoder = knn.Unit(in_channels, z_channels, vgg.make_layers(vgg.vgg_cfg[model_type], **leaky))
    decoder = knn.Unit(z_channels + keypoints, in_channels, vgg.make_layers(vgg.decoder_cfg[model_type]))
    keypoint = knn.Unit(in_channels, keypoints, vgg.make_layers(vgg.vgg_cfg[model_type], **leaky))
    return KeyNet(encoder, keypoint, knn.GaussianLike(sigma=sigma), decoder, init_weights=True)


def make(args):
    """keynet.make(args) (keynet.py:50-66; the reference's own version raises NameError, SURVEY W4 — this
    builds what it intends)."""
    net = build(args.model_type, args.model_in_channels, args.model_z_channels, args.model_keypoints)
    if getattr(args, 'load', None) is not None:
        net.load(args.load)
    if getattr(args, 'transfer_load', None) is not None:
        net.load_from_autoencoder(args.transfer_load)
    return net
