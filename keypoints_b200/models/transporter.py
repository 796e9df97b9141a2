"""Transporter (Kulkarni et al. 2019) with the reference's API (keypoints/models/transporter.py:8-138)."""
import torch
from torch import nn

from . import functional as MF
from . import knn, vgg


class TransporterNet(knn.Container):
    def __init__(self, feature_cnn, keypoint_cnn, key2map, decoder, init_weights=True, combine_method='loop'):
        super().__init__()
        self.feature = feature_cnn
        self.keypoint = keypoint_cnn
        self.ssm = knn.SpatialLogSoftmax()
        self.key2map = key2map
        self.decoder = decoder
        self.combine_method = combine_method
        if init_weights:
            self._initialize_weights()

    def extract(self, x):
        phi = self.feature(x)
        heatmap = self.keypoint(x)
        k, p = self.ssm(heatmap, probs=True)
        m = self.key2map(k, height=phi.size(2), width=phi.size(3))
        return phi, heatmap, k, p, m

    def forward(self, xs, xt):
        """-> (x_t, phi, k_xt, m_xt, (p_h,p_w), heatmap_xt, mask_xs, mask_xt), transporter.py:34-64.
        The source frame is a constant (no_grad) but still updates the BatchNorm running statistics."""
        with torch.no_grad():
            phi_xs, _, k_xs, _, _ = self.extract(xs)
        phi_xt, heatmap_xt, k_xt, p_xt, m_xt = self.extract(xt)
        sigma = getattr(self.key2map, 'sigma', 0.1)
        # 'max' (default of make()), 'sum_and_clamp', 'loop' (:41-60); 'pretrained_network' raises (no MaskMaker in make())
        phi, mask_xs, mask_xt = MF.transport(phi_xs, phi_xt, k_xs, k_xt, mode=self.combine_method, sigma=sigma)
        x_t = self.decoder(phi)
        return x_t, phi, k_xt, m_xt, p_xt, heatmap_xt, mask_xs, mask_xt

    def load(self, directory, map_device=None):
        self.feature.load(directory + '/encoder', map_device=map_device)
        self.keypoint.load(directory + '/keypoint', map_device=map_device)
        self.decoder.load(directory + '/decoder', map_device=map_device)

    def load_from_autoencoder(self, directory):
        self._initialize_weights()
        self.feature.load(directory + '/encoder', out_block=False)
        self.keypoint.load(directory + '/encoder', in_block=True, core=True, out_block=False)
        self.decoder.load(directory + '/decoder', in_block=False, core=True, out_block=True)

    def save(self, directory):
        self.feature.save(directory + '/encoder')
        self.keypoint.save(directory + '/keypoint')
        self.decoder.save(directory + '/decoder')


def make(type, in_channels, z_channels, keypoints, combine_mode='max', load=None, transfer_load=None, map_device=None):
    leaky = dict(nonlinearity=nn.LeakyReLU, nonlinearity_kwargs={'inplace': True})
    encoder = knn.Unit(in_channels, z_channels, vgg.make_layers(vgg.vgg_cfg[type], **leaky))
    decoder = knn.Unit(z_channels, in_channels, vgg.make_layers(vgg.decoder_cfg[type]))
    keypoint = knn.Unit(in_channels, keypoints, vgg.make_layers(vgg.vgg_cfg[type], **leaky))
    net = TransporterNet(encoder, keypoint, knn.GaussianLike(sigma=0.1), decoder, init_weights=True,
                         combine_method=combine_mode)
    if load is not None:
        net.load(load, map_device)
    if transfer_load is not None:
        net.load_from_autoencoder(transfer_load)
    return net
