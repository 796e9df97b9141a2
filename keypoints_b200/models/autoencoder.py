"""Pre-training auto-encoder with the reference's API (keypoints/models/autoencoder.py:4-24, built by autoencode.py:59-66
from two ``knn.Unit`` stacks; its checkpoints feed ``TransporterNet.load_from_autoencoder`` / ``--transfer_load``)."""
from . import knn


class AutoEncoder(knn.Container):
    def __init__(self, encoder, decoder, init_weights=True):
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder
        if init_weights:
            self._initialize_weights()

    def forward(self, x):
        z = self.encoder(x)
        x = self.decoder(z)
        return z, x

    def load(self, root_path):
        self.encoder.load(root_path + '/encoder')
        self.decoder.load(root_path + '/decoder')

    def save(self, root_path):
        self.encoder.save(root_path + '/encoder')
        self.decoder.save(root_path + '/decoder')


def make(type, in_channels, z_channels, load=None):
    """The model autoencode.py:59-66 assembles: LeakyReLU cores on both sides."""
    from torch import nn
    from . import vgg
    leaky = dict(nonlinearity=nn.LeakyReLU, nonlinearity_kwargs={'inplace': True})
    encoder = knn.Unit(in_channels, z_channels, vgg.make_layers(vgg.vgg_cfg[type], **leaky))
    decoder = knn.Unit(z_channels, in_channels, vgg.make_layers(vgg.decoder_cfg[type], **leaky))
    net = AutoEncoder(encoder, decoder, init_weights=load is None)
    if load is not None:
        net.load(load)
    return net
