"""Make the reference's import paths resolve to keypoints_b200.

After ``keypoints_b200.dropin.install()`` the reference training scripts' own imports —
``from keypoints.models import transporter``, ``from tps import ...``, ``from data_augments import TpsAndRotate``,
``from apex import amp`` — bind to this package, so ``transporter.py`` / ``keypoints.py`` drive the CUDA path unchanged.
"""
import contextlib
import sys
import types

import keypoints_b200
from keypoints_b200 import data_augments, tps
from keypoints_b200.models import autoencoder, functional, keynet, knn, transporter, vgg


def _amp_shim():
    """apex.amp is unconditional in the reference scripts (transporter.py:9,50-51,84-86) and not installable here;
    mixed precision is our 'bf16' mode, so initialize / scale_loss are pass-throughs."""
    amp = types.ModuleType('apex.amp')

    def initialize(model, optim, opt_level='O0', **_):
        keypoints_b200.set_precision('fp32' if opt_level == 'O0' else 'bf16')
        return model, optim

    @contextlib.contextmanager
    def scale_loss(loss, optim, **_):
        yield loss

    amp.initialize, amp.scale_loss = initialize, scale_loss
    apex = types.ModuleType('apex')
    apex.amp = amp
    return apex, amp


def install(precision=None, shim_apex=True):
    pkg, models = types.ModuleType('keypoints'), types.ModuleType('keypoints.models')
    pkg.__path__, models.__path__ = [], []
    models.Container = knn.Container        # models/autoencoder.py:1 imports it from the package (the reference's __init__ forgot it)
    for name, mod in dict(knn=knn, vgg=vgg, keynet=keynet, transporter=transporter, functional=functional,
                          autoencoder=autoencoder).items():
        setattr(models, name, mod)
        sys.modules[f'keypoints.models.{name}'] = mod
    pkg.models = models
    sys.modules.update({'keypoints': pkg, 'keypoints.models': models, 'tps': tps, 'data_augments': data_augments})
    if shim_apex and 'apex' not in sys.modules:
        try:
            import apex  # noqa: F401
        except ImportError:
            apex, amp = _amp_shim()
            sys.modules.update({'apex': apex, 'apex.amp': amp})
    if precision is not None:
        keypoints_b200.set_precision(precision)
