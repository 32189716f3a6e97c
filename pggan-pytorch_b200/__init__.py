"""pggan-pytorch_b200: the Progressive-GAN G + D + WGAN-GP training step as hand-written sm_100a CUDA kernels
(libpgk.so, C ABI in include/pgk.h) behind the reference's Python surface.

    from pggan_b200 import Generator, Discriminator, wgan_gp_D_loss, wgan_gp_G_loss, Trainer, DepthManager
"""
from .checkpoint import from_reference, load_snapshot, resume      # noqa: F401
from .dataset import prepare_reals                                 # noqa: F401
from .network import Discriminator, Generator, PGConv2d            # noqa: F401
from .optim import FusedAdam                                       # noqa: F401
from .plugins import AsyncLossMonitor, DepthManager, LRScheduler, Plugin, lr_rampup, schedule  # noqa: F401
from .trainer import Trainer                                       # noqa: F401
from .utils import device_random_latents, random_latents           # noqa: F401
from .wgan_gp_loss import wgan_gp_D_loss, wgan_gp_G_loss           # noqa: F401
from . import _lib, wgan_gp_loss                                   # noqa: F401
