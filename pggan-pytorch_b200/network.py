"""Generator / Discriminator with the reference's constructor surface (network.py:75-116, 190-223), parameter
names (state_dict compatible: block0.c1.conv.weight, blocks.{i}.fromRGB.conv.bias, linear.weight, ...) and
mutable ``depth`` / ``alpha`` attributes (written by DepthManager, plugins.py:66,76; read at forward time).

The modules only HOLD parameters; all arithmetic runs in libpgk.so through ``engine.GEngine`` / ``engine.DEngine``.
``forward`` is inference-only (what generate.py / OutputGenerator use); training goes through
``wgan_gp_loss.wgan_gp_D_loss`` / ``wgan_gp_G_loss`` which run the fused forward+backward chains.
"""
import math

import torch
from torch import nn

from . import _lib

# number of bf16 planes per activation / weight tensor (see include/pgk.h).  'fp32' keeps 24 mantissa bits (three
# planes, six tensor-core products): LeakyReLU masks make the gradients discontinuous in the forward values, so the
# reference's fp32 results are only reproduced to 1e-3 when the forward pass carries (almost) fp32 precision.
PRECISIONS = {'fp32': 3, 'bf16x2': 2, 'bf16': 1}


def nf_fn(fmap_base, fmap_decay, fmap_max):
    def nf(stage):
        return min(int(fmap_base / (2.0 ** (stage * fmap_decay))), fmap_max)
    return nf


class PGConv2d(nn.Module):
    """Parameter holder of one equalised-LR conv (reference network.py:7-30).  `c` is the measured RMS of the
    He-initialised weight, kept as a plain attribute exactly as the reference does (not in the state_dict)."""

    def __init__(self, ch_in, ch_out, ksize=3, stride=1, pad=1, pixelnorm=True, wscale=True, act='lrelu'):
        super().__init__()
        if act not in ('lrelu', None):
            raise NotImplementedError('libpgk implements LeakyReLU(0.2) only (leakyrelu=False is not supported)')
        if stride != 1:
            raise NotImplementedError('stride 1 only')
        self.conv = nn.Conv2d(ch_in, ch_out, ksize, stride, pad)
        if wscale:
            nn.init.kaiming_normal_(self.conv.weight)
            with torch.no_grad():
                c = torch.sqrt(torch.mean(self.conv.weight ** 2))
                self.conv.weight /= c
            self.c = float(c)
        else:
            self.c = 1.0
        self.eps = 1e-8
        self.pixelnorm = pixelnorm
        self.act = act

    @property
    def cf(self):
        return float(self.c)

    def forward(self, x):
        raise RuntimeError('PGConv2d is a parameter holder; run the owning Generator / Discriminator')

    def extra_repr(self):
        return 'c=%.6g, pixelnorm=%s, act=%s' % (self.cf, self.pixelnorm, self.act)


class GFirstBlock(nn.Module):
    def __init__(self, ch_in, ch_out, num_channels, **layer_settings):
        super().__init__()
        self.c1 = PGConv2d(ch_in, ch_out, 4, 1, 3, **layer_settings)
        self.c2 = PGConv2d(ch_out, ch_out, **layer_settings)
        self.toRGB = PGConv2d(ch_out, num_channels, ksize=1, pad=0, pixelnorm=False, act=None)


class GBlock(nn.Module):
    def __init__(self, ch_in, ch_out, num_channels, **layer_settings):
        super().__init__()
        self.c1 = PGConv2d(ch_in, ch_out, **layer_settings)
        self.c2 = PGConv2d(ch_out, ch_out, **layer_settings)
        self.toRGB = PGConv2d(ch_out, num_channels, ksize=1, pad=0, pixelnorm=False, act=None)


class DBlock(nn.Module):
    def __init__(self, ch_in, ch_out, num_channels, **layer_settings):
        super().__init__()
        self.fromRGB = PGConv2d(num_channels, ch_in, ksize=1, pad=0, pixelnorm=False)
        self.c1 = PGConv2d(ch_in, ch_in, **layer_settings)
        self.c2 = PGConv2d(ch_in, ch_out, **layer_settings)


class MinibatchStddev(nn.Module):
    """Marker module (the scalar statistic is computed inside libpgk: pgk_stddev_stats)."""

    def __init__(self):
        super().__init__()
        self.eps = 1.0


class DLastBlock(nn.Module):
    def __init__(self, ch_in, ch_out, num_channels, **layer_settings):
        super().__init__()
        self.fromRGB = PGConv2d(num_channels, ch_in, ksize=1, pad=0, pixelnorm=False)
        self.stddev = MinibatchStddev()
        self.c1 = PGConv2d(ch_in + 1, ch_in, **layer_settings)
        self.c2 = PGConv2d(ch_in, ch_out, 4, 1, 0, **layer_settings)


def _check_channels(widths, what):
    bad = [w for w in widths if w % 8]
    if bad:
        raise ValueError('%s: feature-map counts must be multiples of 8 for the channels-innermost kernels, got %s'
                         % (what, bad))


class _EngineOwner(nn.Module):
    """Common plumbing: lazily built engine, dropped on pickling (SaverPlugin pickles whole modules,
    plugins.py:158-166)."""
    precision = 'fp32'

    def __getstate__(self):
        st = self.__dict__.copy()
        st.pop('_engine', None)
        st.pop('_param_list', None)
        return st

    def zero_grad(self, set_to_none=True):
        """The loss functions call this three times per iteration (reference wgan_gp_loss.py:42-43,69), and
        nn.Module.zero_grad walks the whole module tree each time (~0.4 ms for these models, of an iteration whose host
        side takes a few milliseconds).  The parameter set of these modules never changes: the list is built once."""
        if not set_to_none:
            return super().zero_grad(set_to_none=False)
        ps = self.__dict__.get('_param_list')
        if ps is None:
            ps = self.__dict__['_param_list'] = list(self.parameters())
        for p in ps:
            p.grad = None

    @property
    def planes(self):
        return PRECISIONS[self.precision]

    def _input(self, x):
        if not x.is_cuda:
            raise _lib.PgkError('%s runs on sm_100a CUDA tensors only; got a %s tensor (there is no CPU path)'
                                % (type(self).__name__, x.device))
        _lib.check_device(x.device)
        return x.detach().contiguous().float()

    def set_wscale(self, c_by_name):
        """Install equalised-LR constants, e.g. exported from a reference model ({'block0.c1': c, ...})."""
        mods = dict(self.named_modules())
        for name, c in c_by_name.items():
            mods[name].c = float(c)


class Generator(_EngineOwner):
    def __init__(self, dataset_shape, fmap_base=4096, fmap_decay=1.0, fmap_max=512, latent_size=512,
                 normalize_latents=True, wscale=True, pixelnorm=True, leakyrelu=True):
        super().__init__()
        resolution, num_channels = dataset_shape[-1], dataset_shape[1]
        R = int(math.log2(resolution))
        assert resolution == 2 ** R and resolution >= 4
        nf = nf_fn(fmap_base, fmap_decay, fmap_max)
        if latent_size is None:
            latent_size = nf(0)
        self._construct([nf(i) for i in range(1, R)], num_channels, latent_size, normalize_latents, wscale, pixelnorm,
                        leakyrelu)

    def _construct(self, widths, num_channels, latent_size, normalize_latents=True, wscale=True, pixelnorm=True,
                   leakyrelu=True):
        """widths[i-1] = feature maps of stage i (nf(i), i = 1..R-1).  Shared by __init__ and checkpoint.from_reference
        (a snapshot carries its widths, not the fmap_* arguments that produced them)."""
        _check_channels([latent_size] + list(widths), 'Generator')
        self.normalize_latents = normalize_latents
        self.pixelnorm = pixelnorm
        settings = dict(wscale=wscale, pixelnorm=pixelnorm, act='lrelu' if leakyrelu else 'relu')
        self.block0 = GFirstBlock(latent_size, widths[0], num_channels, **settings)
        self.blocks = nn.ModuleList([GBlock(widths[i - 1], widths[i], num_channels, **settings)
                                     for i in range(1, len(widths))])
        self.depth = 0
        self.alpha = 1.0
        self.eps = 1e-8
        self.latent_size = latent_size
        self.max_depth = len(self.blocks)

    @property
    def engine(self):
        if '_engine' not in self.__dict__:
            from .engine import GEngine
            self.__dict__['_engine'] = GEngine(self)
        return self.__dict__['_engine']

    def forward(self, x):
        """(N, latent) -> (N, C, r, r) fp32, r = 4 * 2**depth.  Inference only (no autograd graph)."""
        with torch.no_grad():
            img, _ = self.engine.forward(self._input(x), self.planes)
        return img


class Discriminator(_EngineOwner):
    def __init__(self, dataset_shape, fmap_base=4096, fmap_decay=1.0, fmap_max=512, wscale=True, pixelnorm=False,
                 leakyrelu=True):
        super().__init__()
        if pixelnorm:
            raise NotImplementedError('the discriminator kernels implement pixelnorm=False (the reference default)')
        resolution, num_channels = dataset_shape[-1], dataset_shape[1]
        R = int(math.log2(resolution))
        assert resolution == 2 ** R and resolution >= 4
        nf = nf_fn(fmap_base, fmap_decay, fmap_max)
        self._construct([nf(i) for i in range(0, R)], num_channels, wscale, pixelnorm, leakyrelu)

    def _construct(self, widths, num_channels, wscale=True, pixelnorm=False, leakyrelu=True):
        """widths[i] = feature maps of stage i (nf(i), i = 0..R-1); see Generator._construct."""
        if pixelnorm:
            raise NotImplementedError('the discriminator kernels implement pixelnorm=False (the reference default)')
        R = len(widths)
        self.R = R
        _check_channels(list(widths), 'Discriminator')
        settings = dict(wscale=wscale, pixelnorm=pixelnorm, act='lrelu' if leakyrelu else 'relu')
        self.blocks = nn.ModuleList([DBlock(widths[i], widths[i - 1], num_channels, **settings)
                                     for i in range(R - 1, 1, -1)]
                                    + [DLastBlock(widths[1], widths[0], num_channels, **settings)])
        self.linear = nn.Linear(widths[0], 1)
        self.depth = 0
        self.alpha = 1.0
        self.eps = 1e-8
        self.max_depth = len(self.blocks) - 1

    @property
    def engine(self):
        if '_engine' not in self.__dict__:
            from .engine import DEngine
            self.__dict__['_engine'] = DEngine(self)
        return self.__dict__['_engine']

    def forward(self, x):
        """(N, C, r, r) -> (N, 1) scores.  Inference only (no autograd graph)."""
        with torch.no_grad():
            x = self._input(x)
            T = self.engine.forward(x, 1, x.shape[0], self.planes)
        return T.scores.view(-1, 1)
