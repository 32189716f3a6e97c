"""Reading the reference's snapshots into the libpgk-backed modules.

The reference's SaverPlugin pickles WHOLE modules -- ``torch.save(model, 'network-snapshot-{G,D}-{epoch}.dat')``
(plugins.py:142-174) -- and ``train.py:60-64`` / ``generate.py:18-25`` read them back with ``torch.load``.  Such a
pickle names the classes of the reference's ``network`` module (Generator, GBlock, PGConv2d, ...), so plain
``torch.load`` needs the reference on ``sys.path``.  ``load_snapshot`` does not: an unpickler maps every ``network.*``
class to an empty ``nn.Module`` shell, and ``from_reference`` rebuilds a ``pggan_b200`` Generator / Discriminator from
what the shell (or a live reference module) carries:

  * the layer widths and image channels, read off the weight shapes (a snapshot does not record fmap_base / fmap_max);
  * every parameter, under the same ``state_dict`` names (``block0.c1.conv.weight``, ``blocks.3.fromRGB.conv.bias``,
    ``linear.weight`` ...);
  * the equalised-LR constant ``c`` of every PGConv2d -- a plain attribute, NOT part of the state_dict
    (network.py:19; a Python float under torch 0.2, a 0-d tensor under current torch);
  * ``depth``, ``alpha``, ``normalize_latents``, ``latent_size``.

Optimizer state is not in the reference's snapshots (plugins.py:158-166) and is not expected here.
"""
import pickle
import types

import torch
from torch import nn

from .network import Discriminator, Generator

_SHELLS = {}


def _shell_class(name):
    if name not in _SHELLS:
        _SHELLS[name] = type(name, (nn.Module,), {'__module__': __name__, '_pgk_reference_shell': True})
    return _SHELLS[name]


class _RefUnpickler(pickle.Unpickler):
    """network.<Class> (the reference's top-level module, train.py:11) -> an nn.Module shell of the same name;
    everything else (torch tensors, nn.Conv2d, nn.LeakyReLU, collections, and this package's own
    `pggan-pytorch_b200.network` classes) resolves normally."""

    def find_class(self, module, name):
        if module == 'network':
            return _shell_class(name)
        return super().find_class(module, name)


def _pickle_module():
    m = types.ModuleType('pgk_reference_pickle')
    m.__dict__.update({k: getattr(pickle, k) for k in dir(pickle) if not k.startswith('__')})
    m.Unpickler = _RefUnpickler
    m.load = lambda f, **kw: _RefUnpickler(f, **kw).load()
    m.loads = lambda b, **kw: _RefUnpickler(__import__('io').BytesIO(b), **kw).load()
    return m


def _convs(ref):
    """{qualified name: module} of the equalised-LR conv wrappers (reference class PGConv2d)."""
    return {n: m for n, m in ref.named_modules() if type(m).__name__ == 'PGConv2d'}


def _check_act(ref):
    for n, m in _convs(ref).items():
        act = m._modules.get('act') if hasattr(m, '_modules') else None
        if act is None:
            act = getattr(m, 'act', None)
        if act is not None and not isinstance(act, (nn.LeakyReLU, str)):
            raise NotImplementedError('%s uses %s; libpgk implements LeakyReLU(0.2) only' % (n, type(act).__name__))


def from_reference(ref, device=None):
    """ref: a reference ``network.Generator`` / ``network.Discriminator`` (live, or the shell tree produced by
    ``load_snapshot``).  Returns the equivalent pggan_b200 module with parameters, c constants, depth and alpha."""
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    convs = _convs(ref)
    _check_act(ref)
    wscale = False    # no He re-initialisation: the weights are overwritten below and every c comes from the snapshot
    if 'block0.c1.conv.weight' in sd:
        nblocks = len({k.split('.')[1] for k in sd if k.startswith('blocks.')})
        w0 = sd['block0.c1.conv.weight']                       # (nf(1), latent, 4, 4)
        widths = [w0.shape[0]] + [sd['blocks.%d.c1.conv.weight' % i].shape[0] for i in range(nblocks)]
        num_channels = sd['block0.toRGB.conv.weight'].shape[0]
        pixelnorm = bool(getattr(convs['block0.c1'], 'pixelnorm', True))
        out = Generator.__new__(Generator)
        nn.Module.__init__(out)
        out._construct(widths, num_channels, w0.shape[1], bool(getattr(ref, 'normalize_latents', True)), wscale, pixelnorm)
    elif 'linear.weight' in sd:
        nblocks = len({k.split('.')[1] for k in sd if k.startswith('blocks.')})
        # blocks[j] = DBlock(nf(R-1-j), nf(R-2-j)) for j < nblocks-1, blocks[-1] = DLastBlock(nf(1), nf(0))
        R = nblocks + 1
        widths = [0] * R
        for j in range(nblocks - 1):
            widths[R - 1 - j] = sd['blocks.%d.c1.conv.weight' % j].shape[0]
        widths[1] = sd['blocks.%d.c1.conv.weight' % (nblocks - 1)].shape[0]
        widths[0] = sd['blocks.%d.c2.conv.weight' % (nblocks - 1)].shape[0]
        num_channels = sd['blocks.0.fromRGB.conv.weight'].shape[1]
        out = Discriminator.__new__(Discriminator)
        nn.Module.__init__(out)
        out._construct(widths, num_channels, wscale, bool(getattr(convs['blocks.0.c1'], 'pixelnorm', False)))
    else:
        raise ValueError('not a reference Generator / Discriminator: state_dict has neither block0.* nor linear.*')
    missing = set(out.state_dict()) ^ set(sd)
    if missing:
        raise ValueError('parameter names differ from the reference layout: %s' % sorted(missing))
    out.load_state_dict(sd)
    out.set_wscale({n: float(getattr(m, 'c', 1.0)) for n, m in convs.items()})
    out.depth = int(getattr(ref, 'depth', 0))
    out.alpha = float(getattr(ref, 'alpha', 1.0))
    if getattr(ref, 'precision', None) in ('fp32', 'bf16x2', 'bf16'):     # (a pggan_b200 module passed through here)
        out.precision = ref.precision
    if device is not None:
        out.to(device)
    return out


def load_snapshot(path, device=None):
    """A ``network-snapshot-{G,D}-*.dat`` written by the reference's SaverPlugin (or by ``torch.save`` of a reference
    module) -> pggan_b200 module.  Snapshots of pggan_b200 modules themselves (our Trainer + the same SaverPlugin) are
    returned as they are."""
    obj = torch.load(path, map_location='cpu', pickle_module=_pickle_module(), weights_only=False)
    if isinstance(obj, (Generator, Discriminator)):
        return obj.to(device) if device is not None else obj
    if isinstance(obj, dict):
        raise ValueError('%s holds a dict, not a pickled module; use module.load_state_dict + set_wscale' % path)
    return from_reference(obj, device)


def resume(g_path, d_path, device=None):
    """train.py:60-64 (load_models): (G, D) from a pair of snapshots."""
    G, D = load_snapshot(g_path, device), load_snapshot(d_path, device)
    if not isinstance(G, Generator) or not isinstance(D, Discriminator):
        raise ValueError('expected a generator snapshot and a discriminator snapshot, got %s / %s'
                         % (type(G).__name__, type(D).__name__))
    assert G.max_depth == D.max_depth, 'train.py:126'
    return G, D
