"""Generate the golden vectors under tests/golden/ by EXECUTING the unmodified
reference (imported from /root/reference; only possible in the authoring
container -- the reference does not travel to the GPU box).

    python tests/golden/make_golden.py

Outputs (committed):
  step_<cfg>.npz   weights (+ equalised-LR constants c), seeded inputs, and what the
                   reference's Generator / Discriminator / wgan_gp_D_loss /
                   wgan_gp_G_loss / backward() produce for them
  trainer_tiny3.npz  parameters after 2 reference Trainer.train() iterations
  schedule.json    DepthManager (depth, alpha, minibatch, tick) for a list of cur_nimg

Shims: the reference hard-codes .cuda(); on this CPU-only host
they are neutralised.  plugins.py imports the long-removed torch.utils.trainer
package, for which a 12-line stand-in is injected.
"""
import json
import os
import sys
import types

import numpy as np
import torch
from torch import nn

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))


def install_shims():
    nn.Module.cuda = lambda self, *a, **k: self
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.FloatTensor = torch.FloatTensor
    tr = types.ModuleType('torch.utils.trainer')
    pl = types.ModuleType('torch.utils.trainer.plugins')
    pp = types.ModuleType('torch.utils.trainer.plugins.plugin')

    class Plugin(object):
        def __init__(self, interval=None):
            self.trigger_interval = interval

        def register(self, trainer):
            raise NotImplementedError

    class LossMonitor(Plugin):
        def __init__(self, *a, **k):
            super().__init__([(1, 'iteration'), (1, 'epoch')])

    class Logger(Plugin):
        def __init__(self, fields=None, interval=None):
            super().__init__(interval)

    pp.Plugin = Plugin
    pl.LossMonitor = LossMonitor
    pl.Logger = Logger
    pl.plugin = pp
    tr.plugins = pl
    sys.modules['torch.utils.trainer'] = tr
    sys.modules['torch.utils.trainer.plugins'] = pl
    sys.modules['torch.utils.trainer.plugins.plugin'] = pp
    sys.path.insert(0, REF)


def export_params(model):
    """state_dict + '<conv>.c' for every PGConv2d (c is a plain attribute, network.py:19)."""
    out = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
    for name, m in model.named_modules():
        if m.__class__.__name__ == 'PGConv2d':
            out[name + '.c'] = np.float32(float(m.c))
    return out


CONFIGS = {
    # name: (resolution, channels, fmap_base, fmap_max, latent, N, depth, alpha)
    'tiny3_d0_a1':   (16, 3, 128, 32, 32, 4, 0, 1.0),
    'tiny3_d1_a05':  (16, 3, 128, 32, 32, 4, 1, 0.5),
    'tiny3_d2_a03':  (16, 3, 128, 32, 32, 4, 2, 0.3),
    'tiny3_d2_a1':   (16, 3, 128, 32, 32, 4, 2, 1.0),
    'tiny1_d1_a025': (8, 1, 64, 16, 16, 6, 1, 0.25),
}


def build(cfg):
    import network
    res, ch, fb, fm, lat, n, depth, alpha = cfg
    torch.manual_seed(1337)
    shape = (1000, ch, res, res)
    G = network.Generator(shape, fmap_base=fb, fmap_max=fm, latent_size=lat)
    D = network.Discriminator(shape, fmap_base=fb, fmap_max=fm)
    G.depth = D.depth = depth
    G.alpha = D.alpha = alpha
    return G, D


# cases that pin the oracle only (CPU tests): file prefix 'ostep_', not picked up by the GPU parity tests
ORACLE_ONLY = {
    # name: (resolution, channels, fmap_base, fmap_max, latent, N, depth, alpha)
    'd3_a0':      (32, 3, 256, 32, 32, 3, 3, 0.0),    # the first iteration of a new depth: alpha == 0.0 exactly
    'd2_n1':      (16, 1, 128, 32, 16, 1, 2, 0.6),    # a batch of one (minibatch stddev over a single sample)
    'd1_a1_wide': (8, 3, 256, 48, 40, 5, 1, 1.0),     # widths that are not powers of two
}


def make_step(name, cfg, prefix='step_'):
    import wgan_gp_loss
    res, ch, fb, fm, lat, n, depth, alpha = cfg
    G, D = build(cfg)
    r = 4 * 2 ** depth
    rng = np.random.RandomState(7)
    z1 = torch.from_numpy(rng.randn(n, lat).astype(np.float32))
    z2 = torch.from_numpy(rng.randn(n, lat).astype(np.float32))
    real = torch.from_numpy(rng.randn(n, ch, r, r).astype(np.float32))
    out = {}
    for k, v in export_params(G).items():
        out['G.' + k] = v
    for k, v in export_params(D).items():
        out['D.' + k] = v
    out.update(z1=z1.numpy(), z2=z2.numpy(), real=real.numpy(),
               meta=np.array([res, ch, fb, fm, lat, n, depth], dtype=np.int64), alpha=np.float64(alpha))

    # plain forwards
    with torch.no_grad():
        fake = G(z1)
        out['fake'] = fake.numpy()
        out['d_real_scores'] = D(real).numpy()
        out['d_fake_scores'] = D(fake).numpy()

    # D step: the mixing factors are the first (N,1) uniform_ draw after the seed
    seed = 4242
    torch.manual_seed(seed)
    mixing = torch.empty(n, 1).uniform_()
    out['mixing'] = mixing.numpy()
    wgan_gp_loss.mixing_factors = None
    wgan_gp_loss.grad_outputs = None
    torch.manual_seed(seed)
    d_cost, d_real_loss, d_fake_loss = wgan_gp_loss.wgan_gp_D_loss(D, G, real, z1)
    assert np.array_equal(wgan_gp_loss.mixing_factors.numpy(), mixing.numpy())
    d_cost.backward()
    out['d_cost'] = d_cost.detach().numpy()
    out['d_real_loss'] = d_real_loss.detach().numpy()
    out['d_fake_loss'] = d_fake_loss.detach().numpy()
    for k, p_ in D.named_parameters():
        if p_.grad is not None:
            out['Dgrad.' + k] = p_.grad.detach().clone().numpy()

    # G step (reference order: after the D step, on fresh latents; D not updated here)
    g_cost = wgan_gp_loss.wgan_gp_G_loss(G, D, z2)
    g_cost.backward()
    out['g_cost'] = g_cost.detach().numpy()
    for k, p_ in G.named_parameters():
        if p_.grad is not None:
            out['Ggrad.' + k] = p_.grad.detach().clone().numpy()
    np.savez_compressed(os.path.join(HERE, '%s%s.npz' % (prefix, name)), **out)
    print(name, 'd_cost', float(d_cost), 'g_cost', float(g_cost), 'keys', len(out))


def make_trainer():
    """Two unmodified Trainer.train() iterations (trainer.py:85-115) with Adam as
    train.py:148-149,195 wires it."""
    import wgan_gp_loss
    from trainer import Trainer
    from functools import partial
    cfg = CONFIGS['tiny3_d1_a05']
    res, ch, fb, fm, lat, n, depth, alpha = cfg
    G, D = build(cfg)
    out = {}
    for k, v in export_params(G).items():
        out['G0.' + k] = v
    for k, v in export_params(D).items():
        out['D0.' + k] = v
    rng = np.random.RandomState(11)
    r = 4 * 2 ** depth
    reals = [torch.from_numpy(rng.randn(n, ch, r, r).astype(np.float32)) for _ in range(2)]
    lats = [torch.from_numpy(rng.randn(n, lat).astype(np.float32)) for _ in range(4)]
    lat_iter = iter(lats)
    opt_g = torch.optim.Adam(G.parameters(), 1e-3, betas=(0.0, 0.99))
    opt_d = torch.optim.Adam(D.parameters(), 1e-3, betas=(0.0, 0.99))
    wgan_gp_loss.mixing_factors = None
    wgan_gp_loss.grad_outputs = None
    t = Trainer(D, G, partial(wgan_gp_loss.wgan_gp_D_loss, return_all=True), wgan_gp_loss.wgan_gp_G_loss,
                opt_d, opt_g, None, iter(reals), lambda: next(lat_iter))
    mix = []
    for it in range(2):
        torch.manual_seed(100 + it)
        mix.append(torch.empty(n, 1).uniform_().numpy())
        torch.manual_seed(100 + it)
        t.train()
    for k, v in export_params(G).items():
        out['G2.' + k] = v
    for k, v in export_params(D).items():
        out['D2.' + k] = v
    out['reals'] = np.stack([x.numpy() for x in reals])
    out['latents'] = np.stack([x.numpy() for x in lats])
    out['mixing'] = np.stack(mix)
    out['meta'] = np.array([res, ch, fb, fm, lat, n, depth], dtype=np.int64)
    out['alpha'] = np.float64(alpha)
    out['cur_nimg'] = np.int64(t.cur_nimg)
    np.savez_compressed(os.path.join(HERE, 'trainer_tiny3.npz'), **out)
    print('trainer: cur_nimg', t.cur_nimg)


def make_schedule():
    import plugins

    class Obj(object):
        pass

    points = [0, 1, 15, 16, 99999, 100000, 100016, 149999, 150000, 199984, 199999, 200000, 299999, 300000,
              300003, 350000, 400000, 499999, 500000, 700000, 900000, 1100000, 1100014, 1299996, 1300000,
              1300006, 1499999, 1500000, 1500003, 1530000, 1599999, 1600000, 1700000, 1700001, 5000000]
    rows = []
    for max_depth in (8, 5, 2):
        for nimg in points:
            tr = Obj()
            tr.cur_nimg = nimg
            tr.stats = {}
            tr.D, tr.G, tr.dataset = Obj(), Obj(), Obj()
            made = []
            dm = plugins.DepthManager(lambda mb: made.append(mb) or [], lambda mb: (lambda: None), max_depth)
            dm.register(tr)
            rows.append(dict(max_depth=max_depth, cur_nimg=nimg, depth=int(dm.depth), alpha=repr(float(dm.alpha)),
                             minibatch=int(tr.stats['minibatch_size']), tick_nimg=int(tr.tick_duration_nimg),
                             D_depth=int(tr.D.depth), G_alpha=repr(float(tr.G.alpha)),
                             dataset_depth=int(tr.dataset.model_depth)))
    # a walk: the sequence of (depth, alpha, dataloader re-creations) over consecutive iterations
    tr = Obj()
    tr.cur_nimg = 99968
    tr.stats = {}
    tr.D, tr.G, tr.dataset = Obj(), Obj(), Obj()
    made = []
    dm = plugins.DepthManager(lambda mb: made.append(mb) or [], lambda mb: (lambda: None), 8)
    dm.register(tr)
    walk = []
    for _ in range(8):
        tr.cur_nimg += tr.stats['minibatch_size']
        dm.iteration()
        walk.append(dict(cur_nimg=tr.cur_nimg, depth=int(dm.depth), alpha=repr(float(dm.alpha)), loaders=len(made)))
    with open(os.path.join(HERE, 'schedule.json'), 'w') as f:
        json.dump(dict(points=rows, walk=walk), f, indent=0)
    print('schedule rows', len(rows))


if __name__ == '__main__':
    install_shims()
    if len(sys.argv) > 1 and sys.argv[1] == 'oracle-only':     # adds the ostep_* files without touching the others
        for name, cfg in ORACLE_ONLY.items():
            make_step(name, cfg, prefix='ostep_')
        sys.exit(0)
    for name, cfg in CONFIGS.items():
        make_step(name, cfg)
    for name, cfg in ORACLE_ONLY.items():
        make_step(name, cfg, prefix='ostep_')
    make_trainer()
    make_schedule()
