"""Shared helpers for the test-suite: golden loading, error metric."""
import glob
import json
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, 'golden')

STEP_CASES = sorted(os.path.basename(f)[5:-4] for f in glob.glob(os.path.join(GOLDEN, 'step_*.npz')))
# reference vectors that pin the oracle only (CPU tests; larger / odd configurations the GPU suite covers through
# the oracle on fresh inputs instead)
ORACLE_ONLY_CASES = sorted(os.path.basename(f)[6:-4] for f in glob.glob(os.path.join(GOLDEN, 'ostep_*.npz')))


def load_step(name, dtype=torch.float32, prefix='step_'):
    z = np.load(os.path.join(GOLDEN, '%s%s.npz' % (prefix, name)))
    res, ch, fb, fm, lat, n, depth = [int(v) for v in z['meta']]
    g = dict(resolution=res, channels=ch, fmap_base=fb, fmap_max=fm, latent=lat, n=n, depth=depth,
             alpha=float(z['alpha']))
    t = lambda a: torch.from_numpy(np.asarray(a)).to(dtype)
    g['pg'] = {k[2:]: t(z[k]) for k in z.files if k.startswith('G.')}
    g['pd'] = {k[2:]: t(z[k]) for k in z.files if k.startswith('D.')}
    g['dgrad'] = {k[6:]: t(z[k]) for k in z.files if k.startswith('Dgrad.')}
    g['ggrad'] = {k[6:]: t(z[k]) for k in z.files if k.startswith('Ggrad.')}
    for k in ('z1', 'z2', 'real', 'mixing', 'fake', 'd_real_scores', 'd_fake_scores', 'd_cost', 'd_real_loss',
              'd_fake_loss', 'g_cost'):
        g[k] = t(z[k])
    return g


def load_trainer(dtype=torch.float32):
    z = np.load(os.path.join(GOLDEN, 'trainer_tiny3.npz'))
    res, ch, fb, fm, lat, n, depth = [int(v) for v in z['meta']]
    g = dict(resolution=res, channels=ch, fmap_base=fb, fmap_max=fm, latent=lat, n=n, depth=depth,
             alpha=float(z['alpha']), cur_nimg=int(z['cur_nimg']))
    t = lambda a: torch.from_numpy(np.asarray(a)).to(dtype)
    for pre in ('G0', 'D0', 'G2', 'D2'):
        g[pre] = {k[3:]: t(z[k]) for k in z.files if k.startswith(pre + '.')}
    g['reals'], g['latents'], g['mixing'] = t(z['reals']), t(z['latents']), t(z['mixing'])
    return g


def load_schedule():
    with open(os.path.join(GOLDEN, 'schedule.json')) as f:
        return json.load(f)


def rel_err(a, b):
    """||a-b||_2 / ||b||_2 (SURVEY.md 8c tolerance definition); b is the reference."""
    a = torch.as_tensor(a).detach().double().cpu().reshape(-1)
    b = torch.as_tensor(b).detach().double().cpu().reshape(-1)
    den = float(b.norm())
    num = float((a - b).norm())
    if den == 0.0:
        return num
    return num / den
