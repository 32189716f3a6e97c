"""Run the UNMODIFIED reference (deepsound-project/pggan-pytorch) as a baseline: its own network.Generator /
Discriminator, wgan_gp_loss.* and trainer.Trainer.train() (trainer.py:85-115) on synthetic inputs.

This file is measurement / test infrastructure, not product: only bench.py's baseline legs (`--impl reference`,
`cpu_baseline`, `gpu_eager_reference`), tests/ and tests/dev/ import it.  Nothing under pggan-pytorch_b200/ does.

The reference is nine plain Python files without a setup.py, so "installing" it is copying them:
`__graft_entry__.build()` does that into baseline/_ref/ whenever /root/reference is present (git-ignored, shipped to
the GPU box with the snapshot).  On a host without CUDA the reference's hard-coded `.cuda()` calls (network.py:30,
trainer.py:86,92,103, wgan_gp_loss.py:16,22) are neutralised; on a GPU box it runs as it is.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = [os.path.join(HERE, '_ref'), '/root/reference']


def ref_dir():
    for d in CANDIDATES:
        if os.path.exists(os.path.join(d, 'network.py')) and os.path.exists(os.path.join(d, 'trainer.py')):
            return d
    return None


def available():
    return ref_dir() is not None


_mods = {}


def load(device='cuda'):
    """Import the reference's network / wgan_gp_loss / trainer modules (once).  device='cpu' installs the .cuda()
    shims first -- process-wide, so a process that has loaded the reference for the CPU must not use CUDA afterwards."""
    if _mods:
        return _mods
    d = ref_dir()
    if d is None:
        raise RuntimeError('the reference is not installed: neither baseline/_ref/ nor /root/reference exists')
    if device == 'cpu':
        from torch import nn
        nn.Module.cuda = lambda self, *a, **k: self
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.cuda.FloatTensor = torch.FloatTensor
    import importlib.util
    import warnings
    warnings.filterwarnings('ignore')
    for name in ('network', 'wgan_gp_loss', 'trainer'):
        spec = importlib.util.spec_from_file_location('pggan_reference_' + name, os.path.join(d, name + '.py'))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        _mods[name] = m
    _mods['dir'] = d
    return _mods


def make_trainer(res, ch, depth, alpha, n, device='cuda', seed=1337, pinned_host_inputs=False):
    """The reference's own objects wired as train.py:123-165 does, on synthetic data: Adam(lr 1e-3, betas (0, .99)),
    reals ~ N(0,1) of the depth's resolution, latents from numpy (utils.py:56-57)."""
    m = load(device)
    import contextlib
    import io
    torch.manual_seed(seed)
    np.random.seed(seed)
    shape = (1000, ch, res, res)
    with contextlib.redirect_stdout(io.StringIO()):      # network.py:90 prints the dataset shape
        G = m['network'].Generator(shape)
        D = m['network'].Discriminator(shape)
    if device != 'cpu':
        G.cuda()
        D.cuda()
    G.depth = D.depth = depth
    G.alpha = D.alpha = alpha
    opt_g = torch.optim.Adam(G.parameters(), 1e-3, betas=(0.0, 0.99))
    opt_d = torch.optim.Adam(D.parameters(), 1e-3, betas=(0.0, 0.99))
    r = 4 * 2 ** depth
    reals = [torch.randn(n, ch, r, r) for _ in range(2)]
    if device != 'cpu' and not pinned_host_inputs:
        reals = [x.cuda() for x in reals]                # inputs resident in HBM: .cuda() in train() is then a no-op
    elif pinned_host_inputs:
        reals = [x.pin_memory() for x in reals]

    def dataiter():
        i = 0
        while True:
            i += 1
            yield reals[i % 2]

    def latents():
        z = torch.from_numpy(np.random.randn(n, 512).astype(np.float32))
        return z

    t = m['trainer'].Trainer(D, G, m['wgan_gp_loss'].wgan_gp_D_loss, m['wgan_gp_loss'].wgan_gp_G_loss, opt_d, opt_g,
                             None, dataiter(), latents)
    return t


def time_train(res, ch, depth, alpha, n, steps, warmup, device='cuda', tf32=False, budget_s=None):
    """images/sec of the reference's Trainer.train().  On CUDA: device-timed with events; on the CPU: wall clock.
    budget_s bounds the CPU leg: the number of timed steps is cut so that the leg ends within it."""
    if device != 'cpu':
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True
    t = make_trainer(res, ch, depth, alpha, n, device)
    if device == 'cpu':
        t0 = time.perf_counter()
        t.train()
        first = time.perf_counter() - t0
        k, w = steps, max(0, warmup - 1)
        if budget_s is not None:
            k = max(1, min(steps, int(budget_s / max(first, 1e-3))))
            w = max(0, min(w, int(0.3 * budget_s / max(first, 1e-3))))
        for _ in range(w):
            t.train()
        t0 = time.perf_counter()
        for _ in range(k):
            t.train()
        dt = (time.perf_counter() - t0) / k
        return n / dt, dt * 1e3, k, w + 1
    for _ in range(warmup):
        t.train()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        t.train()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return n / (ms * 1e-3), ms, steps, warmup
