"""Compact golden vectors of BASELINE.json's configurations AT THEIR REAL WIDTHS, made by EXECUTING the unmodified
reference (/root/reference; authoring container only, CPU, fp32).

    python tests/golden/make_golden_full.py [name ...]

The full 1024x1024 networks have 18.4 M parameters each and the gradients are as large, so neither is stored.
* Parameters and inputs are regenerated from seeds on both sides: oracle.make_*_params(seed) (torch CPU generator:
  bit-reproducible for a given torch build) are loaded into the reference's modules here and into the CUDA-path
  modules in the tests; inputs come from torch.Generator().manual_seed(SEED).
* Of every result the file keeps: losses and scores in full; for the fake image and for every parameter gradient its
  L2 norm and its values at up to 2048 seeded positions (`sample_index`), all of it when it is smaller -- an unbiased sample
  of the per-tensor relative error ||a-b|| / ||b|| the tolerance is defined on.
* Everything is stored twice: as the reference computes it (fp32, keys as above) and from the SAME reference modules
  converted with .double() on the same inputs (keys prefixed 'f64/').  LeakyReLU makes every gradient a discontinuous
  function of the forward values, so the reference's fp32 result has a noise floor of its own -- units whose
  pre-activation lies within fp32 rounding of zero take the other slope under another summation order.  The distance
  between the two copies, per tensor, IS that floor, measured on the reference itself; the GPU tests hold the CUDA
  path to max(1e-3, 3 x floor) against the fp64 copy, and report both numbers.

Files: tests/golden/full_<name>.npz (a few hundred KB each).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))

SAMPLES = 2048
G_SEED, D_SEED = 1337, 1338

# name: (model resolution, channels, depth, alpha, batch, input seed)
CONFIGS = {
    'd4_a05_n2': (1024, 3, 4, 0.5, 2, 21),      # c2's network and phase (64x64 fade-in), 512-channel K=4608 layers
    'd4_a05_n4': (1024, 3, 4, 0.5, 4, 22),
    'd5_c1_n2': (128, 1, 5, 1.0, 2, 23),        # c5's 1-channel 128x128 model
    'd6_a1_n1': (1024, 3, 6, 1.0, 1, 24),       # c3's phase
    'd8_a03_n1': (1024, 3, 8, 0.3, 1, 25),      # c4, the headline: 1024x1024 fade-in, 8/16-channel layers
}


def sample_index(name, numel):
    """The positions kept of a tensor of `numel` elements (all of them when numel <= SAMPLES)."""
    if numel <= SAMPLES:
        return torch.arange(numel)
    seed = sum((i + 1) * ord(c) for i, c in enumerate(name)) % (2 ** 31)
    return torch.randperm(numel, generator=torch.Generator().manual_seed(seed))[:SAMPLES].sort().values


def inputs(cfg):
    res, ch, depth, alpha, n, seed = cfg
    gen = torch.Generator().manual_seed(seed)
    r = 4 * 2 ** depth
    z1, z2 = torch.randn(n, 512, generator=gen), torch.randn(n, 512, generator=gen)
    real = torch.randn(n, ch, r, r, generator=gen)
    mix = torch.rand(n, 1, generator=gen)
    return z1, z2, real, mix


def compact(out, key, t, prefix=''):
    t = t.detach().reshape(-1)
    out[prefix + key + '/norm'] = np.float64(t.double().norm())
    out[prefix + key + '/samples'] = t[sample_index(key, t.numel())].numpy()


def load_into(module, params):
    sd = {k: v for k, v in params.items() if not k.endswith('.c')}
    module.load_state_dict(sd)
    for name, m in module.named_modules():
        if m.__class__.__name__ == 'PGConv2d':
            m.c = params[name + '.c'].clone()      # a 0-d fp32 tensor, as network.py:19 leaves it


def run_reference(cfg, dtype, out, prefix):
    """The reference's modules (converted to `dtype`) on the seeded parameters / inputs; results into out[prefix + key]."""
    import network
    import wgan_gp_loss
    import pggan_oracle as O
    res, ch, depth, alpha, n, seed = cfg
    shape = (1000, ch, res, res)
    G, D = network.Generator(shape), network.Discriminator(shape)
    load_into(G, O.make_generator_params(res, ch, seed=G_SEED))
    load_into(D, O.make_discriminator_params(res, ch, seed=D_SEED))
    G.to(dtype)
    D.to(dtype)
    G.depth = D.depth = depth
    G.alpha = D.alpha = alpha
    z1, z2, real, mix = [t.to(dtype) for t in inputs(cfg)]
    with torch.no_grad():
        fake = G(z1)
        compact(out, 'fake', fake, prefix)
        out[prefix + 'd_real_scores'] = D(real).numpy()
        out[prefix + 'd_fake_scores'] = D(fake).numpy()
    # the reference draws the mixing factors with uniform_() into a module global (wgan_gp_loss.py:15-17): hand it a
    # buffer whose uniform_() leaves our seeded factors in place (and a seed gradient of the right dtype, :21-23)
    class Fixed(torch.Tensor):
        def uniform_(self, *a, **k):
            return self
    wgan_gp_loss.mixing_factors = mix.clone().as_subclass(Fixed)
    wgan_gp_loss.grad_outputs = torch.ones(n, 1, dtype=dtype)
    d_cost, d_real_loss, d_fake_loss = wgan_gp_loss.wgan_gp_D_loss(D, G, real, z1)
    assert torch.equal(torch.Tensor(wgan_gp_loss.mixing_factors), mix)
    d_cost.backward()
    out[prefix + 'd_cost'] = d_cost.detach().numpy()
    out[prefix + 'd_real_loss'] = d_real_loss.detach().numpy()
    out[prefix + 'd_fake_loss'] = d_fake_loss.detach().numpy()
    for k, p_ in D.named_parameters():
        if p_.grad is not None:
            compact(out, 'Dgrad.' + k, p_.grad, prefix)
    g_cost = wgan_gp_loss.wgan_gp_G_loss(G, D, z2)
    g_cost.backward()
    out[prefix + 'g_cost'] = g_cost.detach().numpy()
    for k, p_ in G.named_parameters():
        if p_.grad is not None:
            compact(out, 'Ggrad.' + k, p_.grad, prefix)
    return float(d_cost), float(g_cost)


def make(name):
    cfg = CONFIGS[name]
    res, ch, depth, alpha, n, seed = cfg
    out = {'meta': np.array([res, ch, depth, n, seed, G_SEED, D_SEED], dtype=np.int64), 'alpha': np.float64(alpha)}
    dc, gc = run_reference(cfg, torch.float32, out, '')
    dc64, gc64 = run_reference(cfg, torch.float64, out, 'f64/')
    keys = [k[:-5] for k in out if k.endswith('/norm') and not k.startswith('f64/')]
    floor = {}
    for k in keys:
        a, b = out[k + '/samples'].astype(np.float64), out['f64/' + k + '/samples']
        floor[k] = float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
    worst = max(floor, key=floor.get)
    np.savez_compressed(os.path.join(HERE, 'full_%s.npz' % name), **out)
    print(name, 'd_cost %.6f (f64 %.6f)  g_cost %.6f (f64 %.6f)  keys %d  fp32-vs-fp64 floor: median %.2e, worst %.2e (%s)'
          % (dc, dc64, gc, gc64, len(out), float(np.median(list(floor.values()))), floor[worst], worst), flush=True)


if __name__ == '__main__':
    import make_golden
    make_golden.install_shims()
    torch.set_num_threads(os.cpu_count() or 1)
    for nm in (sys.argv[1:] or list(CONFIGS)):
        make(nm)
