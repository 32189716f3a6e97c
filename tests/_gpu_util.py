"""Helpers for the -m gpu tests: build the CUDA-path modules from golden / oracle parameter dictionaries."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))

import pggan_b200 as pg  # noqa: E402
import pggan_oracle as O  # noqa: E402


def load_params(module, params):
    """params: reference state_dict names + '<conv>.c'."""
    sd = {k: v for k, v in params.items() if not k.endswith('.c')}
    module.load_state_dict(sd)
    module.set_wscale({k[:-2]: float(v) for k, v in params.items() if k.endswith('.c')})


def build_pair(g, precision='fp32', device='cuda'):
    """g: a golden dict (tests/_util.load_step) or any dict with resolution/channels/fmap_base/fmap_max/latent/pg/pd."""
    shape = (1000, g['channels'], g['resolution'], g['resolution'])
    G = pg.Generator(shape, fmap_base=g['fmap_base'], fmap_max=g['fmap_max'], latent_size=g['latent'])
    D = pg.Discriminator(shape, fmap_base=g['fmap_base'], fmap_max=g['fmap_max'])
    load_params(G, g['pg'])
    load_params(D, g['pd'])
    G.cuda()
    D.cuda()
    G.precision = D.precision = precision
    if 'depth' in g:
        G.depth = D.depth = g['depth']
        G.alpha = D.alpha = g['alpha']
    return G, D


def named_grads(module):
    return {k: p.grad.detach().float().cpu() for k, p in module.named_parameters() if p.grad is not None}


def relu_margin(fn):
    """Run fn() (oracle code) and return the smallest |pre-activation| / rms over the LeakyReLU layers with fewer
    than 50k units.  LeakyReLU makes every gradient a discontinuous function of the forward values: a unit whose
    pre-activation lies within float rounding noise of zero (|v| ~ 1e-6 rms) can take the other slope under a
    different summation order -- on the CUDA path, but equally in the reference itself on another BLAS (the fp32 and
    fp64 oracles disagree on such units too).  In the small layers one flipped unit moves all upstream gradients by
    percent, so gradient-parity tests draw inputs whose margin is comfortably above that noise."""
    import torch.nn.functional as F
    margins = []
    orig = F.leaky_relu

    def hooked(x, *a, **k):
        if x.numel() < 50000:
            v = x.detach().double()
            margins.append(float((v.abs() / v.pow(2).mean().sqrt()).min()))
        return orig(x, *a, **k)

    F.leaky_relu = hooked
    try:
        fn()
    finally:
        F.leaky_relu = orig
    return min(margins) if margins else float('inf')


# ---- imposing the kernels' LeakyReLU decisions on the oracle ------------------------------------------------------
def d_masks(T, group, n):
    """The LeakyReLU decisions (stored activation > 0) of the CUDA D forward pass for sample group `group` of a tape,
    in the order in which oracle.discriminator_forward calls leaky_relu (network.py:225-240)."""
    sl = lambda t: (t.sl(group * n, (group + 1) * n).float() > 0).cpu()
    out = [sl(T.t0)]
    if T.depth > 0:
        out += [sl(T.t1), sl(T.t2)]
        if T.fade:
            out.append(sl(T.f))
        for rec in T.blocks:
            out += [sl(rec.a), sl(rec.b)]
    return out + [sl(T.l1), sl(T.l2)]


def g_masks(TG):
    """The same for the generator's tape (the pixel norm keeps the sign), in oracle.generator_forward's order."""
    out = []
    for _, _, a, b in TG.acts:
        out += [(a.float() > 0).cpu(), (b.float() > 0).cpu()]
    return out


class ForcedMasks(object):
    """Context manager: inside it the oracle's leaky_relu calls take their decisions from `queue` (one bool tensor per
    call, None = decide for yourself) instead of from the sign of their input.  LeakyReLU makes every gradient a
    discontinuous function of the forward values; with the decisions of the CUDA forward pass imposed, what is left
    between the two sides is rounding alone.  Counts the units whose own decision differs (`flips` of `units`)."""

    def __init__(self, queue):
        self.queue, self.i, self.flips, self.units = list(queue), 0, 0, 0

    def __enter__(self):
        import torch.nn.functional as F
        self.F, self.orig = F, F.leaky_relu

        def hooked(x, negative_slope=0.01, *a, **k):
            m = self.queue[self.i] if self.i < len(self.queue) else None
            self.i += 1
            if m is None:
                return self.orig(x, negative_slope, *a, **k)
            assert tuple(m.shape) == tuple(x.shape), (self.i, tuple(m.shape), tuple(x.shape))
            self.flips += int(((x.detach() > 0) != m).sum())
            self.units += m.numel()
            one = torch.ones((), dtype=x.dtype)
            return x * torch.where(m, one, one * negative_slope)

        F.leaky_relu = hooked
        return self

    def __exit__(self, *exc):
        self.F.leaky_relu = self.orig
        return False
