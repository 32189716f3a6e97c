"""Prints a table of relative errors (CUDA path vs golden vectors / oracle) for every golden case.
Development aid for the GPU box:  python tools/gpu_diag.py [case ...]"""
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from _util import STEP_CASES, load_step, rel_err  # noqa: E402
from _gpu_util import O, build_pair, named_grads, pg  # noqa: E402


def row(name, a, b):
    e = rel_err(a, b)
    flag = '' if e < 1e-3 else '   <-- FAIL'
    print('  %-40s %.3e%s' % (name, e, flag))
    return e


def oracle_case(spec):
    """'o:depth,alpha,n,ch[,res,fmap_base,fmap_max,latent]' -> a golden-like dict computed by the oracle (fp64 when
    the spec ends in ',64')."""
    v = spec[2:].split(',')
    depth, alpha, n, ch = int(v[0]), float(v[1]), int(v[2]), int(v[3])
    res, fb, fm, lat = (int(x) for x in (v[4:8] if len(v) >= 8 else (32, 512, 64, 64)))
    dt = torch.float64 if len(v) > 8 and v[8] == '64' else torch.float32
    sg, sd, sx = (int(x) for x in v[9:12]) if len(v) >= 12 else (3, 4, 99)
    pgp = O.make_generator_params(res, ch, fmap_base=fb, fmap_max=fm, latent_size=lat, seed=sg)
    pdp = O.make_discriminator_params(res, ch, fmap_base=fb, fmap_max=fm, seed=sd)
    gen = torch.Generator().manual_seed(sx)
    r = 4 * 2 ** depth
    z1, z2 = torch.randn(n, lat, generator=gen), torch.randn(n, lat, generator=gen)
    real = torch.randn(n, ch, r, r, generator=gen)
    mix = torch.rand(n, 1, generator=gen)
    nb = O.n_blocks_for(res)
    c = lambda d: {k: t.to(dt) for k, t in d.items()}
    cost, rl, fl, gd = O.d_step_grads(c(pdp), c(pgp), real.to(dt), z1.to(dt), mix.to(dt), depth, alpha, nb)
    gcost, gg = O.g_step_grads(c(pgp), c(pdp), z2.to(dt), depth, alpha, nb)
    fake = O.generator_forward(c(pgp), z1.to(dt), depth, alpha)
    return dict(resolution=res, channels=ch, fmap_base=fb, fmap_max=fm, latent=lat, n=n, depth=depth, alpha=alpha,
                pg=pgp, pd=pdp, z1=z1, z2=z2, real=real, mixing=mix, fake=fake,
                d_real_scores=O.discriminator_forward(c(pdp), real.to(dt), depth, alpha, nb),
                d_fake_scores=O.discriminator_forward(c(pdp), fake, depth, alpha, nb),
                d_cost=cost, d_real_loss=rl, d_fake_loss=fl, dgrad=gd, g_cost=gcost, ggrad=gg)


def run_case(case, precision):
    g = oracle_case(case) if case.startswith('o:') else load_step(case)
    print('== %s  precision=%s  depth=%d alpha=%g N=%d' % (case, precision, g['depth'], g['alpha'], g['n']))
    G, D = build_pair(g, precision)
    worst = 0.0
    try:
        fake = G(g['z1'].cuda())
        worst = max(worst, row('G(z1)', fake, g['fake']))
        worst = max(worst, row('D(real)', D(g['real'].cuda()), g['d_real_scores']))
        worst = max(worst, row('D(fake_ref)', D(g['fake'].float().cuda()), g['d_fake_scores']))
    except Exception:
        traceback.print_exc()
    try:
        pg.wgan_gp_loss.mixing_factors_override = g['mixing']
        cost, rl, fl = pg.wgan_gp_D_loss(D, G, g['real'].cuda(), g['z1'].cuda())
        cost.backward()
        torch.cuda.synchronize()
        worst = max(worst, row('D_cost', cost, g['d_cost']))
        worst = max(worst, row('D_real_loss', rl, g['d_real_loss']))
        worst = max(worst, row('D_fake_loss', fl, g['d_fake_loss']))
        grads = named_grads(D)
        missing = set(g['dgrad']) - set(grads)
        extra = set(grads) - set(g['dgrad'])
        if missing or extra:
            print('  grad set mismatch: missing %s extra %s' % (sorted(missing), sorted(extra)))
        for k in sorted(g['dgrad']):
            if k in grads:
                worst = max(worst, row('dD/' + k, grads[k], g['dgrad'][k]))
    except Exception:
        traceback.print_exc()
    try:
        cost = pg.wgan_gp_G_loss(G, D, g['z2'].cuda())
        cost.backward()
        torch.cuda.synchronize()
        worst = max(worst, row('G_cost', cost, g['g_cost']))
        grads = named_grads(G)
        missing = set(g['ggrad']) - set(grads)
        extra = set(grads) - set(g['ggrad'])
        if missing or extra:
            print('  grad set mismatch: missing %s extra %s' % (sorted(missing), sorted(extra)))
        for k in sorted(g['ggrad']):
            if k in grads:
                worst = max(worst, row('dG/' + k, grads[k], g['ggrad'][k]))
    except Exception:
        traceback.print_exc()
    print('  worst: %.3e' % worst)
    return worst


if __name__ == '__main__':
    cases = sys.argv[1:] or STEP_CASES
    print(torch.cuda.get_device_name(0), torch.__version__)
    for prec in (os.environ.get('PGK_DIAG_PRECISION', 'fp32'),):
        for c in cases:
            run_case(c, prec)
