"""Development aid: d(sum of -scores/n)/d(image) through the discriminator's u-chain vs oracle autograd (fp64)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from _util import rel_err  # noqa: E402
from _gpu_util import O, build_pair, pg  # noqa: E402
from importlib import import_module

engine = import_module('pggan-pytorch_b200.engine')


def run(depth, alpha, n, ch, res, fb, fm, lat, prec='fp32'):
    pgp = O.make_generator_params(res, ch, fmap_base=fb, fmap_max=fm, latent_size=lat, seed=3)
    pdp = O.make_discriminator_params(res, ch, fmap_base=fb, fmap_max=fm, seed=4)
    G, D = build_pair(dict(resolution=res, channels=ch, fmap_base=fb, fmap_max=fm, latent=lat, pg=pgp, pd=pdp,
                           depth=depth, alpha=alpha), precision=prec)
    gen = torch.Generator().manual_seed(7)
    r = 4 * 2 ** depth
    x = torch.randn(n, ch, r, r, generator=gen)
    nb = O.n_blocks_for(res)
    xd = x.double().requires_grad_(True)
    p64 = {k: v.double() for k, v in pdp.items()}
    sc = O.discriminator_forward(p64, xd, depth, alpha, nb)
    g_ref, = torch.autograd.grad((-sc).mean(), xd)
    ed = D.engine
    T = ed.forward(x.cuda(), 1, n, D.planes)
    seed = torch.full((n,), -1.0 / n, device='cuda')
    ed.backward_head(T, seed, None, None)
    ed.backward_body(T, T.d_hin, 0, n, 0)
    dimg = torch.empty(n, ch, r, r, device='cuda')
    ed.image_grad(T, 0, n, dimg)
    torch.cuda.synchronize()
    print('depth %d alpha %.2f n %d fm %d %s: scores %.2e  dimg %.2e' % (depth, alpha, n, fm, prec, rel_err(T.scores.view(-1, 1), sc),
                                                                rel_err(dimg, g_ref)))
    # intermediate: d_hin vs autograd at the last block's input
    return


if __name__ == '__main__':
    ns = [int(a) for a in sys.argv[1:]] or [4]
    for n in ns:
        for fm in (16, 64):
            for depth in (0, 1, 2, 3):
                for alpha in (1.0, 0.5):
                    if depth == 0 and alpha < 1:
                        continue
                    run(depth, alpha, n, 3, 32, 8 * fm, fm, fm)
