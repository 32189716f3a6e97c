"""Generate tests/golden/ref_snapshot_{generator,discriminator}.dat + ref_snapshot.npz by EXECUTING the unmodified
reference (authoring container only): two whole-module pickles exactly as the reference's SaverPlugin writes them
(plugins.py:158-166, torch.save(model)), and what those modules compute for seeded inputs.

    python tests/golden/make_golden_snapshot.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import export_params, install_shims  # noqa: E402


def main():
    install_shims()
    import network
    torch.manual_seed(77)
    shape = (1000, 3, 16, 16)
    kw = dict(fmap_base=64, fmap_max=16)
    G = network.Generator(shape, latent_size=16, **kw)
    D = network.Discriminator(shape, **kw)
    G.depth = D.depth = 2
    G.alpha = D.alpha = 0.7
    torch.save(G, os.path.join(HERE, 'ref_snapshot_generator.dat'))
    torch.save(D, os.path.join(HERE, 'ref_snapshot_discriminator.dat'))
    gen = torch.Generator().manual_seed(78)
    z = torch.randn(5, 16, generator=gen)
    real = torch.randn(5, 3, 16, 16, generator=gen)
    with torch.no_grad():
        fake = G(z)
        scores = D(real)
    out = {'z': z.numpy(), 'real': real.numpy(), 'fake': fake.numpy(), 'scores': scores.numpy(),
           'depth': np.int64(2), 'alpha': np.float64(0.7)}
    out.update({'G.' + k: v for k, v in export_params(G).items()})
    out.update({'D.' + k: v for k, v in export_params(D).items()})
    np.savez_compressed(os.path.join(HERE, 'ref_snapshot.npz'), **out)
    print('wrote snapshots: G %d params, D %d params' % (sum(p.numel() for p in G.parameters()),
                                                        sum(p.numel() for p in D.parameters())))


if __name__ == '__main__':
    main()
