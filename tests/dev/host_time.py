"""Development aid (CPU, no GPU needed): how long does the HOST take to sequence one iteration?

Every libpgk call is replaced by a no-op and the torch memory operations that are asynchronous launches on the device
(zero fills, copies, the multi-tensor multiply of the gradient deposit) are switched off, so what remains is the Python
work of wgan_gp_D_loss + wgan_gp_G_loss: tensor allocation, argument marshalling, the engine's control flow.  On the
device each call additionally costs a ctypes transition and a launch (a few microseconds each).  If this time
approaches the GPU time of an iteration, the step is host-bound and only CUDA-graph replay (wgan_gp_loss.cuda_graphs)
or less Python per launch helps.

    python tests/dev/host_time.py [--profile]
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import pggan_b200 as pg  # noqa: E402
from importlib import import_module  # noqa: E402

E = import_module('pggan-pytorch_b200.engine')
L = import_module('pggan-pytorch_b200.wgan_gp_loss')
ncalls = [0]


def fake_call(name, *a):
    ncalls[0] += 1


for m in (E, L, pg._lib):
    setattr(m, 'call', fake_call)
for cls in (pg.Generator, pg.Discriminator):
    cls._input = lambda self, x: x
torch.zeros = lambda *a, **k: torch.empty(*a, **k)
torch.Tensor.zero_ = lambda self: self
torch.Tensor.copy_ = lambda self, src, non_blocking=False: self
torch.Tensor.clone = lambda self, *a, **k: self
torch._foreach_mul = lambda ts, s: list(ts)
torch._foreach_zero_ = lambda ts: None
torch.set_num_threads(1)


def main():
    G, D = pg.Generator((None, 3, 1024, 1024)), pg.Discriminator((None, 3, 1024, 1024))
    print('host sequencing time per iteration (kernels stubbed, device-side memory operations off):')
    for name, depth, alpha, n in (('c4', 8, 0.3, 4), ('c3', 6, 1.0, 32), ('c2', 4, 0.5, 128), ('c1', 0, 1.0, 16)):
        G.depth = D.depth = depth
        G.alpha = D.alpha = alpha
        r = 4 * 2 ** depth
        real, z = torch.empty(n, 3, r, r), torch.empty(n, 512)
        best = 1e9
        for _ in range(6):
            ncalls[0] = 0
            t0 = time.perf_counter()
            pg.wgan_gp_D_loss(D, G, real, z)
            pg.wgan_gp_G_loss(G, D, z)
            best = min(best, time.perf_counter() - t0)
        print('  %s depth %d batch %3d: %6.2f ms for %d libpgk calls (%.1f us per call)'
              % (name, depth, n, 1e3 * best, ncalls[0], 1e6 * best / ncalls[0]))
    if '--profile' in sys.argv:
        import cProfile
        import pstats
        G.depth = D.depth = 8
        G.alpha = D.alpha = 0.3
        real, z = torch.empty(4, 3, 1024, 1024), torch.empty(4, 512)
        pr = cProfile.Profile()
        pr.enable()
        pg.wgan_gp_D_loss(D, G, real, z)
        pg.wgan_gp_G_loss(G, D, z)
        pr.disable()
        pstats.Stats(pr).sort_stats('tottime').print_stats(25)


if __name__ == '__main__':
    main()
