"""Development aid (GPU box): time single conv / wgrad launches of the tensor-core kernels on the layer shapes of
a config.   python tools/tc_bench.py [P] [n]      (P planes, n = samples of the forward batch)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pggan_b200 as pg  # noqa: E402
from importlib import import_module  # noqa: E402

E = import_module('pggan-pytorch_b200.engine')
call = pg._lib.call
BF16 = torch.bfloat16


def time_ms(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def bench_conv(N, H, Cin, Cout, P, fwd):
    K = 9 * Cin
    x = E.PT.empty(N, H, H, Cin, P, 'cuda')
    x.t.normal_()
    wf = torch.randn(K, Cout, device='cuda')
    wt = torch.randn(3, Cout, K, device='cuda').to(BF16)
    o = E.PT.empty(N, H, H, Cout, P, 'cuda')
    m = None
    if not fwd:
        m = E.PT.empty(N, H, H, Cout, P, 'cuda')
        m.t.normal_()
    b = torch.randn(Cout, device='cuda') if fwd else None
    ms = time_ms(lambda: E.conv(x, (wf, wt), Cout, 3, o, bias=b, act=1 if fwd else 0, mask=m, fwd=fwd))
    fl = 2.0 * N * H * H * Cout * K
    nprod = {1: 1, 2: 3, 3: 6}[P if fwd else min(P, 2)]
    print('conv %-4s N%-4d %3dx%-3d %3d->%-3d P%d: %7.3f ms  %6.1f TF/s algorithmic  %6.1f TF/s tensor-core products'
          % ('fwd' if fwd else 'bwd', N, H, H, Cin, Cout, P, ms, fl / ms / 1e9, fl * nprod / ms / 1e9))


def bench_wgrad(N, H, Cin, Cout, P, ngroups=1):
    K = 9 * Cin
    x = E.PT.empty(N * ngroups, H, H, Cin, P, 'cuda')
    x.t.normal_()
    g = E.PT.empty(N * ngroups, H, H, Cout, P, 'cuda')
    g.t.normal_()
    dwp = torch.zeros(K, Cout, device='cuda')
    groups = [(i * N, i * N) for i in range(ngroups)]
    ms = time_ms(lambda: E.wgrad(x, g, H, H, Cin, Cout, 3, 0, groups, N, dwp))
    fl = 2.0 * N * ngroups * H * H * Cout * K
    nprod = {1: 1, 2: 3, 3: 3}[P]
    print('wgrad    N%-4d %3dx%-3d %3d->%-3d P%d: %7.3f ms  %6.1f TF/s algorithmic  %6.1f TF/s tensor-core products'
          % (N * ngroups, H, H, Cin, Cout, P, ms, fl / ms / 1e9, fl * nprod / ms / 1e9))


if __name__ == '__main__':
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 384
    what = sys.argv[3] if len(sys.argv) > 3 else 'all'
    shapes = [(64, 128, 128), (64, 128, 256), (32, 256, 256), (32, 256, 512), (16, 512, 512), (8, 512, 512)]
    for H, ci, co in shapes:
        if what in ('all', 'conv'):
            bench_conv(n, H, ci, co, P, True)
            bench_conv(n, H, co, ci, P, False)
        if what in ('all', 'wgrad'):
            bench_wgrad(n // 3, H, ci, co, P, 4)
