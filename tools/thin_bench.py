"""Development aid: time the thin-layer kernels on the 1024^2 / 512^2 / 256^2 shapes of depth 8.
   python tools/thin_bench.py [planes] [batch]        (PGK_THIN_OCC=1|2 selects the CTAs-per-SM plan)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pggan_b200 as pg  # noqa: E402
from importlib import import_module  # noqa: E402

E = import_module('pggan-pytorch_b200.engine')
lib = pg._lib.load()
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


P = int(sys.argv[1]) if len(sys.argv) > 1 else 1
N = int(sys.argv[2]) if len(sys.argv) > 2 else 12
print('PGK_THIN_OCC=%s planes %d batch %d' % (os.environ.get('PGK_THIN_OCC', 'default'), P, N))
for H, ci, co in [(1024, 8, 8), (1024, 8, 16), (1024, 16, 8), (512, 16, 16), (512, 16, 32), (512, 32, 16),
                  (256, 32, 32), (256, 32, 64)]:
    K = 9 * ci
    x = E.PT.empty(N, H, H, ci, P, 'cuda')
    x.t.normal_()
    wf = torch.randn(K, co, device='cuda')
    wt = torch.empty(3, lib.pgk_pack_thin_plane_elems(ci, co), dtype=BF16, device='cuda')
    call('pgk_pack_thin', wf.data_ptr(), ci, co, wt.data_ptr(), wt.stride(0), 3)
    o = E.PT.empty(N, H, H, co, P, 'cuda')
    b = torch.randn(co, device='cuda')
    ms = time_ms(lambda: E.conv(x, (wf, wt), co, 3, o, bias=b, act=1, fwd=True))
    by = N * H * H * (ci + co) * 2 * P
    m = E.PT.empty(N, H, H, co, P, 'cuda')
    m.t.normal_()
    msm = time_ms(lambda: E.conv(x, (wf, wt), co, 3, o, act=0, mask=m, fwd=True))
    bym = by + N * H * H * co * 2
    g = E.PT.empty(N, H, H, co, P, 'cuda')
    g.t.normal_()
    dwp = torch.zeros(K, co, device='cuda')
    db = torch.zeros(co, device='cuda')
    ms2 = time_ms(lambda: E.wgrad(x, g, H, H, ci, co, 3, 0, [(0, 0)], N, dwp, db, [0]))
    print('%4dx%-4d %2d->%-2d: conv %.3f ms (%4.0f GB/s)  masked conv %.3f ms (%4.0f GB/s)  wgrad+bias %.3f ms (%4.0f GB/s)'
          % (H, H, ci, co, ms, by / ms / 1e6, msm, bym / msm / 1e6, ms2, by / ms2 / 1e6))
