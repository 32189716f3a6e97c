"""Development aid: one launch of each thin-layer kernel shape of interest, for `ncu --set full` captures.
   ncu --set full --clock-control none --import-source on -k regex:'conv_thin|wgrad_direct|wgrad_thin' -o gpurun_out/thin python tools/thin_ncu.py"""
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
N, P = 12, 1
for H, ci, co, masked, do_wgrad in [(1024, 8, 8, False, True), (1024, 16, 8, True, False), (512, 16, 16, True, True),
                                    (256, 32, 64, False, False)]:
    K = 9 * ci
    x = E.PT.empty(N, H, H, ci, P, 'cuda')
    x.t.normal_()
    wf = torch.randn(K, co, device='cuda')
    wt = torch.empty(3, lib.pgk_pack_thin_plane_elems(ci, co), dtype=BF16, device='cuda')
    call('pgk_pack_thin', wf.data_ptr(), ci, co, wt.data_ptr(), wt.stride(0), 3)
    o = E.PT.empty(N, H, H, co, P, 'cuda')
    b = torch.randn(co, device='cuda')
    m = E.PT.empty(N, H, H, co, P, 'cuda')
    m.t.normal_()
    for _ in range(2):   # the second launch of each is the warm one
        if masked:
            E.conv(x, (wf, wt), co, 3, o, act=0, mask=m, fwd=True)
        else:
            E.conv(x, (wf, wt), co, 3, o, bias=b, act=1, fwd=True)
    if do_wgrad:
        g = E.PT.empty(N, H, H, co, P, 'cuda')
        g.t.normal_()
        dwp = torch.zeros(K, co, device='cuda')
        db = torch.zeros(co, device='cuda')
        for _ in range(2):
            E.wgrad(x, g, H, H, ci, co, 3, 0, [(0, 0)], N, dwp, db, [0])
    torch.cuda.synchronize()
print('done')
