"""Development aid: per k-slice clock trace of CTA 0 of one conv launch (producer waits / MMA-issue waits)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
buf = torch.zeros(2048, dtype=torch.int64, device='cuda')
os.environ['PGK_TRACE'] = str(buf.data_ptr())
import pggan_b200 as pg  # noqa: E402
from importlib import import_module  # noqa: E402

E = import_module('pggan-pytorch_b200.engine')
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1
N, H, Cin, Cout = 384, 64, 128, int(sys.argv[2]) if len(sys.argv) > 2 else 256
K = 9 * Cin
x = E.PT.empty(N, H, H, Cin, P, 'cuda')
x.t.normal_()
wf = torch.randn(K, Cout, device='cuda')
wt = torch.randn(3, Cout, K, device='cuda').to(torch.bfloat16)
o = E.PT.empty(N, H, H, Cout, P, 'cuda')
E.conv(x, (wf, wt), Cout, 3, o, fwd=True)
torch.cuda.synchronize()
buf.zero_()
E.conv(x, (wf, wt), Cout, 3, o, fwd=True)
torch.cuda.synchronize()
b = buf.cpu().tolist()
t0 = b[0]
print('MMA thread: chunk  wait_start  wait_end(+wait)  issued(+issue)   | producer: wait_start wait_end')
for i in range(60):
    m = b[4 * i:4 * i + 4]
    p = b[1024 + 2 * i:1024 + 2 * i + 2]
    print('%3d  %8d  %8d (+%5d)  mma_issued +%5d  %8d (+%4d)   | %8d %8d (+%5d)' % (i, m[0] - t0, m[1] - t0, m[1] - m[0], m[3] - m[1], m[2] - t0, m[2] - m[1],
                                                             p[0] - t0, p[1] - t0, p[1] - p[0]))
