"""Development aid (GPU box): tensor-core kernels vs the CUDA-core kernels of the same library on random operands.
   python tools/tc_test.py [conv|fp16|thin|thin1|wthin|wgrad|all]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pggan_b200 as pg  # noqa: E402
from importlib import import_module  # noqa: E402

E = import_module('pggan-pytorch_b200.engine')
lib = pg._lib.load()
call = pg._lib.call
BF16 = torch.bfloat16
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
REF = torch.float32


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def conv_case(N, H, W, Cin, Cout, KS, P, act=1, mask=False, bias=True, scale=1.0, pos=False, fwd=True):
    g = torch.Generator(device='cuda').manual_seed(N * 1000 + H + Cin + Cout)
    x = torch.randn(N, Cin, H, W, device='cuda', generator=g)
    K = KS * KS * Cin
    wf = (torch.randn(K, Cout, device='cuda', generator=g) / K ** 0.5).contiguous()
    if KS == 3 and Cin in (8, 16, 32) and Cout in (8, 16, 32, 64):
        wt = torch.empty(3, lib.pgk_pack_thin_plane_elems(Cin, Cout), dtype=BF16, device='cuda')
        call('pgk_pack_thin', wf.data_ptr(), Cin, Cout, wt.data_ptr(), wt.stride(0), 3)
    else:
        wt = torch.empty(3, Cout, K, dtype=BF16, device='cuda')
        call('pgk_pack_operand', wf.data_ptr(), K, Cout, wt.data_ptr(), wt.stride(0), 3)
    w_tc = (wf, wt)
    if KS == 3 and Cin == 64 and Cout in (32, 64) and P == 1 and W % 128 == 0:
        # the one-plane mode's 64-channel layers on the thin kernel (engine.THIN64): fourth element = thin packing
        w64 = torch.zeros(1, lib.pgk_pack_thin_plane_elems(Cin, Cout), dtype=BF16, device='cuda')
        call('pgk_pack_thin', wf.data_ptr(), Cin, Cout, w64.data_ptr(), w64.stride(0), 1)
        w_tc = (wf, wt, None, w64)
    b = torch.randn(Cout, device='cuda', generator=g) if bias else None
    xp = E.PT.from_float(x, P)
    m = E.PT.from_float(torch.randn(N, Cout, H, W, device='cuda', generator=g), P) if mask else None
    posT = torch.randn(H * W * Cout, device='cuda', generator=g) if pos else None
    pos_s = torch.randn(N, device='cuda', generator=g) if pos else None
    outs = []
    for tcon in (0, 1):
        lib.pgk_set_tc(tcon)
        o = E.PT.empty(N, H, W, Cout, P, 'cuda')
        o.t.fill_(float('nan'))
        E.conv(xp, w_tc if tcon else (wf, wt), Cout, KS, o, bias=b, posT=posT, pos_s=pos_s, act=act, mask=m, scale=scale,
               fwd=fwd)
        torch.cuda.synchronize()
        outs.append(o.float())
    lib.pgk_set_tc(1)
    # fp64 reference
    w4 = wf.view(KS, KS, Cin, Cout).permute(3, 2, 0, 1).to(REF)
    ref = torch.nn.functional.conv2d(xp.float().to(REF), w4, b.to(REF) if bias else None, padding=KS // 2)
    if pos:
        ref = ref + pos_s.to(REF).view(N, 1, 1, 1) * posT.to(REF).view(H, W, Cout).permute(2, 0, 1)
    if act:
        ref = torch.nn.functional.leaky_relu(ref, 0.2)
    if mask:
        ref = ref * torch.where(m.float().to(REF) > 0, 1.0, 0.2)
    ref = ref * scale
    e_tc, e_simt = rel(outs[1], ref), rel(outs[0], ref)
    tol = {1: 2e-2, 2: 1e-4, 3: 2e-5}[P if fwd else min(P, 2)]
    flag = 'ok ' if e_tc < tol else 'BAD'
    print('%s conv N%d %dx%d %d->%d k%d P%d act%d mask%d pos%d: tc %.2e simt %.2e' % (flag, N, H, W, Cin, Cout, KS, P, act, mask, pos, e_tc, e_simt))
    return e_tc < tol


def conv_fp16_case(N, H, W, Cin, Cout, KS, P=3, pos=False):
    """pgk_cvt_fp16x2 + pgk_pack_operand_fp16 + pgk_conv_fp16 (forward conv on two IEEE-half operand planes, three
    products) against an fp64 reference, next to the six-product bf16 path on the same operands: the pre-activation
    error decides LeakyReLU masks, so it is printed before the activation (act=0) as max |err| / rms(ref) too."""
    g = torch.Generator(device='cuda').manual_seed(N * 999 + H + Cin + Cout)
    x = torch.randn(N, Cin, H, W, device='cuda', generator=g)
    K = KS * KS * Cin
    wf = (torch.randn(K, Cout, device='cuda', generator=g) * (2.0 / K) ** 0.5).contiguous()   # equalised-LR scale
    wt = torch.empty(3, Cout, K, dtype=BF16, device='cuda')
    call('pgk_pack_operand', wf.data_ptr(), K, Cout, wt.data_ptr(), wt.stride(0), 3)
    wh = torch.empty(2, Cout, K, dtype=torch.float16, device='cuda')
    call('pgk_pack_operand_fp16', wf.data_ptr(), K, Cout, wh.data_ptr(), wh.stride(0), 2)
    b = torch.randn(Cout, device='cuda', generator=g)
    posT = torch.randn(H * W * Cout, device='cuda', generator=g) if pos else None
    pos_s = torch.randn(N, device='cuda', generator=g) if pos else None
    xp = E.PT.from_float(x, P)
    xh = torch.empty((2, xp.N * xp.per), dtype=torch.float16, device='cuda')
    call('pgk_cvt_fp16x2', xp.ptr, xp.ps, xp.P, xp.N * xp.per, xh.data_ptr(), xh.stride(0))
    torch.cuda.synchronize()
    # the two half planes must reproduce the bf16 planes' value to 2^-22
    xv = xp.float().permute(0, 2, 3, 1).reshape(-1).double()
    e_cvt = float(((xh[0].double() + xh[1].double()) - xv).abs().max() / xv.abs().max())
    w_back = (wh[0].double() + wh[1].double()) / float(1 << 6)
    e_pack = float((w_back - wf.t().double()).abs().max() / wf.abs().max())
    o16 = E.PT.empty(N, H, W, Cout, P, 'cuda')
    o16.t.fill_(float('nan'))
    call('pgk_conv_fp16', xh.data_ptr(), xh.stride(0), N, H, W, Cin, Cout, KS, wh.data_ptr(), wh.stride(0), b.data_ptr(),
         None if posT is None else posT.data_ptr(), None if pos_s is None else pos_s.data_ptr(), 0, o16.ptr, P, o16.ps)
    o6 = E.PT.empty(N, H, W, Cout, P, 'cuda')
    o6.t.fill_(float('nan'))
    E.conv(xp, (wf, wt), Cout, KS, o6, bias=b, posT=posT, pos_s=pos_s, act=0, fwd=True)
    torch.cuda.synchronize()
    w4 = wf.view(KS, KS, Cin, Cout).permute(3, 2, 0, 1).double()
    ref = torch.nn.functional.conv2d(xp.float().double(), w4, b.double(), padding=KS // 2)
    if pos:
        ref = ref + pos_s.double().view(N, 1, 1, 1) * posT.double().view(H, W, Cout).permute(2, 0, 1)
    rms = float(ref.pow(2).mean().sqrt())
    e16, e6 = rel(o16.float(), ref), rel(o6.float(), ref)
    m16 = float((o16.float().double() - ref).abs().max()) / rms
    m6 = float((o6.float().double() - ref).abs().max()) / rms
    ok = e16 < 2e-5 and e_cvt < 2.0 ** -21 and e_pack < 2.0 ** -21
    print('%s conv_fp16 N%d %dx%d %d->%d k%d P%d pos%d: fp16x2 rel %.2e max/rms %.2e | bf16x3 rel %.2e max/rms %.2e | cvt %.1e pack %.1e'
          % ('ok ' if ok else 'BAD', N, H, W, Cin, Cout, KS, P, pos, e16, m16, e6, m6, e_cvt, e_pack))
    return ok


def pixelnorm_case(N, H, W, Cin, Cout, P):
    """pgk_conv with pn_r (bias + LeakyReLU + pixel norm, network.py:34-40) against conv2d -> leaky_relu -> pixel norm in
    PyTorch fp32; also checks the stored per-pixel factor."""
    g = torch.Generator(device='cuda').manual_seed(N * 77 + H + Cin + Cout)
    x = torch.randn(N, Cin, H, W, device='cuda', generator=g)
    K = 9 * Cin
    wf = (torch.randn(K, Cout, device='cuda', generator=g) / K ** 0.5).contiguous()
    if Cin in (8, 16, 32) and Cout in (8, 16, 32, 64):
        wt = torch.empty(3, lib.pgk_pack_thin_plane_elems(Cin, Cout), dtype=BF16, device='cuda')
        call('pgk_pack_thin', wf.data_ptr(), Cin, Cout, wt.data_ptr(), wt.stride(0), 3)
    else:
        wt = torch.empty(3, Cout, K, dtype=BF16, device='cuda')
        call('pgk_pack_operand', wf.data_ptr(), K, Cout, wt.data_ptr(), wt.stride(0), 3)
    b = torch.randn(Cout, device='cuda', generator=g)
    xp = E.PT.from_float(x, P)
    o = E.PT.empty(N, H, W, Cout, P, 'cuda')
    o.t.fill_(float('nan'))
    r = torch.full((N * H * W,), float('nan'), device='cuda')
    E.conv(xp, (wf, wt), Cout, 3, o, bias=b, act=1, fwd=True, pn_r=r)
    torch.cuda.synchronize()
    w4 = wf.view(3, 3, Cin, Cout).permute(3, 2, 0, 1).to(REF)
    h = torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(xp.float().to(REF), w4, b.to(REF), padding=1), 0.2)
    rr = torch.rsqrt(torch.mean(h * h, 1, keepdim=True) + 1e-8)
    e_y, e_r = rel(o.float(), h * rr), rel(r.view(N, 1, H, W), rr)
    tol = {1: 2e-2, 2: 1e-4, 3: 2e-5}[P]
    flag = 'ok ' if e_y < tol and e_r < tol else 'BAD'
    print('%s conv+pixelnorm N%d %dx%d %d->%d P%d: y %.2e r %.2e' % (flag, N, H, W, Cin, Cout, P, e_y, e_r))
    return e_y < tol and e_r < tol


def wgrad_case(N, H, W, Cin, Cout, KS, P, ngroups=1):
    # (the engine reads min(P, 2) planes for weight gradients)
    g = torch.Generator(device='cuda').manual_seed(N * 1000 + H + Cin + Cout)
    ntot = N * ngroups + 2
    x = torch.randn(ntot, Cin, H, W, device='cuda', generator=g)
    gg = torch.randn(ntot, Cout, H, W, device='cuda', generator=g)
    xp, gp = E.PT.from_float(x, P), E.PT.from_float(gg, P)
    K = KS * KS * Cin
    groups = [(1 + i * N, (ngroups - 1 - i) * N) for i in range(ngroups)]
    outs = []
    for tcon in (0, 1):
        lib.pgk_set_tc(tcon)
        dwp = torch.zeros(K, Cout, device='cuda')
        db = torch.zeros(Cout, device='cuda')
        E.wgrad(xp, gp, H, W, Cin, Cout, KS, 0, groups, N, dwp, db, [go for _, go in groups[:max(1, ngroups - 1)]])
        torch.cuda.synchronize()
        outs.append(torch.cat([dwp.flatten(), db]))
    lib.pgk_set_tc(1)
    xd, gd = xp.float().to(REF), gp.float().to(REF)
    ref = torch.zeros(K, Cout, dtype=REF, device='cuda')
    for xo, go in groups:
        xs = torch.nn.functional.pad(xd[xo:xo + N], (KS // 2,) * 4)
        gs = gd[go:go + N]
        for ky in range(KS):
            for kx in range(KS):
                patch = xs[:, :, ky:ky + H, kx:kx + W]
                tap = ky * KS + kx
                ref[tap * Cin:(tap + 1) * Cin] += torch.einsum('nchw,nohw->co', patch, gs)
    dbref = sum(gd[go:go + N].sum(dim=(0, 2, 3)) for _, go in groups[:max(1, ngroups - 1)])
    ref = torch.cat([ref.flatten(), dbref])
    e_tc, e_simt = rel(outs[1], ref), rel(outs[0], ref)
    tol = {1: 2e-2, 2: 1e-4, 3: 1e-4}[P]
    flag = 'ok ' if e_tc < tol else 'BAD'
    print('%s wgrad N%d x%d groups %dx%d %d->%d k%d P%d: tc %.2e simt %.2e' % (flag, N, ngroups, H, W, Cin, Cout, KS, P, e_tc, e_simt))
    return e_tc < tol


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else 'all'
    ok = True
    if what in ('conv', 'all'):
        ok &= conv_case(2, 16, 16, 64, 64, 3, 1)
        ok &= conv_case(2, 16, 16, 64, 64, 3, 3)
        ok &= conv_case(3, 4, 4, 128, 64, 3, 2, pos=True)
        ok &= conv_case(5, 8, 8, 64, 128, 3, 3, mask=True, act=0, bias=False, scale=0.25)
        ok &= conv_case(2, 32, 32, 128, 256, 3, 3)
        ok &= conv_case(1, 64, 64, 256, 512, 3, 1)
        ok &= conv_case(1, 64, 64, 256, 512, 3, 3)
        ok &= conv_case(2, 128, 128, 64, 16, 3, 3)
        ok &= conv_case(1, 256, 256, 64, 32, 3, 1, mask=True)
        ok &= conv_case(130, 1, 1, 512, 2048, 1, 3)
        ok &= conv_case(7, 1, 1, 1024, 64, 1, 3, mask=True, act=0)
        ok &= conv_case(3, 32, 32, 256, 512, 3, 3, mask=True, act=0, bias=False, fwd=False)
        ok &= conv_case(40, 16, 16, 512, 512, 3, 3)
        ok &= conv_case(40, 16, 16, 512, 512, 3, 1)
        ok &= conv_case(300, 4, 4, 64, 64, 3, 1)
    if what in ('fp16', 'all'):
        ok &= conv_fp16_case(2, 16, 16, 64, 64, 3)
        ok &= conv_fp16_case(3, 4, 4, 512, 512, 3, pos=True)
        ok &= conv_fp16_case(2, 32, 32, 256, 256, 3)
        ok &= conv_fp16_case(2, 64, 64, 128, 256, 3)
        ok &= conv_fp16_case(1, 128, 128, 64, 128, 3, P=2)
        ok &= conv_fp16_case(9, 1, 1, 512, 8192, 1)
        ok &= conv_fp16_case(5, 1, 1, 8192, 512, 1)
        ok &= conv_fp16_case(40, 16, 16, 512, 512, 3)
    if what in ('thin', 'all'):
        ok &= conv_case(1, 128, 128, 16, 16, 3, 1)
        ok &= conv_case(2, 256, 256, 16, 16, 3, 3)
        ok &= conv_case(1, 256, 256, 8, 8, 3, 1)
        ok &= conv_case(1, 256, 256, 8, 16, 3, 3, mask=True, act=0, bias=False, scale=0.25, fwd=False)
        ok &= conv_case(2, 128, 256, 32, 32, 3, 1, mask=True)
        ok &= conv_case(1, 256, 256, 32, 64, 3, 3)
        ok &= conv_case(1, 512, 512, 16, 8, 3, 1, mask=True, act=0, fwd=False)
        ok &= conv_case(3, 128, 128, 32, 16, 3, 2)
    if what in ('thin1', 'all'):
        # every (Cin, Npad) instance of the thin conv in the one-plane mode (the input-row-stationary flavour)
        for cin in (8, 16, 32):
            for cout in (8, 16, 32, 64):
                ok &= conv_case(2, 32, 256, cin, cout, 3, 1)
        ok &= conv_case(1, 64, 128, 8, 8, 3, 1, mask=True, act=0, bias=False, scale=0.25, fwd=False)
        ok &= conv_case(3, 8, 128, 16, 32, 3, 1, mask=True)
        ok &= conv_case(1, 1024, 1024, 8, 16, 3, 1)
    if what in ('wthin', 'all'):
        ok &= wgrad_case(1, 128, 128, 16, 16, 3, 1)
        ok &= wgrad_case(2, 256, 256, 16, 16, 3, 2, ngroups=2)
        ok &= wgrad_case(1, 256, 256, 8, 8, 3, 1, ngroups=3)
        ok &= wgrad_case(1, 128, 256, 8, 16, 3, 3)
        ok &= wgrad_case(2, 128, 128, 32, 32, 3, 1, ngroups=4)
        ok &= wgrad_case(1, 256, 256, 32, 64, 3, 1)
        ok &= wgrad_case(1, 512, 512, 16, 8, 3, 1)
        ok &= wgrad_case(2, 256, 256, 64, 32, 3, 1, ngroups=2)
        ok &= wgrad_case(1, 128, 128, 64, 16, 3, 2)
    if what in ('wgrad', 'all'):
        ok &= wgrad_case(4, 16, 16, 64, 64, 3, 1)
        ok &= wgrad_case(4, 16, 16, 64, 64, 3, 3)
        ok &= wgrad_case(16, 4, 4, 128, 64, 3, 3, ngroups=2)
        ok &= wgrad_case(8, 8, 8, 64, 128, 3, 2, ngroups=3)
        ok &= wgrad_case(2, 32, 32, 128, 256, 3, 3, ngroups=4)
        ok &= wgrad_case(2, 64, 64, 256, 512, 3, 1)
        ok &= wgrad_case(1, 128, 128, 64, 64, 3, 3, ngroups=2)
    print('ALL OK' if ok else 'FAILURES')
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
