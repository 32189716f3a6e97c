"""Development aid (GPU box): per-SHAPE timing of every pgk_conv / pgk_wgrad launch of one training iteration.

bench.py's roofline aggregates per kernel family; this lists every distinct (entry point, N, H, W, Cin, Cout, KS,
planes read, mask) signature with its launch count, mean device time (CUDA events on the launching stream, around the
C-ABI call), algorithmic TFLOP/s and GB/s, and the share of the iteration -- the table from which the next kernel to
work on is picked (which shapes sit far below the tensor / HBM roofline, and how much of the step they are).

    python tools/shape_profile.py [--config c4] [--steps 3] [--warmup 2] [--batch N] [--top 40]

Honours every PGK_* switch (e.g. PGK_CONV_NT=64 python tools/shape_profile.py --config c3), so the same command A/Bs
a knob shape by shape.  Events around single launches include ~2-3 us of launch gap; shapes that small are marked.
"""
import argparse
import collections
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pggan_b200 as pg  # noqa: E402
from importlib import import_module  # noqa: E402

E = import_module('pggan-pytorch_b200.engine')


def signature(name, a):
    """Shape signature + algorithmic (flops, bytes) of one call, from the positional arguments of include/pgk.h."""
    if name == 'pgk_conv':
        (x, P, Pr, x_ps, N, H, W, Cin, Cout, KS, ups, wf, wt, wt_ps, bias, posT, pos_s, act, mask, mask_ps, scale, out,
         out_ps, pn_r) = a
        fl = 2.0 * N * H * W * Cout * KS * KS * Cin
        by = 2.0 * N * H * W * (Cin * Pr / (4 if ups else 1) + Cout * P + (Cout if mask else 0))
        return ('conv', N, H, W, Cin, Cout, KS, Pr, 'mask' if mask else ('pn' if pn_r else ('act' if act else '-')),
                'ups' if ups else ''), fl, by
    if name == 'pgk_conv_fp16':
        (xh, xh_ps, N, H, W, Cin, Cout, KS, wth, wth_ps, bias, posT, pos_s, act, out, P, out_ps) = a
        fl = 2.0 * N * H * W * Cout * KS * KS * Cin
        by = 2.0 * N * H * W * (Cin * 2 + Cout * P)
        return ('conv', N, H, W, Cin, Cout, KS, 2, 'act' if act else '-', 'fp16'), fl, by
    (x, x_ps, g, g_ps, P, Pr, H, W, Cin, Cout, KS, ups, ngroups, group_n, xoff, goff, dwp, db, bmask) = a
    n = ngroups * group_n
    fl = 2.0 * n * H * W * Cout * KS * KS * Cin
    by = 2.0 * n * H * W * (Cin / (4 if ups else 1) + Cout) * Pr
    return ('wgrad', n, H, W, Cin, Cout, KS, Pr, 'bias' if db else '-', 'ups' if ups else ''), fl, by


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='c4', choices=sorted(bench.CONFIGS))
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=2)
    ap.add_argument('--batch', type=int, default=0)
    ap.add_argument('--top', type=int, default=40)
    ap.add_argument('--json', default='', help='also write the table to this file')
    ap.add_argument('--others', action='store_true',
                    help='second table: every other libpgk entry point by (name, small integer arguments)')
    args = ap.parse_args()
    cfg = dict(bench.CONFIGS[args.config])
    if args.batch:
        cfg['n'] = args.batch
    depth, alpha, n, ch = cfg['depth'], cfg['alpha'], cfg['n'], cfg['ch']
    dev = torch.device('cuda', 0)
    torch.manual_seed(1337)
    np.random.seed(1337)
    shape = (1000, ch, cfg['res'], cfg['res'])
    G, D = pg.Generator(shape).to(dev), pg.Discriminator(shape).to(dev)
    G.precision = D.precision = cfg['precision']
    G.depth = D.depth = depth
    G.alpha = D.alpha = alpha
    opt_g = pg.FusedAdam(G.parameters(), 1e-3, betas=(0.0, 0.99))
    opt_d = pg.FusedAdam(D.parameters(), 1e-3, betas=(0.0, 0.99))
    gen = torch.Generator(device=dev).manual_seed(1337)
    r = 4 * 2 ** depth
    real = torch.randn(n, ch, r, r, device=dev, generator=gen)
    z1, z2 = torch.randn(n, 512, device=dev, generator=gen), torch.randn(n, 512, device=dev, generator=gen)

    def step():
        cost, _, _ = pg.wgan_gp_D_loss(D, G, real, z1)
        cost.backward()
        opt_d.step()
        gcost = pg.wgan_gp_G_loss(G, D, z2)
        gcost.backward()
        opt_g.step()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    # whole-iteration time without the per-launch events
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    step_ms = e0.elapsed_time(e1) / args.steps

    records = []
    real_call = pg._lib.call

    others = []

    def timed_call(name, *a):
        if name not in ('pgk_conv', 'pgk_wgrad', 'pgk_conv_fp16'):
            if not args.others:
                return real_call(name, *a)
            # signature: the integer arguments that are sizes / flags (pointers and strides are huge or None)
            sig = (name,) + tuple(v for v in a if isinstance(v, int) and not isinstance(v, bool) and 0 <= v < (1 << 16))
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = real_call(name, *a)
            e.record()
            others.append((sig, s, e))
            return r
        sig, fl, by = signature(name, a)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        real_call(name, *a)
        e.record()
        records.append((sig, fl, by, s, e))

    pg._lib.call = E.call = timed_call
    try:
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
    finally:
        pg._lib.call = E.call = real_call

    agg = collections.OrderedDict()
    for sig, fl, by, s, e in records:
        t = agg.setdefault(sig, [0, 0.0, 0.0, 0.0])
        t[0] += 1
        t[1] += s.elapsed_time(e)
        t[2] += fl
        t[3] += by
    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak_tf, peak_gb = float(peaks.get('bf16_tflops_sustained', 1400.0)), float(peaks.get('hbm_gbs', 6650.0))
    rows = []
    for sig, (cnt, ms, fl, by) in agg.items():
        tf, gb = fl / (ms * 1e-3) / 1e12, by / (ms * 1e-3) / 1e9
        prods = sig[7] * (sig[7] + 1) // 2
        rows.append(dict(sig=' '.join(str(v) for v in sig if v != ''), launches_per_step=cnt / args.steps,
                         ms_per_step=ms / args.steps, us_per_launch=1e3 * ms / cnt, tflops=tf, issued_tflops=tf * prods,
                         gbs=gb, frac_tensor=tf * prods / peak_tf, frac_hbm=gb / peak_gb,
                         share=ms / args.steps / step_ms))
    rows.sort(key=lambda d: -d['ms_per_step'])
    tot = sum(d['ms_per_step'] for d in rows)
    print('%s: %.2f ms / iteration; pgk_conv + pgk_wgrad %.2f ms (%.0f %%) in %d shapes'
          % (args.config, step_ms, tot, 100 * tot / step_ms, len(rows)))
    print('%-52s %6s %9s %9s %9s %8s %7s %7s %6s' % ('op N H W Cin Cout KS Pr epilogue', 'n/step', 'ms/step', 'us/launch',
                                                     'TFLOP/s', 'GB/s', 'f_tens', 'f_hbm', 'share'))
    for d in rows[:args.top]:
        print('%-52s %6.1f %9.3f %9.1f%s %8.1f %8.0f %7.2f %7.2f %5.1f%%'
              % (d['sig'], d['launches_per_step'], d['ms_per_step'], d['us_per_launch'],
                 '*' if d['us_per_launch'] < 12 else ' ', d['tflops'], d['gbs'], d['frac_tensor'], d['frac_hbm'],
                 100 * d['share']))
    print("(* = under 12 us per launch: the ~2-3 us gap between back-to-back launches is part of the figure;"
          " f_tens counts issued bf16 products against %.0f TF/s, f_hbm algorithmic bytes against %.0f GB/s)"
          % (peak_tf, peak_gb))
    if args.others:
        oa = collections.OrderedDict()
        for sig, s_, e_ in others:
            t = oa.setdefault(sig, [0, 0.0])
            t[0] += 1
            t[1] += s_.elapsed_time(e_)
        orows = sorted(oa.items(), key=lambda kv: -kv[1][1])
        print('other entry points: %.2f ms / iteration' % (sum(v[1] for v in oa.values()) / args.steps))
        print('%-70s %6s %9s %9s' % ('entry point + small integer arguments (include/pgk.h order)', 'n/step', 'ms/step', 'us/launch'))
        for sig, (cnt, ms) in orows[:args.top]:
            print('%-70s %6.1f %9.3f %9.1f' % (' '.join(str(v) for v in sig), cnt / args.steps, ms / args.steps, 1e3 * ms / cnt))
    if args.json:
        with open(args.json, 'w') as f:
            json.dump({'config': args.config, 'step_ms': step_ms, 'rows': rows}, f, indent=1)


if __name__ == '__main__':
    main()
