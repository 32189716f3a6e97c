"""What would PERFECT layer-by-layer kernels reach?  The roofline ceiling of one iteration (and of its D step) when every
3x3 / 1x1 / dense layer is one kernel that reads its operands and writes its result exactly once (bf16 activations,
fp32 accumulate): time = sum over layer passes of max(FLOPs / tensor peak, bytes / HBM peak).

    python tools/roofline_ceiling.py [--depth 8] [--batch 4] [--ch 3] [--fade 1]

Passes per iteration as the engine runs them: D step = forward of [real|fake|mixed] (3N samples), u-chain
data gradients (3N), v-chain forward (N), w-chain data gradients (N), weight gradients over four sample groups (4N),
plus one G forward (N); G step = G forward + D forward + D data gradients + G data / weight gradients (N each).
Elementwise passes (pooling, masks, pixel norm, ...) are NOT counted: a perfect implementation fuses them away, so the
figure is an upper bound for any implementation that keeps one kernel per conv layer and stores every activation once.
"""
import argparse
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def nf(stage, fmap_base=4096, fmap_max=512):
    return min(int(fmap_base / (2.0 ** stage)), fmap_max)


def layers(depth, ch, fade):
    """(name, res, cin, cout, k) of D's and G's conv layers active at `depth` (dense 4x4 layers as k = 4 at res 1)."""
    D, G = [], []
    G.append(('G.block0.c1', 1, nf(0), nf(1) * 16, 1))
    G.append(('G.block0.c2', 4, nf(1), nf(1), 3))
    for j in range(1, depth + 1):
        r = 4 * 2 ** j
        G.append(('G.b%d.c1' % j, r, nf(j), nf(j + 1), 3))
        G.append(('G.b%d.c2' % j, r, nf(j + 1), nf(j + 1), 3))
    r = 4 * 2 ** depth
    G.append(('G.toRGB', r, nf(depth + 1), ch, 1))
    D.append(('D.fromRGB', r, ch, nf(depth + 1), 1))
    if fade:
        G.append(('G.toRGB_prev', r // 2, nf(depth), ch, 1))
        D.append(('D.fromRGB_prev', r // 2, ch, nf(depth), 1))
    for j in range(depth, 0, -1):
        r = 4 * 2 ** j
        D.append(('D.b%d.c1' % j, r, nf(j + 1), nf(j + 1), 3))
        D.append(('D.b%d.c2' % j, r, nf(j + 1), nf(j), 3))
    D.append(('D.last.c1', 4, nf(1), nf(1), 3))
    D.append(('D.last.c2', 1, nf(1) * 16, nf(0), 1))
    return D, G


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--depth', type=int, default=8)
    ap.add_argument('--batch', type=int, default=4)
    ap.add_argument('--ch', type=int, default=3)
    ap.add_argument('--fade', type=int, default=1)
    ap.add_argument('--tensor-tflops', type=float, default=0.0)
    ap.add_argument('--hbm-gbs', type=float, default=0.0)
    args = ap.parse_args()
    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
    except Exception:
        pass
    tf = (args.tensor_tflops or float(peaks.get('bf16_tflops_sustained', 1400.0))) * 1e12
    bw = (args.hbm_gbs or float(peaks.get('hbm_gbs', 6650.0))) * 1e9
    n = args.batch
    D, G = layers(args.depth, args.ch, args.fade)

    def t(name, res, cin, cout, k, samples, passes):
        """one kind of pass over a layer: flops and bytes for `samples` samples; x and y in bf16, weights negligible
        at high resolution but counted (bf16)."""
        px = res * res * samples
        flops = 2.0 * px * cin * cout * k * k
        by = 2.0 * px * (cin + cout) + 2.0 * cin * cout * k * k
        return passes * max(flops / tf, by / bw), passes * flops, passes * by

    rows, d_time, d_flops, tot_time, tot_flops = [], 0.0, 0.0, 0.0, 0.0
    for (name, res, cin, cout, k) in D:
        # D step: fwd 3N, dgrad 3N (the input gradient of fromRGB only for the mixed third: counted as N), v fwd N,
        # w dgrad N, wgrad 4N (reads x and g, writes nothing big);  G step: fwd N, dgrad N
        a = t(name, res, cin, cout, k, 3 * n, 1)
        b = t(name, res, cin, cout, k, (n if 'fromRGB' in name else 3 * n), 1)
        c = t(name, res, cin, cout, k, n, 2)
        w = t(name, res, cin, cout, k, 4 * n, 1)
        g = t(name, res, cin, cout, k, n, 2)
        dt = a[0] + b[0] + c[0] + w[0]
        df = a[1] + b[1] + c[1] + w[1]
        d_time += dt
        d_flops += df
        tot_time += dt + g[0]
        tot_flops += df + g[1]
        rows.append((name, res, cin, cout, dt + g[0], (df + g[1]) / (dt + g[0]) / tf))
    for (name, res, cin, cout, k) in G:
        f1 = t(name, res, cin, cout, k, n, 1)            # D step: forward only
        f2 = t(name, res, cin, cout, k, n, 3)            # G step: forward, data gradient, weight gradient
        d_time += f1[0]
        d_flops += f1[1]
        tot_time += f1[0] + f2[0]
        tot_flops += f1[1] + f2[1]
        rows.append((name, res, cin, cout, f1[0] + f2[0], (f1[1] + f2[1]) / (f1[0] + f2[0]) / tf))
    print('depth %d, batch %d/GPU, tensor peak %.0f TF/s, HBM %.0f GB/s' % (args.depth, n, tf / 1e12, bw / 1e9))
    print('%-18s %5s %4s %5s %10s %8s' % ('layer', 'res', 'cin', 'cout', 'us/iter', 'f_tensor'))
    for r in rows:
        print('%-18s %5d %4d %5d %10.1f %8.2f' % (r[0], r[1], r[2], r[3], r[4] * 1e6, r[5]))
    print('iteration: %.3f ms at the layer-wise roofline = %.0f images/s per GPU; tensor fraction of that ceiling %.2f'
          % (tot_time * 1e3, n / tot_time, tot_flops / tot_time / tf))
    print('D step:    %.3f ms at the layer-wise roofline; tensor fraction of that ceiling %.2f'
          % (d_time * 1e3, d_flops / d_time / tf))


if __name__ == '__main__':
    main()
