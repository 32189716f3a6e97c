"""Development aid (CPU): what operand precision do the FORWARD passes need?  A model of the planes arithmetic on top
of the fp64 oracle.

LeakyReLU makes every gradient a discontinuous function of the forward values ("noise floor"): the
forward operands of the tensor-core kernels must carry enough bits that almost no unit lands on the other side of zero
than in the reference.  This script measures that directly: the oracle runs in fp64, but every PGConv2d's
pre-activation is REPLACED IN VALUE (not in the autograd graph) by what a given operand format would produce --

    value  = sum over the plane products of the scheme of  conv(plane_i(x), plane_j(c * w))   (exact fp64 products)
    planes = successive roundings of the value to the format:  p0 = rn(v), p1 = rn(v - p0), ...

so the LeakyReLU decisions (and the stored activations) are those of the scheme while the gradients flow through the
exact graph.  The D-step / G-step parameter gradients are then compared with the all-fp64 run:
||g - g64|| / ||g64|| per tensor, worst over tensors.  Schemes:

    fp32          operands rounded to fp32 (what the reference itself computes with; its own noise floor)
    bf16x3 (6)    three bf16 planes, products i + j <= 2           -- the fp32-faithful mode of the kernels today
    bf16x2 (3)    two bf16 planes, products i + j <= 1             -- measured too coarse on the GPU (2-3e-3)
    fp16x2 (3)    two fp16 planes, products i + j <= 1, weights scaled by 2^k into the normal range
    fp16x2 raw    the same without scaling the weights (their low plane falls into fp16 subnormals)
    bf16 / fp16   one plane (the bf16 mode, and what one fp16 plane would give instead)

    python tests/dev/precision_model.py [--res 32] [--depth 3] [--fmap-base 1024] [--fmap-max 128] [--n 4] [--seeds 3]
"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import pggan_oracle as O  # noqa: E402


def planes(v, dtype, n):
    """Successive roundings of an fp64 tensor to `dtype`: [p0, p1, ...], each returned as fp64."""
    out, rest = [], v
    for _ in range(n):
        p = rest.to(dtype).to(torch.float64)
        out.append(p)
        rest = rest - p
    return out


class Scheme(object):
    def __init__(self, name, dtype, nplanes, max_order, wscale_log2=0):
        self.name, self.dtype, self.nplanes, self.max_order, self.k = name, dtype, nplanes, max_order, wscale_log2

    def conv(self, x, wc, pad):
        """x: fp64 activation, wc: fp64 c * weight.  Returns the value the scheme's products add up to."""
        if self.dtype is None:
            return F.conv2d(x, wc, None, padding=pad)
        xs = planes(x, self.dtype, self.nplanes)
        ws = planes(wc * (2.0 ** self.k), self.dtype, self.nplanes)
        acc = None
        for i, xp in enumerate(xs):
            for j, wp in enumerate(ws):
                if i + j <= self.max_order:
                    t = F.conv2d(xp, wp, None, padding=pad)
                    acc = t if acc is None else acc + t
        return acc * (2.0 ** -self.k)

    def store(self, h):
        """The activation as the next layer reads it back: the sum of its stored planes."""
        if self.dtype is None:
            return h
        return sum(planes(h, self.dtype, self.nplanes))


SCHEMES = [
    Scheme('fp32', torch.float32, 1, 0),
    Scheme('bf16x3 (6 products)', torch.bfloat16, 3, 2),
    Scheme('bf16x2 (3 products)', torch.bfloat16, 2, 1),
    Scheme('fp16x2 (3 products), w*2^6', torch.float16, 2, 1, 6),
    Scheme('fp16x2 (3 products), raw w', torch.float16, 2, 1, 0),
    Scheme('fp16 (1 product), w*2^6', torch.float16, 1, 0, 6),
    Scheme('bf16 (1 product)', torch.bfloat16, 1, 0),
]


def install(scheme):
    """Replace oracle.pgconv by a version whose pre-activation VALUE is the scheme's (graph: exact fp64)."""
    orig = O.pgconv

    def pgconv(p, name, x, pad, act=True, pixelnorm=False):
        c, w, b = p[name + '.c'], p[name + '.conv.weight'], p[name + '.conv.bias']
        h = F.conv2d(x * c, w, b, stride=1, padding=pad)
        if scheme is not None and w.shape[-1] > 1:      # the 1x1 fromRGB / toRGB layers run in fp32 on CUDA cores
            with torch.no_grad():
                he = scheme.conv(x.detach(), (w * c).detach(), pad) + b.detach().view(1, -1, 1, 1)
            h = h + (he - h).detach()
        if act:
            h = F.leaky_relu(h, O.LRELU_SLOPE)
        if pixelnorm:
            h = O.pixel_norm(h)
        if scheme is not None and w.shape[-1] > 1:
            with torch.no_grad():
                hs = scheme.store(h.detach())
            h = h + (hs - h).detach()
        return h

    O.pgconv = pgconv
    return orig


def worst(ga, gb):
    out = 0.0
    for k, v in gb.items():
        den = float(v.norm())
        if den > 0:
            out = max(out, float((ga[k] - v).norm()) / den)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--res', type=int, default=32)
    ap.add_argument('--depth', type=int, default=3)
    ap.add_argument('--alpha', type=float, default=0.5)
    ap.add_argument('--fmap-base', type=int, default=1024)
    ap.add_argument('--fmap-max', type=int, default=128)
    ap.add_argument('--latent', type=int, default=128)
    ap.add_argument('--n', type=int, default=4)
    ap.add_argument('--seeds', type=int, default=3)
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    nb = O.n_blocks_for(args.res)
    rows = {s.name: [0.0, 0.0, 0.0] for s in SCHEMES}
    for seed in range(args.seeds):
        f64 = lambda d: {k: (v.double() if torch.is_tensor(v) else v) for k, v in d.items()}
        pgp = f64(O.make_generator_params(args.res, 3, args.fmap_base, 1.0, args.fmap_max, args.latent, seed=10 + seed))
        pdp = f64(O.make_discriminator_params(args.res, 3, args.fmap_base, 1.0, args.fmap_max, seed=20 + seed))
        gen = torch.Generator().manual_seed(seed)
        r = 4 * 2 ** args.depth
        real = torch.randn(args.n, 3, r, r, generator=gen).double()
        z1 = torch.randn(args.n, args.latent, generator=gen).double()
        z2 = torch.randn(args.n, args.latent, generator=gen).double()
        mix = torch.rand(args.n, 1, generator=gen).double()

        def run():
            cost, _, _, gd = O.d_step_grads(pdp, pgp, real, z1, mix, args.depth, args.alpha, nb)
            gcost, gg = O.g_step_grads(pgp, pdp, z2, args.depth, args.alpha, nb)
            return float(cost), gd, gg

        c64, gd64, gg64 = run()
        for s in SCHEMES:
            orig = install(s)
            try:
                c, gd, gg = run()
            finally:
                O.pgconv = orig
            row = rows[s.name]
            row[0] = max(row[0], abs(c - c64) / max(abs(c64), 1e-12))
            row[1] = max(row[1], worst(gd, gd64))
            row[2] = max(row[2], worst(gg, gg64))
    print('model %dx%d depth %d alpha %g, fmap_base %d max %d, batch %d, worst over %d seeds and all parameter tensors'
          % (args.res, args.res, args.depth, args.alpha, args.fmap_base, args.fmap_max, args.n, args.seeds))
    print('%-32s %12s %14s %14s' % ('forward operand scheme', 'D_cost', 'D-step grads', 'G-step grads'))
    for s in SCHEMES:
        print('%-32s %12.2e %14.2e %14.2e' % ((s.name,) + tuple(rows[s.name])))


if __name__ == '__main__':
    main()
