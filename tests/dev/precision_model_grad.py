"""Development aid (CPU): what operand precision do the GRADIENT CHAINS need?  The companion of precision_model.py
(which answers the question for the forward passes).

The data-gradient convolutions of the u-, v- and w-chains and the weight gradients read two bf16 planes
of both operands today and issue the three products of planes i + j <= 1.  Gradients are continuous in these
operands (the LeakyReLU masks come from the stored forward activations), so the question is only how the operand
rounding adds up along a chain.  This script measures it on top of the fp64 oracle: every 3x3 / 4x4 PGConv2d keeps its
exact fp64 forward, but its BACKWARD is computed from rounded operands --

    d_x = conv_transpose(planes(d_y), planes(c * w))        d_w = correlate(planes(x), planes(d_y))

with the operand VALUES replaced by what a scheme's plane products add up to and the autograd graph left exact, so
that the double backward of the gradient penalty (autograd.grad(create_graph=True), wgan_gp_loss.py:25-28) flows
through the same rounded operands, as it does through the kernels.  Compared: every parameter gradient of the D step
and of the G step against the all-fp64 run, ||g - g64|| / ||g64||, worst over tensors and seeds.  Schemes
(data-gradient operands d_y x w | weight-gradient operands x x d_y):

    bf16x2 x bf16x2 (3)   | bf16x2 x bf16x2 (3)     what the kernels do today
    bf16x2 x fp16 (2)     | bf16x2 x bf16x2 (3)     weights as ONE IEEE-half plane (times 2^6): two products
    bf16x2 x bf16x2 (3)   | fp16 x bf16x2 (2)       only the weight gradient's activations as one half plane
    bf16x2 x bf16 (2)     | bf16x2 x bf16x2 (3)     weights as one bf16 plane: two products, 8 bits
    bf16x2 x fp16 (2)     | fp16 x bf16x2 (2)       and the activations of the weight gradient as one half plane
    bf16 x bf16 (1)       | bf16 x bf16 (1)         what the bf16 mode does (for scale)

kind::f16 takes the A and B formats independently, so bf16 gradient planes against half weights is one instruction.

    python tests/dev/precision_model_grad.py [--res 32] [--depth 3] [--fmap-base 1024] [--fmap-max 128] [--n 4] [--seeds 3]
"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import pggan_oracle as O  # noqa: E402


def planes(v, dtype, n, k=0):
    """Successive roundings of (v * 2^k) to `dtype`, scaled back: [p0, p1, ...] as fp64."""
    out, rest = [], v * (2.0 ** k)
    for _ in range(n):
        p = rest.to(dtype).to(torch.float64)
        out.append(p * (2.0 ** -k))
        rest = rest - p
    return out


class Operand(object):
    """How one operand of a gradient product is read: `n` planes of `dtype` (scaled by 2^k before rounding)."""

    def __init__(self, dtype, n, k=0):
        self.dtype, self.n, self.k = dtype, n, k

    def __call__(self, v):
        return planes(v, self.dtype, self.n, self.k) if self.dtype is not None else [v]


def ste(t, value):
    """A tensor with `value` as its value and t's autograd graph (identity derivative)."""
    return t + (value - t).detach()


class Scheme(object):
    def __init__(self, name, dgrad_g, dgrad_w, wgrad_x, wgrad_g):
        self.name, self.dg, self.dw, self.wx, self.wg = name, dgrad_g, dgrad_w, wgrad_x, wgrad_g

    @staticmethod
    def _pairs(a_planes, b_planes):
        """The plane pairs the kernels multiply: i + j < max(len) (1, 2 -> all of a against b0; 2, 2 -> i + j <= 1)."""
        order = max(len(a_planes), len(b_planes)) - 1
        return [(a, b) for i, a in enumerate(a_planes) for j, b in enumerate(b_planes) if i + j <= order]


def make_qconv(s):
    class QConv(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w, pad):
            ctx.save_for_backward(x, w)
            ctx.pad = pad
            return F.conv2d(x, w, None, padding=pad)

        @staticmethod
        def backward(ctx, g):
            x, w = ctx.saved_tensors
            pad = ctx.pad
            # exact graph (so that a second backward sees conv_transpose / correlate of its inputs) ...
            gx = torch.nn.grad.conv2d_input(x.shape, w, g, padding=pad)
            gw = torch.nn.grad.conv2d_weight(x, w.shape, g, padding=pad)
            # ... with the VALUES of what the scheme's plane products add up to
            with torch.no_grad():
                vx = sum(torch.nn.grad.conv2d_input(x.shape, b, a, padding=pad)
                         for a, b in Scheme._pairs(s.dg(g), s.dw(w)))
                vw = sum(torch.nn.grad.conv2d_weight(a, w.shape, b, padding=pad)
                         for a, b in Scheme._pairs(s.wx(x), s.wg(g)))
            return ste(gx, vx), ste(gw, vw), None

    return QConv


def install(scheme):
    orig = O.pgconv
    if scheme is None:
        return orig
    Q = make_qconv(scheme)

    def pgconv(p, name, x, pad, act=True, pixelnorm=False):
        c, w, b = p[name + '.c'], p[name + '.conv.weight'], p[name + '.conv.bias']
        if w.shape[-1] > 1:      # the 1x1 fromRGB / toRGB layers run in fp32 on CUDA cores
            h = Q.apply(x, w * c, pad) + b.view(1, -1, 1, 1)
        else:
            h = F.conv2d(x * c, w, b, stride=1, padding=pad)
        if act:
            h = F.leaky_relu(h, O.LRELU_SLOPE)
        if pixelnorm:
            h = O.pixel_norm(h)
        return h

    O.pgconv = pgconv
    return orig


BF, HF = torch.bfloat16, torch.float16
SCHEMES = [
    Scheme('bf16x2 x bf16x2 (3) | bf16x2 x bf16x2 (3)  [today]', Operand(BF, 2), Operand(BF, 2), Operand(BF, 2), Operand(BF, 2)),
    Scheme('bf16x2 x fp16*2^6 (2) | bf16x2 x bf16x2 (3)', Operand(BF, 2), Operand(HF, 1, 6), Operand(BF, 2), Operand(BF, 2)),
    Scheme('bf16x2 x bf16x2 (3) | fp16 x bf16x2 (2)', Operand(BF, 2), Operand(BF, 2), Operand(HF, 1), Operand(BF, 2)),
    Scheme('bf16x2 x bf16 (2) | bf16x2 x bf16x2 (3)', Operand(BF, 2), Operand(BF, 1), Operand(BF, 2), Operand(BF, 2)),
    Scheme('bf16x2 x fp16*2^6 (2) | fp16 x bf16x2 (2)', Operand(BF, 2), Operand(HF, 1, 6), Operand(HF, 1), Operand(BF, 2)),
    Scheme('bf16 x bf16 (1) | bf16 x bf16 (1)  [bf16 mode]', Operand(BF, 1), Operand(BF, 1), Operand(BF, 1), Operand(BF, 1)),
]


def worst(ga, gb):
    out, where = 0.0, ''
    for k, v in gb.items():
        den = float(v.norm())
        if den > 0:
            e = float((ga[k] - v).norm()) / den
            if e > out:
                out, where = e, k
    return out, where


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
    rows = {s.name: [0.0, (0.0, ''), (0.0, '')] for s in SCHEMES}
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
    print('%-54s %10s %12s %12s   %s' % ('dgrad operands (products) | wgrad operands (products)', 'D_cost', 'D grads',
                                        'G grads', 'worst tensors'))
    for s in SCHEMES:
        r_ = rows[s.name]
        print('%-54s %10.2e %12.2e %12.2e   %s / %s' % (s.name, r_[0], r_[1][0], r_[2][0], r_[1][1], r_[2][1]))


if __name__ == '__main__':
    main()
