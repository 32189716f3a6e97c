"""CPU oracle for the bf16 mode of the CUDA path (precision='bf16': one bf16 plane per stored tensor).

TEST INFRASTRUCTURE ONLY (see oracle/pggan_oracle.py): imported by tests/ and tests/dev/ alone.

The reference (network.py, wgan_gp_loss.py) is fp32 only; BASELINE.json nevertheless quotes configs c3-c5 in bf16.
What "bf16" means on the CUDA path is stated here as arithmetic, so that it can be held to a real tolerance instead of
a direction check: the reference's algorithm (the functions of pggan_oracle.py, same citations) with a
round-to-nearest-even to bfloat16 at exactly the places where the kernels store a tensor:

  * the operands of every 3x3 / 4x4 convolution: activations as stored (below) and weights bf16(c * w) -- the
    equalised-LR constant is folded in before the rounding (csrc/pgk_elem.cu prep_weight_kernel + pgk_pack_operand);
    the 1x1 fromRGB / toRGB convolutions, the linear head and the constant minibatch-stddev channel of the last block
    use fp32 weights (c * w in fp32); every product is accumulated in fp32;
  * every stored activation, once, after the whole PGConv2d (bias, LeakyReLU, pixel norm) -- or twice where the pixel
    norm is a second in-place pass over the stored tensor (`pn_fused(name) == False`: the dense first layer and the
    layers of the wide kernel whose tile does not hold all channels of a pixel);
  * the pooled (and fade-in blended) tensor between two D blocks, the normalised latents;
  * every stored gradient: the gradient w.r.t. each conv's PRE-activation (`ua`), w.r.t. a block's input, and their
    adjoints in the penalty's double backward (the v / w chains) -- autograd places those automatically because the
    backward of a rounding node is again a rounding node;
  * images, scores, losses, the per-sample statistics, all parameter gradients: fp32, never rounded.

The rounding is straight-through for differentiation.  With the forward roundings at the kernels' own places the
LeakyReLU decisions of both sides are taken on (nearly) bit-identical pre-activations, which is what makes a tight
tolerance possible at all: against the fp32 oracle a bf16 run differs by per cent, because ~0.2 % of the units fall on
the other side of zero.  Where a stored gradient is rounded once on the GPU and at a slightly different place here
(before instead of after the multiplication by the LeakyReLU slope), the difference is one more bf16 rounding of the
same value: noise of relative size 2^-9 per element, not a discontinuity.
"""
import torch
import torch.nn.functional as F

import pggan_oracle as O


ROUND = True     # tests set this to False to check the restructured graph against pggan_oracle.py (must then be identical)


class _Round(torch.autograd.Function):
    """y = bf16(x) if fwd else x;  dx = bf16(dy) if bwd else dy  (and so on for higher derivatives)."""

    @staticmethod
    def forward(ctx, x, fwd, bwd):
        ctx.bwd = bwd
        return x.to(torch.bfloat16).to(x.dtype) if (fwd and ROUND) else x.clone()

    @staticmethod
    def backward(ctx, g):
        return (_Round.apply(g, ctx.bwd, ctx.bwd) if ctx.bwd else g), None, None


def r_fb(x):
    """stored activation whose gradient is stored too"""
    return _Round.apply(x, True, True)


def r_f(x):
    """rounded in the forward pass only (operands whose gradient stays fp32: weights; activations whose incoming
    gradient is consumed by a fused mask multiplication before it is stored)"""
    return _Round.apply(x, True, False)


def r_b(x):
    """identity whose gradient is stored in bf16 (a pre-activation)"""
    return _Round.apply(x, False, True)


def conv3(p, name, x, pad, pixelnorm=False, pn_fused=True, grad_of_output_stored=False, extra=None):
    """PGConv2d.forward (network.py:32-41) for the tensor-core layers: bf16(c*w) weights, fp32 accumulate, + bias
    (+ `extra`, the constant-channel term of the last D block), LeakyReLU, [pixel norm], store."""
    wc = r_f(p[name + '.conv.weight'] * p[name + '.c'])
    a = F.conv2d(x, wc, p[name + '.conv.bias'], stride=1, padding=pad)
    if extra is not None:
        a = a + extra
    h = F.leaky_relu(r_b(a), O.LRELU_SLOPE)
    store = r_fb if grad_of_output_stored else r_f
    if pixelnorm:
        if not pn_fused:
            h = r_f(h)                  # stored, then normalised in place by a second pass over the stored values
        h = O.pixel_norm(h)
    return store(h)


def conv1(p, name, x, act):
    """1x1 fromRGB (+LeakyReLU) / toRGB on the image surface: fp32 weights c*w (network.py:49,65,145,160)."""
    a = F.conv2d(x, p[name + '.conv.weight'] * p[name + '.c'], p[name + '.conv.bias'])
    if act:
        return r_f(F.leaky_relu(r_b(a), O.LRELU_SLOPE))
    return a


def generator_forward(p, z, depth, alpha, pn_fused=lambda name: True, normalize_latents=True, pixelnorm=True):
    """Generator.forward (network.py:118-139), bf16 mode."""
    h = z.unsqueeze(2).unsqueeze(3)
    if normalize_latents:
        h = O.pixel_norm(h)
    h = r_fb(h)
    h = conv3(p, 'block0.c1', h, 3, pixelnorm, False, True)       # dense GEMM, pixel norm as a second pass
    h = conv3(p, 'block0.c2', h, 1, pixelnorm, pn_fused('block0.c2'), True)
    if depth == 0:
        return conv1(p, 'block0.toRGB', h, False)
    hprev = h
    for i in range(1, depth + 1):
        b = O.g_block_name(i)
        hprev = h
        hu = r_b(O.upsample2(h))                                   # exact copy; its gradient is stored before the 2x2 sum
        u = conv3(p, b + '.c1', hu, 1, pixelnorm, pn_fused(b + '.c1'), True)
        h = conv3(p, b + '.c2', u, 1, pixelnorm, pn_fused(b + '.c2'), True)
    ult = conv1(p, O.g_block_name(depth) + '.toRGB', h, False)
    if alpha < 1.0:
        # toRGB_{d-1} of the upsampled features (network.py:131-135) == upsampled toRGB_{d-1}: a 1x1 conv commutes
        # with nearest-neighbour upsampling, the values are identical
        prev = O.upsample2(conv1(p, O.g_block_name(depth - 1) + '.toRGB', hprev, False))
    else:
        prev = 0
    return prev * (1 - alpha) + ult * alpha


def discriminator_forward(p, x, depth, alpha, n_blocks):
    """Discriminator.forward (network.py:225-240), bf16 mode."""
    blk = lambda k: 'blocks.%d' % (n_blocks - k)
    top = blk(depth + 1)
    h = conv1(p, top + '.fromRGB', x, True)
    if depth > 0:
        h = conv3(p, top + '.c1', h, 1)
        h = conv3(p, top + '.c2', h, 1)
        h = F.avg_pool2d(h, 2)
        if alpha < 1.0:
            lo = conv1(p, blk(depth) + '.fromRGB', F.avg_pool2d(x, 2), True)
            h = h * alpha + (1 - alpha) * lo
        h = r_fb(h)
        for k in range(depth, 1, -1):
            h = conv3(p, blk(k) + '.c1', h, 1)
            h = conv3(p, blk(k) + '.c2', h, 1)
            h = r_fb(F.avg_pool2d(h, 2))
    else:
        h = r_fb(h)          # depth 0: the fromRGB output is the last block's input (its gradient is stored)
    last = blk(1)
    # MinibatchStddev (network.py:174-187): one fp32 scalar over the stored values; its channel meets fp32 weights
    s = torch.sqrt(((h - h.mean()) ** 2).mean() + O.STD_EPS)
    w = p[last + '.c1.conv.weight'] * p[last + '.c1.c']
    C = w.shape[1] - 1
    sch = s.expand(h.size(0), 1, h.size(2), h.size(3))
    extra = F.conv2d(sch, w[:, C:], None, padding=1)
    wc = r_f(w[:, :C])
    a = F.conv2d(h, wc, p[last + '.c1.conv.bias'], padding=1) + extra
    h = r_f(F.leaky_relu(r_b(a), O.LRELU_SLOPE))
    h = conv3(p, last + '.c2', h, 0)
    h = h.squeeze(-1).squeeze(-1)
    return F.linear(h, p['linear.weight'], p['linear.bias'])


def _leaves(params):
    names = [k for k in params if not k.endswith('.c')]
    leaves = {k: params[k].detach().clone().requires_grad_(True) for k in names}
    p = dict(params)
    p.update(leaves)
    return names, leaves, p


def d_step_grads(pd, pg, real, latents, mixing, depth, alpha, n_blocks, pn_fused=lambda name: True,
                 iwass_lambda=10.0, iwass_epsilon=0.001, iwass_target=1.0):
    """wgan_gp_D_loss + backward (wgan_gp_loss.py:36-65, trainer.py:97-98) in the bf16 mode's arithmetic."""
    names, leaves, p = _leaves(pd)
    d_real = discriminator_forward(p, real, depth, alpha, n_blocks)
    d_real_loss = -d_real + d_real ** 2 * iwass_epsilon
    with torch.no_grad():
        fake = generator_forward(pg, latents, depth, alpha, pn_fused)
    d_fake_loss = discriminator_forward(p, fake, depth, alpha, n_blocks)
    n = real.size(0)
    mixed = (real.reshape(n, -1) * (1 - mixing) + fake.reshape(n, -1) * mixing).reshape(real.shape)
    mixed = mixed.detach().requires_grad_(True)
    scores = discriminator_forward(p, mixed, depth, alpha, n_blocks)
    g, = torch.autograd.grad(scores, mixed, torch.ones_like(scores), create_graph=True, retain_graph=True)
    gp = ((g.reshape(n, -1).norm(2, dim=1) - iwass_target) ** 2) * iwass_lambda / (iwass_target ** 2)
    cost = (d_fake_loss + d_real_loss + gp).mean()
    grads = torch.autograd.grad(cost, [leaves[k] for k in names], allow_unused=True)
    return (cost.detach(), d_real_loss.detach(), d_fake_loss.detach(),
            {k: v for k, v in zip(names, grads) if v is not None}, fake)


def g_step_grads(pg, pd, latents, depth, alpha, n_blocks, pn_fused=lambda name: True):
    """wgan_gp_G_loss + backward (wgan_gp_loss.py:68-74, trainer.py:110-111) in the bf16 mode's arithmetic."""
    names, leaves, p = _leaves(pg)
    cost = (-discriminator_forward(pd, generator_forward(p, latents, depth, alpha, pn_fused), depth, alpha,
                                   n_blocks)).mean()
    grads = torch.autograd.grad(cost, [leaves[k] for k in names], allow_unused=True)
    return cost.detach(), {k: v for k, v in zip(names, grads) if v is not None}
