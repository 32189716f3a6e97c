"""CPU oracle for the Progressive-GAN G + D + WGAN-GP training step.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker (or the timed CPU baseline), never as the thing shipped.

It is a *functional restatement* (state-dict in, tensors out; torch CPU ops,
fp32 or fp64) of the reference's algorithm for the hot path.  The reference
keeps the arithmetic in ``nn.Module`` objects that call PyTorch; here the same
arithmetic is written as plain functions over a parameter dictionary so that
the CUDA path, which shares the parameter names, can be compared tensor by
tensor.  Each function cites the reference lines it follows.

Parity pin: the reference ships no tests or golden vectors ("parity unpinned"
by the reference's own suite), so the pin is produced by *executing* the
unmodified reference here: ``tests/golden/make_golden.py`` imports
``/root/reference`` modules, runs them on seeded inputs and stores the results
in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this file
against those vectors.

Parameter dictionary keys (identical to the reference ``state_dict`` names,
plus ``<conv>.c`` for the equalised-LR constant that the reference keeps as a
plain attribute, network.py:19):

  G: block0.{c1,c2,toRGB}.conv.{weight,bias}, blocks.{i}.{c1,c2,toRGB}.conv.*
  D: blocks.{i}.{fromRGB,c1,c2}.conv.{weight,bias}, linear.{weight,bias}
"""
import math

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.2      # network.py:27
PN_EPS = 1e-8          # network.py:23,38-39
STD_EPS = 1.0e-8       # network.py:175


# --------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------
def nf(stage, fmap_base=4096, fmap_decay=1.0, fmap_max=512):
    """Feature-map count of a stage (network.py:94-95, 207-208)."""
    return min(int(fmap_base / (2.0 ** (stage * fmap_decay))), fmap_max)


def pixel_norm(h, eps=PN_EPS):
    """h * rsqrt(mean_c(h^2) + eps)  (network.py:37-40, 119-123)."""
    return h * torch.rsqrt(torch.mean(h * h, 1, keepdim=True) + eps)


def pgconv(p, name, x, pad, act=True, pixelnorm=False):
    """PGConv2d.forward (network.py:32-41): scale the INPUT by c, conv + bias,
    LeakyReLU(0.2), optional pixel norm -- in that order."""
    c = p[name + '.c']
    h = F.conv2d(x * c, p[name + '.conv.weight'], p[name + '.conv.bias'], stride=1, padding=pad)
    if act:
        h = F.leaky_relu(h, LRELU_SLOPE)
    if pixelnorm:
        h = pixel_norm(h)
    return h


def upsample2(h):
    """F.upsample(h, scale_factor=2), default nearest mode (network.py:127,129)."""
    return h.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def minibatch_stddev(x):
    """MinibatchStddev.forward + Tstdeps (network.py:174-187): ONE scalar
    sqrt(mean((x - mean x)^2) + 1e-8) over the entire tensor, appended as an
    extra constant channel."""
    s = torch.sqrt(((x - x.mean()) ** 2).mean() + STD_EPS)
    return torch.cat((x, s.expand(x.size(0), 1, x.size(2), x.size(3))), dim=1)


# --------------------------------------------------------------------------
# networks
# --------------------------------------------------------------------------
def g_block_name(i):
    """Block i of the generator: 0 is block0, i>=1 is blocks[i-1] (network.py:108-112)."""
    return 'block0' if i == 0 else 'blocks.%d' % (i - 1)


def generator_forward(p, z, depth, alpha, normalize_latents=True, pixelnorm=True):
    """Generator.forward (network.py:118-139)."""
    h = z.unsqueeze(2).unsqueeze(3)
    if normalize_latents:
        h = pixel_norm(h)
    # GFirstBlock (network.py:52-57): c1 is 4x4 with pad 3 on the 1x1 input.
    h = pgconv(p, 'block0.c1', h, 3, pixelnorm=pixelnorm)
    h = pgconv(p, 'block0.c2', h, 1, pixelnorm=pixelnorm)
    if depth == 0:
        return pgconv(p, 'block0.toRGB', h, 0, act=False)
    for i in range(1, depth):                      # network.py:126-128
        h = upsample2(h)
        h = pgconv(p, g_block_name(i) + '.c1', h, 1, pixelnorm=pixelnorm)
        h = pgconv(p, g_block_name(i) + '.c2', h, 1, pixelnorm=pixelnorm)
    h = upsample2(h)                               # network.py:129
    last = g_block_name(depth)
    u = pgconv(p, last + '.c1', h, 1, pixelnorm=pixelnorm)
    u = pgconv(p, last + '.c2', u, 1, pixelnorm=pixelnorm)
    ult = pgconv(p, last + '.toRGB', u, 0, act=False)
    if alpha < 1.0:                                # network.py:131-137 (strict <)
        prev = pgconv(p, g_block_name(depth - 1) + '.toRGB', h, 0, act=False)
    else:
        prev = 0
    return prev * (1 - alpha) + ult * alpha        # network.py:138


def discriminator_forward(p, x, depth, alpha, n_blocks):
    """Discriminator.forward (network.py:225-240).  ``n_blocks`` = len(D.blocks);
    python index -(k) maps to n_blocks-k."""
    def blk(k):                                    # blocks[-k]
        return 'blocks.%d' % (n_blocks - k)

    def run_block(k, h, first):
        b = blk(k)
        if first:                                  # fromRGB: 1x1 + LeakyReLU (network.py:145,160)
            h = pgconv(p, b + '.fromRGB', h, 0)
        if k == 1:                                 # DLastBlock (network.py:165-171)
            h = minibatch_stddev(h)
            h = pgconv(p, b + '.c1', h, 1)
            return pgconv(p, b + '.c2', h, 0)      # 4x4 valid -> 1x1
        h = pgconv(p, b + '.c1', h, 1)
        return pgconv(p, b + '.c2', h, 1)

    h = run_block(depth + 1, x, True)
    if depth > 0:
        h = F.avg_pool2d(h, 2)
        if alpha < 1.0:                            # network.py:230-233
            lo = pgconv(p, blk(depth) + '.fromRGB', F.avg_pool2d(x, 2), 0)
            h = h * alpha + (1 - alpha) * lo
    for i in range(depth, 0, -1):                  # network.py:235-238
        h = run_block(i, h, False)
        if i > 1:
            h = F.avg_pool2d(h, 2)
    h = h.squeeze(-1).squeeze(-1)
    return F.linear(h, p['linear.weight'], p['linear.bias'])


# --------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------
def gradient_penalty(pd, real, fake, mixing, depth, alpha, n_blocks, iwass_lambda, iwass_target):
    """calc_gradient_penalty (wgan_gp_loss.py:13-33) with the per-sample mixing
    factors passed in (the reference draws them with uniform_(), :15-17)."""
    n = real.size(0)
    mixed = (real.reshape(n, -1) * (1 - mixing) + fake.reshape(n, -1) * mixing).reshape(real.shape)
    mixed = mixed.detach().requires_grad_(True)
    scores = discriminator_forward(pd, mixed, depth, alpha, n_blocks)
    g, = torch.autograd.grad(scores, mixed, torch.ones_like(scores), create_graph=True, retain_graph=True)
    g = g.reshape(n, -1)
    return ((g.norm(2, dim=1) - iwass_target) ** 2) * iwass_lambda / (iwass_target ** 2)


def d_loss(pd, pg, real, latents, mixing, depth, alpha, n_blocks,
           iwass_lambda=10.0, iwass_epsilon=0.001, iwass_target=1.0):
    """wgan_gp_D_loss (wgan_gp_loss.py:36-65).  Returns (D_cost, D_real_loss (N,1),
    D_fake_loss (N,1)).  D_cost keeps the reference's (N,1)+(N,1)+(N,) -> (N,N)
    broadcast before .mean() (:62)."""
    d_real = discriminator_forward(pd, real, depth, alpha, n_blocks)
    d_real_loss = -d_real + d_real ** 2 * iwass_epsilon
    with torch.no_grad():                          # fake is detached via .data (:51-52)
        fake = generator_forward(pg, latents, depth, alpha)
    d_fake_loss = discriminator_forward(pd, fake, depth, alpha, n_blocks)
    gp = gradient_penalty(pd, real, fake, mixing, depth, alpha, n_blocks, iwass_lambda, iwass_target)
    d_cost = (d_fake_loss + d_real_loss + gp).mean()
    return d_cost, d_real_loss, d_fake_loss


def g_loss(pg, pd, latents, depth, alpha, n_blocks):
    """wgan_gp_G_loss (wgan_gp_loss.py:68-74): mean(-D(G(z)))."""
    return (-discriminator_forward(pd, generator_forward(pg, latents, depth, alpha), depth, alpha, n_blocks)).mean()


def d_step_grads(pd, pg, real, latents, mixing, depth, alpha, n_blocks, **kw):
    """D_cost.backward() as in trainer.py:97-98: gradients of D_cost w.r.t. every
    D parameter that takes part (others stay None, as in the reference)."""
    names = [k for k in pd if not k.endswith('.c')]
    leaves = {k: pd[k].detach().clone().requires_grad_(True) for k in names}
    p = dict(pd)
    p.update(leaves)
    cost, rl, fl = d_loss(p, pg, real, latents, mixing, depth, alpha, n_blocks, **kw)
    grads = torch.autograd.grad(cost, [leaves[k] for k in names], allow_unused=True)
    return cost.detach(), rl.detach(), fl.detach(), {k: g for k, g in zip(names, grads) if g is not None}


def g_step_grads(pg, pd, latents, depth, alpha, n_blocks):
    """G_cost.backward() as in trainer.py:110-111, G parameters only (the D grads
    the reference also fills are discarded by the next D.zero_grad())."""
    names = [k for k in pg if not k.endswith('.c')]
    leaves = {k: pg[k].detach().clone().requires_grad_(True) for k in names}
    p = dict(pg)
    p.update(leaves)
    cost = g_loss(p, pd, latents, depth, alpha, n_blocks)
    grads = torch.autograd.grad(cost, [leaves[k] for k in names], allow_unused=True)
    return cost.detach(), {k: g for k, g in zip(names, grads) if g is not None}


# --------------------------------------------------------------------------
# parameter construction (same distributions as the reference's __init__)
# --------------------------------------------------------------------------
def _init_conv(p, name, cin, cout, k, gen, wscale=True):
    """PGConv2d.__init__ (network.py:8-30): kaiming-normal weight (fan_in, gain
    sqrt 2), c = sqrt(mean(w^2)) measured, w /= c; bias keeps Conv2d's default
    U(+-1/sqrt(fan_in))."""
    fan_in = cin * k * k
    w = torch.randn(cout, cin, k, k, generator=gen) * math.sqrt(2.0 / fan_in)
    c = torch.sqrt(torch.mean(w ** 2)) if wscale else torch.tensor(1.0)
    p[name + '.conv.weight'] = w / c
    p[name + '.conv.bias'] = (torch.rand(cout, generator=gen) * 2 - 1) / math.sqrt(fan_in)
    p[name + '.c'] = c.clone()


def make_generator_params(resolution, num_channels, fmap_base=4096, fmap_decay=1.0, fmap_max=512,
                          latent_size=512, seed=0):
    """Generator.__init__ layer table (network.py:76-116)."""
    gen = torch.Generator().manual_seed(seed)
    R = int(math.log2(resolution))
    assert resolution == 2 ** R and resolution >= 4
    f = lambda s: nf(s, fmap_base, fmap_decay, fmap_max)
    p = {}
    _init_conv(p, 'block0.c1', latent_size, f(1), 4, gen)
    _init_conv(p, 'block0.c2', f(1), f(1), 3, gen)
    _init_conv(p, 'block0.toRGB', f(1), num_channels, 1, gen)
    for i in range(2, R):
        b = 'blocks.%d' % (i - 2)
        _init_conv(p, b + '.c1', f(i - 1), f(i), 3, gen)
        _init_conv(p, b + '.c2', f(i), f(i), 3, gen)
        _init_conv(p, b + '.toRGB', f(i), num_channels, 1, gen)
    return p


def make_discriminator_params(resolution, num_channels, fmap_base=4096, fmap_decay=1.0, fmap_max=512, seed=1):
    """Discriminator.__init__ layer table (network.py:191-223)."""
    gen = torch.Generator().manual_seed(seed)
    R = int(math.log2(resolution))
    assert resolution == 2 ** R and resolution >= 4
    f = lambda s: nf(s, fmap_base, fmap_decay, fmap_max)
    p = {}
    j = 0
    for i in range(R - 1, 1, -1):
        b = 'blocks.%d' % j
        _init_conv(p, b + '.fromRGB', num_channels, f(i), 1, gen)
        _init_conv(p, b + '.c1', f(i), f(i), 3, gen)
        _init_conv(p, b + '.c2', f(i), f(i - 1), 3, gen)
        j += 1
    b = 'blocks.%d' % j
    _init_conv(p, b + '.fromRGB', num_channels, f(1), 1, gen)
    _init_conv(p, b + '.c1', f(1) + 1, f(1), 3, gen)
    _init_conv(p, b + '.c2', f(1), f(0), 4, gen)
    k = 1.0 / math.sqrt(f(0))                       # nn.Linear default init (network.py:219)
    p['linear.weight'] = (torch.rand(1, f(0), generator=gen) * 2 - 1) * k
    p['linear.bias'] = (torch.rand(1, generator=gen) * 2 - 1) * k
    return p


def n_blocks_for(resolution):
    return int(math.log2(resolution)) - 1


# --------------------------------------------------------------------------
# depth / alpha schedule (integer arithmetic: must be bit exact)
# --------------------------------------------------------------------------
def depth_schedule(cur_nimg, max_depth, lod_training_nimg=100 * 1000, lod_transition_nimg=100 * 1000,
                   minibatch_default=16, minibatch_overrides=None,
                   tick_kimg_default=20, tick_kimg_overrides=None):
    """DepthManager.iteration (plugins.py:57-81) as a pure function.
    Returns (depth, alpha, minibatch_size, tick_duration_nimg)."""
    if minibatch_overrides is None:
        minibatch_overrides = {6: 14, 7: 6, 8: 3}
    if tick_kimg_overrides is None:
        tick_kimg_overrides = {3: 10, 4: 10, 5: 5, 6: 2, 7: 2, 8: 1}
    full, rem = divmod(cur_nimg, lod_training_nimg + lod_transition_nimg)
    trans, rem = divmod(rem, lod_training_nimg)
    depth = min(max_depth, full + trans)
    alpha = rem / lod_transition_nimg if (trans > 0 and full + trans == depth) else 1.0
    return (depth, alpha, minibatch_overrides.get(depth, minibatch_default),
            tick_kimg_overrides.get(depth, tick_kimg_default) * 1000)


def lr_rampup(cur_nimg, lr_rampup_kimg=40):
    """train.py:151-156."""
    if cur_nimg < lr_rampup_kimg * 1000:
        q = max(0.0, 1 - cur_nimg / (lr_rampup_kimg * 1000))
        return math.exp(-q * q * 5.0)
    return 1.0


def adam_step(params, grads, state, lr, beta1=0.0, beta2=0.99, eps=1e-8):
    """torch.optim.Adam as wired by train.py:148-149,195 (betas (0, .99), eps 1e-8,
    no weight decay); parameters without a gradient are skipped (trainer.py:100)."""
    for k, g in grads.items():
        st = state.setdefault(k, {'step': 0, 'm': torch.zeros_like(g), 'v': torch.zeros_like(g)})
        st['step'] += 1
        st['m'] = beta1 * st['m'] + (1 - beta1) * g
        st['v'] = beta2 * st['v'] + (1 - beta2) * g * g
        bc1 = 1 - beta1 ** st['step']
        bc2 = 1 - beta2 ** st['step']
        denom = st['v'].sqrt() / math.sqrt(bc2) + eps
        params[k] = params[k] - (lr / bc1) * st['m'] / denom
    return params


# --------------------------------------------------------------------------
# real-image preparation (the step just before the path)
# --------------------------------------------------------------------------
def alpha_fade(datapoint, alpha):
    """DepthDataset.alpha_fade (dataset.py:109-113, 238-242), numpy: 2x2 box mean, nearest upsample, lerp by 1-alpha."""
    import numpy as np
    c, h, w = datapoint.shape
    t = datapoint.reshape(c, h // 2, 2, w // 2, 2).mean((2, 4)).repeat(2, 1).repeat(2, 2)
    return datapoint + (t - datapoint) * (1 - alpha)


def adjust_dynamic_range(data, range_in, range_out):
    """utils.adjust_dynamic_range (utils.py:24-30)."""
    if range_in != range_out:
        (min_in, max_in) = range_in
        (min_out, max_out) = range_out
        scale_factor = (max_out - min_out) / (max_in - min_in)
        data = (data - min_in) * scale_factor + min_out
    return data


def prepare_real(datapoint, alpha, range_in, range_out):
    """DepthDataset.__getitem__ after the pyramid lookup (dataset.py:60-67): fade iff alpha < 1, range, float32."""
    if alpha < 1.0:
        datapoint = alpha_fade(datapoint, alpha)
    return adjust_dynamic_range(datapoint, range_in, range_out).astype('float32')
