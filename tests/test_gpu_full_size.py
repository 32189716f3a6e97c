"""-m gpu: BASELINE.json's full sizes, through a property that needs no oracle.

The CPU oracle cannot run the full-width model at 64^2 x 128 or 1024^2 in test time, so the parity tests proper run
narrower models.  What can be checked at the real widths (512 ... 8 feature maps, all nine levels of the 1024^2
model) is that the hand-written backward passes -- including the gradient penalty's double backward through D -- are
the derivative of the loss the forward kernels compute: for a parameter tensor p with deposited gradient g and a
direction d,

        loss(p + eps d) - loss(p - eps d)  =  2 eps <g, d>  +  O(eps^3)

Both sides come from the CUDA path (fp32-faithful mode); the identity ties every backward kernel on the way to p --
wide and thin tensor-core convolutions, their weight gradients, pooling, masks, the stddev layer, fromRGB / toRGB with
the fade-in lerps -- to the forward kernels at the shapes bench.py measures.  Tolerance 1e-1 (measured 0.003 ... 2.5 %):
LeakyReLU makes the loss
piecewise smooth (units whose pre-activation crosses zero inside the +-eps segment bend it) and the loss value carries
fp32 rounding noise; a wrong or missing term shows up as O(1).

The gradient penalty is excluded from the finite differences (iwass_lambda = 0 there): it is a function of grad_x D,
which is piecewise CONSTANT in x, so the penalty JUMPS whenever a perturbation of the weights flips a LeakyReLU unit;
autograd (the reference) and the hand-written double backward both differentiate between the jumps, and a finite
difference sees them (measured: 0.6 ... 34 % apart, independent of eps).  The penalty's double backward is tied down
by an identity that involves no flips instead: D is positively homogeneous of degree one along
(W_l, b_l, every later bias [, the fade-in's low-resolution fromRGB]) -- scaling those by (1+e) scales D(x), and
grad_x D(x), by (1+e) for every x without changing any pre-activation's sign -- so Euler's theorem gives, for EVERY
layer l of the active path,

    <dL/dW_l, W_l> + sum_{j >= l} <dL/db_j, b_j> [+ low fromRGB terms]
        = mean( D(fake) - D(real) + 2 eps_drift D(real)^2 + 2 lambda (|g| - T) |g| / T^2 )

whose right-hand side needs only forward quantities.  Likewise every generator conv is followed by the pixel norm, so
G_cost does not change when (W_l, b_l) are scaled: <dG/dW_l, W_l> + <dG/db_l, b_l> = 0.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-1


@pytest.fixture(scope='module')
def pg():
    from _gpu_util import pg
    return pg


def _direction(g, gen):
    """Positively correlated with the gradient (so <g, d> is far above the rounding noise of the loss) but with random
    weights on every component (so an error in any component moves <g, d> at first order)."""
    w = torch.rand(g.shape, device=g.device, generator=gen) + 0.25
    d = torch.sign(g) * w
    return d / d.norm()


def _check(loss_fn, params, names, gen):
    """params: name -> parameter; loss_fn() -> (float loss after a fresh forward+backward, grads by name)."""
    base, grads = loss_fn(True)
    worst = {}
    for name in names:
        p, g = params[name], grads[name]
        assert g is not None and float(g.norm()) > 0, name
        d = _direction(g, gen)
        slope = float((g.double() * d.double()).sum())
        # step: 0.3 % of the tensor's norm, but no more than what changes the loss by ~2 % (keeps the segment short
        # enough for the piecewise-linear network, long enough for the fp32 loss value)
        eps = min(3e-3 * float(p.detach().norm()), 2e-2 * max(abs(base), 1e-3) / max(abs(slope), 1e-12))
        with torch.no_grad():
            p.add_(d, alpha=eps)
        lp, _ = loss_fn(False)
        with torch.no_grad():
            p.add_(d, alpha=-2 * eps)
        lm, _ = loss_fn(False)
        with torch.no_grad():
            p.add_(d, alpha=eps)
        fd = (lp - lm) / (2 * eps)
        worst[name] = abs(fd - slope) / max(abs(slope), 1e-12)
    return worst


def _models(pg, depth, alpha):
    torch.manual_seed(1234)
    shape = (1000, 3, 1024, 1024)
    G, D = pg.Generator(shape).cuda(), pg.Discriminator(shape).cuda()
    G.precision = D.precision = 'fp32'
    G.depth = D.depth = depth
    G.alpha = D.alpha = alpha
    return G, D


@pytest.mark.parametrize('depth,alpha,n', [(4, 0.5, 16), (8, 0.3, 1), (6, 1.0, 2)])
def test_d_step_gradients_are_the_derivative_of_d_cost(pg, depth, alpha, n):
    G, D = _models(pg, depth, alpha)
    gen = torch.Generator(device='cuda').manual_seed(99 + depth)
    r = 4 * 2 ** depth
    real = torch.randn(n, 3, r, r, device='cuda', generator=gen)
    z = torch.randn(n, 512, device='cuda', generator=gen)
    mix = torch.rand(n, 1, device='cuda', generator=gen).cpu()
    params = dict(D.named_parameters())

    def loss_fn(want_grads):
        pg.wgan_gp_loss.mixing_factors_override = mix
        try:
            cost, _, _ = pg.wgan_gp_D_loss(D, G, real, z, iwass_lambda=0.0)
            if want_grads:
                cost.backward()
        finally:
            pg.wgan_gp_loss.mixing_factors_override = None
        grads = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in params.items()} if want_grads else None
        return float(cost), grads

    nb = len(D.blocks)
    top, mid = nb - 1 - depth, nb - 1 - max(depth - 2, 0)     # blocks[-(depth+1)] and a block two levels below
    names = ['blocks.%d.fromRGB.conv.weight' % top, 'blocks.%d.c1.conv.weight' % top, 'blocks.%d.c2.conv.bias' % top,
             'blocks.%d.c2.conv.weight' % mid, 'blocks.%d.c1.conv.weight' % (nb - 1), 'linear.weight']
    if alpha < 1.0:
        names.append('blocks.%d.fromRGB.conv.weight' % (top + 1))
    worst = _check(loss_fn, params, names, gen)
    print({k: round(v, 5) for k, v in worst.items()})
    assert max(worst.values()) < TOL, worst


@pytest.mark.parametrize('depth,alpha,n', [(4, 0.5, 16), (8, 0.3, 1)])
def test_g_step_gradients_are_the_derivative_of_g_cost(pg, depth, alpha, n):
    G, D = _models(pg, depth, alpha)
    gen = torch.Generator(device='cuda').manual_seed(7 + depth)
    z = torch.randn(n, 512, device='cuda', generator=gen)
    params = dict(G.named_parameters())

    def loss_fn(want_grads):
        cost = pg.wgan_gp_G_loss(G, D, z)
        if want_grads:
            cost.backward()
        grads = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in params.items()} if want_grads else None
        return float(cost), grads

    names = ['block0.c1.conv.weight', 'block0.c2.conv.bias', 'blocks.%d.c1.conv.weight' % (depth - 1),
             'blocks.%d.c2.conv.weight' % (depth - 1), 'blocks.%d.toRGB.conv.weight' % (depth - 1),
             'blocks.%d.c2.conv.weight' % max(depth - 3, 0)]
    if alpha < 1.0:
        names.append(('blocks.%d.toRGB.conv.weight' % (depth - 2)) if depth >= 2 else 'block0.toRGB.conv.weight')
    worst = _check(loss_fn, params, names, gen)
    print({k: round(v, 5) for k, v in worst.items()})
    assert max(worst.values()) < TOL, worst


def _dot(a, b):
    return float((a.double() * b.double()).sum())


@pytest.mark.parametrize('depth,alpha,n', [(4, 0.5, 16), (8, 0.3, 1), (6, 1.0, 2), (5, 0.0, 3)])
def test_d_step_gradients_satisfy_eulers_identity_with_the_penalty(pg, depth, alpha, n):
    lam, eps_drift, target = 10.0, 0.001, 1.0
    G, D = _models(pg, depth, alpha)
    gen = torch.Generator(device='cuda').manual_seed(31 + depth)
    r = 4 * 2 ** depth
    real = torch.randn(n, 3, r, r, device='cuda', generator=gen)
    z = torch.randn(n, 512, device='cuda', generator=gen)
    mix = torch.rand(n, 1, device='cuda', generator=gen).cpu()
    with torch.no_grad():   # biases away from their small default init, so that their terms carry weight
        for name, p in D.named_parameters():
            if name.endswith('bias'):
                p.add_(0.1 * torch.randn(p.shape, device='cuda', generator=gen))
    pg.wgan_gp_loss.mixing_factors_override = mix
    try:
        cost, _, _ = pg.wgan_gp_D_loss(D, G, real, z, iwass_lambda=lam, iwass_epsilon=eps_drift, iwass_target=target)
        cost.backward()
    finally:
        pg.wgan_gp_loss.mixing_factors_override = None
    norms = pg.wgan_gp_loss.last_aux['grad_norms'].double()
    d_real, d_fake = D(real).double().view(-1), D(G(z)).double().view(-1)
    rhs = float((d_fake - d_real + 2 * eps_drift * d_real ** 2 + 2 * lam * (norms - target) * norms / target ** 2).mean())

    nb = len(D.blocks)
    fade = depth > 0 and alpha < 1.0
    top = nb - 1 - depth
    mods = [('blocks.%d.fromRGB' % top, True), ('blocks.%d.c1' % top, True), ('blocks.%d.c2' % top, True)]
    for k in range(top + 1, nb):
        mods += [('blocks.%d.c1' % k, False), ('blocks.%d.c2' % k, False)]
    par = dict(D.named_parameters())
    wname = lambda m: m + ('.conv.weight' if m != 'linear' else '.weight')
    bname = lambda m: m + ('.conv.bias' if m != 'linear' else '.bias')
    mods.append(('linear', False))
    bias_terms = [_dot(par[bname(m)].grad, par[bname(m)]) for m, _ in mods]
    low = 0.0
    if fade:
        lowm = 'blocks.%d.fromRGB' % (top + 1)
        low = _dot(par[wname(lowm)].grad, par[wname(lowm)]) + _dot(par[bname(lowm)].grad, par[bname(lowm)])
    scale = max(abs(rhs), max(abs(_dot(par[wname(m)].grad, par[wname(m)])) for m, _ in mods))
    worst = {}
    for i, (m, in_top) in enumerate(mods):
        lhs = _dot(par[wname(m)].grad, par[wname(m)]) + sum(bias_terms[i:]) + (low if (fade and in_top) else 0.0)
        worst[m] = abs(lhs - rhs) / scale
    print('rhs %.6f' % rhs, {k: round(v, 6) for k, v in worst.items()})
    assert max(worst.values()) < 2e-3, (rhs, worst)


@pytest.mark.parametrize('depth,alpha,n', [(4, 0.5, 16), (8, 0.3, 1)])
def test_g_step_gradients_are_orthogonal_to_the_scaling_the_pixel_norm_removes(pg, depth, alpha, n):
    G, D = _models(pg, depth, alpha)
    gen = torch.Generator(device='cuda').manual_seed(17 + depth)
    z = torch.randn(n, 512, device='cuda', generator=gen)
    with torch.no_grad():
        for name, p in G.named_parameters():
            if name.endswith('bias'):
                p.add_(0.1 * torch.randn(p.shape, device='cuda', generator=gen))
    pg.wgan_gp_G_loss(G, D, z).backward()
    par = dict(G.named_parameters())
    worst = {}
    for i in range(depth + 1):
        blk = 'block0' if i == 0 else 'blocks.%d' % (i - 1)
        for c in ('c1', 'c2'):
            w, b = par['%s.%s.conv.weight' % (blk, c)], par['%s.%s.conv.bias' % (blk, c)]
            lhs = _dot(w.grad, w) + _dot(b.grad, b)
            worst['%s.%s' % (blk, c)] = abs(lhs) / (float(w.grad.norm()) * float(w.norm()) + 1e-30)
    print({k: round(v, 6) for k, v in worst.items()})
    assert max(worst.values()) < 2e-3, worst
