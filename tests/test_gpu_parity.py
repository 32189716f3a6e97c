"""-m gpu: the CUDA path (through the C ABI) against the golden vectors produced by the reference and against
the oracle on fresh seeded inputs.  Tolerance (SURVEY.md 8c / BASELINE.json north_star): per-tensor
||a-b||_2 / ||b||_2 <= 1e-3 in the fp32-faithful mode; bf16 mode tolerances are stated per test."""
import os

import pytest
import torch

from _util import STEP_CASES, load_step, load_trainer, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-3


@pytest.fixture(scope='module')
def gpu():
    from _gpu_util import O, build_pair, named_grads, pg
    return dict(O=O, build_pair=build_pair, named_grads=named_grads, pg=pg)


@pytest.mark.parametrize('case', STEP_CASES)
def test_forward_golden(gpu, case):
    g = load_step(case)
    G, D = gpu['build_pair'](g)
    assert rel_err(G(g['z1'].cuda()), g['fake']) < TOL
    assert rel_err(D(g['real'].cuda()), g['d_real_scores']) < TOL
    assert rel_err(D(g['fake'].cuda()), g['d_fake_scores']) < TOL


@pytest.mark.parametrize('case', STEP_CASES)
def test_d_step_golden(gpu, case):
    g = load_step(case)
    pg = gpu['pg']
    G, D = gpu['build_pair'](g)
    pg.wgan_gp_loss.mixing_factors_override = g['mixing']
    try:
        cost, rl, fl = pg.wgan_gp_D_loss(D, G, g['real'].cuda(), g['z1'].cuda())
        cost.backward()
    finally:
        pg.wgan_gp_loss.mixing_factors_override = None
    assert rel_err(cost, g['d_cost']) < TOL
    assert rl.shape == (g['n'], 1) and rel_err(rl, g['d_real_loss']) < TOL
    assert fl.shape == (g['n'], 1) and rel_err(fl, g['d_fake_loss']) < TOL
    grads = gpu['named_grads'](D)
    assert set(grads) == set(g['dgrad']), 'same parameters receive a gradient as in the reference'
    for k, v in g['dgrad'].items():
        assert rel_err(grads[k], v) < TOL, k
    assert all(p.grad is None for p in G.parameters()), 'G gets no gradient in the D step'


@pytest.mark.parametrize('case', STEP_CASES)
def test_g_step_golden(gpu, case):
    g = load_step(case)
    pg = gpu['pg']
    G, D = gpu['build_pair'](g)
    cost = pg.wgan_gp_G_loss(G, D, g['z2'].cuda())
    cost.backward()
    assert rel_err(cost, g['g_cost']) < TOL
    grads = gpu['named_grads'](G)
    assert set(grads) == set(g['ggrad'])
    for k, v in g['ggrad'].items():
        assert rel_err(grads[k], v) < TOL, k


@pytest.mark.parametrize('fused', [False, True])
def test_two_trainer_iterations_golden(gpu, fused):
    """Trainer.train() twice with Adam(betas=(0,.99)) -- parameters match the reference's after the same.
    fused=True: pg.FusedAdam, which updates the parameters through raw pointers -- the second iteration must see the
    first one's update in every re-laid weight operand (engine.ConvW caches them by Tensor._version)."""
    g = load_trainer()
    pg = gpu['pg']
    Adam = pg.FusedAdam if fused else torch.optim.Adam
    g2 = dict(g, pg=g['G0'], pd=g['D0'])
    G, D = gpu['build_pair'](g2)
    G.depth = D.depth = g['depth']
    G.alpha = D.alpha = g['alpha']
    opt_g = Adam(G.parameters(), 1e-3, betas=(0.0, 0.99))
    opt_d = Adam(D.parameters(), 1e-3, betas=(0.0, 0.99))
    lats = iter(list(g['latents']))
    t = pg.Trainer(D, G, pg.wgan_gp_D_loss, pg.wgan_gp_G_loss, opt_d, opt_g, None, iter(list(g['reals'])),
                   lambda: next(lats))
    for it in range(2):
        pg.wgan_gp_loss.mixing_factors_override = g['mixing'][it]
        t.train()
    pg.wgan_gp_loss.mixing_factors_override = None
    assert t.cur_nimg == g['cur_nimg']
    sd_d = {k: v.detach().cpu() for k, v in D.state_dict().items()}
    sd_g = {k: v.detach().cpu() for k, v in G.state_dict().items()}
    # Adam with beta1=0 normalises every gradient to +-lr on the first steps, so parameter agreement is a
    # sign-level check of all gradients: ||dp|| ~ lr*sqrt(numel).  Compare the UPDATE, not the parameter.
    for k, v in g['D2'].items():
        if k.endswith('.c'):
            continue
        upd_ref = v - g['D0'][k]
        upd = sd_d[k] - g['D0'][k]
        if float(upd_ref.norm()) > 0:
            assert rel_err(upd, upd_ref) < 5e-2, k
        else:
            assert float(upd.norm()) == 0, k
    for k, v in g['G2'].items():
        if k.endswith('.c'):
            continue
        upd_ref = v - g['G0'][k]
        upd = sd_g[k] - g['G0'][k]
        if float(upd_ref.norm()) > 0:
            assert rel_err(upd, upd_ref) < 5e-2, k
        else:
            assert float(upd.norm()) == 0, k


def test_prefetched_reals_give_the_same_parameters(gpu):
    """trainer.prefetch_reals = True (look-ahead H2D copy of pinned real batches on a copy stream) must not change a
    single bit of what two Trainer.train() iterations do."""
    g = load_trainer()
    pg = gpu['pg']
    out = []
    for prefetch in (False, True):
        g2 = dict(g, pg=g['G0'], pd=g['D0'])
        G, D = gpu['build_pair'](g2)
        G.depth = D.depth = g['depth']
        G.alpha = D.alpha = g['alpha']
        opt_g = torch.optim.Adam(G.parameters(), 1e-3, betas=(0.0, 0.99))
        opt_d = torch.optim.Adam(D.parameters(), 1e-3, betas=(0.0, 0.99))
        lats = iter(list(g['latents']))
        reals = [r.clone().pin_memory() for r in g['reals']]
        t = pg.Trainer(D, G, pg.wgan_gp_D_loss, pg.wgan_gp_G_loss, opt_d, opt_g, None, iter(reals), lambda: next(lats))
        t.prefetch_reals = prefetch
        for it in range(2):
            pg.wgan_gp_loss.mixing_factors_override = g['mixing'][it]
            t.train()
        pg.wgan_gp_loss.mixing_factors_override = None
        torch.cuda.synchronize()
        out.append({k: v.detach().clone() for k, v in list(D.state_dict().items()) + list(G.state_dict().items())})
    for k in out[0]:
        assert rel_err(out[1][k], out[0][k]) < 1e-5, k     # weight-gradient atomics: last bits are order dependent


@pytest.mark.parametrize('depth,alpha,n,ch', [(3, 0.5, 4, 3), (3, 1.0, 3, 3), (2, 0.0, 5, 1)])
def test_step_vs_oracle_fresh_inputs(gpu, depth, alpha, n, ch):
    """Larger channel counts / other depths than the golden files: CUDA path vs the oracle on the same seeded inputs."""
    O, pg = gpu['O'], gpu['pg']
    res, fb, fm, lat = 32, 512, 64, 64
    pgp = O.make_generator_params(res, ch, fmap_base=fb, fmap_max=fm, latent_size=lat, seed=3)
    pdp = O.make_discriminator_params(res, ch, fmap_base=fb, fmap_max=fm, seed=4)
    g = dict(resolution=res, channels=ch, fmap_base=fb, fmap_max=fm, latent=lat, pg=pgp, pd=pdp, depth=depth,
             alpha=alpha)
    G, D = gpu['build_pair'](g)
    r = 4 * 2 ** depth
    nb = O.n_blocks_for(res)
    from _gpu_util import relu_margin
    for seed in range(99, 140):   # first seed whose LeakyReLU margins are clear of float noise (see relu_margin)
        gen = torch.Generator().manual_seed(seed)
        z1, z2 = torch.randn(n, lat, generator=gen), torch.randn(n, lat, generator=gen)
        real = torch.randn(n, ch, r, r, generator=gen)
        mix = torch.rand(n, 1, generator=gen)
        res_o = {}

        def run_oracle():
            res_o['d'] = O.d_step_grads(pdp, pgp, real, z1, mix, depth, alpha, nb)
            res_o['g'] = O.g_step_grads(pgp, pdp, z2, depth, alpha, nb)
        if relu_margin(run_oracle) > 2e-5:
            break
    cost_o, rl_o, fl_o, gd_o = res_o['d']
    gcost_o, gg_o = res_o['g']
    pg.wgan_gp_loss.mixing_factors_override = mix
    try:
        cost, rl, fl = pg.wgan_gp_D_loss(D, G, real.cuda(), z1.cuda())
        cost.backward()
    finally:
        pg.wgan_gp_loss.mixing_factors_override = None
    assert rel_err(cost, cost_o) < TOL and rel_err(rl, rl_o) < TOL and rel_err(fl, fl_o) < TOL
    grads = gpu['named_grads'](D)
    assert set(grads) == set(gd_o)
    for k, v in gd_o.items():
        assert rel_err(grads[k], v) < TOL, k
    gcost = pg.wgan_gp_G_loss(G, D, z2.cuda())
    gcost.backward()
    assert rel_err(gcost, gcost_o) < TOL
    grads = gpu['named_grads'](G)
    assert set(grads) == set(gg_o)
    for k, v in gg_o.items():
        assert rel_err(grads[k], v) < TOL, k


@pytest.mark.parametrize('depth,alpha,n,precision', [(6, 0.5, 2, 'fp32'), (5, 1.0, 2, 'fp32'), (6, 1.0, 2, 'bf16')])
def test_step_vs_oracle_high_resolution_thin_layers(gpu, depth, alpha, n, precision):
    """128^2 / 256^2 levels with 16..64 feature maps: the layers served by the row-streaming thin-layer tensor-core
    kernels (csrc/pgk_conv_thin.cu, pgk_wgrad_thin.cu) inside the full D step + G step, against the oracle."""
    O, pg = gpu['O'], gpu['pg']
    res, ch, fb, fm, lat = 256, 3, 2048, 64, 64
    pgp = O.make_generator_params(res, ch, fmap_base=fb, fmap_max=fm, latent_size=lat, seed=5)
    pdp = O.make_discriminator_params(res, ch, fmap_base=fb, fmap_max=fm, seed=6)
    g = dict(resolution=res, channels=ch, fmap_base=fb, fmap_max=fm, latent=lat, pg=pgp, pd=pdp, depth=depth,
             alpha=alpha)
    G, D = gpu['build_pair'](g, precision=precision)
    r = 4 * 2 ** depth
    nb = O.n_blocks_for(res)
    # Fixed inputs.  With millions of LeakyReLU units some pre-activation always lies within float noise of zero, and
    # a unit that takes the other slope (here or in the reference on another BLAS) moves gradients by ~1e-3; the
    # seeds below were picked (scanning on the CPU, see tests/_gpu_util.relu_margin) so that the smallest layers,
    # where one flipped unit matters most, have margins clear of that noise.
    seed = {6: 61, 5: 99}[depth]
    gen = torch.Generator().manual_seed(seed)
    z1, z2 = torch.randn(n, lat, generator=gen), torch.randn(n, lat, generator=gen)
    real = torch.randn(n, ch, r, r, generator=gen)
    mix = torch.rand(n, 1, generator=gen)
    res_o = {'d': O.d_step_grads(pdp, pgp, real, z1, mix, depth, alpha, nb),
             'g': O.g_step_grads(pgp, pdp, z2, depth, alpha, nb)}
    cost_o, rl_o, fl_o, gd_o = res_o['d']
    gcost_o, gg_o = res_o['g']
    tol_v, tol_g = (TOL, TOL) if precision == 'fp32' else (5e-2, 2.5e-1)
    pg._lib.prof_reset()
    pg._lib.prof_enable(True)
    pg.wgan_gp_loss.mixing_factors_override = mix
    try:
        cost, rl, fl = pg.wgan_gp_D_loss(D, G, real.cuda(), z1.cuda())
        cost.backward()
        gcost = pg.wgan_gp_G_loss(G, D, z2.cuda())
        gcost.backward()
    finally:
        pg.wgan_gp_loss.mixing_factors_override = None
        pg._lib.prof_enable(False)
    thin_conv, thin_wgrad = pg._lib.prof_read(4)[3], pg._lib.prof_read(5)[3]
    pg._lib.prof_reset()
    # (two-plane 32-channel weight gradients do not fit the thin kernel's shared memory and take the generic path)
    assert thin_conv > 0 and (thin_wgrad > 0 or depth == 5), 'the thin-layer tensor-core kernels were not exercised'
    assert rel_err(cost, cost_o) < tol_v and rel_err(rl, rl_o) < tol_v and rel_err(fl, fl_o) < tol_v
    assert rel_err(gcost, gcost_o) < tol_v
    gd, gg = gpu['named_grads'](D), gpu['named_grads'](G)
    assert set(gd) == set(gd_o) and set(gg) == set(gg_o)
    if precision == 'fp32':
        for k, v in gd_o.items():
            assert rel_err(gd[k], v) < tol_g, k
        for k, v in gg_o.items():
            assert rel_err(gg[k], v) < tol_g, k
    else:
        # bf16 mode has no counterpart in the (fp32-only) reference.  Activations carry 8 mantissa bits, so ~0.2% of
        # the LeakyReLU units take the other slope and the gradient-penalty double backward amplifies that: per-tensor
        # deviations of 2..30% from the fp32 oracle are the precision, not a defect.  What is asserted is the
        # direction: cosine similarity >= 0.9 with the oracle's gradient for every tensor.
        cos = lambda a, b: float(torch.nn.functional.cosine_similarity(a.double().flatten(), b.double().flatten(), dim=0))
        for k, v in gd_o.items():
            if k != 'linear.bias':   # a sum of +-1/N seeds that cancels to ~0
                assert cos(gd[k], v) > 0.9, k
        for k, v in gg_o.items():
            assert cos(gg[k], v) > 0.9, k


def test_bf16_mode_close_to_oracle(gpu):
    """bf16 mode (one plane): activations and gradients are rounded to 8 mantissa bits at every layer boundary, so
    the tolerance against the fp32 oracle is 5e-2 on losses / outputs and 1e-1 on gradients."""
    g = load_step('tiny3_d2_a03')
    pg = gpu['pg']
    G, D = gpu['build_pair'](g, precision='bf16')
    assert rel_err(G(g['z1'].cuda()), g['fake']) < 5e-2
    pg.wgan_gp_loss.mixing_factors_override = g['mixing']
    try:
        cost, rl, fl = pg.wgan_gp_D_loss(D, G, g['real'].cuda(), g['z1'].cuda())
        cost.backward()
    finally:
        pg.wgan_gp_loss.mixing_factors_override = None
    assert rel_err(cost, g['d_cost']) < 5e-2
    grads = gpu['named_grads'](D)
    for k, v in g['dgrad'].items():
        assert rel_err(grads[k], v) < 1e-1, k


def test_no_cpu_path(gpu):
    g = load_step('tiny3_d0_a1')
    G, D = gpu['build_pair'](g)
    with pytest.raises(RuntimeError):
        G.cpu()(g['z1'])


def test_real_preparation_golden(gpu):
    """pgk_real_prep (on-device alpha_fade + adjust_dynamic_range) against the vectors produced by the
    reference's own code.  uint8 data (float64 arithmetic in numpy, exact box sums): bit exact; float32 data (float32
    arithmetic in numpy, whose 4-term summation order is an implementation detail): <= 1 float32 ulp."""
    import os
    import numpy as np
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'fade.npz'))
    pg = gpu['pg']
    for i, (alpha, a0, a1, b0, b1) in enumerate(z['cases']):
        x = torch.from_numpy(z['in%d' % i])
        ref = torch.from_numpy(z['out%d' % i])
        got = pg.prepare_reals(x.pin_memory(), alpha=float(alpha), range_in=(a0, a1), range_out=(b0, b1)).cpu()
        assert got.dtype == torch.float32 and got.shape == ref.shape
        if x.dtype == torch.uint8:
            assert torch.equal(got, ref), i
        else:
            assert float((got - ref).abs().max()) <= 1.2e-7 * max(1.0, float(ref.abs().max())), i


def test_fused_adam_matches_torch_adam(gpu):
    """pgk_adam_multi vs torch.optim.Adam(betas=(0, .99)) over 5 steps, parameters without a
    gradient skipped, LambdaLR driving the learning rate (train.py:148-158): <= 1e-6 relative on every parameter."""
    pg = gpu['pg']
    torch.manual_seed(0)
    shapes = [(64, 33, 3, 3), (64,), (3, 64, 1, 1), (1, 512), (1,)]
    pa = [torch.nn.Parameter(torch.randn(s, device='cuda')) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    oa = torch.optim.Adam(pa, 1e-3, betas=(0.0, 0.99))
    ob = pg.FusedAdam(pb, 1e-3, betas=(0.0, 0.99))
    sa = torch.optim.lr_scheduler.LambdaLR(oa, pg.lr_rampup)
    sb = torch.optim.lr_scheduler.LambdaLR(ob, pg.lr_rampup)
    for it in range(5):
        for i, (a, b) in enumerate(zip(pa, pb)):
            if i == 2 and it % 2 == 1:           # an inactive block: no gradient this step
                a.grad = b.grad = None
                continue
            g = torch.randn_like(a) * (10.0 ** (it - 2))
            a.grad, b.grad = g.clone(), g.clone()
        oa.step()
        ob.step()
        sa.step(it * 8000)
        sb.step(it * 8000)
    for a, b in zip(pa, pb):
        assert rel_err(b, a) < 1e-6
    assert ob.state[pb[2]]['step'] == oa.state[pa[2]]['step'] == 3


def test_fused_adam_many_unsynchronised_steps(gpu):
    """50 FusedAdam steps enqueued without a single host synchronisation (the host runs far ahead of the GPU: a slow
    kernel is queued first), fresh gradient tensors and a different learning rate every step, against torch Adam.
    The pointer / step-size table of step i must not be overwritten by step i+k before its copy has run."""
    pg = gpu['pg']
    torch.manual_seed(1)
    shapes = [(128, 64, 3, 3), (128,), (3, 128, 1, 1), (1, 512)]
    pa = [torch.nn.Parameter(torch.randn(s, device='cuda')) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    oa = torch.optim.Adam(pa, 1e-3, betas=(0.0, 0.99))
    ob = pg.FusedAdam(pb, 1e-3, betas=(0.0, 0.99))
    grads = [[torch.randn_like(a) * (1.0 + it) for a in pa] for it in range(50)]
    big = torch.randn(8192, 8192, device='cuda')
    torch.cuda.synchronize()
    for _ in range(20):
        big = big @ big * 1e-4            # ~20 x 1 ms of queued GPU work: every step below is enqueued behind it
    for it in range(50):
        for a, b, g in zip(pa, pb, grads[it]):
            a.grad, b.grad = g, g.clone()
        for o in (oa, ob):
            o.param_groups[0]['lr'] = 1e-3 * (1 + it % 7)
        oa.step()
        ob.step()
        assert all(b._version > 0 for b in pb)
    torch.cuda.synchronize()
    for a, b in zip(pa, pb):
        assert rel_err(b, a) < 1e-6


@pytest.mark.parametrize('fading', [False, True])
def test_cuda_graph_replay_equals_eager(gpu, fading):
    """wgan_gp_loss.cuda_graphs: the captured + replayed kernel sequence gives the eager results, including after the
    optimizer changed the weights (the weight re-layout kernels are part of the graph).  fading=True: alpha changes
    EVERY iteration, as DepthManager does during a transition phase (plugins.py:57-81) -- one graph serves them all,
    the kernels read alpha from the engine's device-side scalar pair."""
    pg = gpu['pg']
    g = load_step('tiny3_d2_a03')

    def run(graphs):
        G, D = gpu['build_pair'](g)
        od = torch.optim.SGD(D.parameters(), 1e-2)
        og = torch.optim.SGD(G.parameters(), 1e-2)
        pg.wgan_gp_loss.cuda_graphs = graphs
        pg.wgan_gp_loss._graphs.clear()
        out = []
        try:
            for it in range(6):
                if fading:
                    G.alpha = D.alpha = 0.0 if it == 3 else 0.1 + 0.17 * it      # (0.0: the first iteration of a level)
                gen = torch.Generator().manual_seed(100 + it)
                real = torch.randn(g['n'], g['channels'], g['real'].shape[-1], g['real'].shape[-1], generator=gen).cuda()
                z1 = torch.randn(g['n'], g['latent'], generator=gen).cuda()
                z2 = torch.randn(g['n'], g['latent'], generator=gen).cuda()
                pg.wgan_gp_loss.mixing_factors_override = torch.rand(g['n'], 1, generator=gen)
                cost, rl, fl = pg.wgan_gp_D_loss(D, G, real, z1)
                cost.backward()
                od.step()
                gcost = pg.wgan_gp_G_loss(G, D, z2)
                gcost.backward()
                og.step()
                out.append((float(cost), float(gcost), rl.clone(), fl.clone()))
        finally:
            pg.wgan_gp_loss.cuda_graphs = False
            pg.wgan_gp_loss.mixing_factors_override = None
            pg.wgan_gp_loss._graphs.clear()
        return out, [p.detach().clone() for p in list(D.parameters()) + list(G.parameters())]

    eager, pe = run(False)
    graph, pgr = run(True)
    for (c0, g0, r0, f0), (c1, g1, r1, f1) in zip(eager, graph):
        assert abs(c0 - c1) <= 1e-5 * max(1.0, abs(c0)) and abs(g0 - g1) <= 1e-5 * max(1.0, abs(g0))
        assert rel_err(r1, r0) < 1e-5 and rel_err(f1, f0) < 1e-5
    for a, b in zip(pe, pgr):
        assert rel_err(b, a) < 1e-5      # atomics make the last bits of the weight gradients order dependent


@pytest.mark.parametrize('depth,alpha', [(2, 1.0), (3, 0.4)])
def test_fused_half_planes_equal_the_conversion_pass(gpu, depth, alpha):
    """fp32-faithful mode with fp16 forward operands: the half planes written by from_rgb / pool2 / the upsample / the
    pixel norm (include/pgk.h: out16) against those of the separate pgk_cvt_fp16x2 pass -- the forward results (no
    atomics on that path) must be equal bit for bit, at the reference's channel widths (512: the wide kernel)."""
    import importlib
    pg = gpu['pg']
    E = importlib.import_module('pggan-pytorch_b200.engine')
    if not E.FWD_FP16:
        pytest.skip('PGK_FWD_FP16=0')
    torch.manual_seed(7)
    shape = (1000, 3, 1024, 1024)
    G, D = pg.Generator(shape).cuda(), pg.Discriminator(shape).cuda()
    G.depth = D.depth = depth
    G.alpha = D.alpha = alpha
    n, r = 4, 4 * 2 ** depth
    z = torch.randn(n, 512, device='cuda')
    x = torch.randn(n, 3, r, r, device='cuda')
    seen = []
    real_call = E.call

    def spy(name, *a):
        seen.append(name)
        return real_call(name, *a)

    def run(fuse):
        old = E.FUSE_CVT
        E.FUSE_CVT = fuse
        E.call = spy
        del seen[:]
        try:
            with torch.no_grad():
                img = G(z).clone()
                score = D(x).clone()
            torch.cuda.synchronize()
        finally:
            E.FUSE_CVT = old
            E.call = real_call
        return img, score, seen.count('pgk_cvt_fp16x2'), seen.count('pgk_conv_fp16')

    img0, s0, cvt0, conv0 = run(False)
    img1, s1, cvt1, conv1 = run(True)
    assert conv0 == conv1 and conv0 > 0 and cvt0 == conv0        # without fusion: one conversion per half-operand conv
    assert cvt1 < cvt0                                           # with it: only conv -> conv links still convert
    assert torch.equal(img0, img1) and torch.equal(s0, s1)
