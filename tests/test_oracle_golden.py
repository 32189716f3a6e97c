"""CPU: pin oracle/pggan_oracle.py against vectors produced by executing the
unmodified reference (tests/golden/make_golden.py)."""
import math
import os
import sys

import pytest
import torch

from _util import ORACLE_ONLY_CASES, STEP_CASES, load_schedule, load_step, load_trainer, rel_err

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
import pggan_oracle as O  # noqa: E402

TOL = 2e-5  # same fp32 torch ops in a different association order
CASES = [(c, 'step_') for c in STEP_CASES] + [(c, 'ostep_') for c in ORACLE_ONLY_CASES]
IDS = [p + c for c, p in CASES]


@pytest.mark.parametrize('case', CASES, ids=IDS)
def test_forward_matches_reference(case):
    g = load_step(case[0], prefix=case[1])
    nb = O.n_blocks_for(g['resolution'])
    fake = O.generator_forward(g['pg'], g['z1'], g['depth'], g['alpha'])
    assert fake.shape == g['fake'].shape
    assert rel_err(fake, g['fake']) < TOL
    assert rel_err(O.discriminator_forward(g['pd'], g['real'], g['depth'], g['alpha'], nb), g['d_real_scores']) < TOL
    assert rel_err(O.discriminator_forward(g['pd'], g['fake'], g['depth'], g['alpha'], nb), g['d_fake_scores']) < TOL


@pytest.mark.parametrize('case', CASES, ids=IDS)
def test_d_step_matches_reference(case):
    g = load_step(case[0], prefix=case[1])
    nb = O.n_blocks_for(g['resolution'])
    cost, rl, fl, grads = O.d_step_grads(g['pd'], g['pg'], g['real'], g['z1'], g['mixing'], g['depth'], g['alpha'], nb)
    assert rel_err(cost, g['d_cost']) < TOL
    assert rl.shape == g['d_real_loss'].shape and rel_err(rl, g['d_real_loss']) < TOL
    assert fl.shape == g['d_fake_loss'].shape and rel_err(fl, g['d_fake_loss']) < TOL
    assert set(grads) == set(g['dgrad']), 'same set of parameters receives a gradient'
    for k in grads:
        assert rel_err(grads[k], g['dgrad'][k]) < 2e-4, k


@pytest.mark.parametrize('case', CASES, ids=IDS)
def test_g_step_matches_reference(case):
    g = load_step(case[0], prefix=case[1])
    nb = O.n_blocks_for(g['resolution'])
    cost, grads = O.g_step_grads(g['pg'], g['pd'], g['z2'], g['depth'], g['alpha'], nb)
    assert rel_err(cost, g['g_cost']) < TOL
    assert set(grads) == set(g['ggrad'])
    for k in grads:
        assert rel_err(grads[k], g['ggrad'][k]) < 2e-4, k


def test_two_trainer_iterations_match_reference():
    """D step -> Adam -> G step (through the UPDATED D) -> Adam, twice (trainer.py:85-115)."""
    g = load_trainer()
    nb = O.n_blocks_for(g['resolution'])
    pd, pg = dict(g['D0']), dict(g['G0'])
    sd, sg = {}, {}
    li = 0
    for it in range(2):
        _, _, _, gd = O.d_step_grads(pd, pg, g['reals'][it], g['latents'][li], g['mixing'][it], g['depth'], g['alpha'], nb)
        pd = O.adam_step(pd, gd, sd, 1e-3)
        li += 1
        _, gg = O.g_step_grads(pg, pd, g['latents'][li], g['depth'], g['alpha'], nb)
        pg = O.adam_step(pg, gg, sg, 1e-3)
        li += 1
    for k, v in g['D2'].items():
        assert rel_err(pd[k], v) < 1e-4, k
    for k, v in g['G2'].items():
        assert rel_err(pg[k], v) < 1e-4, k


def test_schedule_bit_exact():
    s = load_schedule()
    for row in s['points']:
        d, a, mb, tick = O.depth_schedule(row['cur_nimg'], row['max_depth'])
        assert d == row['depth'] and repr(float(a)) == row['alpha'], row
        assert mb == row['minibatch'] and tick == row['tick_nimg'], row


def test_param_tables_match_reference_shapes():
    g = load_step('tiny3_d2_a03')
    pg = O.make_generator_params(16, 3, fmap_base=128, fmap_max=32, latent_size=32)
    pd = O.make_discriminator_params(16, 3, fmap_base=128, fmap_max=32)
    assert {k: tuple(v.shape) for k, v in pg.items()} == {k: tuple(v.shape) for k, v in g['pg'].items()}
    assert {k: tuple(v.shape) for k, v in pd.items()} == {k: tuple(v.shape) for k, v in g['pd'].items()}
    # unit-RMS weights, measured c close to sqrt(2/fan_in)
    w = pd['blocks.0.c1.conv.weight']
    assert abs(float((w ** 2).mean()) - 1.0) < 1e-5
    assert abs(float(pd['blocks.0.c1.c']) / math.sqrt(2.0 / (w.shape[1] * 9)) - 1) < 0.1


def test_lr_rampup_values():
    assert abs(1e-3 * O.lr_rampup(0) - 6.7379e-6) < 1e-9
    assert abs(1e-3 * O.lr_rampup(20000) - 2.8650e-4) < 1e-8
    assert O.lr_rampup(40000) == 1.0


def test_real_preparation_matches_reference():
    """oracle alpha_fade / adjust_dynamic_range vs vectors produced by the reference's own function bodies
    (tests/golden/make_golden_fade.py): bit exact, same numpy arithmetic."""
    import numpy as np
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'fade.npz'))
    for i, (alpha, a0, a1, b0, b1) in enumerate(z['cases']):
        rin = (int(a0), int(a1)) if float(a0).is_integer() else (a0, a1)
        rout = (int(b0), int(b1))
        for d, ref in zip(z['in%d' % i], z['out%d' % i]):
            got = O.prepare_real(d, float(alpha), rin, rout)
            assert got.dtype == np.float32 and np.array_equal(got, ref), i


@pytest.mark.parametrize('name', ['d4_a05_n2', 'd5_c1_n2'])
def test_oracle_at_baseline_widths_matches_reference(name):
    """The oracle on BASELINE's own networks (full 1024x1024 model at depth 4 with its 512-channel K = 4608 layers;
    the 1-channel 128x128 model at depth 5) against compact vectors of the reference executed at these widths
    (tests/golden/make_golden_full.py): parameters / inputs regenerated from the same seeds, every loss and every
    parameter gradient (norm + up to 2048 sampled entries per tensor)."""
    import numpy as np
    from _util import GOLDEN
    sys.path.insert(0, GOLDEN)
    import make_golden_full as MF
    z = np.load(os.path.join(GOLDEN, 'full_%s.npz' % name))
    cfg = MF.CONFIGS[name]
    res, ch, depth, alpha, n, seed = cfg
    pgp, pdp = O.make_generator_params(res, ch, seed=MF.G_SEED), O.make_discriminator_params(res, ch, seed=MF.D_SEED)
    z1, z2, real, mix = MF.inputs(cfg)
    nb = O.n_blocks_for(res)
    cost, rl, fl, gd = O.d_step_grads(pdp, pgp, real, z1, mix, depth, alpha, nb)
    gcost, gg = O.g_step_grads(pgp, pdp, z2, depth, alpha, nb)
    assert rel_err(cost, z['d_cost']) < TOL and rel_err(rl, z['d_real_loss']) < TOL and rel_err(fl, z['d_fake_loss']) < TOL
    assert rel_err(gcost, z['g_cost']) < TOL

    def check(prefix, grads):
        keys = sorted(k[len(prefix):-5] for k in z.files if k.startswith(prefix) and k.endswith('/norm'))
        assert set(keys) == set(grads)
        for k in keys:
            t = grads[k].reshape(-1)
            key = prefix + k
            assert rel_err(t[MF.sample_index(key, t.numel())], z[key + '/samples']) < 5e-4, key
            assert abs(float(t.double().norm()) / float(z[key + '/norm']) - 1) < 5e-4, key
    check('Dgrad.', gd)
    check('Ggrad.', gg)


def test_bf16_oracle_without_rounding_is_the_oracle():
    """oracle/pggan_oracle_bf16.py restructures the forward graphs (folded c, split constant channel, rounding nodes
    whose backward is again a rounding node).  With the roundings switched off it must BE the oracle: same losses and
    gradients, including the penalty's double backward through the rounding nodes."""
    import pggan_oracle_bf16 as B
    for (res, ch, fb, fm, lat, n, depth, alpha) in [(16, 3, 128, 32, 32, 4, 2, 0.3), (16, 3, 128, 32, 32, 4, 0, 1.0),
                                                    (32, 1, 256, 32, 32, 3, 3, 1.0)]:
        pgp = O.make_generator_params(res, ch, fmap_base=fb, fmap_max=fm, latent_size=lat, seed=3)
        pdp = O.make_discriminator_params(res, ch, fmap_base=fb, fmap_max=fm, seed=4)
        nb, r = O.n_blocks_for(res), 4 * 2 ** depth
        gen = torch.Generator().manual_seed(5)
        z1, z2 = torch.randn(n, lat, generator=gen), torch.randn(n, lat, generator=gen)
        real, mix = torch.randn(n, ch, r, r, generator=gen), torch.rand(n, 1, generator=gen)
        c0, rl0, fl0, gd0 = O.d_step_grads(pdp, pgp, real, z1, mix, depth, alpha, nb)
        gc0, gg0 = O.g_step_grads(pgp, pdp, z2, depth, alpha, nb)
        B.ROUND = False
        try:
            c1, rl1, fl1, gd1, _ = B.d_step_grads(pdp, pgp, real, z1, mix, depth, alpha, nb)
            gc1, gg1 = B.g_step_grads(pgp, pdp, z2, depth, alpha, nb)
        finally:
            B.ROUND = True
        assert rel_err(c1, c0) < TOL and rel_err(gc1, gc0) < TOL and rel_err(rl1, rl0) < TOL and rel_err(fl1, fl0) < TOL
        assert set(gd1) == set(gd0) and set(gg1) == set(gg0)
        for k in gd0:
            assert rel_err(gd1[k], gd0[k]) < TOL, k
        for k in gg0:
            assert rel_err(gg1[k], gg0[k]) < TOL, k
        # and with them on it is a different (coarser) arithmetic
        c2, _, _, gd2, _ = B.d_step_grads(pdp, pgp, real, z1, mix, depth, alpha, nb)
        assert max(rel_err(gd2[k], gd0[k]) for k in gd0) > 1e-3
