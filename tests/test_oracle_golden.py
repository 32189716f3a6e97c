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
