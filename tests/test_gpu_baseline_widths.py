"""-m gpu: parity of the CUDA path on BASELINE.json's own networks -- the full 1024x1024 model (512-channel K = 4608
layers, 8 / 16-channel 1024^2 layers) and the 1-channel 128x128 model -- through the C ABI.

What is compared with what, and why there are two kinds of gradient test:

  * LOSSES, SCORES, IMAGES against compact golden vectors made by executing the unmodified REFERENCE at these widths
    (tests/golden/make_golden_full.py -> full_*.npz; parameters and inputs regenerated from seeds on both sides):
    <= 1e-3 per tensor (measured <= 6e-5).
  * GRADIENTS are a discontinuous function of the forward values: a LeakyReLU unit whose pre-activation lies within
    rounding noise of zero takes the other slope under another summation order, and ONE such unit in a 512-channel 8x8
    layer moves that layer's bias gradient by ~3e-3 at batch 1 (everything upstream of it follows).  The reference has
    this floor against itself (its fp32 and fp64 runs, both stored in the golden files, differ by up to 1.5e-3 per
    tensor), and the tensor-core arithmetic of the fp32-faithful mode (three bf16 planes, truncating fp32 accumulate:
    ~1e-6 relative on a pre-activation against ~1e-7 for fp32 FMA) flips a few more units than fp32 does.  So:
      - `..._vs_reference_golden`: every parameter gradient against the reference's, reported next to the reference's
        own fp32-vs-fp64 floor; gated at 2e-2 (a wrong kernel is off by O(1), a flipped unit by O(1e-3));
      - `..._under_the_kernels_lrelu_decisions`: the oracle re-run with the LeakyReLU decisions the KERNELS took
        (read from the stored activations) imposed on it -- what remains is rounding alone: every gradient <= 1e-3
        (measured ~1e-4), and the units whose decision differs from the oracle's own are counted (<= 1e-4 of all).
  * precision='bf16' (BASELINE c3-c5; the reference has no such mode) against the bf16 oracle
    (oracle/pggan_oracle_bf16.py: the reference's algorithm with a bf16 rounding wherever the kernels store a tensor),
    again under the kernels' own LeakyReLU decisions -- bf16 activations flip ~0.2 % of the units against ANY other
    implementation, which alone moves gradients by per cent (measured 5 % median against the free-running bf16 oracle
    and 6 % against the fp32 oracle: both reported).
  * 20 UNSCREENED seeds on a small network against the fp64 oracle, with the fp32 oracle's own distance beside it.

`PGK_PARITY_REPORT=<file>` appends the measured numbers as JSON lines.
"""
import json
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from _util import GOLDEN, ORACLE_ONLY_CASES, load_step, rel_err

sys.path.insert(0, GOLDEN)
import make_golden_full as MF  # noqa: E402  (seeds, sample positions; importing it does not touch the reference)

pytestmark = pytest.mark.gpu

TOL = 1e-3
FULL = sorted(MF.CONFIGS)


def report(**row):
    path = os.environ.get('PGK_PARITY_REPORT')
    if path:
        with open(path, 'a') as f:
            f.write(json.dumps(row) + '\n')


@pytest.fixture(scope='module')
def gpu():
    from _gpu_util import O, build_pair, named_grads, pg
    return dict(O=O, build_pair=build_pair, named_grads=named_grads, pg=pg)


_params = {}


def full_params(O, res, ch):
    """oracle.make_*_params(seed): the 18 M-parameter networks regenerated from the seeds the golden files were made with"""
    key = (res, ch)
    if key not in _params:
        _params.clear()         # one model at a time: 2 x 73 MB of host memory each
        _params[key] = (O.make_generator_params(res, ch, seed=MF.G_SEED),
                        O.make_discriminator_params(res, ch, seed=MF.D_SEED))
    return _params[key]


def build_full(gpu, name, precision='fp32'):
    res, ch, depth, alpha, n, seed = MF.CONFIGS[name]
    pgp, pdp = full_params(gpu['O'], res, ch)
    G, D = gpu['build_pair'](dict(resolution=res, channels=ch, fmap_base=4096, fmap_max=512, latent=512, pg=pgp,
                                   pd=pdp, depth=depth, alpha=alpha), precision=precision)
    return G, D, pgp, pdp


def run_step(gpu, G, D, z1, z2, real, mix):
    """D step + G step through the reference-facing API.  Returns losses, the fake image and both gradient dicts."""
    pg = gpu['pg']
    from _gpu_util import d_masks, g_masks
    n = real.shape[0]
    with torch.no_grad():
        fake = G(z1.cuda())
    pg.wgan_gp_loss.mixing_factors_override = mix
    pg.wgan_gp_loss.keep_tapes = True
    try:
        cost, rl, fl = pg.wgan_gp_D_loss(D, G, real.cuda(), z1.cuda())
        cost.backward()
        gd = gpu['named_grads'](D)
        T = pg.wgan_gp_loss.last_aux.pop('d_tape')
        # oracle order of the D step: D(real), G(z) (no graph: its own decisions), D(fake), D(mixed)
        masks_d = d_masks(T, 0, n) + [None] * (2 * (T.depth + 1)) + d_masks(T, 1, n) + d_masks(T, 2, n)
        del T
        gcost = pg.wgan_gp_G_loss(G, D, z2.cuda())
        gcost.backward()
        gg = gpu['named_grads'](G)
        TG, T = pg.wgan_gp_loss.last_aux.pop('g_tape')
        masks_g = g_masks(TG) + d_masks(T, 0, n)
        del T, TG
    finally:
        pg.wgan_gp_loss.mixing_factors_override = None
        pg.wgan_gp_loss.keep_tapes = False
    return dict(cost=cost.detach().cpu(), rl=rl.detach().cpu(), fl=fl.detach().cpu(), gcost=gcost.detach().cpu(),
                fake=fake.cpu(), gd=gd, gg=gg, masks_d=masks_d, masks_g=masks_g)


def sampled_err(key, t, z, prefix=''):
    """Relative error of tensor t against the golden's sample of it (and of its norm): both must hold the tolerance."""
    t = t.detach().reshape(-1).cpu()
    ref = torch.from_numpy(z[prefix + key + '/samples'])
    got = t[MF.sample_index(key, t.numel())]
    e_s = rel_err(got, ref)
    nref = float(z[prefix + key + '/norm'])
    e_n = abs(float(t.double().norm()) - nref) / nref if nref > 0 else float(t.double().norm())
    return max(e_s, e_n)


GRAD_SANITY = 2e-2


@pytest.mark.parametrize('name', FULL)
def test_full_width_step_vs_reference_golden(gpu, name):
    """Losses, scores and the fake image of one D step + G step at the real widths against what the reference itself
    computed: 1e-3.  Every parameter gradient against the reference's fp64 run, reported next to the reference's own
    fp32-vs-fp64 distance for the same tensor (see the module docstring), gated at 2e-2."""
    z = np.load(os.path.join(GOLDEN, 'full_%s.npz' % name))
    cfg = MF.CONFIGS[name]
    G, D, _, _ = build_full(gpu, name)
    z1, z2, real, mix = MF.inputs(cfg)
    with torch.no_grad():
        e_scores = rel_err(D(real.cuda()), z['d_real_scores'])
    out = run_step(gpu, G, D, z1, z2, real, mix)
    errs = {'d_real_scores': e_scores, 'fake': sampled_err('fake', out['fake'], z),
            'd_cost': rel_err(out['cost'], z['d_cost']), 'd_real_loss': rel_err(out['rl'], z['d_real_loss']),
            'd_fake_loss': rel_err(out['fl'], z['d_fake_loss']), 'g_cost': rel_err(out['gcost'], z['g_cost'])}
    gkeys = sorted(k[:-5] for k in z.files if k[1:6] == 'grad.' and k.endswith('/norm'))
    assert {k[6:] for k in gkeys if k[0] == 'D'} == set(out['gd']) and {k[6:] for k in gkeys if k[0] == 'G'} == set(out['gg']), \
        'same parameters receive a gradient as in the reference'
    gerr, floor = {}, {}
    for k in gkeys:
        t = out['gd'][k[6:]] if k[0] == 'D' else out['gg'][k[6:]]
        gerr[k] = sampled_err(k, t, z, 'f64/')
        floor[k] = rel_err(z[k + '/samples'].astype(np.float64), z['f64/' + k + '/samples'])
    worst = max(gerr, key=gerr.get)
    report(test='full_width_vs_reference_golden', case=name, values={k: float('%.3g' % v) for k, v in errs.items()},
           grad_worst=worst, grad_worst_err=gerr[worst], grad_median_err=float(np.median(list(gerr.values()))),
           reference_fp32_vs_fp64_worst=max(floor.values()), reference_fp32_vs_fp64_median=float(np.median(list(floor.values()))),
           grads_over_1e_3=sum(v > TOL for v in gerr.values()), grads=len(gerr))
    bad = {k: v for k, v in errs.items() if not v < TOL}
    bad.update({k: v for k, v in gerr.items() if not v < GRAD_SANITY})
    assert not bad, bad


@pytest.mark.parametrize('name', ['d4_a05_n2', 'd5_c1_n2', 'd6_a1_n1', 'd8_a03_n1'])
def test_full_width_gradients_under_the_kernels_lrelu_decisions(gpu, name):
    """Every loss and every parameter gradient, full tensors, against the oracle run on this host's CPU (10 .. 30 s
    each) with the LeakyReLU decisions of the CUDA forward passes imposed on it: 1e-3 on everything.  The number of
    units whose decision differs from the oracle's own is reported and bounded (1e-4 of all units)."""
    from _gpu_util import ForcedMasks
    O = gpu['O']
    cfg = MF.CONFIGS[name]
    res, ch, depth, alpha, n, seed = cfg
    G, D, pgp, pdp = build_full(gpu, name)
    z1, z2, real, mix = MF.inputs(cfg)
    nb = O.n_blocks_for(res)
    out = run_step(gpu, G, D, z1, z2, real, mix)
    with ForcedMasks(out['masks_d']) as fd:
        cost_o, rl_o, fl_o, gd_o = O.d_step_grads(pdp, pgp, real, z1, mix, depth, alpha, nb)
    with ForcedMasks(out['masks_g']) as fg:
        gcost_o, gg_o = O.g_step_grads(pgp, pdp, z2, depth, alpha, nb)
    assert fd.i == len(out['masks_d']) and fg.i == len(out['masks_g']), 'one decision tensor per leaky_relu call'
    errs = {'d_cost': rel_err(out['cost'], cost_o), 'd_real_loss': rel_err(out['rl'], rl_o),
            'd_fake_loss': rel_err(out['fl'], fl_o), 'g_cost': rel_err(out['gcost'], gcost_o)}
    assert set(out['gd']) == set(gd_o) and set(out['gg']) == set(gg_o)
    errs.update({'Dgrad.' + k: rel_err(out['gd'][k], v) for k, v in gd_o.items()})
    errs.update({'Ggrad.' + k: rel_err(out['gg'][k], v) for k, v in gg_o.items()})
    worst = max(errs, key=errs.get)
    flips, units = fd.flips + fg.flips, fd.units + fg.units
    report(test='full_width_under_kernel_decisions', case=name, worst=worst, worst_err=errs[worst],
           median_err=float(np.median(list(errs.values()))), flipped_units=flips, units=units)
    bad = {k: v for k, v in errs.items() if not v < TOL}
    assert not bad, bad
    assert flips <= 1e-4 * units, (flips, units)


# ---- bf16 mode ---------------------------------------------------------------------------------------------------
BF16_TOL_VALUE, BF16_TOL_GRAD = 3e-2, 2e-2


def pn_fused_fn(pg, G, n):
    """Which generator layers normalise inside the conv epilogue (one rounding) and which as a second pass over the
    stored tensor (two): asked of the library itself, per layer shape (one-plane mode)."""
    import ctypes
    lib = pg._lib.load()
    lib.pgk_conv_thin_supported.argtypes = [ctypes.c_int] * 7
    lib.pgk_conv_thin_fuses_pixelnorm.argtypes = [ctypes.c_int]
    lib.pgk_conv_tc_fuses_pixelnorm.argtypes = [ctypes.c_int] * 2
    table = {}
    res = 4
    for i in range(0, G.max_depth + 1):
        b = G.block0 if i == 0 else G.blocks[i - 1]
        nm = 'block0' if i == 0 else 'blocks.%d' % (i - 1)
        if i > 0:
            res *= 2
        for c in ('c1', 'c2'):
            w = getattr(b, c).conv.weight
            cout, cin, ks = w.shape[0], w.shape[1], w.shape[2]
            thin = ks == 3 and lib.pgk_conv_thin_supported(n, res, res, cin, cout, 3, 0)
            # (engine.THIN64: 64 -> 32 / 64 layers of the one-plane mode run on the thin kernel at W >= 128)
            E_ = importlib.import_module('pggan-pytorch_b200.engine')
            thin = thin or (E_.THIN64 and ks == 3 and cin == 64 and cout in (32, 64) and res % 128 == 0)
            if thin:
                fused = lib.pgk_conv_thin_fuses_pixelnorm(cout)
            else:
                fused = ks == 3 and lib.pgk_conv_tc_supported(n, res, res, cin, cout, 3, 0) and \
                    lib.pgk_conv_tc_fuses_pixelnorm(cout, 0)
            table['%s.%s' % (nm, c)] = bool(fused)
    return lambda name: table[name]


BF16_CASES = {
    # name: (resolution, channels, fmap_base, fmap_max, latent, depth, alpha, n, seed) -- or a full-width golden name
    'thin256_d6_a05': (256, 3, 2048, 64, 64, 6, 0.5, 2, 61),
    'thin256_d5_a1': (256, 3, 2048, 64, 64, 5, 1.0, 2, 99),
    'd4_a05_n2': None,
    'd5_c1_n2': None,
}


@pytest.mark.parametrize('case', sorted(BF16_CASES))
def test_bf16_mode_vs_bf16_oracle(gpu, case):
    """precision='bf16' (BASELINE c3-c5) against the bf16 oracle -- the reference's algorithm with a bf16 rounding at
    every place the kernels store a tensor -- run with the LeakyReLU decisions of the CUDA forward passes imposed on it.
    Tolerance: 3e-2 on losses / images (bf16 carries 8 mantissa bits, 4e-3 per stored element, through up to 20
    layers: measured 0.5 .. 1.8 %), 2e-2 per parameter-gradient tensor (measured: worst 0.7 .. 1.4 %, median 0.4 .. 0.6 %).  The distances to the free-running bf16 oracle and to the fp32 oracle (per cent: ~0.2 % of the
    units decide differently between any two bf16 implementations) are measured and reported, not gated."""
    from _gpu_util import ForcedMasks
    O, pg = gpu['O'], gpu['pg']
    import pggan_oracle_bf16 as B
    spec = BF16_CASES[case]
    if spec is None:
        res, ch, depth, alpha, n, seed = MF.CONFIGS[case]
        G, D, pgp, pdp = build_full(gpu, case, precision='bf16')
        z1, z2, real, mix = MF.inputs(MF.CONFIGS[case])
    else:
        res, ch, fb, fm, lat, depth, alpha, n, seed = spec
        pgp = O.make_generator_params(res, ch, fmap_base=fb, fmap_max=fm, latent_size=lat, seed=5)
        pdp = O.make_discriminator_params(res, ch, fmap_base=fb, fmap_max=fm, seed=6)
        G, D = gpu['build_pair'](dict(resolution=res, channels=ch, fmap_base=fb, fmap_max=fm, latent=lat, pg=pgp,
                                      pd=pdp, depth=depth, alpha=alpha), precision='bf16')
        gen = torch.Generator().manual_seed(seed)
        r = 4 * 2 ** depth
        z1, z2 = torch.randn(n, lat, generator=gen), torch.randn(n, lat, generator=gen)
        real = torch.randn(n, ch, r, r, generator=gen)
        mix = torch.rand(n, 1, generator=gen)
    nb = O.n_blocks_for(res)
    fused = pn_fused_fn(pg, G, n)
    pg._lib.prof_reset()
    pg._lib.prof_enable(True)
    try:
        out = run_step(gpu, G, D, z1, z2, real, mix)
    finally:
        pg._lib.prof_enable(False)
    simt_convs = pg._lib.prof_read(2)[3]
    pg._lib.prof_reset()
    assert simt_convs == 0, 'a conv ran on the CUDA-core kernel, which multiplies by fp32 weights: not the bf16 arithmetic'
    with ForcedMasks(out['masks_d']) as fd:
        cost_b, rl_b, fl_b, gd_b, fake_b = B.d_step_grads(pdp, pgp, real, z1, mix, depth, alpha, nb, fused)
    with ForcedMasks(out['masks_g']) as fg:
        gcost_b, gg_b = B.g_step_grads(pgp, pdp, z2, depth, alpha, nb, fused)
    assert fd.i == len(out['masks_d']) and fg.i == len(out['masks_g'])
    errs = {'fake': rel_err(out['fake'], fake_b), 'd_cost': rel_err(out['cost'], cost_b),
            'd_real_loss': rel_err(out['rl'], rl_b), 'd_fake_loss': rel_err(out['fl'], fl_b),
            'g_cost': rel_err(out['gcost'], gcost_b)}
    assert set(out['gd']) == set(gd_b) and set(out['gg']) == set(gg_b)
    gerrs = {'Dgrad.' + k: rel_err(out['gd'][k], v) for k, v in gd_b.items()}
    gerrs.update({'Ggrad.' + k: rel_err(out['gg'][k], v) for k, v in gg_b.items()})
    worst_v, worst_g = max(errs, key=errs.get), max(gerrs, key=gerrs.get)
    row = dict(test='bf16_under_kernel_decisions', case=case, worst_value=worst_v, worst_value_err=errs[worst_v],
               worst_grad=worst_g, worst_grad_err=gerrs[worst_g], median_grad_err=float(np.median(list(gerrs.values()))),
               flipped_units=fd.flips + fg.flips, units=fd.units + fg.units)
    if depth <= 5:      # for the record (not gates): the free-running bf16 oracle and the fp32 oracle
        _, _, _, gd_f, _ = B.d_step_grads(pdp, pgp, real, z1, mix, depth, alpha, nb, fused)
        _, _, _, gd_o = O.d_step_grads(pdp, pgp, real, z1, mix, depth, alpha, nb)
        med = lambda ref: float(np.median([rel_err(out['gd'][k], v) for k, v in ref.items() if float(v.norm()) > 0]))
        row.update(median_dgrad_vs_free_bf16_oracle=med(gd_f), median_dgrad_vs_fp32_oracle=med(gd_o),
                   median_free_bf16_oracle_vs_fp32_oracle=float(np.median(
                       [rel_err(gd_f[k], v) for k, v in gd_o.items() if float(v.norm()) > 0])))
    report(**row)
    bad = {k: v for k, v in errs.items() if not v < BF16_TOL_VALUE}
    bad.update({k: v for k, v in gerrs.items() if not v < BF16_TOL_GRAD})
    assert not bad, bad


# ---- edge cases the reference-made vectors hold: a batch of one, widths that are not powers of two ---------------
@pytest.mark.parametrize('case', ORACLE_ONLY_CASES)
def test_edge_case_goldens_on_the_gpu(gpu, case):
    g = load_step(case, prefix='ostep_')
    pg = gpu['pg']
    G, D = gpu['build_pair'](g)
    assert rel_err(G(g['z1'].cuda()), g['fake']) < TOL
    assert rel_err(D(g['real'].cuda()), g['d_real_scores']) < TOL
    pg.wgan_gp_loss.mixing_factors_override = g['mixing']
    try:
        cost, rl, fl = pg.wgan_gp_D_loss(D, G, g['real'].cuda(), g['z1'].cuda())
        cost.backward()
    finally:
        pg.wgan_gp_loss.mixing_factors_override = None
    assert rel_err(cost, g['d_cost']) < TOL and rel_err(rl, g['d_real_loss']) < TOL and rel_err(fl, g['d_fake_loss']) < TOL
    grads = gpu['named_grads'](D)
    assert set(grads) == set(g['dgrad'])
    errs = {k: rel_err(grads[k], v) for k, v in g['dgrad'].items()}
    gcost = pg.wgan_gp_G_loss(G, D, g['z2'].cuda())
    gcost.backward()
    assert rel_err(gcost, g['g_cost']) < TOL
    grads = gpu['named_grads'](G)
    assert set(grads) == set(g['ggrad'])
    errs.update({'G.' + k: rel_err(grads[k], v) for k, v in g['ggrad'].items()})
    worst = max(errs, key=errs.get)
    report(test='edge_case_golden', case=case, worst=worst, worst_err=errs[worst])
    bad = {k: v for k, v in errs.items() if not v < TOL}
    assert not bad, bad


# ---- gradient error over unscreened inputs -------------------------------------------------------------------------
def test_gradient_error_over_20_unscreened_seeds(gpu):
    """No input screening: 20 consecutive seeds, depth 3 of a 32x32 model with 64..32 feature maps, N = 4.  For every
    seed the worst per-tensor gradient error of the CUDA path and of the fp32 ORACLE, both against the fp64 oracle.
    A LeakyReLU unit whose pre-activation lies within fp32 rounding noise of zero takes the other slope under another
    summation order -- in the reference's own fp32 arithmetic too -- so single seeds may exceed 1e-3 on either side;
    gated: the median over the seeds holds 1e-3, and no seed is worse than 3x what the fp32 oracle shows at worst."""
    O, pg = gpu['O'], gpu['pg']
    res, ch, fb, fm, lat, depth, alpha, n = 32, 3, 512, 64, 64, 3, 0.5, 4
    pgp = O.make_generator_params(res, ch, fmap_base=fb, fmap_max=fm, latent_size=lat, seed=3)
    pdp = O.make_discriminator_params(res, ch, fmap_base=fb, fmap_max=fm, seed=4)
    f64 = lambda d: {k: v.double() for k, v in d.items()}
    pgp64, pdp64 = f64(pgp), f64(pdp)
    G, D = gpu['build_pair'](dict(resolution=res, channels=ch, fmap_base=fb, fmap_max=fm, latent=lat, pg=pgp, pd=pdp,
                                   depth=depth, alpha=alpha))
    nb = O.n_blocks_for(res)
    r = 4 * 2 ** depth
    e_cuda, e_o32 = [], []
    for seed in range(1000, 1020):
        gen = torch.Generator().manual_seed(seed)
        z1, z2 = torch.randn(n, lat, generator=gen), torch.randn(n, lat, generator=gen)
        real = torch.randn(n, ch, r, r, generator=gen)
        mix = torch.rand(n, 1, generator=gen)
        _, _, _, gd64 = O.d_step_grads(pdp64, pgp64, real.double(), z1.double(), mix.double(), depth, alpha, nb)
        _, gg64 = O.g_step_grads(pgp64, pdp64, z2.double(), depth, alpha, nb)
        _, _, _, gd32 = O.d_step_grads(pdp, pgp, real, z1, mix, depth, alpha, nb)
        _, gg32 = O.g_step_grads(pgp, pdp, z2, depth, alpha, nb)
        out = run_step(gpu, G, D, z1, z2, real, mix)
        worst = lambda gd, gg: max([rel_err(gd[k], v) for k, v in gd64.items()] + [rel_err(gg[k], v) for k, v in gg64.items()])
        e_cuda.append(worst(out['gd'], out['gg']))
        e_o32.append(worst(gd32, gg32))
    e_cuda, e_o32 = np.array(e_cuda), np.array(e_o32)
    report(test='unscreened_seeds', seeds=20, cuda_vs_fp64=[float('%.3g' % v) for v in e_cuda],
           fp32_oracle_vs_fp64=[float('%.3g' % v) for v in e_o32], cuda_median=float(np.median(e_cuda)),
           cuda_max=float(e_cuda.max()), fp32_oracle_median=float(np.median(e_o32)), fp32_oracle_max=float(e_o32.max()))
    assert np.median(e_cuda) < TOL
    assert e_cuda.max() < max(TOL, 3.0 * e_o32.max())
