"""-m gpu: single-kernel numerics.  Every tensor-core kernel family of libpgk (pgk_conv / pgk_wgrad through the C ABI)
against a plain PyTorch fp32 reference of the same op on the same random operands (tools/tc_test.py holds the
case builders; they also run the library's CUDA-core kernel on the same inputs for comparison).

Tolerances are those of the operand format, stated in tools/tc_test.py: 2e-2 with one bf16 plane (8 mantissa bits),
1e-4 with two planes, 2e-5 with three (the fp32-faithful mode).  The shapes cover what the full-step tests cannot
reach cheaply: every (Cin, Cout) pair of the row-streaming thin kernels, several units per CTA and both CTA-per-SM
plans (PGK_THIN_OCC), masks with every accumulator width, and the small-reduction weight gradients."""
import importlib
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def T():
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    return importlib.import_module('tc_test')


# (N, H, W, Cin, Cout, planes, kwargs)
THIN_CONV = [
    (1, 128, 128, 16, 16, 1, {}),
    (2, 256, 256, 16, 16, 3, {}),
    (1, 256, 256, 8, 8, 1, {}),
    (1, 256, 256, 8, 16, 3, dict(mask=True, act=0, bias=False, scale=0.25, fwd=False)),
    (2, 128, 256, 32, 32, 1, dict(mask=True)),
    (1, 256, 256, 32, 64, 3, {}),
    (1, 512, 512, 16, 8, 1, dict(mask=True, act=0, fwd=False)),
    (3, 128, 128, 32, 16, 2, {}),
    # several units per CTA, every chunk size the launcher can pick, masks on the widest accumulators
    (5, 512, 512, 8, 16, 1, dict(mask=True)),
    (3, 1024, 1024, 16, 8, 1, dict(mask=True, act=0, bias=False, fwd=False)),
    (7, 256, 256, 32, 64, 1, dict(mask=True, act=0, bias=False)),
    (4, 256, 256, 32, 64, 2, dict(mask=True, act=0, bias=False, fwd=False)),
    (2, 8, 128, 8, 32, 1, dict(mask=True)),
    (1, 24, 384, 16, 64, 3, {}),
    (9, 64, 128, 32, 8, 2, dict(mask=True, fwd=False)),
]


@pytest.mark.parametrize('case', THIN_CONV, ids=lambda c: 'N%d_%dx%d_%d-%d_P%d%s' % (c[0], c[1], c[2], c[3], c[4], c[5], '_mask' if c[6].get('mask') else ''))
def test_thin_conv(T, case):
    n, h, w, ci, co, p, kw = case
    assert T.conv_case(n, h, w, ci, co, 3, p, **kw)


WIDE_CONV = [
    (2, 16, 16, 64, 64, 3, 1, {}),
    (2, 16, 16, 64, 64, 3, 3, {}),
    (3, 4, 4, 128, 64, 3, 2, dict(pos=True)),
    (5, 8, 8, 64, 128, 3, 3, dict(mask=True, act=0, bias=False, scale=0.25)),
    (2, 32, 32, 128, 256, 3, 3, {}),
    (1, 64, 64, 256, 512, 3, 1, {}),
    (2, 128, 128, 64, 16, 3, 3, {}),
    (1, 256, 256, 64, 32, 3, 1, dict(mask=True)),
    (130, 1, 1, 512, 2048, 1, 3, {}),
    (7, 1, 1, 1024, 64, 1, 3, dict(mask=True, act=0)),
    (3, 32, 32, 256, 512, 3, 3, dict(mask=True, act=0, bias=False, fwd=False)),
    (300, 4, 4, 64, 64, 3, 1, {}),
]


@pytest.mark.parametrize('case', WIDE_CONV, ids=lambda c: 'N%d_%dx%d_%d-%d_k%d_P%d' % c[:7])
def test_wide_conv(T, case):
    n, h, w, ci, co, ks, p, kw = case
    assert T.conv_case(n, h, w, ci, co, ks, p, **kw)


# (N, H, W, Cin, Cout, planes, groups)
WGRAD = [
    (1, 128, 128, 16, 16, 1, 1),
    (2, 256, 256, 16, 16, 2, 2),
    (1, 256, 256, 8, 8, 1, 3),
    (1, 128, 256, 8, 16, 3, 1),
    (2, 128, 128, 32, 32, 1, 4),
    (1, 256, 256, 32, 64, 1, 1),
    (1, 512, 512, 16, 8, 1, 1),
    (2, 256, 256, 64, 32, 1, 2),
    (1, 128, 128, 64, 16, 2, 1),
    (3, 512, 512, 8, 8, 1, 4),
    (4, 16, 16, 64, 64, 1, 1),
    (4, 16, 16, 64, 64, 3, 1),
    (16, 4, 4, 128, 64, 3, 2),
    (8, 8, 8, 64, 128, 2, 3),
    (2, 32, 32, 128, 256, 3, 4),
    (2, 64, 64, 256, 512, 1, 1),
    (1, 128, 128, 64, 64, 3, 2),
    # small reductions (the 4x4 ... 16x16 levels at batch 4): tensor-core path from PGK_WGRAD_TC_MIN pixels up
    (4, 4, 4, 512, 512, 1, 4),
    (4, 8, 8, 512, 512, 2, 1),
    (4, 16, 16, 256, 512, 1, 1),
]


@pytest.mark.parametrize('case', WGRAD, ids=lambda c: 'N%dx%d_%dx%d_%d-%d_P%d' % (c[0], c[6], c[1], c[2], c[3], c[4], c[5]))
def test_wgrad(T, case):
    n, h, w, ci, co, p, groups = case
    assert T.wgrad_case(n, h, w, ci, co, 3, p, ngroups=groups)


# conv + bias + LeakyReLU + pixel norm in one call (the generator's PGConv2d.forward, network.py:32-40): fused into the
# thin kernel's epilogue for Cout <= 32, a second in-place pass on the other kernels
PIXELNORM = [(2, 128, 128, 16, 8, 1), (1, 256, 256, 8, 8, 3), (3, 128, 256, 32, 16, 2), (2, 256, 128, 32, 32, 1),
             (1, 128, 128, 16, 64, 3), (2, 32, 32, 64, 32, 3), (3, 8, 8, 128, 128, 1)]


@pytest.mark.parametrize('case', PIXELNORM, ids=lambda c: 'N%d_%dx%d_%d-%d_P%d' % c)
def test_conv_pixelnorm(T, case):
    assert T.pixelnorm_case(*case)


# ---- opt-in paths written at the end of round 1 without a GPU: run with PGK_TEST_EXPERIMENTAL=1 until they have been
# verified once (the env-switched kernel flavours -- PGK_THIN_ATM, PGK_WTHIN_ATM, PGK_WTHIN_SW128, PGK_WGRAD_RED4,
# PGK_PREP_TILED -- are read once per process: run this whole file with the switch set, tools/gpu_call_r2_first.sh)
experimental = pytest.mark.skipif(os.environ.get('PGK_TEST_EXPERIMENTAL') != '1',
                                  reason='opt-in path not yet run on a GPU (PGK_TEST_EXPERIMENTAL=1 enables it)')

CONV_FP16 = [(2, 16, 16, 64, 64, 3, 3, False), (3, 4, 4, 512, 512, 3, 3, True), (2, 64, 64, 128, 256, 3, 3, False),
             (1, 128, 128, 64, 128, 3, 2, False), (9, 1, 1, 512, 8192, 1, 3, False), (5, 1, 1, 8192, 512, 1, 3, False)]


@experimental
@pytest.mark.parametrize('case', CONV_FP16, ids=lambda c: 'N%d_%dx%d_%d-%d_k%d_P%d' % c[:7])
def test_conv_on_half_operand_planes(T, case):
    n, h, w, ci, co, ks, p, pos = case
    assert T.conv_fp16_case(n, h, w, ci, co, ks, P=p, pos=pos)
