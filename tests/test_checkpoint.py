"""Reference snapshots (SaverPlugin, plugins.py:142-174: torch.save of whole modules) -> pggan_b200 modules.
The fixtures under tests/golden/ref_snapshot_* were written by the unmodified reference
(tests/golden/make_golden_snapshot.py); loading them here must not need the reference on sys.path."""
import io
import os
import sys

import numpy as np
import pytest
import torch

from _util import GOLDEN, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pggan_b200 as pg  # noqa: E402

G_PATH = os.path.join(GOLDEN, 'ref_snapshot_generator.dat')
D_PATH = os.path.join(GOLDEN, 'ref_snapshot_discriminator.dat')


@pytest.fixture(scope='module')
def golden():
    return np.load(os.path.join(GOLDEN, 'ref_snapshot.npz'))


def _check_params(module, golden, prefix):
    sd = module.state_dict()
    mods = dict(module.named_modules())
    seen = 0
    for k in golden.files:
        if not k.startswith(prefix):
            continue
        name = k[len(prefix):]
        if name.endswith('.c'):
            assert mods[name[:-2]].c == pytest.approx(float(golden[k]), rel=0, abs=1e-9), name
        else:
            assert torch.equal(sd[name], torch.from_numpy(golden[k])), name
            seen += 1
    assert seen == len(sd), 'every parameter of the rebuilt module comes from the snapshot'


def test_snapshots_load_without_the_reference(golden):
    assert 'network' not in sys.modules, 'the reference must not be importable in this test'
    G, D = pg.resume(G_PATH, D_PATH)
    assert isinstance(G, pg.Generator) and isinstance(D, pg.Discriminator)
    assert (G.depth, D.depth) == (2, 2) and G.alpha == 0.7 and D.alpha == 0.7          # plugins.py:66,76
    assert G.max_depth == D.max_depth == 2 and G.latent_size == 16 and G.normalize_latents is True
    _check_params(G, golden, 'G.')
    _check_params(D, golden, 'D.')


def test_generator_and_discriminator_are_told_apart():
    with pytest.raises(ValueError):
        pg.resume(D_PATH, G_PATH)


def test_own_snapshots_round_trip(golden):
    """Our modules under the reference's SaverPlugin: torch.save(model) / torch.load give the module back,
    c constants included (they are plain attributes, as in the reference)."""
    G = pg.load_snapshot(G_PATH)
    buf = io.BytesIO()
    torch.save(G, buf)
    path = os.path.join(os.environ.get('TMPDIR', '/tmp'), 'pgk_own_snapshot_%d.dat' % os.getpid())
    with open(path, 'wb') as f:
        f.write(buf.getvalue())
    try:
        G2 = pg.load_snapshot(path)
    finally:
        os.remove(path)
    assert isinstance(G2, pg.Generator) and G2.depth == 2 and G2.alpha == 0.7
    _check_params(G2, golden, 'G.')


def test_state_dict_is_not_a_snapshot(tmp_path):
    p = tmp_path / 'sd.dat'
    torch.save(pg.load_snapshot(G_PATH).state_dict(), p)
    with pytest.raises(ValueError):
        pg.load_snapshot(str(p))


@pytest.mark.gpu
def test_loaded_snapshot_reproduces_the_reference_outputs(golden):
    """generate.py:18-25 / OutputGenerator: G(z) from the loaded generator, and D(real), equal what the pickled
    reference modules computed (1e-3 rel, the fp32-faithful mode)."""
    G, D = pg.resume(G_PATH, D_PATH, device='cuda')
    fake = G(torch.from_numpy(golden['z']).cuda())
    assert rel_err(fake, torch.from_numpy(golden['fake'])) < 1e-3
    scores = D(torch.from_numpy(golden['real']).cuda())
    assert rel_err(scores, torch.from_numpy(golden['scores'])) < 1e-3
