"""CPU: host-side logic of the product package -- C ABI surface, schedule, module surface, trainer plumbing and the
data-parallel gradient exchange (gloo, world_size 2).  No kernels are launched here."""
import os
import pickle
import re
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _util import load_schedule

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import pggan_b200 as pg  # noqa: E402


def test_library_exports_every_declared_symbol():
    """include/pgk.h is the contract: every function it declares is exported by libpgk.so and bound in _lib.py."""
    hdr = open(os.path.join(ROOT, 'include', 'pgk.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(pgk_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) > 30
    lib = pg._lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), '%s declared in pgk.h but not exported' % name
    assert declared == set(pg._lib.exported_symbols()), declared ^ set(pg._lib.exported_symbols())
    assert lib.pgk_version() >= 100


def test_no_cpu_fallback():
    G = pg.Generator((1, 3, 16, 16), fmap_base=64, fmap_max=16, latent_size=16)
    D = pg.Discriminator((1, 3, 16, 16), fmap_base=64, fmap_max=16)
    with pytest.raises(RuntimeError):
        G(torch.randn(2, 16))
    with pytest.raises(RuntimeError):
        D(torch.randn(2, 3, 4, 4))
    with pytest.raises(RuntimeError):
        pg.wgan_gp_G_loss(G, D, torch.randn(2, 16))


def test_module_surface_matches_reference():
    """constructor kwargs / attributes / state_dict keys of network.py:75-116,190-223 (names from SURVEY.md 5)."""
    G = pg.Generator((1, 3, 32, 32))
    D = pg.Discriminator((1, 3, 32, 32))
    assert (G.depth, G.alpha, G.latent_size, G.max_depth) == (0, 1.0, 512, 3)
    assert (D.depth, D.alpha, D.max_depth) == (0, 1.0, 3)
    gk, dk = set(G.state_dict()), set(D.state_dict())
    assert {'block0.c1.conv.weight', 'block0.toRGB.conv.bias', 'blocks.2.c2.conv.weight'} <= gk
    assert {'blocks.0.fromRGB.conv.weight', 'blocks.3.c1.conv.weight', 'linear.weight', 'linear.bias'} <= dk
    assert G.block0.c1.conv.weight.shape == (512, 512, 4, 4)
    assert D.blocks[3].c1.conv.weight.shape == (512, 513, 3, 3)
    assert D.blocks[3].c2.conv.weight.shape == (512, 512, 4, 4)
    assert not any(k.endswith('.c') for k in gk | dk), 'c is a plain attribute, not in the state_dict (network.py:19)'
    # 1024x1024 parameter counts probed on the reference
    assert sum(p.numel() for p in pg.Generator((1, 3, 1024, 1024)).parameters()) == 18359731
    assert sum(p.numel() for p in pg.Discriminator((1, 3, 1024, 1024)).parameters()) == 18367369
    # SaverPlugin pickles whole modules (plugins.py:158-166)
    G2 = pickle.loads(pickle.dumps(G))
    assert set(G2.state_dict()) == gk and abs(G2.block0.c1.c - G.block0.c1.c) == 0


def test_depth_manager_drives_modules_like_the_reference():
    pts = [p for p in load_schedule()['points'] if p['max_depth'] == 8]
    G = pg.Generator((1, 3, 1024, 1024), fmap_base=4096, fmap_max=8, latent_size=8)
    D = pg.Discriminator((1, 3, 1024, 1024), fmap_base=4096, fmap_max=8)
    assert G.max_depth == D.max_depth == 8

    class DS(object):
        model_depth, alpha = 0, 1.0

    made = []
    t = pg.Trainer(D, G, None, None, None, None, DS(), None, None)
    dm = pg.DepthManager(lambda mb: made.append(mb) or iter(()), lambda mb: (lambda: None), G.max_depth)
    t.register_plugin(dm)
    assert (G.depth, G.alpha, D.depth, D.alpha) == (0, 1.0, 0, 1.0) and made == [16]
    for p in pts:
        t.cur_nimg = p['cur_nimg']
        dm.iteration()
        assert (G.depth, D.depth, t.dataset.model_depth) == (p['depth'],) * 3
        assert repr(G.alpha) == p['alpha'] and D.alpha == G.alpha == t.dataset.alpha      # bit exact (repr of the float)
        assert t.tick_duration_nimg == p['tick_nimg']
        assert t.stats['minibatch_size'] == p['minibatch']
    assert len(pts) > 10


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from importlib import import_module
        engine = import_module('pggan-pytorch_b200.engine')
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.zeros(3, 5)), torch.nn.Parameter(torch.zeros(7))]
        gs = engine.GradSet(params, 'cpu')
        gs[params[0]].fill_(float(rank + 1))
        gs[params[1]].copy_(torch.arange(7.0) * (rank + 1))
        pg.wgan_gp_loss._allreduce(gs)          # ONE all-reduce of the flat buffer, averaged
        loss = pg.wgan_gp_loss._deposit(torch.tensor(2.0), gs)
        loss.backward()
        out[rank] = (params[0].grad.clone(), params[1].grad.clone(), float(loss))
    finally:
        dist.destroy_process_group()


def test_data_parallel_gradient_exchange_gloo_world2():
    """Exact DP oracle: every rank ends with the mean over ranks of the per-rank gradients."""
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_dp_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    for rank in range(world):
        g0, g1, loss = res[rank]
        assert torch.equal(g0, torch.full((3, 5), 1.5))
        assert torch.equal(g1, torch.arange(7.0) * 1.5)
        assert loss == 2.0


def test_deposit_scales_with_upstream_gradient():
    from importlib import import_module
    engine = import_module('pggan-pytorch_b200.engine')
    p = torch.nn.Parameter(torch.zeros(4))
    gs = engine.GradSet([p], 'cpu')
    gs[p].copy_(torch.tensor([1.0, 2.0, 3.0, 4.0]))
    (pg.wgan_gp_loss._deposit(torch.tensor(1.0), gs) * 0.5).backward()
    assert torch.equal(p.grad, torch.tensor([0.5, 1.0, 1.5, 2.0]))


def _fake_trainer(batches, prefetch):
    """A Trainer whose losses / optimizers are stubs and whose device transfer is the identity (host logic only)."""
    seen = []

    class Opt(object):
        def step(self):
            pass

    def d_loss(D, G, real, lat):
        seen.append(int(real[0]))
        return (torch.zeros((), requires_grad=True) * 1.0, torch.zeros(1), torch.zeros(1))

    def g_loss(G, D, lat):
        return torch.zeros((), requires_grad=True) * 1.0

    t = pg.Trainer(None, None, d_loss, g_loss, Opt(), Opt(), None, iter(batches), lambda: torch.zeros(2, 4))
    t._to_device = lambda x: x
    t.prefetch_reals = prefetch
    return t, seen


def test_real_batch_prefetch_keeps_the_reference_order_and_drops_stale_batches():
    """Opt-in look-ahead of the real batch (trainer.prefetch_reals): same batches in the same order, a batch fetched
    from an iterator that a plugin has replaced is dropped, StopIteration surfaces in the same train() call."""
    mk = lambda lo, hi: [torch.full((3,), float(i)) for i in range(lo, hi)]
    for prefetch in (False, True):
        t, seen = _fake_trainer(mk(0, 4), prefetch)
        t.train(), t.train()
        assert seen == [0, 1] and t.cur_nimg == 6
        # a plugin swaps the loader between iterations (DepthManager on a depth change, plugins.py:65-77)
        t.dataiter = iter(mk(100, 102))
        t.train(), t.train()
        assert seen == [0, 1, 100, 101]
        with pytest.raises(StopIteration):      # the new iterator is exhausted: raised by THIS call, as in the reference
            t.train()
        assert t.cur_nimg == 12 and t.iterations == 4
    # the look-ahead batch really is fetched before the plugins run (its copy has to be in flight under the step)
    t, seen = _fake_trainer(mk(0, 3), True)
    t.train()
    assert t._ahead is not None and int(t._ahead[1][0]) == 1


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md's table maps every function of include/pgk.h to the reference lines it replaces."""
    hdr = re.sub(r'/\*.*?\*/', '', open(os.path.join(ROOT, 'include', 'pgk.h')).read(), flags=re.S)
    declared = set(re.findall(r'\b(pgk_[a-z0-9_]+)\s*\(', hdr))
    doc = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    missing = sorted(d for d in declared if d not in doc and not d.startswith('pgk_prof_'))
    assert not missing, missing


def test_forward_conv_takes_the_fp16_operands_only_where_it_may(monkeypatch):
    """engine.conv under PGK_FWD_FP16 (host logic only, kernels stubbed): a forward conv of the fp32-faithful mode on a
    wide-kernel shape makes the fp16 copy and runs pgk_conv_fp16 (pixel norm chained); gradient-chain convs (mask),
    scaled outputs, the bf16 mode and layers without an fp16 weight packing keep pgk_conv."""
    from importlib import import_module
    E = import_module('pggan-pytorch_b200.engine')
    calls = []
    monkeypatch.setattr(E, 'call', lambda name, *a: calls.append((name, a)))

    class FakeLib(object):
        @staticmethod
        def pgk_conv_tc_supported(n, h, w, cin, cout, ks, ups):
            return int(cin % 64 == 0 and not ups)
    monkeypatch.setattr(E._lib, 'load', lambda: FakeLib)
    monkeypatch.setattr(E, 'FWD_FP16', True)
    x3, o3 = E.PT.empty(4, 8, 8, 64, 3, 'cpu'), E.PT.empty(4, 8, 8, 128, 3, 'cpu')
    wf, wt, wh = torch.zeros(9 * 64 * 128), torch.zeros(3, 128, 9 * 64, dtype=torch.bfloat16), \
        torch.zeros(2, 128, 9 * 64, dtype=torch.float16)
    r = torch.zeros(4 * 8 * 8)
    E.conv(x3.sl(0, 3), (wf, wt, wh), 128, 3, o3.sl(0, 3), act=1, fwd=True, pn_r=r)
    assert [c[0] for c in calls] == ['pgk_cvt_fp16x2', 'pgk_conv_fp16', 'pgk_pixelnorm']
    assert calls[0][1][3] == 3 * x3.per                       # the copy covers the three samples of the slice
    assert calls[1][1][8] == wh.data_ptr()
    for kw, w in ((dict(fwd=False, mask=o3), (wf, wt, wh)),      # a gradient-chain conv
                  (dict(fwd=True, scale=0.5), (wf, wt, wh)),     # scaled output
                  (dict(fwd=True), (wf, wt))):                   # no fp16 packing of the weights (thin layers)
        calls.clear()
        E.conv(x3, w, 128, 3, o3, **kw)
        assert [c[0] for c in calls] == ['pgk_conv'], kw
    calls.clear()
    x1, o1 = E.PT.empty(4, 8, 8, 64, 1, 'cpu'), E.PT.empty(4, 8, 8, 128, 1, 'cpu')
    E.conv(x1, (wf, wt, wh), 128, 3, o1, fwd=True)            # the bf16 mode has one plane: nothing to gain
    assert [c[0] for c in calls] == ['pgk_conv']



@pytest.mark.parametrize('fp16', [False, True])
def test_python_sequencing_of_one_iteration_with_stubbed_kernels(monkeypatch, fp16):
    """The host side of a D step and a G step (engine.py / wgan_gp_loss.py) runs end to end on CPU tensors when every
    libpgk call is replaced by a recorder: no kernel runs, but every tensor is allocated, every launch is sequenced and
    autograd receives the deposited gradients -- exactly the parameters of the active set get a .grad
    (network.py:118-139, 225-240; inactive blocks keep None, as in the reference).  With fp16=True the same under the
    two opt-in fp16 switches (their routing code is host logic too)."""
    import collections
    from importlib import import_module
    E = import_module('pggan-pytorch_b200.engine')
    L = import_module('pggan-pytorch_b200.wgan_gp_loss')
    calls = collections.Counter()

    def fake_call(name, *a):
        calls[name] += 1
    for m in (E, L, pg._lib):
        monkeypatch.setattr(m, 'call', fake_call)
    monkeypatch.setattr(E, 'FWD_FP16', fp16)
    for cls in (pg.Generator, pg.Discriminator):
        monkeypatch.setattr(cls, '_input', lambda self, x: x.contiguous().float())
    for depth, alpha, nd_expect, ng_expect in ((0, 1.0, 8, 6), (2, 0.5, 18, 16), (3, 1.0, 20, 18)):
        G = pg.Generator((None, 3, 32, 32), fmap_base=1024, fmap_max=128, latent_size=64)
        D = pg.Discriminator((None, 3, 32, 32), fmap_base=1024, fmap_max=128)
        G.depth = D.depth = depth
        G.alpha = D.alpha = alpha
        r = 4 * 2 ** depth
        calls.clear()
        cost, d_real, d_fake = pg.wgan_gp_D_loss(D, G, torch.randn(4, 3, r, r), torch.randn(4, 64))
        assert tuple(d_real.shape) == (4, 1) and tuple(d_fake.shape) == (4, 1) and cost.dim() == 0
        cost.backward()
        assert sum(p.grad is not None for p in D.parameters()) == nd_expect
        assert all(p.grad is None for p in G.parameters())
        assert calls['pgk_gp_penalty'] == 1 and calls['pgk_stddev_bwd2'] == 1
        if fp16:
            # (inputs written by from_rgb / pool2 / the upsample / the pixel norm arrive with their half planes: fewer
            # conversion passes than half-operand convs)
            assert calls['pgk_conv_fp16'] > 0 and 0 < calls['pgk_cvt_fp16x2'] < calls['pgk_conv_fp16']
        else:
            assert calls['pgk_conv_fp16'] == 0 and calls['pgk_pack_operand_fp16'] == 0
        gcost = pg.wgan_gp_G_loss(G, D, torch.randn(4, 64))
        gcost.backward()
        assert sum(p.grad is not None for p in G.parameters()) == ng_expect


def test_async_loss_monitor_reads_once_per_epoch():
    """AsyncLossMonitor: running sums stay tensors during the iterations; `trainer.stats` gets the epoch mean and the
    last value at the epoch boundary, with EfficientLossMonitor's conventions (plugins.py:102-111: losses 0 / 1 are
    scalars, 2 / 3 are (N,1) tensors whose mean is logged)."""
    class T(object):
        stats = {}
    m0, m2 = pg.AsyncLossMonitor(0, 'G_loss'), pg.AsyncLossMonitor(2, 'D_real')
    t = T()
    m0.register(t), m2.register(t)
    for it in range(1, 4):
        args = (torch.tensor(float(it)), torch.tensor(0.0), torch.full((4, 1), 2.0 * it), torch.zeros(4, 1))
        m0.iteration(it, *args), m2.iteration(it, *args)
        assert 'epoch_mean' not in t.stats['G_loss']
    m0.epoch(1), m2.epoch(1)
    assert t.stats['G_loss']['epoch_mean'] == 2.0 and t.stats['G_loss']['last'] == 3.0
    assert t.stats['D_real']['epoch_mean'] == 4.0 and t.stats['D_real']['last'] == 6.0
    m0.epoch(2)     # an epoch without iterations leaves the stats alone
    assert t.stats['G_loss']['epoch_mean'] == 2.0


def test_device_random_latents_has_the_shape_and_statistics_of_random_latents():
    draw = pg.device_random_latents(64, 512, device='cpu', seed=3)
    a, b = draw(), draw()
    assert tuple(a.shape) == (64, 512) and a.dtype == torch.float32 and not torch.equal(a, b)
    assert abs(float(a.mean())) < 0.05 and abs(float(a.std()) - 1.0) < 0.05
    ref = pg.random_latents(64, 512)
    assert tuple(ref.shape) == tuple(a.shape) and ref.dtype == a.dtype


def test_first_gradient_bucket_is_a_prefix_of_the_flat_buffer():
    """DEngine orders parameters (= the flat gradient buffer) low resolution first, and first_bucket_elems() is the
    size of a PREFIX of it: linear + blocks[-1..-4] -- what the D step all-reduces while the high-resolution weight
    gradients still run.  The active set itself is unchanged (same parameters as the reference gives a gradient)."""
    D = pg.Discriminator((None, 3, 1024, 1024))
    e = D.engine
    for depth in range(0, 9):
        for fade in (False, True):
            if fade and depth == 0:
                continue
            ps = e.active_params(depth, fade)
            assert len({id(p) for p in ps}) == len(ps) == 2 + 4 * (depth + 1) + 2 + (2 if fade else 0)
            split = e.first_bucket_elems(depth)
            if depth < 4:
                assert split == 0
                continue
            first = [D.linear.weight, D.linear.bias]
            for k in range(1, 5):
                b = D.blocks[len(D.blocks) - k]
                first += [b.c1.conv.weight, b.c1.conv.bias, b.c2.conv.weight, b.c2.conv.bias]
            assert [id(p) for p in ps[:len(first)]] == [id(p) for p in first]
            assert split == sum(p.numel() for p in first)
            assert split > 0.85 * sum(p.numel() for p in ps)       # the 512-channel layers: most of the bytes
