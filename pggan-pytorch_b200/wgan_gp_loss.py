"""WGAN-GP losses with the reference's signatures (wgan_gp_loss.py:36-74), computed by the fused libpgk chains.

Both functions run forward AND backward immediately (Trainer.train() always calls ``.backward()`` on the returned
scalar, trainer.py:98,111) and return a scalar whose ``grad_fn`` deposits the pre-computed parameter gradients, so
``loss.backward(); optimizer.step()`` behaves exactly as with the reference.

Data parallel: when torch.distributed is initialised with world_size > 1, the flat gradient buffer of the step is
all-reduced (ONE NCCL call per optimizer step) and averaged before it is handed to autograd.
"""
import torch
import torch.distributed as dist

from . import _lib
from .engine import GradSet

call = _lib.call

# test hook: a (N,1) tensor used instead of drawing U(0,1) mixing factors (wgan_gp_loss.py:15-17)
mixing_factors_override = None
# last step's auxiliary outputs (per-sample gradient norms, penalty) for monitoring / tests
last_aux = {}
# test hook: keep the tapes (stored activations) of the last D / G step in last_aux['d_tape'] / ['g_tape'], so that a test
# can read the LeakyReLU decisions the kernels took (tests/test_gpu_baseline_widths.py imposes them on the oracle)
keep_tapes = False


class _Deposit(torch.autograd.Function):
    """value -> value, with d(value)/d(param_i) := grads_i (already computed by the kernels)."""

    @staticmethod
    def forward(ctx, value, n, *params_and_grads):
        ctx.n = n
        ctx.grads = params_and_grads[n:]
        return value.clone()

    @staticmethod
    def backward(ctx, gout):
        grads = torch._foreach_mul(list(ctx.grads), gout)
        return (None, None) + tuple(grads) + (None,) * ctx.n


def _deposit(value, gs):
    return _Deposit.apply(value, len(gs.params), *(gs.params + gs.grads()))


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _allreduce(gs):
    """Average the flat gradient buffer over the ranks (one collective), unless the step did it in buckets already."""
    if _world() > 1 and not getattr(gs, 'reduced', False):
        dist.all_reduce(gs.flat)
        gs.flat.div_(dist.get_world_size())


# ---- CUDA graphs -----------------------------------------------------------------------------
# cuda_graphs = True: the kernel sequence of a loss is captured once per (networks, depth, fading or not, batch shape,
# precision, loss constants) and replayed afterwards: several hundred launches per step become one graph launch, which
# is what bounds the low-resolution phases.  alpha is NOT part of the key: while a level fades in it changes every
# iteration (plugins.py:57-81), so the five kernels it enters read it from a two-float device buffer of the engine
# (engine.Fade, include/pgk.h "device-side fade-in scalars") that is rewritten before every replay -- one graph serves
# the whole transition phase.  A key is captured on its third use.  Parameters must keep their storage (Adam updates
# in place).
cuda_graphs = False
_graphs = {}


class _Captured(object):
    __slots__ = ('uses', 'graph', 'inputs', 'outputs', 'launches')

    def __init__(self):
        self.uses, self.graph, self.inputs, self.outputs, self.launches = 0, None, None, None, 0


def _fade_scalars(*models):
    """Graph mode: give every engine its device-side (alpha, 1 - alpha) pair and write the current values into it (two
    tiny launches per engine, outside the graph).  Eager mode: no device scalars, alpha travels as a kernel argument."""
    from .engine import Fade
    for m in models:
        e = m.engine
        if not cuda_graphs:
            e.fade_dev = None
            continue
        if e.fade_dev is None or e.fade_dev.device != next(m.parameters()).device:
            e.fade_dev = torch.empty(2, dtype=torch.float32, device=next(m.parameters()).device)
        Fade(float(m.alpha), e.fade_dev).write()


def _phase(m):
    """what a captured graph depends on: the depth and whether the level is fading in -- not alpha itself"""
    return int(m.depth), bool(int(m.depth) > 0 and float(m.alpha) < 1.0)


def _run(key, body, inputs):
    """body(*inputs) -> tuple of tensors / objects.  Eager for the first two uses of a key, then captured + replayed."""
    if not cuda_graphs:
        return body(*inputs)
    c = _graphs.get(key)
    if c is None:
        if len(_graphs) > 16:           # bounded cache: phases change only a handful of times per run
            for k in [k for k, v in _graphs.items() if v.graph is None] or list(_graphs):
                del _graphs[k]
        c = _graphs[key] = _Captured()
    c.uses += 1
    if c.uses <= 2:
        return body(*inputs)
    if c.graph is None:
        from . import engine
        c.inputs = [t.clone() for t in inputs]
        torch.cuda.synchronize()
        before = _lib.launch_count()
        g = torch.cuda.CUDAGraph()
        _run.epoch += 1
        engine.CAPTURE_EPOCH = _run.epoch   # the weight re-layout must be part of the graph whatever the cache says
        pdl = _lib.pdl_state()
        _lib.pdl_state(False)               # programmatic launch edges inside a graph measured slower than plain ones
        try:
            with torch.cuda.graph(g):
                c.outputs = body(*c.inputs)
        finally:
            engine.CAPTURE_EPOCH = 0
            _lib.pdl_state(pdl)
        c.launches = _lib.launch_count() - before
        c.graph = g
    else:
        for dst, src in zip(c.inputs, inputs):
            dst.copy_(src, non_blocking=True)
        _lib.add_launches(c.launches)
    c.graph.replay()
    return c.outputs


_run.epoch = 0


def _d_body(D, G, iwass_lambda, iwass_epsilon, iwass_target):
    def body(real, z, eps):
        n, C, r = real.shape[0], real.shape[1], real.shape[-1]
        dev = real.device
        P = D.planes
        ed, eg = D.engine, G.engine
        per = C * r * r
        # input batch of D: [real | fake | mixed]
        ximg = torch.empty((3 * n, C, r, r), dtype=torch.float32, device=dev)
        ximg[:n].copy_(real)
        eg.forward(z, G.planes, out=ximg[n:2 * n])                  # fake = G(z), no graph (wgan_gp_loss.py:51-52)
        call('pgk_interpolate', ximg.data_ptr(), ximg[n:].data_ptr(), eps.data_ptr(), n, per, ximg[2 * n:].data_ptr())

        T = ed.forward(ximg, 3, n, P, slots=n)
        f32 = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        d_real_loss, d_fake_loss, seed, wseed = f32(n), f32(n), f32(3 * n), f32(3 * n)
        call('pgk_d_loss_seed', T.scores.data_ptr(), n, iwass_epsilon, d_real_loss.data_ptr(), d_fake_loss.data_ptr(),
             seed.data_ptr(), wseed.data_ptr())

        gs = GradSet(ed.active_params(T.depth, T.fade), dev)
        # u-chain for all three groups at once; the mixed group's chain is the create_graph gradient of the penalty
        ed.backward_head(T, seed, wseed, gs)
        ed.backward_body(T, T.d_hin, 0, 3 * n, 0)
        g_img = f32(n, C, r, r)
        ed.image_grad(T, 2 * n, n, g_img)
        norms, gp, v0, cost = f32(2 * n), f32(n), f32(n, C, r, r), f32(1)
        call('pgk_gp_penalty', g_img.data_ptr(), n, per, iwass_lambda, iwass_target, d_real_loss.data_ptr(),
             d_fake_loss.data_ptr(), norms.data_ptr(), gp.data_ptr(), v0.data_ptr(), cost.data_ptr())
        # second order: v-chain (adjoint of the u-chain) and the w-chain entering through MinibatchStddev
        ed.v_chain(T, v0, 2 * n, 3 * n)
        v_l2 = T.l2.sl(3 * n, 4 * n).g()
        call('pgk_colsum', v_l2.ptr, v_l2.ps, v_l2.P, n, v_l2.C, 1.0, gs[D.linear.weight].data_ptr())
        ed.backward_body(T, T.w_h, 2 * n, 3 * n, 3 * n)
        top_pairs = [(ximg, 0, 0, True), (ximg, n, n, True), (ximg, 2 * n, 3 * n, True), (v0, 0, 2 * n, False)]
        low_pairs = None
        if T.fade:
            low_pairs = [(T.xlow, 0, 0, True), (T.xlow, n, n, True), (T.xlow, 2 * n, 3 * n, True),
                         (T.v0low, 0, 2 * n, False)]
        # Data parallel: the gradients of the low-resolution blocks (~90 % of the bytes) are final first; their
        # all-reduce runs on NCCL's stream underneath the long weight gradients of the high-resolution blocks, the
        # small remainder follows at the end (two collectives instead of one exposed one).  Not while capturing a graph.
        split = ed.first_bucket_elems(T.depth)
        works, hook = [], None
        if split and _world() > 1 and not cuda_graphs and gs.flat.is_cuda:
            def hook():
                works.append(dist.all_reduce(gs.flat[:split], op=dist.ReduceOp.AVG, async_op=True))
        ed.param_grads(T, gs,
                       groups=[(0, 0), (n, n), (2 * n, 3 * n), (3 * n, 2 * n)], bias_goffs=[0, n, 3 * n],
                       head_groups=[(0, 0), (n, n), (3 * n, 2 * n)], head_bias_goffs=[0, n],
                       img_pairs=dict(top=top_pairs, low=low_pairs), ev_pair=(T.ev, 2 * n, 3 * n),
                       first_bucket_done=hook)
        if hook is not None:
            works.append(dist.all_reduce(gs.flat[split:], op=dist.ReduceOp.AVG, async_op=True))
            for w in works:
                w.wait()            # the current stream waits for NCCL's; the host does not
            gs.reduced = True
        if keep_tapes:
            last_aux['d_tape'] = T
        return cost, d_real_loss, d_fake_loss, gs, norms, gp
    return body


def wgan_gp_D_loss(D, G, real_images_in, fake_latents_in, iwass_lambda=10.0, iwass_epsilon=0.001, iwass_target=1.0,
                   return_all=True):
    D.zero_grad()
    G.zero_grad()
    real = D._input(real_images_in)
    z = G._input(fake_latents_in)
    n, dev = real.shape[0], real.device
    if mixing_factors_override is not None:
        eps = mixing_factors_override.to(dev, torch.float32).contiguous().view(-1)
    else:
        eps = torch.rand(n, device=dev, dtype=torch.float32)       # wgan_gp_loss.py:15-17
    _fade_scalars(D, G)
    key = ('D', id(D), id(G), _phase(D), _phase(G), tuple(real.shape),
           tuple(z.shape), D.precision, G.precision, float(iwass_lambda), float(iwass_epsilon), float(iwass_target))
    cost, d_real_loss, d_fake_loss, gs, norms, gp = _run(
        key, _d_body(D, G, iwass_lambda, iwass_epsilon, iwass_target), (real, z, eps))
    _allreduce(gs)
    last_aux.update(grad_norms=norms[:n], gradient_penalty=gp, mixing=eps)
    D_cost = _deposit(cost.view(()), gs)
    if return_all:
        return D_cost, d_real_loss.view(n, 1).clone(), d_fake_loss.view(n, 1).clone()
    return D_cost


def _g_body(G, D):
    def body(z):
        n, dev = z.shape[0], z.device
        ed, eg = D.engine, G.engine
        img, TG = eg.forward(z, G.planes, tape=True)
        T = ed.forward(img, 1, n, D.planes)
        seed = torch.empty(n, dtype=torch.float32, device=dev)
        call('pgk_fill', seed.data_ptr(), n, -1.0 / n)
        ed.backward_head(T, seed, None, None)
        ed.backward_body(T, T.d_hin, 0, n, 0)
        dimg = torch.empty_like(img)
        ed.image_grad(T, 0, n, dimg)
        gs = GradSet(eg.active_params(TG.depth, TG.fade), dev)
        eg.backward(TG, dimg, gs)
        cost = torch.empty(1, dtype=torch.float32, device=dev)
        call('pgk_mean_scale', T.scores.data_ptr(), n, -1.0, cost.data_ptr())
        if keep_tapes:
            last_aux['g_tape'] = (TG, T)
        return cost, gs
    return body


def wgan_gp_G_loss(G, D, fake_latents_in):
    G.zero_grad()
    z = G._input(fake_latents_in)
    _fade_scalars(D, G)
    key = ('G', id(G), id(D), _phase(D), _phase(G), tuple(z.shape), D.precision, G.precision)
    cost, gs = _run(key, _g_body(G, D), (z,))
    _allreduce(gs)
    return _deposit(cost.view(()), gs)
