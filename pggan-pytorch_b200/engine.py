"""Kernel sequencing for the G + D + WGAN-GP step (host side, Python; all arithmetic is in libpgk.so).

Everything here is stream-ordered enqueueing of libpgk kernels on torch's current stream -- no host
synchronisation, no torch arithmetic on activations.  Tensors are allocated through torch's caching
allocator (plumbing) and handed to the C ABI as raw pointers.

Notation used below (see DESIGN.md "the gradient-penalty double backward"):
  x_l   activation after layer l (post LeakyReLU)            -- "planes" tensors, N,H,W,C
  ua_l  gradient w.r.t. the PRE-activation of layer l         = (incoming gradient) * lrelu'(x_l)
  u-chain  the ordinary data-gradient chain through D (network.py:225-240 backwards).  For the mixed samples it
           is the `autograd.grad(..., create_graph=True)` of wgan_gp_loss.py:25-28.
  v-chain  the adjoint of the u-chain: D's forward applied to v0 = dPenalty/dg with biases dropped and
           LeakyReLU replaced by the stored masks.  Gives dPenalty/dW as wgrad(v_{l-1}, ua_l).
  w-chain  the second-order term that enters the forward graph through MinibatchStddev (network.py:174-187).
"""
import ctypes
import os
from types import SimpleNamespace

import torch

from . import _lib

call = _lib.call
BF16 = torch.bfloat16
CAPTURE_EPOCH = 0    # > 0 while a CUDA graph is being captured: every weight re-layout must be recorded once
GRAD_PLANES = 2   # planes carried by gradient tensors / read by the gradient chains (include/pgk.h, pgk_conv: Pr)
W_CONV, W_GFIRST, W_DLAST = 0, 1, 2
# Forward convolutions of the fp32-faithful mode that run on the wide tensor-core kernel read a two-plane fp16 copy of
# their input and an fp16 packing of their weights (include/pgk.h "forward convolution on IEEE-half operand planes"):
# three products per FLOP instead of six at the same 22 operand bits.  Measured (round 2, B200): depth 4 / batch 128
# 88.96 -> 77.64 ms per iteration; every parity test (full widths, kernel-decision-conditioned gradients, unscreened
# seeds) is unchanged with it.  Default since the fault of its first version was found (the tensor maps declared a third
# operand plane behind the two-plane allocations: DESIGN.md 7c-4); PGK_FWD_FP16=0 keeps the six-product bf16 path (A/B).
FWD_FP16 = os.environ.get('PGK_FWD_FP16', '1') == '1'
# Producers of a tensor that such a conv reads next (from_rgb, pool2, the materialised upsample, the pixel norm) write
# the fp16 planes themselves (include/pgk.h: out16), which saves the pgk_cvt_fp16x2 pass; PGK_FUSE_CVT=0 for A/B runs.
FUSE_CVT = os.environ.get('PGK_FUSE_CVT', '1') == '1'
# 3x3 layers with 64 input channels and 32 / 64 output channels at W >= 128 (the 256^2 / 512^2 levels' 64 -> 32 and the
# 128^2 / 256^2 levels' 64 -> 64 convs and data gradients), one-plane mode: the wide kernel runs them at a quarter of
# the tensor rate (a tcgen05.mma of N = 32 ... 64 costs what one of N = 128 does; c3: 0.26-0.44 of the peak), the
# row-streaming thin kernel stacks the three filter rows along N and is bound by HBM.  Fourth element of the operand
# tuples: the layer's pgk_pack_thin packing.  PGK_THIN64=0 for A/B runs.
THIN64 = os.environ.get('PGK_THIN64', '1') == '1'


def _ints(vals):
    vals = list(vals)
    return (ctypes.c_int * 4)(*(vals + [0] * (4 - len(vals))))


class PT(object):
    """A 'planes' tensor: P planes of bf16 in N,H,W,C order; value = plane0 (+ plane1)."""
    __slots__ = ('t', 'N', 'H', 'W', 'C', 'P', 'off', 'ntot', 'aux')

    def __init__(self, t, N, H, W, C, P, off, ntot, aux=None):
        self.t, self.N, self.H, self.W, self.C, self.P, self.off, self.ntot = t, N, H, W, C, P, off, ntot
        self.aux = {} if aux is None else aux    # shared by every slice / view of the same storage

    @staticmethod
    def empty(N, H, W, C, P, device):
        t = torch.empty((P, N, H, W, C), dtype=BF16, device=device)
        return PT(t, N, H, W, C, P, 0, N)

    @property
    def per(self):
        return self.H * self.W * self.C

    @property
    def ps(self):
        return self.ntot * self.per

    @property
    def ptr(self):
        return self.t.data_ptr() + 2 * self.off * self.per

    def sl(self, n0, n1):
        return PT(self.t, n1 - n0, self.H, self.W, self.C, self.P, self.off + n0, self.ntot, self.aux)

    def view(self, H, W, C):
        assert H * W * C == self.per
        return PT(self.t, self.N, H, W, C, self.P, self.off, self.ntot, self.aux)

    def g(self):
        """The same storage seen through the planes the gradient chains use (GRAD_PLANES: 16 mantissa bits)."""
        return PT(self.t, self.N, self.H, self.W, self.C, min(self.P, GRAD_PLANES), self.off, self.ntot, self.aux)

    def float(self):
        """fp32 N,C,H,W copy (tests / debugging only)."""
        v = self.t.view(self.P, self.ntot, self.H, self.W, self.C)[:, self.off:self.off + self.N].float().sum(0)
        return v.permute(0, 3, 1, 2).contiguous()

    @staticmethod
    def from_float(x, P):
        """fp32 N,C,H,W -> planes (tests only; the product path never converts through torch)."""
        n, c, h, w = x.shape
        v = x.permute(0, 2, 3, 1).contiguous().float()
        out = PT.empty(n, h, w, c, P, x.device)
        for p in range(P):
            out.t[p] = v.to(BF16)
            v = v - out.t[p].float()
        return out


def _mask(m):
    return (m.ptr, m.ps) if m is not None else (None, 0)


def _h16_out(out, want):
    """(pointer, plane stride) of the IEEE-half companion a producer kernel fills next to `out`, or (None, 0).  The
    companion covers the whole storage (2 planes x ntot samples); aux['h16_range'] remembers which samples are valid."""
    if not (want and FWD_FP16 and FUSE_CVT and out.P >= 2 and out.C % 64 == 0):
        return None, 0
    h = out.aux.get('h16')
    if h is None or h.shape[1] != out.ntot * out.per:
        h = out.aux['h16'] = torch.empty((2, out.ntot * out.per), dtype=torch.float16, device=out.t.device)
    out.aux['h16_range'] = (out.off, out.off + out.N)
    return h.data_ptr() + 2 * out.off * out.per, out.ntot * out.per


def _h16_in(x):
    """(pointer, plane stride) of a valid companion of x, or None."""
    h, rng = x.aux.get('h16'), x.aux.get('h16_range')
    if h is None or rng is None or x.off < rng[0] or x.off + x.N > rng[1] or h.shape[1] != x.ntot * x.per:
        return None
    return h.data_ptr() + 2 * x.off * x.per, x.ntot * x.per


class Fade(object):
    """The fade-in factors alpha and 1 - alpha as the kernels take them: (host scalar, device pointer or None).
    Eager mode: the host scalar carries everything.  CUDA-graph mode (wgan_gp_loss.cuda_graphs): the host scalar is the
    alpha-free part and the kernel multiplies by the float at the device pointer, which the loss function rewrites
    before every replay -- alpha changes every iteration of a transition phase (plugins.py:57-81), the graph does not."""
    __slots__ = ('alpha', 'dev')

    def __init__(self, alpha, dev=None):
        self.alpha, self.dev = float(alpha), dev

    def a(self, k=1.0):
        return (k, self.dev.data_ptr()) if self.dev is not None else (k * self.alpha, None)

    def b(self, k=1.0):
        return (k, self.dev.data_ptr() + 4) if self.dev is not None else (k * (1.0 - self.alpha), None)

    def write(self):
        """(graph mode) put the current factors where the captured kernels read them"""
        if self.dev is not None:
            call('pgk_fill', self.dev.data_ptr(), 1, self.alpha)
            call('pgk_fill', self.dev.data_ptr() + 4, 1, 1.0 - self.alpha)


def _sc(v):
    """scalar argument -> (host scalar, device multiplier pointer or None)"""
    return v if isinstance(v, tuple) else (float(v), None)


# ---------------------------------------------------------------------------------------------
# thin op wrappers (argument marshalling only)
# ---------------------------------------------------------------------------------------------


def conv(x, w, cout, ks, out, ups=0, bias=None, posT=None, pos_s=None, act=0, mask=None, scale=1.0, fwd=False,
         pn_r=None, out_h16=False):
    """out <- pgk_conv(x); H, W are taken from `out`.  w = (fp32 [K][Cout] operand, bf16 planes [3][Cout][K] operand).
    fwd=True: a forward pass whose values decide LeakyReLU masks -- all planes are read.
    pn_r: fp32 (N*H*W) tensor -- apply the pixel norm after the activation and store its per-pixel factor there.
    out_h16: the next reader of `out` is a wide forward conv: let the pixel-norm pass write its fp16 planes too."""
    mp, mps = _mask(mask)
    wf, wt = w[0], w[1]
    if (FWD_FP16 and fwd and len(w) > 2 and w[2] is not None and x.P >= 2 and mask is None and not ups and scale == 1.0
            and _lib.load().pgk_conv_tc_supported(out.N, out.H, out.W, x.C, cout, ks, 0)):
        # fp16 two-plane copy of the input (one extra pass), then the conv on half operands
        comp = _h16_in(x)
        if comp is not None:
            xh_ptr, xh_ps = comp             # written by the kernel that produced x
        else:
            xh = torch.empty((2, x.N * x.per), dtype=torch.float16, device=x.t.device)
            call('pgk_cvt_fp16x2', x.ptr, x.ps, x.P, x.N * x.per, xh.data_ptr(), xh.stride(0))
            xh_ptr, xh_ps = xh.data_ptr(), xh.stride(0)
        call('pgk_conv_fp16', xh_ptr, xh_ps, out.N, out.H, out.W, x.C, cout, ks, w[2].data_ptr(),
             w[2].stride(0), None if bias is None else bias.data_ptr(), None if posT is None else posT.data_ptr(),
             None if pos_s is None else pos_s.data_ptr(), act, out.ptr, out.P, out.ps)
        # the half planes have one reader: hand them back to the (stream-ordered) caching allocator now rather than
        # when the tape dies
        x.aux.pop('h16', None)
        x.aux.pop('h16_range', None)
        if pn_r is not None:
            o16, o16_ps = _h16_out(out, out_h16)
            call('pgk_pixelnorm', out.ptr, out.ps, out.P, out.N * out.H * out.W, out.C, out.ptr, out.ps,
                 pn_r.data_ptr(), o16, o16_ps)
        return out
    if (len(w) > 3 and w[3] is not None and x.P == 1 and out.P == 1 and x.C == 64 and ks == 3 and not ups and posT is None
            and out.W % 128 == 0 and out.H % 8 == 0 and out.H >= 8 and (mask is None or mask.P == 1)):
        fuse_pn = pn_r is not None and cout <= 32 and mask is None and scale == 1.0
        call('pgk_conv_thin', x.ptr, 1, 1, x.ps, out.N, out.H, out.W, 64, cout, w[3].data_ptr(), w[3].stride(0),
             None if bias is None else bias.data_ptr(), act, mp, mps, scale, out.ptr, out.ps,
             pn_r.data_ptr() if fuse_pn else None)
        if pn_r is not None and not fuse_pn:
            call('pgk_pixelnorm', out.ptr, out.ps, out.P, out.N * out.H * out.W, out.C, out.ptr, out.ps, pn_r.data_ptr(),
                 None, 0)
        return out
    call('pgk_conv', x.ptr, x.P, x.P if fwd else min(x.P, GRAD_PLANES), x.ps, out.N, out.H, out.W, x.C, cout, ks, ups,
         None if wf is None else wf.data_ptr(), wt.data_ptr(),
         wt.stride(0), None if bias is None else bias.data_ptr(), None if posT is None else posT.data_ptr(),
         None if pos_s is None else pos_s.data_ptr(), act, mp, mps, scale, out.ptr, out.ps,
         None if pn_r is None else pn_r.data_ptr())
    return out


def wgrad(x, g, H, W, cin, cout, ks, ups, groups, group_n, dwp, db=None, bias_goffs=()):
    """dwp += sum over the (x offset, g offset) sample groups; db (optional, zeroed by the caller) += the bias gradient
    over the groups whose g offset is listed in bias_goffs (fused into the same launch on the thin-layer path)."""
    xoff, goff = _ints([a for a, _ in groups]), _ints([b for _, b in groups])
    mask = sum(1 << i for i, (_, go) in enumerate(groups) if go in bias_goffs)
    pg_ = min(x.P, g.P, GRAD_PLANES)
    call('pgk_wgrad', x.ptr, x.ps, g.ptr, g.ps, pg_, pg_, H, W, cin, cout, ks, ups, len(groups),
         group_n, xoff, goff, dwp.data_ptr(), None if db is None else db.data_ptr(), mask)


def bias_grad(g, hw, cout, goffs, group_n, db, scale=1.0, accumulate=0):
    call('pgk_bias_grad', g.ptr, g.ps, g.P, hw, cout, len(goffs), group_n, _ints(goffs), scale, db.data_ptr(),
         accumulate)


def from_rgb(img, mod, out, act=1, bias=True, mask=None, h16=False):
    n, c, h, w = img.shape
    mp, mps = _mask(mask)
    o16, o16_ps = _h16_out(out, h16)
    call('pgk_from_rgb', img.data_ptr(), n, c, h, w, out.C, mod.conv.weight.data_ptr(), mod.cf,
         mod.conv.bias.data_ptr() if bias else None, act, mp, mps, out.ptr, out.P, out.ps, o16, o16_ps)
    return out


def from_rgb_dgrad(g, mod, dimg, scale=1.0, ups=0, accumulate=0):
    n, c, h, w = dimg.shape
    call('pgk_from_rgb_dgrad', g.ptr, g.P, g.ps, n, c, h, w, g.C, mod.conv.weight.data_ptr(), mod.cf, scale, ups,
         accumulate, dimg.data_ptr())


def rgb_wgrad(img, img_n0, t, n, c, h, w, pool, scale_w, scale_b, dw, sa, sk, colsum, imgsum, dscale=None):
    call('pgk_rgb_wgrad', img.data_ptr(), img_n0, t.ptr, t.P, t.ps, 0, n, c, h, w, t.C, pool, scale_w, scale_b,
         None if dw is None else dw.data_ptr(), sa, sk, None if colsum is None else colsum.data_ptr(),
         None if imgsum is None else imgsum.data_ptr(), dscale)


def pool2(src, out, avg=1, a=1.0, other=None, b=0.0, h16=False):
    op, ops = _mask(other)
    (a, da), (b, db) = _sc(a), _sc(b)
    o16, o16_ps = _h16_out(out, h16)
    call('pgk_pool2', src.ptr, src.ps, src.P, out.N, out.H, out.W, out.C, avg, a, op, ops, b, out.ptr, out.ps, da, db,
         o16, o16_ps)
    return out


def mask_mul(src, out, ref=None, ups=0, scale=1.0, h16=False):
    rp, rps = _mask(ref)
    scale, dscale = _sc(scale)
    o16, o16_ps = _h16_out(out, h16)
    call('pgk_mask_mul', src.ptr, src.ps, src.P, out.N, out.H, out.W, out.C, ups, scale, rp, rps, out.ptr, out.ps,
         dscale, o16, o16_ps)
    return out


def pool_img(img, avg=1, scale=1.0):
    n, c, h, w = img.shape
    out = torch.empty((n, c, h // 2, w // 2), dtype=torch.float32, device=img.device)
    call('pgk_pool_img', img.data_ptr(), n, c, h // 2, w // 2, avg, scale, out.data_ptr())
    return out


# ---------------------------------------------------------------------------------------------
# prepared weights
# ---------------------------------------------------------------------------------------------
def _tc_channels(cin, cout, ks):
    """Do these channel counts run on the wide tensor-core kernel at every (power-of-two) resolution?  Then the fp32
    [K][Cout] operands, which only the CUDA-core kernels read, need not be produced."""
    if os.environ.get('PGK_TC', '1') == '0':
        return False
    return bool(_lib.load().pgk_conv_tc_supported(1, 8, 8, cin, cout, 1 if ks == 4 else ks, 0))


class ConvW(object):
    """Kernel-side operands of one equalised-LR conv (network.py:8-41): c folded in, re-laid for the GEMMs.  All stale
    layers of a network are refreshed together by ONE pgk_prep_multi launch (`prepare`)."""

    def __init__(self, mod, kind, cin, cout, ks, cin_stride=None, pos_hw=None, need_wb=True):
        self.mod, self.kind, self.cin, self.cout, self.ks = mod, kind, cin, cout, ks
        self.cin_stride = cin_stride if cin_stride is not None else cin
        self.pos_hw, self.need_wb = pos_hw, need_wb
        self.wf = self.wb = self.posT = self.bias16 = self.dwp = None
        self.F = self.B = None   # (fp32 [K][N] or None, bf16 planes [3][N][K]) operand pairs: forward / data gradient
        self.version = None
        self.cap_epoch = 0
        taps = ks * ks
        if kind == W_CONV:
            self.geom = (taps * cin, cout, taps * cout, cin)           # kf, nf, kb, nb
            gf, gb = (cin, cout, ks), (cout, cin, ks)
        elif kind == W_GFIRST:
            self.geom = (cin, 16 * cout, 16 * cout, cin)
            gf, gb = (cin, 16 * cout, 1), (16 * cout, cin, 1)
        else:
            self.geom = (16 * cin, cout, cout, 16 * cin)
            gf, gb = (16 * cin, cout, 1), (cout, 16 * cin, 1)
        # thin layers (csrc/pgk_conv_thin.cu) take their own packing; the data-gradient operand swaps the roles
        thin = lambda ci, co: kind == W_CONV and ks == 3 and ci in (8, 16, 32) and co in (8, 16, 32, 64)
        self.thin_f, self.thin_b = thin(cin, cout), thin(cout, cin)
        # 64-input-channel layers the thin kernel takes over in the one-plane mode (THIN64): a second, thin packing
        t64 = lambda ci, co: (THIN64 and os.environ.get('PGK_TC', '1') != '0' and kind == W_CONV and ks == 3 and ci == 64
                              and co in (32, 64))
        self.t64_f, self.t64_b = t64(cin, cout), need_wb and t64(cout, cin)
        # the fp32 operands are read by the CUDA-core kernels only: shapes that always take a tensor-core kernel skip
        # them (the thin packing of the 64-channel layers is made from them)
        self.need_wf_f = self.t64_f or not _tc_channels(*gf)
        self.need_wf_b = need_wb and (self.t64_b or not _tc_channels(*gb))

    @property
    def weight(self):
        return self.mod.conv.weight

    @property
    def bias(self):
        return self.mod.conv.bias

    def stale(self, planes):
        w = self.weight
        ver = (w._version, w.data_ptr(), self.bias._version, self.mod.cf, planes)
        return ver != self.version or (CAPTURE_EPOCH != 0 and self.cap_epoch != CAPTURE_EPOCH)

    def _allocate(self, dev):
        n = self.cout * self.cin * self.ks * self.ks
        kf, nf_, kb, nb = self.geom
        lib = _lib.load()
        self.wf = torch.empty(n, dtype=torch.float32, device=dev) if self.need_wf_f else None
        self.wb = torch.empty(n, dtype=torch.float32, device=dev) if self.need_wf_b else None
        if self.pos_hw is not None:
            self.posT = torch.empty(self.pos_hw[0] * self.pos_hw[1] * self.cout, dtype=torch.float32, device=dev)
        # tensor-core operands: [plane][output channel][K], K-major (the thin packing has padding entries that are
        # never written: zeroed once here)
        if self.thin_f:
            ft = torch.zeros((3, lib.pgk_pack_thin_plane_elems(self.cin, self.cout)), dtype=BF16, device=dev)
        else:
            ft = torch.empty((3, nf_, kf), dtype=BF16, device=dev)
        self.F = (self.wf, ft)
        if FWD_FP16 and not self.thin_f:
            # third element of the forward operand tuple: [2 planes][output channel][K] IEEE half, times 2^PGK_FP16_WSHIFT
            self.F = (self.wf, ft, torch.empty((2, nf_, kf), dtype=torch.float16, device=dev))
        if self.t64_f:
            f64 = torch.zeros((1, lib.pgk_pack_thin_plane_elems(self.cin, self.cout)), dtype=BF16, device=dev)
            self.F = (self.F + (None,))[:3] + (f64,)
        self.B = None
        if self.need_wb:
            if self.thin_b:
                bt = torch.zeros((3, lib.pgk_pack_thin_plane_elems(self.cout, self.cin)), dtype=BF16, device=dev)
            else:
                bt = torch.empty((3, nb, kb), dtype=BF16, device=dev)
            self.B = (self.wb, bt)
            if self.t64_b:
                b64 = torch.zeros((1, lib.pgk_pack_thin_plane_elems(self.cout, self.cin)), dtype=BF16, device=dev)
                self.B = (self.wb, bt, None, b64)

    def fill(self, d, planes):
        """Fill one PgkPrepLayer (include/pgk.h) for this layer; allocates the operand buffers on first use."""
        w = self.weight
        if self.F is None or self.F[1].device != w.device:
            self._allocate(w.device)
        d.w, d.c, d.kind = w.data_ptr(), self.mod.cf, self.kind
        d.cin, d.cin_stride, d.cout, d.ks, d.planes = self.cin, self.cin_stride, self.cout, self.ks, planes
        d.wf = None if self.wf is None else self.wf.data_ptr()
        d.wb = None if self.wb is None else self.wb.data_ptr()
        d.F, d.F_ps, d.thinF = self.F[1].data_ptr(), self.F[1].stride(0), int(self.thin_f)
        if self.B is not None:
            d.B, d.B_ps, d.thinB = self.B[1].data_ptr(), self.B[1].stride(0), int(self.thin_b)
        else:
            d.B, d.B_ps, d.thinB = None, 0, 0
        if len(self.F) > 2 and self.F[2] is not None and planes >= 2:   # (the half packing: fp32-faithful mode only)
            d.F16, d.F16_ps = self.F[2].data_ptr(), self.F[2].stride(0)
        else:
            d.F16, d.F16_ps = None, 0

    def finish(self, planes):
        """The per-layer extras after the batched launch, and the cache key."""
        w = self.weight
        if self.pos_hw is not None:
            call('pgk_prep_posbias', w.data_ptr(), self.mod.cf, self.cin_stride, self.cin, self.cout, self.pos_hw[0],
                 self.pos_hw[1], self.posT.data_ptr())
        if self.kind == W_GFIRST:
            self.bias16 = self.bias.detach().repeat(16)
        if planes == 1:      # (the thin packing of the 64-channel layers is read by the one-plane mode only)
            if self.t64_f:
                call('pgk_pack_thin', self.wf.data_ptr(), self.cin, self.cout, self.F[3].data_ptr(), self.F[3].stride(0), 1)
            if self.t64_b:
                call('pgk_pack_thin', self.wb.data_ptr(), self.cout, self.cin, self.B[3].data_ptr(), self.B[3].stride(0), 1)
        self.version = (w._version, w.data_ptr(), self.bias._version, self.mod.cf, planes)
        self.cap_epoch = CAPTURE_EPOCH

    def ensure(self, planes=3):
        if self.stale(planes):
            prepare([self], planes)
        return self

    def scratch(self):
        dev = self.weight.device
        if self.dwp is None or self.dwp.device != dev:
            self.dwp = torch.empty(self.cout * self.cin * self.ks * self.ks, dtype=torch.float32, device=dev)
        return self.dwp

    def wgrad_into(self, grad, x, g, H, W, ups, groups, group_n, db=None, bias_goffs=(), pending=None):
        """grad (PyTorch layout, fp32) <- c * sum_groups x (*) g;  db (zero-initialised) += bias gradient.
        pending (a list): the scratch was zeroed by the caller (zero_scratch) and the map back to the PyTorch layout is
        left to one pgk_unprep_multi launch over all layers (unprep_all)."""
        dwp = self.scratch()
        if pending is None:
            dwp.zero_()
        if self.kind == W_CONV:
            wgrad(x, g, H, W, self.cin, self.cout, self.ks, ups, groups, group_n, dwp, db, bias_goffs)
        elif self.kind == W_GFIRST:   # x: (n,1,1,cin), g: (n,1,1,16*cout)
            wgrad(x, g, 1, 1, self.cin, 16 * self.cout, 1, 0, groups, group_n, dwp)
        else:                          # x: (n,1,1,16*cin), g: (n,1,1,cout)
            wgrad(x, g, 1, 1, 16 * self.cin, self.cout, 1, 0, groups, group_n, dwp)
        if pending is not None:
            pending.append((self, grad))
            return
        call('pgk_unprep_grad', dwp.data_ptr(), self.mod.cf, self.kind, self.cin, self.cin_stride, self.cout, self.ks,
             grad.data_ptr(), 0)


def zero_scratch(convws):
    """Zero the weight-gradient scratch of all these layers with one multi-tensor launch."""
    torch._foreach_zero_([w.scratch() for w in convws])


def unprep_all(pending):
    """[(ConvW, grad)] -> one pgk_unprep_multi launch: grad (PyTorch layout) <- c * scratch ([K][Cout] layout)."""
    if not pending:
        return
    table = (_lib.UnprepLayer * len(pending))()
    for d, (w, grad) in zip(table, pending):
        d.dwp, d.c, d.kind = w.dwp.data_ptr(), w.mod.cf, w.kind
        d.cin, d.cin_stride, d.cout, d.ks = w.cin, w.cin_stride, w.cout, w.ks
        d.dw, d.accumulate = grad.data_ptr(), 0
    call('pgk_unprep_multi', ctypes.cast(table, ctypes.c_void_p), len(pending))


def prepare(convws, planes):
    """Refresh the operands of every stale layer of `convws` with one pgk_prep_multi launch (csrc/pgk_prep.cu)."""
    todo = [w for w in convws if w.stale(planes)]
    if not todo:
        return
    table = (_lib.PrepLayer * len(todo))()
    for d, w in zip(table, todo):
        w.fill(d, planes)
    call('pgk_prep_multi', ctypes.cast(table, ctypes.c_void_p), len(todo))
    for w in todo:
        w.finish(planes)


class GradSet(object):
    """One flat fp32 buffer holding the gradients of the parameters that take part in a step (the active set
    depends only on (depth, alpha < 1)); `views[param]` are the per-parameter views handed to autograd.  A single
    flat buffer = a single NCCL all-reduce per optimizer step under data parallelism."""

    def __init__(self, params, device):
        self.params = list(params)
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        self.views = {}
        o = 0
        for p in self.params:
            self.views[id(p)] = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()

    def __getitem__(self, p):
        return self.views[id(p)]

    def grads(self):
        return [self.views[id(p)] for p in self.params]


# ---------------------------------------------------------------------------------------------
# Discriminator
# ---------------------------------------------------------------------------------------------
class DEngine(object):
    """Forward / backward chains of Discriminator.forward (network.py:225-240) at the current (depth, alpha)."""

    def __init__(self, D):
        self.D = D
        self._cw = {}
        self._P = 3
        self.fade_dev = None    # two floats on the device (alpha, 1 - alpha) in CUDA-graph mode, see Fade

    def blk(self, k):
        """blocks[-k] (network.py:227,231,236)."""
        return self.D.blocks[len(self.D.blocks) - k]

    def _get(self, mod, kind=W_CONV):
        key = id(mod)
        if key not in self._cw:
            w = mod.conv.weight
            cout, cin_stride, ks = w.shape[0], w.shape[1], w.shape[2]
            if kind == W_DLAST:
                self._cw[key] = ConvW(mod, W_DLAST, cin_stride, cout, 4)
            elif cin_stride % 8 == 1:   # the stddev layer: 512 real channels + 1 constant channel
                self._cw[key] = ConvW(mod, W_CONV, cin_stride - 1, cout, ks, cin_stride=cin_stride, pos_hw=(4, 4))
            else:
                self._cw[key] = ConvW(mod, W_CONV, cin_stride, cout, ks)
        return self._cw[key]

    def cw(self, mod, kind=W_CONV):
        return self._get(mod, kind).ensure(self._P)

    def prepare(self, depth, P):
        """Refresh the operands of every conv layer active at `depth` in one launch (a no-op while they are current)."""
        self._P = P
        prepare(self.active_convs(depth), P)

    def active_convs(self, depth):
        ws = []
        for k in range(depth + 1, 1, -1):
            b = self.blk(k)
            ws += [self._get(b.c1), self._get(b.c2)]
        last = self.blk(1)
        return ws + [self._get(last.c1), self._get(last.c2, W_DLAST)]

    # Parameters (= order of the flat gradient buffer) and weight gradients go LOW RESOLUTION FIRST: the 4x4 ... 32x32
    # blocks hold ~90 % of the gradient bytes (512-channel layers) and their weight gradients are short, the 64x64-and-up
    # blocks hold few bytes and their weight gradients are the long ones.  Under data parallelism the all-reduce of the
    # first bucket therefore runs underneath the weight gradients of the second (wgan_gp_loss._d_body).
    BUCKET_LEVELS = 4       # blocks[-1] .. blocks[-4] (4x4 .. 32x32) + linear form the first bucket

    def active_params(self, depth, fade):
        ps = [self.D.linear.weight, self.D.linear.bias]
        for k in range(1, depth + 2):
            b = self.blk(k)
            ps += [b.c1.conv.weight, b.c1.conv.bias, b.c2.conv.weight, b.c2.conv.bias]
        top = self.blk(depth + 1)
        ps += [top.fromRGB.conv.weight, top.fromRGB.conv.bias]
        if fade:
            lo = self.blk(depth)
            ps += [lo.fromRGB.conv.weight, lo.fromRGB.conv.bias]
        return ps

    def first_bucket_elems(self, depth):
        """elements of the flat gradient buffer that belong to the first bucket (0 = no split worth making)"""
        if depth + 1 <= self.BUCKET_LEVELS:
            return 0
        n = self.D.linear.weight.numel() + self.D.linear.bias.numel()
        for k in range(1, self.BUCKET_LEVELS + 1):
            b = self.blk(k)
            n += sum(p.numel() for p in (b.c1.conv.weight, b.c1.conv.bias, b.c2.conv.weight, b.c2.conv.bias))
        return n

    # -- forward ---------------------------------------------------------------------------
    def forward(self, ximg, ngroups, group_n, P, slots=0):
        """ximg: fp32 (B,C,r,r), B = ngroups*group_n, r = 4*2^depth.  Activations are allocated with B+slots
        samples so that the v-chain can live next to them (D step).  Returns the tape."""
        D = self.D
        depth, alpha = int(D.depth), float(D.alpha)
        fade = depth > 0 and alpha < 1.0
        B = ximg.shape[0]
        Bt = B + slots
        dev = ximg.device
        r = ximg.shape[-1]
        assert B == ngroups * group_n and r == 4 * 2 ** depth, 'input resolution must match the current depth'
        T = SimpleNamespace(depth=depth, alpha=alpha, fade=fade, B=B, Bt=Bt, P=P, ngroups=ngroups, group_n=group_n,
                            ximg=ximg, blocks=[], fd=Fade(alpha, self.fade_dev if fade else None))
        self.prepare(depth, P)
        new = lambda res, c: PT.empty(Bt, res, res, c, P, dev)
        top = self.blk(depth + 1)
        ctop = top.fromRGB.conv.weight.shape[0]
        T.t0 = new(r, ctop)
        from_rgb(ximg, top.fromRGB, T.t0.sl(0, B), h16=True)
        if depth == 0:
            hin = T.t0
        else:
            w1, w2 = self.cw(top.c1), self.cw(top.c2)
            T.t1 = new(r, w1.cout)
            conv(T.t0.sl(0, B), w1.F, w1.cout, 3, T.t1.sl(0, B), bias=w1.bias, act=1, fwd=True)
            T.t2 = new(r, w2.cout)
            conv(T.t1.sl(0, B), w2.F, w2.cout, 3, T.t2.sl(0, B), bias=w2.bias, act=1, fwd=True)
            h = new(r // 2, w2.cout)
            if fade:
                T.xlow = pool_img(ximg)
                T.f = new(r // 2, w2.cout)
                from_rgb(T.xlow, self.blk(depth).fromRGB, T.f.sl(0, B))
                pool2(T.t2.sl(0, B), h.sl(0, B), avg=1, a=T.fd.a(), other=T.f.sl(0, B), b=T.fd.b(), h16=True)
            else:
                pool2(T.t2.sl(0, B), h.sl(0, B), avg=1, h16=True)
            res = r // 2
            for k in range(depth, 1, -1):
                b = self.blk(k)
                w1, w2 = self.cw(b.c1), self.cw(b.c2)
                a_ = new(res, w1.cout)
                conv(h.sl(0, B), w1.F, w1.cout, 3, a_.sl(0, B), bias=w1.bias, act=1, fwd=True)
                b_ = new(res, w2.cout)
                conv(a_.sl(0, B), w2.F, w2.cout, 3, b_.sl(0, B), bias=w2.bias, act=1, fwd=True)
                hn = new(res // 2, w2.cout)
                pool2(b_.sl(0, B), hn.sl(0, B), avg=1, h16=True)
                T.blocks.append(SimpleNamespace(mod=b, hin=h, a=a_, b=b_, res=res))
                h, res = hn, res // 2
            hin = h
        T.hin = hin
        last = self.blk(1)
        wl1, wl2 = self.cw(last.c1), self.cw(last.c2, W_DLAST)
        C = hin.C
        T.stats = torch.empty(ngroups * 4, dtype=torch.float32, device=dev)
        T.svec = torch.empty(B, dtype=torch.float32, device=dev)
        call('pgk_stddev_stats', hin.ptr, hin.ps, P, ngroups, group_n * 16 * C, T.stats.data_ptr(), T.svec.data_ptr(),
             group_n)
        T.l1 = new(4, wl1.cout)
        conv(hin.sl(0, B), wl1.F, wl1.cout, 3, T.l1.sl(0, B), bias=wl1.bias, posT=wl1.posT, pos_s=T.svec, act=1, fwd=True)
        T.l2 = PT.empty(Bt, 1, 1, wl2.cout, P, dev)
        conv(T.l1.sl(0, B).view(1, 1, 16 * wl1.cout), wl2.F, wl2.cout, 1, T.l2.sl(0, B), bias=wl2.bias, act=1, fwd=True)
        T.scores = torch.empty(B, dtype=torch.float32, device=dev)
        call('pgk_linear_fwd', T.l2.ptr, T.l2.ps, P, B, wl2.cout, D.linear.weight.data_ptr(), D.linear.bias.data_ptr(),
             T.scores.data_ptr())
        return T

    # -- u-chain: head -> hin ---------------------------------------------------------------
    def backward_head(self, T, seed, wseed, gs):
        """Fills T.ua_l2, T.ua_l1, T.q and T.d_hin (gradient w.r.t. the last block's input, stddev term included)
        for all B samples.  gs: GradSet receiving linear.weight/bias grads (None = no parameter grads)."""
        D, B, P = self.D, T.B, min(T.P, GRAD_PLANES)   # gradient tensors carry GRAD_PLANES planes
        dev = T.scores.device
        last = self.blk(1)
        wl1, wl2 = self.cw(last.c1), self.cw(last.c2, W_DLAST)
        T.ua_l2 = PT.empty(T.Bt, 1, 1, wl2.cout, P, dev)
        call('pgk_linear_bwd', T.l2.ptr, T.l2.ps, P, B, wl2.cout, D.linear.weight.data_ptr(), seed.data_ptr(),
             None if wseed is None else wseed.data_ptr(), T.ua_l2.ptr, T.ua_l2.ps,
             None if gs is None else gs[D.linear.weight].data_ptr(),
             None if gs is None else gs[D.linear.bias].data_ptr())
        T.ua_l1 = PT.empty(T.Bt, 4, 4, wl1.cout, P, dev)
        conv(T.ua_l2.sl(0, B), wl2.B, 16 * wl1.cout, 1, T.ua_l1.sl(0, B).view(1, 1, 16 * wl1.cout),
             mask=T.l1.sl(0, B).view(1, 1, 16 * wl1.cout))
        T.q = torch.empty(T.ngroups, dtype=torch.float32, device=dev)
        call('pgk_group_dot_pos', T.ua_l1.ptr, T.ua_l1.ps, P, T.ngroups, T.group_n, 16, wl1.cout, wl1.posT.data_ptr(),
             T.q.data_ptr())
        C = T.hin.C
        T.d_hin = PT.empty(B, 4, 4, C, P, dev)
        conv(T.ua_l1.sl(0, B), wl1.B, C, 3, T.d_hin)
        call('pgk_stddev_bwd', T.hin.ptr, T.hin.ps, T.stats.data_ptr(), T.q.data_ptr(), P, T.ngroups,
             T.group_n * 16 * C, T.d_hin.ptr, T.d_hin.ps)

    # -- the part of the backward chain below the last block (shared by the u- and w-chains) -----
    def backward_body(self, T, d_hin, n0, n1, g0):
        """d_hin: gradient at the last block's input for tape samples [n0, n1).  Writes the pre-activation gradients
        into the `ua` tensors of the tape at sample offset g0 (allocating them with Bt slots on first use)."""
        P, dev, n = min(T.P, GRAD_PLANES), d_hin.t.device, n1 - n0
        depth, alpha, fade = T.depth, T.alpha, T.fade
        top = self.blk(depth + 1)

        def ua_of(name, like):
            if not hasattr(T, name):
                setattr(T, name, PT.empty(T.Bt, like.H, like.W, like.C, P, dev))
            return getattr(T, name).sl(g0, g0 + n)

        if depth == 0:
            ua_t0 = mask_mul(d_hin, ua_of('ua_t0', T.t0), ref=T.t0.sl(n0, n1))
        else:
            d_h = d_hin
            for i, rec in enumerate(reversed(T.blocks)):
                w1, w2 = self.cw(rec.mod.c1), self.cw(rec.mod.c2)
                name = 'ua_blk%d' % (len(T.blocks) - 1 - i)
                ua_b = mask_mul(d_h, ua_of(name + 'b', rec.b), ref=rec.b.sl(n0, n1), ups=1, scale=0.25)
                ua_a = conv(ua_b, w2.B, w1.cout, 3, ua_of(name + 'a', rec.a), mask=rec.a.sl(n0, n1))
                d_h = conv(ua_a, w1.B, w1.cin, 3, PT.empty(n, rec.res, rec.res, w1.cin, P, dev))
            w1, w2 = self.cw(top.c1), self.cw(top.c2)
            ua_t2 = mask_mul(d_h, ua_of('ua_t2', T.t2), ref=T.t2.sl(n0, n1), ups=1,
                             scale=T.fd.a(0.25) if fade else 0.25)
            ua_t1 = conv(ua_t2, w2.B, w1.cout, 3, ua_of('ua_t1', T.t1), mask=T.t1.sl(n0, n1))
            ua_t0 = conv(ua_t1, w1.B, w1.cin, 3, ua_of('ua_t0', T.t0), mask=T.t0.sl(n0, n1))
            if fade:
                ua_f = mask_mul(d_h, ua_of('ua_f', T.f), ref=T.f.sl(n0, n1), scale=T.fd.b())

    def image_grad(self, T, g0, n, dimg):
        """dimg (fp32 n,C,r,r) <- gradient w.r.t. the input image from the ua tensors at sample offset g0
        (fromRGB backwards, plus the avg-pooled low-res branch during a fade, network.py:230-233)."""
        top = self.blk(T.depth + 1)
        from_rgb_dgrad(T.ua_t0.sl(g0, g0 + n), top.fromRGB, dimg)
        if T.fade:
            from_rgb_dgrad(T.ua_f.sl(g0, g0 + n), self.blk(T.depth).fromRGB, dimg, scale=0.25, ups=1, accumulate=1)

    # -- v-chain ------------------------------------------------------------------------------
    def v_chain(self, T, v0, m0, vs):
        """D's forward applied to v0 (fp32 n,C,r,r) with biases dropped and LeakyReLU replaced by the masks of the
        tape samples [m0, m0+n) (the mixed samples).  The v tensors are written into slot `vs` of the tape's
        activation buffers.  Returns the extra-channel value ev and w_h (the w-chain seed)."""
        P, dev = min(T.P, GRAD_PLANES), v0.device
        n = v0.shape[0]
        depth, alpha, fade = T.depth, T.alpha, T.fade
        top = self.blk(depth + 1)
        m = lambda t: t.sl(m0, m0 + n)
        v = lambda t: t.sl(vs, vs + n).g()     # the v tensors live in the spare sample slots, GRAD_PLANES planes
        from_rgb(v0, top.fromRGB, v(T.t0), act=0, bias=False, mask=m(T.t0))
        if depth == 0:
            vh = v(T.t0)
        else:
            w1, w2 = self.cw(top.c1), self.cw(top.c2)
            conv(v(T.t0), w1.F, w1.cout, 3, v(T.t1), mask=m(T.t1))
            conv(v(T.t1), w2.F, w2.cout, 3, v(T.t2), mask=m(T.t2))
            res = T.t2.H // 2
            if fade:
                T.v0low = pool_img(v0)
                from_rgb(T.v0low, self.blk(depth).fromRGB, v(T.f), act=0, bias=False, mask=m(T.f))
                dst = T.blocks[0].hin if T.blocks else T.hin
                pool2(v(T.t2), v(dst), avg=1, a=T.fd.a(), other=v(T.f), b=T.fd.b())
            else:
                dst = T.blocks[0].hin if T.blocks else T.hin
                pool2(v(T.t2), v(dst), avg=1)
            for i, rec in enumerate(T.blocks):
                w1, w2 = self.cw(rec.mod.c1), self.cw(rec.mod.c2)
                conv(v(rec.hin), w1.F, w1.cout, 3, v(rec.a), mask=m(rec.a))
                conv(v(rec.a), w2.F, w2.cout, 3, v(rec.b), mask=m(rec.b))
                dst = T.blocks[i + 1].hin if i + 1 < len(T.blocks) else T.hin
                pool2(v(rec.b), v(dst), avg=1)
            vh = v(T.hin)
        last = self.blk(1)
        wl1, wl2 = self.cw(last.c1), self.cw(last.c2, W_DLAST)
        C = T.hin.C
        g = m0 // T.group_n
        T.ev = torch.empty(n, dtype=torch.float32, device=dev)
        T.w_h = PT.empty(n, 4, 4, C, P, dev)
        scratch = torch.empty(2, dtype=torch.float32, device=dev)
        hm = m(T.hin)
        call('pgk_stddev_bwd2', hm.ptr, hm.ps, vh.ptr, vh.ps, T.stats[4 * g:].data_ptr(), T.q[g:].data_ptr(), P,
             n * 16 * C, T.ev.data_ptr(), n, T.w_h.ptr, T.w_h.ps, scratch.data_ptr())
        conv(vh, wl1.F, wl1.cout, 3, v(T.l1), posT=wl1.posT, pos_s=T.ev, mask=m(T.l1))
        conv(v(T.l1).view(1, 1, 16 * wl1.cout), wl2.F, wl2.cout, 1, v(T.l2), mask=m(T.l2))

    # -- parameter gradients -------------------------------------------------------------------
    def param_grads(self, T, gs, groups, bias_goffs, head_groups, head_bias_goffs, img_pairs, ev_pair=None,
                    first_bucket_done=None):
        """groups: (x sample offset, ua sample offset) pairs of group_n samples each for the layers below the last
        block; head_groups: the same for the last block's c1 / c2.  img_pairs: (image tensor, img_n0, ua offset)
        triples for fromRGB.  ev_pair = (coef vector (n), ua offset) adds the v-chain's extra-channel term.
        Order: low resolution first (see active_params); first_bucket_done() is called once the gradients of the first
        bucket are final in gs (the caller starts their all-reduce there)."""
        n, P = T.group_n, min(T.P, GRAD_PLANES)
        depth, alpha, fade = T.depth, T.alpha, T.fade
        top = self.blk(depth + 1)
        C = T.ximg.shape[1]
        r = T.ximg.shape[-1]
        pending = []
        zero_scratch(self.active_convs(depth))

        def conv_layer(mod, x, ua, res):
            w = self.cw(mod)
            w.wgrad_into(gs[mod.conv.weight], x, ua, res, res, 0, groups, n, db=gs[mod.conv.bias],
                         bias_goffs=bias_goffs, pending=pending)

        # ---- the last block (4x4) and the linear head (its gradients were written by backward_head / pgk_colsum)
        last = self.blk(1)
        wl1, wl2 = self.cw(last.c1), self.cw(last.c2, W_DLAST)
        g1 = gs[last.c1.conv.weight]
        wl1.wgrad_into(g1, T.hin, T.ua_l1, 4, 4, 0, head_groups, n, pending=pending)
        # the constant (stddev) input channel of c1: coefficient s_g for the ordinary terms, e for the v-chain
        for xo, go in head_groups:
            if ev_pair is not None and go == ev_pair[1] and xo == ev_pair[2]:
                coef = ev_pair[0]
            else:
                coef = T.svec[xo:xo + n]
            u = T.ua_l1.sl(go, go + n)
            call('pgk_posbias_wgrad', u.ptr, u.ps, P, n, 4, 4, wl1.cout, coef.data_ptr(), last.c1.cf, wl1.cin_stride,
                 wl1.cin, g1.data_ptr())
        bias_grad(T.ua_l1, 16, wl1.cout, head_bias_goffs, n, gs[last.c1.conv.bias])
        wl2.wgrad_into(gs[last.c2.conv.weight], T.l1.view(1, 1, 16 * wl1.cout), T.ua_l2, 1, 1, 0, head_groups, n,
                       pending=pending)
        bias_grad(T.ua_l2, 1, wl2.cout, head_bias_goffs, n, gs[last.c2.conv.bias])
        # ---- blocks[-2] (8x8) upwards; T.blocks runs from the block below the top one down to blocks[-2]
        level = 1

        def close_bucket():
            unprep_all(pending)
            del pending[:]
            if first_bucket_done is not None:
                first_bucket_done()

        for i in range(len(T.blocks) - 1, -1, -1):
            if level == self.BUCKET_LEVELS:
                close_bucket()
            rec = T.blocks[i]
            conv_layer(rec.mod.c1, rec.hin, getattr(T, 'ua_blk%da' % i), rec.res)
            conv_layer(rec.mod.c2, rec.a, getattr(T, 'ua_blk%db' % i), rec.res)
            level += 1
        if depth > 0:
            if level == self.BUCKET_LEVELS:
                close_bucket()
            conv_layer(top.c1, T.t0, T.ua_t1, r)
            conv_layer(top.c2, T.t1, T.ua_t2, r)

        def rgb(mod, t_ua, res, pairs_):
            gw, gb = gs[mod.conv.weight], gs[mod.conv.bias]
            for img, img_n0, goff, with_bias in pairs_:
                rgb_wgrad(img, img_n0, t_ua.sl(goff, goff + n), n, C, res, res, 0, mod.cf, 1.0, gw, 1, C,
                          gb if with_bias else None, None)

        rgb(top.fromRGB, T.ua_t0, r, img_pairs['top'])
        if fade:
            rgb(self.blk(depth).fromRGB, T.ua_f, r // 2, img_pairs['low'])
        unprep_all(pending)


# ---------------------------------------------------------------------------------------------
# Generator
# ---------------------------------------------------------------------------------------------
class GEngine(object):
    """Generator.forward (network.py:118-139) and its backward."""

    def __init__(self, G):
        self.G = G
        self._cw = {}
        self._P = 3
        self.fade_dev = None

    def block(self, i):
        return self.G.block0 if i == 0 else self.G.blocks[i - 1]

    def _get(self, mod, kind=W_CONV):
        key = id(mod)
        if key not in self._cw:
            w = mod.conv.weight
            if kind == W_GFIRST:
                self._cw[key] = ConvW(mod, W_GFIRST, w.shape[1], w.shape[0], 4, need_wb=False)
            else:
                self._cw[key] = ConvW(mod, W_CONV, w.shape[1], w.shape[0], w.shape[2])
        return self._cw[key]

    def cw(self, mod, kind=W_CONV):
        return self._get(mod, kind).ensure(self._P)

    def prepare(self, depth, P):
        """Refresh the operands of every conv layer active at `depth` in one launch (a no-op while they are current)."""
        self._P = P
        prepare(self.active_convs(depth), P)

    def active_convs(self, depth):
        ws = [self._get(self.G.block0.c1, W_GFIRST), self._get(self.G.block0.c2)]
        for i in range(1, depth + 1):
            b = self.block(i)
            ws += [self._get(b.c1), self._get(b.c2)]
        return ws

    def active_params(self, depth, fade):
        ps = []
        for i in range(depth + 1):
            b = self.block(i)
            ps += [b.c1.conv.weight, b.c1.conv.bias, b.c2.conv.weight, b.c2.conv.bias]
        ps += [self.block(depth).toRGB.conv.weight, self.block(depth).toRGB.conv.bias]
        if fade:
            ps += [self.block(depth - 1).toRGB.conv.weight, self.block(depth - 1).toRGB.conv.bias]
        return ps

    def _pn(self, h, T, name):
        """The per-pixel factor buffer of the pixel norm that pgk_conv applies after the activation
        (network.py:37-40); kept on the tape for the backward pass.  None when the generator has no pixel norm."""
        if not self.G.pixelnorm:
            return None
        r = torch.empty(h.N * h.H * h.W, dtype=torch.float32, device=h.t.device)
        if T is not None:
            setattr(T, 'r_' + name, r)
        return r

    def forward(self, z, P, out=None, tape=False):
        """z: fp32 (N, latent).  Returns (image fp32 N,C,r,r, tape or None); `out` lets the caller place the image
        (e.g. straight into the fake slot of D's input batch)."""
        G = self.G
        depth, alpha = int(G.depth), float(G.alpha)
        fade = depth > 0 and alpha < 1.0
        n, dev = z.shape[0], z.device
        fd = Fade(alpha, self.fade_dev if fade else None)
        T = SimpleNamespace(depth=depth, alpha=alpha, fade=fade, n=n, P=P, acts=[], fd=fd) if tape else None
        self.prepare(depth, P)
        zn = PT.empty(n, 1, 1, z.shape[1], P, dev)
        call('pgk_latent_norm', z.data_ptr(), n, z.shape[1], 1 if G.normalize_latents else 0, zn.ptr, P, zn.ps)
        b0 = G.block0
        w1, w2 = self.cw(b0.c1, W_GFIRST), self.cw(b0.c2)
        h1 = PT.empty(n, 4, 4, w1.cout, P, dev)
        conv(zn, w1.F, 16 * w1.cout, 1, h1.view(1, 1, 16 * w1.cout), bias=w1.bias16, act=1, fwd=True)
        r = self._pn(h1, T, 'b0c1')   # the dense first layer: its 16 output pixels sit side by side in one GEMM row
        if r is not None:
            call('pgk_pixelnorm', h1.ptr, h1.ps, h1.P, h1.N * h1.H * h1.W, h1.C, h1.ptr, h1.ps, r.data_ptr(), None, 0)
        h2 = PT.empty(n, 4, 4, w2.cout, P, dev)
        conv(h1, w2.F, w2.cout, 3, h2, bias=w2.bias, act=1, fwd=True, pn_r=self._pn(h2, T, 'b0c2'))
        if tape:
            T.zn = zn
            T.acts.append((None, None, h1, h2))
        h, res = h2, 4
        hprev = None
        for i in range(1, depth + 1):
            b = self.block(i)
            w1, w2 = self.cw(b.c1), self.cw(b.c2)
            res *= 2
            # nearest-neighbour 2x upsample (network.py:127,129), materialised once: the tensor-core conv and the
            # weight gradient both read it through plain TMA boxes
            hu = mask_mul(h, PT.empty(n, res, res, h.C, P, dev), ups=1, h16=True)
            u1 = PT.empty(n, res, res, w1.cout, P, dev)
            conv(hu, w1.F, w1.cout, 3, u1, bias=w1.bias, act=1, fwd=True, pn_r=self._pn(u1, T, 'b%dc1' % i), out_h16=True)
            u2 = PT.empty(n, res, res, w2.cout, P, dev)
            conv(u1, w2.F, w2.cout, 3, u2, bias=w2.bias, act=1, fwd=True, pn_r=self._pn(u2, T, 'b%dc2' % i))
            if tape:
                T.acts.append((h, hu, u1, u2))
            hprev, h = h, u2
        C = self.block(depth).toRGB.conv.weight.shape[0]
        img = out if out is not None else torch.empty((n, C, res, res), dtype=torch.float32, device=dev)
        hi = self.block(depth).toRGB
        if fade:
            lo = self.block(depth - 1).toRGB
            (a_hi, d_hi), (a_lo, d_lo) = fd.a(), fd.b()
            call('pgk_to_rgb', h.ptr, P, h.ps, n, res, res, h.C, hi.conv.weight.data_ptr(), hi.cf,
                 hi.conv.bias.data_ptr(), a_hi, hprev.ptr, hprev.ps, hprev.C, lo.conv.weight.data_ptr(), lo.cf,
                 lo.conv.bias.data_ptr(), a_lo, C, img.data_ptr(), d_hi, d_lo)
        else:
            # depth 0 returns toRGB(h); depth > 0 with alpha >= 1 returns 0*(1-alpha) + ult*alpha (network.py:136-138)
            a_hi = 1.0 if depth == 0 else alpha
            call('pgk_to_rgb', h.ptr, P, h.ps, n, res, res, h.C, hi.conv.weight.data_ptr(), hi.cf,
                 hi.conv.bias.data_ptr(), a_hi, None, 0, 0, None, 0.0, None, 0.0, C, img.data_ptr(), None, None)
        return img, T

    def backward(self, T, dimg, gs):
        """dimg: fp32 gradient w.r.t. the generated image; fills gs with every active parameter's gradient."""
        G = self.G
        depth, alpha, fade, n, P = T.depth, T.alpha, T.fade, T.n, min(T.P, GRAD_PLANES)
        dev = dimg.device
        C = dimg.shape[1]
        res = dimg.shape[-1]
        groups = [(0, 0)]
        pending = []
        zero_scratch(self.active_convs(depth))
        hi = self.block(depth).toRGB
        hprev, _, u1, u2 = T.acts[depth]
        # toRGB (network.py:49,65) and the fade-in lerp (network.py:138): factor alpha on the new level's branch (1 at
        # depth 0), 1 - alpha on the previous level's
        k_hi, d_hi = T.fd.a() if fade else ((1.0 if depth == 0 else alpha), None)
        rgb_wgrad(dimg, 0, u2, n, C, res, res, 0, hi.cf * k_hi, k_hi, gs[hi.conv.weight], u2.C, 1, None,
                  gs[hi.conv.bias], dscale=d_hi)
        d = PT.empty(n, res, res, u2.C, P, dev)
        call('pgk_to_rgb_dgrad', dimg.data_ptr(), n, C, res, res, u2.C, hi.conv.weight.data_ptr(), hi.cf, k_hi, 0,
             d.ptr, P, d.ps, d_hi)
        d_lo = None
        if fade:
            lo = self.block(depth - 1).toRGB
            k_lo, dp_lo = T.fd.b()
            rgb_wgrad(dimg, 0, hprev, n, C, res // 2, res // 2, 1, lo.cf * k_lo, k_lo,
                      gs[lo.conv.weight], hprev.C, 1, None, gs[lo.conv.bias], dscale=dp_lo)
            d_lo = PT.empty(n, res // 2, res // 2, hprev.C, P, dev)
            call('pgk_to_rgb_dgrad', dimg.data_ptr(), n, C, res // 2, res // 2, hprev.C, lo.conv.weight.data_ptr(),
                 lo.cf, k_lo, 1, d_lo.ptr, P, d_lo.ps, dp_lo)
        for i in range(depth, -1, -1):
            b = self.block(i)
            hprev, hup, u1, u2 = T.acts[i]
            w2 = self.cw(b.c2)
            da2 = self._act_bwd(d, u2, T, 'b%dc2' % i)
            w2.wgrad_into(gs[b.c2.conv.weight], u1, da2, res, res, 0, groups, n, db=gs[b.c2.conv.bias], bias_goffs=[0],
                          pending=pending)
            d1 = conv(da2, w2.B, w2.cin, 3, PT.empty(n, res, res, w2.cin, P, dev))
            da1 = self._act_bwd(d1, u1, T, 'b%dc1' % i)
            if i == 0:
                w1 = self.cw(b.c1, W_GFIRST)
                w1.wgrad_into(gs[b.c1.conv.weight], T.zn, da1.view(1, 1, 16 * w1.cout), 1, 1, 0, groups, n, pending=pending)
                bias_grad(da1, 16, w1.cout, [0], n, gs[b.c1.conv.bias])
                break
            w1 = self.cw(b.c1)
            w1.wgrad_into(gs[b.c1.conv.weight], hup, da1, res, res, 0, groups, n, db=gs[b.c1.conv.bias], bias_goffs=[0],
                          pending=pending)
            d_up = conv(da1, w1.B, w1.cin, 3, PT.empty(n, res, res, w1.cin, P, dev))
            res //= 2
            d = PT.empty(n, res, res, w1.cin, P, dev)
            if d_lo is not None:
                pool2(d_up, d, avg=0, a=1.0, other=d_lo, b=1.0)
                d_lo = None
            else:
                pool2(d_up, d, avg=0, a=1.0)
        unprep_all(pending)

    def _act_bwd(self, d, y, T, name):
        """gradient w.r.t. the conv output given the gradient w.r.t. the (lrelu -> pixelnorm) output y."""
        out = PT.empty(y.N, y.H, y.W, y.C, d.P, y.t.device)
        if self.G.pixelnorm:
            r = getattr(T, 'r_' + name)
            call('pgk_pixelnorm_bwd', d.ptr, d.ps, y.ptr, y.ps, r.data_ptr(), d.P, y.N * y.H * y.W, y.C, out.ptr,
                 out.ps)
        else:
            mask_mul(d, out, ref=y)
        return out
