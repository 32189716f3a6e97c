"""FusedAdam: torch.optim.Adam as the reference wires it (train.py:148-149,195: lr 1e-3, betas (0.0, 0.99), eps 1e-8,
no weight decay) with ONE libpgk launch per step for every parameter that has a gradient.

Drop-in for ``torch.optim.Adam`` in train.py: same constructor keywords, ``param_groups`` (so ``LambdaLR`` keeps
working, train.py:157-158), per-parameter ``state['step'/'exp_avg'/'exp_avg_sq']``, parameters whose ``grad`` is None
are skipped exactly as torch does (the inactive blocks of the current depth, trainer.py:100).
"""
import math
import struct

import numpy as np
import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError('FusedAdam implements the reference configuration: no weight decay, no amsgrad')
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False))
        # Pointer tables go host -> device asynchronously, and the host runs ahead of the GPU (no sync per step): a
        # ring of pinned staging buffers, each guarded by an event recorded after its copy, so that a slot is never
        # rewritten while the DMA that reads it is still pending.  The device table is allocated per launch from
        # the caching allocator (stream-ordered reuse).
        self._ring = []       # [pinned host table, copy-done event or None]
        self._slot = 0
        self._keep = []       # gradient buffers of the launch in flight

    RING = 8

    def _stage(self, n):
        """The next pinned staging table with room for n rows, safe to overwrite."""
        if len(self._ring) < self.RING:
            self._ring.append([torch.empty((max(64, n), 8), dtype=torch.int64).pin_memory(), None])
            self._slot = len(self._ring) - 1
        else:
            self._slot = (self._slot + 1) % self.RING
        host, done = self._ring[self._slot]
        if done is not None:
            done.synchronize()          # returns at once unless the host is more than RING launches ahead
        if host.shape[0] < n:
            host = self._ring[self._slot][0] = torch.empty((n, 8), dtype=torch.int64).pin_memory()
        return host

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            beta1, beta2 = group['betas']
            rows, nblocks, dev, keep, updated = [], 0, None, [], []
            chunk = _lib.adam_chunk()
            for p in group['params']:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise _lib.PgkError('FusedAdam runs on sm_100a CUDA tensors only (there is no CPU path)')
                g = p.grad
                if g.dtype != torch.float32 or p.dtype != torch.float32 or not p.is_contiguous():
                    raise TypeError('FusedAdam takes contiguous fp32 parameters and gradients')
                g = g.contiguous()
                st = self.state[p]
                if not st:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st['step'] += 1
                t = st['step']
                step_size = group['lr'] / (1.0 - beta1 ** t)
                inv_sqrt_bc2 = 1.0 / math.sqrt(1.0 - beta2 ** t)
                rows.append((p.data_ptr(), g.data_ptr(), st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr(), p.numel(),
                             struct.unpack('<I', struct.pack('<f', step_size))[0],
                             struct.unpack('<I', struct.pack('<f', inv_sqrt_bc2))[0], nblocks))
                nblocks += (p.numel() + chunk - 1) // chunk     # (include/pgk.h: one block per PGK_ADAM_CHUNK elements)
                keep.append(g)       # the gradient buffer must outlive the asynchronous launch
                updated.append(p)
                dev = p.device
            if not rows:
                continue
            _lib.check_device(dev)
            n = len(rows)
            host = self._stage(n)
            host[:n] = torch.from_numpy(np.array(rows, dtype=np.uint64).view(np.int64))
            table = torch.empty((n, 8), dtype=torch.int64, device=dev)
            table.copy_(host[:n], non_blocking=True)
            self._ring[self._slot][1] = torch.cuda.Event()
            self._ring[self._slot][1].record()
            _lib.call('pgk_adam_multi', table.data_ptr(), n, nblocks, float(beta1), float(beta2),
                      float(group['eps']))
            # the kernel writes through raw pointers: tell autograd (and every cache keyed on Tensor._version, e.g.
            # engine.ConvW's re-laid weights) that the parameters changed
            torch.autograd.graph.increment_version(updated)
            self._keep = keep
        return loss
