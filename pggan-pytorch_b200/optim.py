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
        self._host = None     # pinned staging for the pointer table
        self._dev = None
        self._keep = []       # gradient buffers of the launch in flight

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            beta1, beta2 = group['betas']
            rows, max_numel, dev, keep = [], 0, None, []
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
                             struct.unpack('<I', struct.pack('<f', inv_sqrt_bc2))[0], 0))
                keep.append(g)       # the gradient buffer must outlive the asynchronous launch
                max_numel = max(max_numel, p.numel())
                dev = p.device
            if not rows:
                continue
            _lib.check_device(dev)
            n = len(rows)
            if self._host is None or self._host.shape[0] < n or self._dev.device != dev:
                cap = max(64, n)
                self._host = torch.empty((cap, 8), dtype=torch.int64).pin_memory()
                self._dev = torch.empty((cap, 8), dtype=torch.int64, device=dev)
            self._host[:n] = torch.from_numpy(np.array(rows, dtype=np.uint64).view(np.int64))
            self._dev[:n].copy_(self._host[:n], non_blocking=True)
            _lib.call('pgk_adam_multi', self._dev.data_ptr(), n, max_numel, float(beta1), float(beta2),
                      float(group['eps']))
            self._keep = keep
        return loss
