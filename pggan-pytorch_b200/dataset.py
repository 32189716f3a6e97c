"""Device-side real-image preparation: the step immediately before the training path.

The reference does this per sample on the CPU inside ``DepthDataset.__getitem__`` (dataset.py:60-67): ``alpha_fade``
(dataset.py:109-113) while a new resolution fades in, then ``adjust_dynamic_range`` (utils.py:24-30), float32 cast,
pageable H2D copy of float data.  Here the raw batch (uint8 as stored, or float) is copied once -- pinned memory makes
it asynchronous -- and one libpgk kernel does fade + range + cast on the device.
"""
import torch

from . import _lib


def prepare_reals(batch, alpha=1.0, range_in=(0, 255), range_out=(-1, 1), out=None):
    """batch: (N, C, H, W) uint8 or float32 tensor at the CURRENT depth's resolution, on the host (ideally pinned) or
    already on the device.  Returns the float32 CUDA tensor the loss functions take as ``real_images_in``.

    Follows the reference's *intended* semantics: alpha is the value of the current iteration (the reference's worker
    processes keep a stale alpha between depth changes, plugins.py:65-77)."""
    if batch.dtype not in (torch.uint8, torch.float32):
        raise TypeError('prepare_reals takes uint8 or float32 batches, got %s' % batch.dtype)
    if batch.dim() != 4:
        raise ValueError('prepare_reals takes (N, C, H, W) batches')
    if not batch.is_cuda:
        if not torch.cuda.is_available():
            raise _lib.PgkError('prepare_reals runs on sm_100a CUDA devices only (there is no CPU path)')
        batch = batch.cuda(non_blocking=True)
    _lib.check_device(batch.device)
    batch = batch.contiguous()
    n, c, h, w = batch.shape
    if out is None:
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=batch.device)
    _lib.call('pgk_real_prep', batch.data_ptr(), 1 if batch.dtype == torch.uint8 else 0, n, c, h, w, float(alpha),
              float(range_in[0]), float(range_in[1]), float(range_out[0]), float(range_out[1]), out.data_ptr())
    return out
