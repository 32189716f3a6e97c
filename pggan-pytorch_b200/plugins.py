"""DepthManager: the progressive-growing schedule of the reference (plugins.py:13-81), host side, integer exact.

``torch.utils.trainer`` (the base classes the reference imports, plugins.py:8-9) no longer exists in PyTorch, so a
minimal ``Plugin`` base is provided here with the same contract: ``trigger_interval`` = [(n, unit)], ``register``.
"""
import math


class Plugin(object):
    def __init__(self, interval=None):
        if interval is None:
            interval = []
        self.trigger_interval = interval

    def register(self, trainer):
        raise NotImplementedError


def schedule(cur_nimg, max_depth, lod_training_nimg=100 * 1000, lod_transition_nimg=100 * 1000):
    """cur_nimg -> (depth, alpha).  Each level is a transition phase followed by a stabilisation phase; depth 0 has
    only the latter.  Pure integer divmod + one true division, bit-identical to plugins.py:57-63."""
    period = lod_training_nimg + lod_transition_nimg
    full_passes, rem = divmod(cur_nimg, period)
    in_transition, rem = divmod(rem, lod_training_nimg)
    level = full_passes + in_transition
    depth = min(max_depth, level)
    if in_transition > 0 and level == depth:
        return depth, rem / lod_transition_nimg
    return depth, 1.0


class DepthManager(Plugin):

    def __init__(self, create_dataloader_fun, create_rlg, max_depth, minibatch_default=16,
                 minibatch_overrides={6: 14, 7: 6, 8: 3}, tick_kimg_default=20,
                 tick_kimg_overrides={3: 10, 4: 10, 5: 5, 6: 2, 7: 2, 8: 1}, lod_training_nimg=100 * 1000,
                 lod_transition_nimg=100 * 1000, max_lod=None, depth_offset=None):
        super().__init__([(1, 'iteration')])
        self.create_dataloader_fun, self.create_rlg = create_dataloader_fun, create_rlg
        self.max_depth = max_depth
        self.minibatch_default, self.minibatch_overrides = minibatch_default, minibatch_overrides
        self.tick_kimg_default, self.tick_kimg_overrides = tick_kimg_default, tick_kimg_overrides
        self.lod_training_nimg, self.lod_transition_nimg = lod_training_nimg, lod_transition_nimg
        self.max_lod, self.depth_offset = max_lod, depth_offset
        self.trainer = None
        self.depth = -1
        self.alpha = -1

    @property
    def _has_lod(self):
        return self.max_lod is not None and self.depth_offset is not None

    @property
    def lod(self):
        if self._has_lod:
            return self.max_lod - self.depth_offset - self.depth - self.alpha + 1
        return -1

    def register(self, trainer):
        self.trainer = trainer
        trainer.stats['minibatch_size'] = self.minibatch_default
        trainer.stats['alpha'] = {'log_name': 'alpha', 'log_epoch_fields': ['{val:.2f}'], 'val': self.alpha}
        if self._has_lod:
            trainer.stats['lod'] = {'log_name': 'lod', 'log_epoch_fields': ['{val:.2f}'], 'val': self.lod}
        self.iteration()

    def iteration(self, *args):
        tr = self.trainer
        depth, alpha = schedule(tr.cur_nimg, self.max_depth, self.lod_training_nimg, self.lod_transition_nimg)
        if depth != self.depth:
            # new resolution: networks, dataset, minibatch size, data loader, latent generator, tick length
            tr.D.depth = tr.G.depth = tr.dataset.model_depth = depth
            self.depth = depth
            mb = self.minibatch_overrides.get(depth, self.minibatch_default)
            tr.dataiter = iter(self.create_dataloader_fun(mb))
            tr.random_latents_generator = self.create_rlg(mb)
            tr.tick_duration_nimg = self.tick_kimg_overrides.get(depth, self.tick_kimg_default) * 1000
            tr.stats['minibatch_size'] = mb
        if alpha != self.alpha:
            tr.D.alpha = tr.G.alpha = tr.dataset.alpha = alpha
            self.alpha = alpha
        tr.stats['depth'] = depth
        tr.stats['alpha']['val'] = alpha
        if self._has_lod:
            tr.stats['lod']['val'] = self.lod


def lr_rampup(cur_nimg, lr_rampup_kimg=40):
    """LambdaLR factor of train.py:151-156."""
    if cur_nimg < lr_rampup_kimg * 1000:
        p = max(0.0, 1 - cur_nimg / (lr_rampup_kimg * 1000))
        return math.exp(-p * p * 5.0)
    return 1.0


class LRScheduler(Plugin):
    """Steps both LambdaLR schedulers with cur_nimg every iteration (plugins.py:84-99)."""

    def __init__(self, lr_scheduler_d, lr_scheduler_g):
        super().__init__([(1, 'iteration')])
        self.lrs_d, self.lrs_g = lr_scheduler_d, lr_scheduler_g

    def register(self, trainer):
        self.trainer = trainer
        self.iteration()

    def iteration(self, *args):
        self.lrs_d.step(self.trainer.cur_nimg)
        self.lrs_g.step(self.trainer.cur_nimg)


class AsyncLossMonitor(Plugin):
    """EfficientLossMonitor (plugins.py:102-111) without the device synchronisation it costs every iteration: the
    reference reads `val.data[0]` on the host after each iteration (four monitors = four D2H syncs per iteration, which
    stop the host from running ahead of the GPU).  Here the running sum stays on the device and the host reads one mean
    per epoch (tick).  Same arguments (`loss_no`: 0 G loss, 1 D cost, 2 D real loss, 3 D fake loss; `stat_name`), same
    `trainer.stats[stat_name]` fields `last` / `epoch_mean` as torch.utils.trainer's LossMonitor wrote; `last` is filled at
    the epoch boundary only."""

    def __init__(self, loss_no, stat_name):
        super().__init__([(1, 'iteration'), (1, 'epoch')])
        self.loss_no, self.stat_name = loss_no, stat_name
        self._sum = None
        self._last = None
        self._count = 0

    def register(self, trainer):
        self.trainer = trainer
        trainer.stats.setdefault(self.stat_name, {'log_format': ':.4f', 'log_epoch_fields': ['{epoch_mean:.4f}']})

    def iteration(self, iteration, *args):
        val = args[self.loss_no]
        val = val.detach() if self.loss_no < 2 else val.detach().mean()
        self._sum = val.clone() if self._sum is None else self._sum.add_(val)      # enqueued on the stream: no sync
        self._last = val
        self._count += 1

    def epoch(self, epoch_index):
        st = self.trainer.stats[self.stat_name]
        if self._count:
            st['epoch_mean'] = float(self._sum) / self._count       # the one device read of the tick
            st['last'] = float(self._last)
        self._sum, self._last, self._count = None, None, 0
