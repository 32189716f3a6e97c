"""The training-iteration surface of the reference (trainer.py:5-116), host side.

``Trainer.train()`` is the unit of work of the whole repository: one discriminator step (loss, backward,
optimizer) per ``D_training_repeats`` and one generator step, then the 'iteration' plugins.  The losses are injected
callables (``wgan_gp_loss.wgan_gp_D_loss`` / ``wgan_gp_G_loss``); nothing here touches activations.
"""
import heapq


class Trainer(object):

    def __init__(self, D, G, D_loss, G_loss, optimizer_d, optimizer_g, dataset, dataiter, random_latents_generator,
                 D_training_repeats=1, tick_nimg_default=2 * 1000, resume_nimg=0):
        self.D, self.G = D, G
        self.D_loss, self.G_loss = D_loss, G_loss
        self.optimizer_d, self.optimizer_g = optimizer_d, optimizer_g
        self.dataset, self.dataiter = dataset, dataiter
        self.random_latents_generator = random_latents_generator
        self.D_training_repeats = D_training_repeats
        self.cur_nimg = resume_nimg
        self.tick_start_nimg = resume_nimg
        self.tick_duration_nimg = tick_nimg_default
        self.iterations = 0
        self.cur_tick = 0
        self.time = 0
        self.stats = {
            'kimg_stat': {'val': self.cur_nimg / 1000., 'log_epoch_fields': ['{val:8.3f}'], 'log_name': 'kimg'},
            'tick_stat': {'val': self.cur_tick, 'log_epoch_fields': ['{val:5}'], 'log_name': 'tick'},
        }
        self.plugin_queues = {'iteration': [], 'epoch': [], 's': [], 'end': []}

    # -- plugin bus (reference trainer.py:47-69): a heap of (next trigger time, registration order, plugin) ------
    def register_plugin(self, plugin):
        plugin.register(self)
        intervals = plugin.trigger_interval
        if not isinstance(intervals, list):
            intervals = [intervals]
        for duration, unit in intervals:
            q = self.plugin_queues[unit]
            q.append((duration, len(q), plugin))

    def call_plugins(self, queue_name, time, *args):
        q = self.plugin_queues[queue_name]
        while q and q[0][0] <= time:
            due, order, plugin = q[0]
            getattr(plugin, queue_name)(time, *args)
            interval = [d for d, unit in plugin.trigger_interval if unit == queue_name][-1]
            heapq.heapreplace(q, (time + interval, order, plugin))

    def run(self, total_kimg=1):
        for q in self.plugin_queues.values():
            heapq.heapify(q)
        total = total_kimg * 1000
        while self.cur_nimg < total:
            self.train()
            if self.cur_nimg >= self.tick_start_nimg + self.tick_duration_nimg or self.cur_nimg >= total:
                self.cur_tick += 1
                self.tick_start_nimg = self.cur_nimg
                self.stats['kimg_stat']['val'] = self.cur_nimg / 1000.
                self.stats['tick_stat']['val'] = self.cur_tick
                self.call_plugins('epoch', self.cur_tick)
        self.call_plugins('end', 1)

    @staticmethod
    def _to_device(t):
        return t.cuda(non_blocking=True)

    def train(self):
        """One iteration (reference trainer.py:85-115)."""
        latents = self._to_device(self.random_latents_generator())
        d_losses = (0, 0, 0)
        for _ in range(self.D_training_repeats):
            real = self._to_device(next(self.dataiter))
            self.cur_nimg += real.size(0)
            d_losses = tuple(self.D_loss(self.D, self.G, real, latents))
            d_losses[0].backward()
            self.optimizer_d.step()
            latents = self._to_device(self.random_latents_generator())   # fresh latents for the next step
        g_losses = self.G_loss(self.G, self.D, latents)
        if isinstance(g_losses, list):
            g_losses = tuple(g_losses)
        elif not isinstance(g_losses, tuple):
            g_losses = (g_losses,)
        g_losses[0].backward()
        self.optimizer_g.step()
        self.iterations += 1
        self.call_plugins('iteration', self.iterations, *(g_losses + d_losses))
