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
        # Opt-in ("pinned async H2D"): fetch the NEXT real batch and start its host-to-device copy on a
        # copy stream as soon as this iteration's kernels are enqueued, so that the copy runs under the iteration's GPU
        # work instead of in front of the next one (at depth 8 the 50 MB batch is ~1 ms of a 14 ms iteration).  Only
        # the real images are fetched ahead -- the latents stay where the reference draws them (the order of numpy
        # RNG draws is observable) -- and a batch fetched from an iterator that a plugin has since replaced
        # (DepthManager on a depth change, plugins.py:65-77) is dropped.
        self.prefetch_reals = False
        self._ahead = None          # (iterator it came from, device tensor, copy-done event or None)
        self._copy_stream = None

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

    def _copy_ahead(self, host):
        """Start the host-to-device copy of `host` on the copy stream; returns (device tensor, event)."""
        import torch
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
        with torch.cuda.stream(self._copy_stream):
            dev = host.cuda(non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._copy_stream)
        return dev, done

    def _claim(self, dev, done):
        """Make a tensor copied on the copy stream usable on the current stream."""
        if done is not None:
            import torch
            cur = torch.cuda.current_stream()
            cur.wait_event(done)
            dev.record_stream(cur)
        return dev

    def _next_real(self):
        ahead, self._ahead = self._ahead, None
        if ahead is not None and ahead[0] is self.dataiter:
            return self._claim(ahead[1], ahead[2])
        return self._to_device(next(self.dataiter))      # the reference's own path (trainer.py:92)

    def _fetch_ahead(self):
        if not self.prefetch_reals:
            return
        it = self.dataiter
        try:
            host = next(it)
        except StopIteration:       # the next train() call raises it, exactly where the reference would
            return
        if getattr(host, 'is_pinned', lambda: False)():
            dev, done = self._copy_ahead(host)
        else:
            dev, done = self._to_device(host), None
        self._ahead = (it, dev, done)

    def train(self):
        """One iteration (reference trainer.py:85-115)."""
        latents = self._to_device(self.random_latents_generator())
        d_losses = (0, 0, 0)
        for _ in range(self.D_training_repeats):
            real = self._next_real()
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
        self._fetch_ahead()         # before the plugins: their loss reads synchronise with the device
        self.iterations += 1
        self.call_plugins('iteration', self.iterations, *(g_losses + d_losses))
