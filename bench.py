"""bench.py -- images/sec of one Progressive-GAN training iteration (D step with gradient penalty + G step +
both Adam updates; reference trainer.py:85-115) on synthetic data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One JSON line on rank 0.  `value` = whole-job images/sec with inputs resident in HBM; `e2e` = the same through
Trainer.train() with pinned HOST inputs (H2D inside the timed region) and a D2H read of the loss every step;
`roofline` = the dominant kernel family (pgk_conv: forward conv / data gradient) timed per launch with CUDA events on
the launching stream; `cpu_baseline` = the CPU oracle (a port of the reference's step) on a bounded sample.
`--impl reference` times that CPU path alone with all host threads.

Configs (BASELINE.json): c1 depth 0 N16 fp32 | c2 depth 4 alpha .5 N128 fp32 (default: the metric's 1-GPU config) |
c3 depth 6 N32 bf16 | c4 depth 8 alpha .3 N4 bf16 | c5 depth 5, 1-channel 128x128 model, N64 bf16.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    'c1': dict(depth=0, alpha=1.0, n=16, precision='fp32', res=1024, ch=3),
    'c2': dict(depth=4, alpha=0.5, n=128, precision='fp32', res=1024, ch=3),
    'c3': dict(depth=6, alpha=1.0, n=32, precision='bf16', res=1024, ch=3),
    'c4': dict(depth=8, alpha=0.3, n=4, precision='bf16', res=1024, ch=3),
    'c5': dict(depth=5, alpha=1.0, n=64, precision='bf16', res=128, ch=1),
}
DTYPE = {'fp32': 'bf16x3 (fp32-faithful: tensors as three bf16 planes = 24 mantissa bits; wide forward convs read two '
                 'fp16 planes = 22 bits; fp32 accumulate)',
         'bf16': 'bf16 (fp32 accumulate)'}


def nf(stage, fmap_base=4096, fmap_max=512):
    return min(int(fmap_base / (2.0 ** stage)), fmap_max)


def flops_per_image(depth, ch, fade):
    """Algorithmic FLOPs of one iteration per image: 2*(14*MAC_D + 4*MAC_G) (SURVEY.md 8d / BASELINE.md 3), the
    4x4 first-G / last-D convs counted as the dense GEMMs they are."""
    mac_g, mac_d = flops_per_image_parts(depth, ch, fade)
    return 2.0 * (14 * mac_d + 4 * mac_g)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(',')])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [nm for i, nm in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == 'Active' for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx[0] if mx else None, 'reasons': reasons,
                'samples': len(sm)}


def cpu_reference_leg(cfg, steps, warmup, budget_s=25.0):
    """The reference's algorithm for the step on the host CPU: oracle/pggan_oracle.py (a functional port of
    network.py / wgan_gp_loss.py / trainer.py:85-115 on torch CPU ops, all host threads), D step + Adam + G step +
    Adam, on a bounded sample (a small batch of the same depth / alpha / resolution).  Returns img/s."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import pggan_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    depth, alpha, ch, res = cfg['depth'], cfg['alpha'], cfg['ch'], cfg['res']
    n = max(1, min(cfg['n'], {0: 16, 1: 16, 2: 8, 3: 4, 4: 2}.get(depth, 1)))
    pgp = O.make_generator_params(res, ch, seed=1337)
    pdp = O.make_discriminator_params(res, ch, seed=1338)
    nb = O.n_blocks_for(res)
    gen = torch.Generator().manual_seed(1337)
    r = 4 * 2 ** depth
    sd, sg = {}, {}

    def step():
        nonlocal pgp, pdp
        real = torch.randn(n, ch, r, r, generator=gen)
        z1, z2 = torch.randn(n, 512, generator=gen), torch.randn(n, 512, generator=gen)
        mix = torch.rand(n, 1, generator=gen)
        _, _, _, gd = O.d_step_grads(pdp, pgp, real, z1, mix, depth, alpha, nb)
        pdp = O.adam_step(dict(pdp), gd, sd, 1e-3)
        _, gg = O.g_step_grads(pgp, pdp, z2, depth, alpha, nb)
        pgp = O.adam_step(dict(pgp), gg, sg, 1e-3)

    t_first = time.perf_counter()
    step()
    t_first = time.perf_counter() - t_first
    # bound the whole leg: as many warm-up / timed steps as requested, but never beyond the budget
    k = max(1, min(steps, int(budget_s / max(t_first, 1e-3))))
    w = max(0, min(warmup - 1, int(0.3 * budget_s / max(t_first, 1e-3))))
    for _ in range(w):
        step()
    t0 = time.perf_counter()
    for _ in range(k):
        step()
    dt = (time.perf_counter() - t0) / k
    return dict(value=n / dt, unit='images/sec', cores=cores, kind='port',
                sample='%d timed iteration(s) of batch %d at depth %d (%dx%d), alpha %g, fp32, torch CPU ops on %d threads'
                       % (k, n, depth, r, r, alpha, cores)), dt, k, w + 1     # the probe step is an untimed warm-up too


FAMILIES = {0: 'conv_tc_kernel (tcgen05 implicit-GEMM conv: forward + data gradient)',
            1: 'wgrad_tc_kernel (tcgen05 weight gradient)', 2: 'conv_simt_kernel (CUDA-core conv)',
            3: 'wgrad_simt_kernel (CUDA-core weight gradient)',
            4: 'conv_thin_kernel (row-streaming tcgen05 conv, Cin 8/16/32)',
            5: 'wgrad_thin_kernel (row-streaming tcgen05 weight gradient, Cin 8/16/32)'}


def assemble_roofline(config, cfg, fam, prod, ksteps, step_s, batch_overridden=False):
    """The `roofline` object of the JSON line from the per-family launch records (pure function; CPU-tested).
    fam[k] = (algorithmic FLOPs, algorithmic bytes, device ms, launches) and prod[k] = bf16 tensor-core FLOPs issued,
    summed over `ksteps` steps of `step_s` seconds each, for the kernel families of include/pgk.h (PGK_PROF_*)."""
    depth, alpha, n, ch = cfg['depth'], cfg['alpha'], cfg['n'], cfg['ch']
    fade = depth > 0 and alpha < 1.0
    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak_tf = float(peaks.get('bf16_tflops_sustained', 1400.0))
    peak_src = 'MEASURED_PEAKS.json bf16_tflops_sustained' if peaks else 'fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)'
    fimg = flops_per_image(depth, ch, fade)
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    dom = max(fam, key=lambda k: fam[k][2])           # the kernel family with the largest share of the step
    fl, by, ms_k, n_k = fam[dom]
    tf = fl / (ms_k * 1e-3) / 1e12 if ms_k > 0 else 0.0
    gbs = by / (ms_k * 1e-3) / 1e9 if ms_k > 0 else 0.0
    if dom in (4, 5):      # thin layers: 36..190 flop/byte, left of the ridge (209): HBM roofline
        roof = {'bound': 'hbm', 'kernel': FAMILIES[dom], 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s',
                'frac': gbs / hbm_peak, 'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback 6.65 TB/s'}
    else:
        roof = {'bound': 'tensor', 'kernel': FAMILIES[dom], 'achieved': tf, 'peak': peak_tf, 'unit': 'TFLOP/s',
                'frac': tf / peak_tf, 'peak_source': peak_src}
    # DRAM bytes per launch of the dominant kernel from a committed ncu capture of this config (tools/ncu_traffic.py)
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            t = json.load(f).get(config, {}).get(FAMILIES[dom].split(' ')[0])
        if t and not batch_overridden:
            traffic, traffic_src = t['bytes_per_launch'], t['source']
    except Exception:
        pass
    if traffic is not None:
        roof['traffic_source'] = traffic_src
        roof['algorithmic_bytes_per_launch'] = by / max(1, n_k)
    # what the tensor pipe itself executed for the dominant kernel: issued bf16 products per second against the same
    # measured peak (in the bf16 mode this equals achieved / frac; in the fp32-faithful mode it is 3..6x higher)
    ptf = prod[dom] / (ms_k * 1e-3) / 1e12 if ms_k > 0 else 0.0
    roof['tensor_pipe'] = {'issued_tflops': ptf, 'frac_of_peak': ptf / peak_tf,
                           'products_per_flop': prod[dom] / fl if fl > 0 else 0.0}
    roof.update(traffic=traffic, launches_timed=n_k, kernel_ms_per_step=ms_k / ksteps,
                share_of_step=ms_k / ksteps / (step_s * 1e3),
                families={FAMILIES[k].split(' ')[0]: {'ms_per_step': v[2] / ksteps, 'launches': v[3],
                                                       'tflops': v[0] / (v[2] * 1e-3) / 1e12 if v[2] > 0 else 0.0,
                                                       'issued_tflops': prod[k] / (v[2] * 1e-3) / 1e12 if v[2] > 0 else 0.0,
                                                       'gbs': v[1] / (v[2] * 1e-3) / 1e9 if v[2] > 0 else 0.0}
                          for k, v in fam.items() if v[3]},
                step_algorithmic={'gflop_per_image': fimg / 1e9, 'achieved': fimg * n / step_s / 1e12,
                                  'frac': fimg * n / step_s / 1e12 / peak_tf})
    if cfg['precision'] == 'fp32':
        roof['note'] = ('fp32-faithful mode: every algorithmic FLOP costs 3 tensor-core products (forward: two fp16 '
                        'planes; gradient chains: two bf16 planes; 6 with PGK_FWD_FP16=0), so frac <= 1/3 by '
                        'construction; tensor_pipe counts the products')
    return roof


def d_step_flops_per_image(depth, ch, fade):
    """Algorithmic FLOPs of the D step alone per image: 2*(12*MAC_D + MAC_G) (324.1 GFLOP at depth 8)."""
    mac_g, mac_d = flops_per_image_parts(depth, ch, fade)
    return 2.0 * (12 * mac_d + mac_g)


def flops_per_image_parts(depth, ch, fade):
    """(MAC_G, MAC_D) of one forward pass (the two terms of flops_per_image)."""
    mac_g = nf(0) * nf(1) * 16 + 16 * nf(1) * nf(1) * 9
    mac_d = 16 * nf(1) * nf(1) * 9 + 16 * nf(1) * nf(0) + nf(0)
    for j in range(1, depth + 1):
        px = (4 * 2 ** j) ** 2
        mac_g += px * 9 * (nf(j) * nf(j + 1) + nf(j + 1) * nf(j + 1))
        mac_d += px * 9 * (nf(j + 1) * nf(j + 1) + nf(j + 1) * nf(j))
    px = (4 * 2 ** depth) ** 2
    mac_g += px * nf(depth + 1) * ch
    mac_d += px * nf(depth + 1) * ch
    if fade:
        mac_g += (px // 4) * nf(depth) * ch
        mac_d += (px // 4) * nf(depth) * ch
    return mac_g, mac_d


def reference_leg_cpu(cfg, steps, warmup, budget_s):
    """The reference arm / cpu_baseline: the UNMODIFIED reference's Trainer.train() (baseline/_ref, installed by
    __graft_entry__.build()) on the host cores when it is there (kind "reference"), else the oracle port (kind "port").
    The batch is a bounded sample of the config's: the largest power of two whose iteration stays near a second."""
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    import ref_harness as R
    if not R.available():
        return cpu_reference_leg(cfg, steps, warmup, budget_s)
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    depth, alpha, ch, res = cfg['depth'], cfg['alpha'], cfg['ch'], cfg['res']
    n = max(1, min(cfg['n'], {0: 16, 1: 16, 2: 16, 3: 16, 4: 8, 5: 4, 6: 2}.get(depth, 1)))
    ips, ms, k, w = R.time_train(res, ch, depth, alpha, n, steps, warmup, device='cpu', budget_s=budget_s)
    r = 4 * 2 ** depth
    return dict(value=ips, unit='images/sec', cores=cores, kind='reference',
                sample="%d timed iteration(s) of the unmodified reference's Trainer.train() (trainer.py:85-115), batch %d of the "
                       "config's %d at depth %d (%dx%d), alpha %g, fp32, torch CPU ops on %d threads"
                       % (k, n, cfg['n'], depth, r, r, alpha, cores)), ms * 1e-3, k, w


def gpu_eager_reference(cfg, steps=3, warmup=2):
    """SURVEY.md 8(d) "GPU reference bar": the unmodified reference's Trainer.train() in PyTorch eager on this GPU at
    the config's own batch, fp32 (cuDNN fp32 kernels) and with TF32 allowed -- the number the hand-written kernels
    have to beat.  None when baseline/_ref is not installed."""
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    import ref_harness as R
    import torch
    if not R.available():
        return None
    out = {'what': "the unmodified reference's Trainer.train() (baseline/_ref: network.py, wgan_gp_loss.py, trainer.py:85-115), "
                   "PyTorch eager + cuDNN on this GPU, inputs resident in HBM", 'batch': cfg['n'], 'steps': steps}
    keep = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    try:
        for name, tf32 in (('fp32', False), ('tf32', True)):
            n = cfg['n']
            while True:
                try:
                    ips, ms, _, _ = R.time_train(cfg['res'], cfg['ch'], cfg['depth'], cfg['alpha'], n, steps, warmup,
                                                 'cuda', tf32)
                    break
                except torch.OutOfMemoryError:
                    torch.cuda.empty_cache()
                    if n == 1:
                        ips, ms = None, None
                        break
                    n //= 2
            out[name] = {'images_per_sec': ips, 'ms_per_step': ms, 'batch': n}
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = keep
    return out


def run_config(name, cfg, args, rank, world, local_rank, steps, warmup, primary):
    """Device-timed value, end-to-end value and the roofline of one BASELINE config.  Returns a dict (rank 0) or None."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import pggan_b200 as pg
    dev = torch.device('cuda', local_rank)
    depth, alpha, n, ch = cfg['depth'], cfg['alpha'], cfg['n'], cfg['ch']
    fade = depth > 0 and alpha < 1.0
    r = 4 * 2 ** depth
    torch.manual_seed(1337)
    np.random.seed(1337 + rank)
    shape = (1000, ch, cfg['res'], cfg['res'])
    G, D = pg.Generator(shape).to(dev), pg.Discriminator(shape).to(dev)
    G.precision = D.precision = cfg['precision']
    pg.wgan_gp_loss.cuda_graphs = args.graphs
    G.depth = D.depth = depth
    G.alpha = D.alpha = alpha
    opt_g = pg.FusedAdam(G.parameters(), 1e-3, betas=(0.0, 0.99))      # train.py:148-149,195
    opt_d = pg.FusedAdam(D.parameters(), 1e-3, betas=(0.0, 0.99))
    gen = torch.Generator(device=dev).manual_seed(1337 + rank)
    nbuf = 4
    reals = [torch.randn(n, ch, r, r, device=dev, generator=gen) for _ in range(nbuf)]
    lats = [torch.randn(n, 512, device=dev, generator=gen) for _ in range(2 * nbuf)]

    def d_step(i):
        cost, _, _ = pg.wgan_gp_D_loss(D, G, reals[i % nbuf], lats[(2 * i) % (2 * nbuf)])
        cost.backward()
        opt_d.step()

    def step_device(i):
        """inputs already resident in HBM"""
        d_step(i)
        gcost = pg.wgan_gp_G_loss(G, D, lats[(2 * i + 1) % (2 * nbuf)])
        gcost.backward()
        opt_g.step()
        return gcost

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        sync_all()
        return float(ms)

    for i in range(warmup):
        step_device(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    torch.cuda.reset_peak_memory_stats(dev)
    pg._lib.reset_launch_count()
    ms = timed(step_device, steps)
    launches = pg._lib.launch_count()
    if not primary:
        # the secondary configs share the process with the ones before them (allocator pools just emptied, clocks and
        # power state left by another workload): the faster of two K-step segments is reported, and the line says so
        ms = min(ms, timed(step_device, steps))
    clocks = sampler.finish() if rank == 0 else None
    peak_mem = torch.cuda.max_memory_allocated(dev)

    if primary and os.environ.get('PGK_BENCH_MAIN_ONLY'):    # profiling runs (ncu): the device-resident leg only
        if rank == 0:
            print(json.dumps({'ms_per_step': ms / steps, 'gpu_launches': launches, 'note': 'main leg only'}))
        return None

    # ---- the D step alone (the north star's 60 % tensor-pipe target is defined on the 1024^2 D step) ----------
    ms_d = timed(d_step, steps) if (depth >= 8 or args.d_step) else None

    # ---- end to end: Trainer.train() fed from pinned host memory, loss read back every step ----------------
    host_reals = [torch.randn(n, ch, r, r).pin_memory() for _ in range(2)]
    host_lats = [torch.from_numpy(np.random.randn(n, 512).astype(np.float32)).pin_memory() for _ in range(4)]
    cnt = {'r': 0, 'l': 0}

    def next_real():
        while True:
            cnt['r'] += 1
            yield host_reals[cnt['r'] % 2]

    def next_lat():
        cnt['l'] += 1
        return host_lats[cnt['l'] % 4]

    got = {}

    class Grab(pg.Plugin):
        def __init__(self):
            super().__init__([(1, 'iteration')])

        def register(self, trainer):
            pass

        def iteration(self, it, g_cost, d_cost, *rest):
            got['v'] = (float(g_cost.detach()), float(d_cost.detach()))   # D2H read of both losses (2 x 4 bytes)

    tr = pg.Trainer(D, G, pg.wgan_gp_D_loss, pg.wgan_gp_G_loss, opt_d, opt_g, None, next_real(), next_lat)
    tr.register_plugin(Grab())
    tr.prefetch_reals = args.prefetch
    import heapq
    for q in tr.plugin_queues.values():
        heapq.heapify(q)
    for _ in range(2):
        tr.train()
    ms_e2e = timed(lambda i: tr.train(), steps)
    h2d = n * ch * r * r * 4 + 2 * n * 512 * 4
    d2h = 8

    # ---- roofline leg: per-launch CUDA-event timing of the conv kernels over the same steps -----------------
    pg._lib.prof_reset()
    pg._lib.prof_enable(True)
    sync_all()
    ksteps = min(steps, 3)
    for i in range(ksteps):
        step_device(i)
    torch.cuda.synchronize()
    pg._lib.prof_enable(False)
    fam = {k: pg._lib.prof_read(k) for k in FAMILIES}
    prod = {k: pg._lib.prof_read_products(k) for k in FAMILIES}    # bf16 tensor-core FLOPs issued (1/3/6 per FLOP)
    pg._lib.prof_reset()
    pg.wgan_gp_loss._graphs.clear()
    del G, D, opt_g, opt_d, tr, reals, lats
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    step_s = ms / 1e3 / steps
    roof = assemble_roofline(name, cfg, fam, prod, ksteps, step_s, bool(args.batch))
    out = {'value': n * world / step_s, 'unit': 'images/sec', 'ms_per_step': step_s * 1e3, 'steps': steps,
           'warmup': warmup, 'best_of_segments': 1 if primary else 2, 'dtype': DTYPE[cfg['precision']],
           'workload': '%s: depth %d (%dx%d), alpha %g, batch %d/GPU, %s, D step (WGAN-GP) + G step + 2x Adam'
                       % (name, depth, r, r, alpha, n, cfg['precision']),
           'e2e': {'value': n * world / (ms_e2e / 1e3 / steps), 'unit': 'images/sec', 'h2d_bytes_per_step': h2d,
                   'd2h_bytes_per_step': d2h,
                   'api': 'Trainer.train() with pinned host reals/latents' + (', Trainer.prefetch_reals = True (the next real batch is copied on a copy stream under the running iteration)' if args.prefetch else '')},
           'gpu_launches': launches, 'clocks': clocks, 'roofline': roof, 'peak_mem_gb': peak_mem / 1e9}
    if ms_d is not None:
        peak_tf = roof['peak'] if roof['bound'] == 'tensor' else None
        try:
            with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
                peak_tf = float(json.load(f).get('bf16_tflops_sustained', 1400.0))
        except Exception:
            peak_tf = peak_tf or 1400.0
        fd = d_step_flops_per_image(depth, ch, fade)
        d_s = ms_d / 1e3 / steps
        out['d_step'] = {'ms': d_s * 1e3, 'gflop_per_image': fd / 1e9, 'achieved_tflops': fd * n / d_s / 1e12,
                         'peak_tflops': peak_tf, 'tensor_frac': fd * n / d_s / 1e12 / peak_tf,
                         'what': 'wgan_gp_D_loss + backward + Adam alone, device-timed; algorithmic FLOPs 2*(12*MAC_D + MAC_G) '
                                 'per image against the measured sustained bf16 peak (north star: >= 0.6 at 1024x1024)'}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--config', default='c2', choices=sorted(CONFIGS))
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true',
                    help='skip the secondary legs of the line (other BASELINE configs, GPU-eager reference)')
    ap.add_argument('--d-step', action='store_true', help='time the D step alone at every config (default: depth 8 only)')
    ap.add_argument('--graphs', action='store_true', help='replay the losses as CUDA graphs (wgan_gp_loss.cuda_graphs)')
    ap.add_argument('--batch', type=int, default=0, help='override the per-GPU batch of the config')
    ap.add_argument('--no-prefetch', dest='prefetch', action='store_false',
                    help='e2e leg: without the look-ahead H2D copy of the next real batch (Trainer.prefetch_reals)')
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.batch:
        cfg['n'] = args.batch
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    depth, alpha, n, ch = cfg['depth'], cfg['alpha'], cfg['n'], cfg['ch']
    r = 4 * 2 ** depth
    workload = {'workload': '%s: depth %d (%dx%d), alpha %g, batch %d/GPU, %s, D step (WGAN-GP) + G step + 2x Adam'
                            % (args.config, depth, r, r, alpha, n, cfg['precision']),
                'global_batch': n * max(world, 1), 'parallelism': 'dp%d' % max(world, 1),
                'model_resolution': cfg['res'], 'channels': ch}

    if args.impl == 'reference':
        if rank != 0:
            return
        base, dt, k, w = reference_leg_cpu(cfg, args.steps, args.warmup, budget_s=float(os.environ.get('PGK_CPU_BUDGET_S', '90')))
        print(json.dumps({'impl': 'reference', 'metric': 'images/sec (G+D+GP step)', 'value': base['value'],
                          'unit': 'images/sec', 'n_gpus': args.gpus, 'steps': k, 'warmup': w, 'ms_per_step': dt * 1e3,
                          'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
                          'data': 'synthetic', 'config': workload, 'cpu_baseline': base,
                          'note': 'host CPU only (rank 0); n_gpus names the launch this line pairs with',
                          'e2e': {'value': base['value'], 'unit': 'images/sec', 'h2d_bytes_per_step': 0,
                                  'd2h_bytes_per_step': 0}}))
        return

    import torch
    import torch.distributed as dist

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (there is no CPU path)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    main_res = run_config(args.config, cfg, args, rank, world, local_rank, args.steps, args.warmup, True)
    if os.environ.get('PGK_BENCH_MAIN_ONLY'):
        return
    # ---- the other BASELINE configs on the same line: c4 is the north star's headline (depth 8), c3 / c5 the other
    # bf16 configs; a few device-timed steps each, same legs (value, e2e, roofline; D step alone at depth 8) ----------
    others = {}
    if not args.no_extras and args.config == 'c2' and not args.batch:
        for nm in (('c4', 'c3', 'c5') if world == 1 else ('c4',)):
            res_o = run_config(nm, dict(CONFIGS[nm]), args, rank, world, local_rank, min(args.steps, 10), 3, False)
            if res_o is not None:
                others[nm] = res_o
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    out = {
        'metric': 'images/sec (G+D+GP step)', 'value': main_res['value'], 'unit': 'images/sec', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': main_res['ms_per_step'], 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': main_res['dtype'], 'data': 'synthetic',
        'config': dict(workload, l2='no flush needed: activations written per step (%.1f GB peak allocated) >> 126 MB L2'
                       % main_res['peak_mem_gb']),
        'e2e': main_res['e2e'], 'gpu_launches': main_res['gpu_launches'], 'clocks': main_res['clocks'],
        'roofline': main_res['roofline'],
    }
    if 'd_step' in main_res:
        out['d_step'] = main_res['d_step']
    if others:
        out['configs'] = others
    switches = {k: v for k, v in sorted(os.environ.items()) if k.startswith('PGK_')}
    if switches or args.graphs or not args.prefetch:
        # a line measured with non-default tuning switches says so (A/B runs; the driver's run has none)
        out['config']['switches'] = dict(switches, **({'--graphs': '1'} if args.graphs else {}),
                                         **({'--no-prefetch': '1'} if not args.prefetch else {}))
    if world == 1 and not args.no_extras:
        # the bar to beat: the reference itself in PyTorch eager on this GPU, for the line's config and the headline
        try:
            ge = {args.config: gpu_eager_reference(cfg)}
            if 'c4' in others:
                ge['c4'] = gpu_eager_reference(dict(CONFIGS['c4']), steps=2, warmup=1)
            if ge[args.config] is not None:
                out['gpu_eager_reference'] = ge
        except Exception as e:           # a baseline leg must never take the line down
            out['gpu_eager_reference'] = {'error': repr(e)[:300]}
    if not args.no_cpu_baseline and world == 1:      # the CPU leg is reported at N = 1 only
        torch.cuda.synchronize()
        base = cpu_baseline_subprocess(args.config, args.batch)
        if base is not None:
            out['cpu_baseline'] = base
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline_subprocess(config, batch):
    """The CPU leg runs in a child process: loading the reference for the CPU replaces torch's .cuda() methods
    process-wide (baseline/ref_harness.py), which must not happen in the process that drives the GPU."""
    cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--config', config, '--steps', '3',
           '--warmup', '1']
    if batch:
        cmd += ['--batch', str(batch)]
    try:
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=240,
                           env=dict(os.environ, CUDA_VISIBLE_DEVICES='', PGK_CPU_BUDGET_S='20'))
        line = [l for l in p.stdout.splitlines() if l.startswith('{')][-1]
        return json.loads(line)['cpu_baseline']
    except Exception:
        return None


if __name__ == '__main__':
    main()
