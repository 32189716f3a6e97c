"""Development aid (GPU box): the "GPU reference bar" of SURVEY.md 8(d) -- the reference's own arithmetic (the oracle:
torch library ops, autograd double backward) run in PyTorch eager mode ON THE GPU, fp32 and TF32-allowed, for the
bench configs.  This is the number the hand-written kernels have to beat; it is NOT part of bench.py's contract (the
reference arm there is the host-CPU path) and nothing in the product imports this file.

    python tests/dev/gpu_eager_bar.py [c1 c2 c3 c4 c5] [--steps 5] [--warmup 2] [--batch N]

Prints one JSON line per (config, precision): images/sec of D step + Adam + G step + Adam with inputs resident in HBM.
Batches that do not fit (cuDNN workspace at depth 8) are halved until they do and the batch used is reported.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'baseline'))
import bench  # noqa: E402
import pggan_oracle as O  # noqa: E402
import ref_harness as R  # noqa: E402


def run(cfg, steps, warmup, tf32, n, device='cuda'):
    if R.available():      # the unmodified reference's own Trainer.train() (baseline/_ref)
        ips, ms, _, _ = R.time_train(cfg['res'], cfg['ch'], cfg['depth'], cfg['alpha'], n, steps, warmup, device, tf32)
        return ips, ms
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True
    dev = torch.device(device)
    cuda = dev.type == 'cuda'
    depth, alpha, ch, res = cfg['depth'], cfg['alpha'], cfg['ch'], cfg['res']
    to = lambda d: {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in d.items()}
    pgp, pdp = to(O.make_generator_params(res, ch, seed=1337)), to(O.make_discriminator_params(res, ch, seed=1338))
    nb = O.n_blocks_for(res)
    gen = torch.Generator(device=dev).manual_seed(1337)
    r = 4 * 2 ** depth
    real = torch.randn(n, ch, r, r, device=dev, generator=gen)
    z1, z2 = torch.randn(n, 512, device=dev, generator=gen), torch.randn(n, 512, device=dev, generator=gen)
    mix = torch.rand(n, 1, device=dev, generator=gen)
    sd, sg = {}, {}
    state = {'g': pgp, 'd': pdp}

    def step():
        _, _, _, gd = O.d_step_grads(state['d'], state['g'], real, z1, mix, depth, alpha, nb)
        state['d'] = O.adam_step(dict(state['d']), gd, sd, 1e-3)
        _, gg = O.g_step_grads(state['g'], state['d'], z2, depth, alpha, nb)
        state['g'] = O.adam_step(dict(state['g']), gg, sg, 1e-3)

    for _ in range(warmup):
        step()
    if not cuda:     # --device cpu exists only to check this script where there is no GPU
        import time
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        ms = (time.perf_counter() - t0) * 1e3 / steps
        return n / (ms * 1e-3), ms
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return n / (ms * 1e-3), ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('configs', nargs='*', default=['c2', 'c3', 'c4', 'c5'])
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=2)
    ap.add_argument('--batch', type=int, default=0)
    ap.add_argument('--device', default='cuda')
    args = ap.parse_args()
    for name in args.configs:
        cfg = bench.CONFIGS[name]
        for tf32 in (False, True):
            n = args.batch or cfg['n']
            while True:
                try:
                    ips, ms = run(cfg, args.steps, args.warmup, tf32, n, args.device)
                    break
                except torch.OutOfMemoryError:
                    torch.cuda.empty_cache()
                    if n == 1:
                        ips, ms = float('nan'), float('nan')
                        break
                    n //= 2
            print(json.dumps({'tool': 'gpu_eager_bar', 'config': name, 'depth': cfg['depth'], 'alpha': cfg['alpha'],
                              'batch': n, 'config_batch': cfg['n'], 'precision': 'tf32' if tf32 else 'fp32',
                              'images_per_sec': ips, 'ms_per_step': ms,
                              'what': ("the unmodified reference's Trainer.train() (baseline/_ref) in PyTorch eager on the GPU"
                                       if R.available() else
                                       "the reference's arithmetic (oracle) in PyTorch eager on the GPU")}), flush=True)
            torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
