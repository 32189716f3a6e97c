"""-m gpu, needs two GPUs (skipped otherwise): data-parallel correctness over NCCL (the exact DP oracle).

Two processes, one per GPU, `torch.distributed` backend nccl: each rank runs the D step and the G step of the CUDA path
on ITS OWN shard (reals, latents, mixing factors; minibatch-stddev statistics and the penalty stay per replica); the
loss functions all-reduce the flat gradient buffer once per step.  Every rank must end with the MEAN over the ranks of
the gradients the oracle computes for each rank's shard on its own -- and both ranks with the same bits.
"""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

from _util import rel_err

pytestmark = pytest.mark.gpu

CFG = dict(res=32, ch=3, fb=512, fm=64, lat=64, depth=3, alpha=0.5, n=4)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def shard_inputs(rank):
    c = CFG
    gen = torch.Generator().manual_seed(500 + rank)
    r = 4 * 2 ** c['depth']
    z1, z2 = torch.randn(c['n'], c['lat'], generator=gen), torch.randn(c['n'], c['lat'], generator=gen)
    real = torch.randn(c['n'], c['ch'], r, r, generator=gen)
    mix = torch.rand(c['n'], 1, generator=gen)
    return z1, z2, real, mix


def make_params():
    from _gpu_util import O
    c = CFG
    pgp = O.make_generator_params(c['res'], c['ch'], fmap_base=c['fb'], fmap_max=c['fm'], latent_size=c['lat'], seed=3)
    pdp = O.make_discriminator_params(c['res'], c['ch'], fmap_base=c['fb'], fmap_max=c['fm'], seed=4)
    return pgp, pdp


def _worker(rank, world, port, out):
    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from _gpu_util import build_pair, named_grads, pg
        c = CFG
        pgp, pdp = make_params()
        G, D = build_pair(dict(resolution=c['res'], channels=c['ch'], fmap_base=c['fb'], fmap_max=c['fm'],
                               latent=c['lat'], pg=pgp, pd=pdp, depth=c['depth'], alpha=c['alpha']),
                          device='cuda')
        z1, z2, real, mix = shard_inputs(rank)
        pg.wgan_gp_loss.mixing_factors_override = mix
        cost, _, _ = pg.wgan_gp_D_loss(D, G, real.cuda(), z1.cuda())
        cost.backward()
        pg.wgan_gp_loss.mixing_factors_override = None
        gd = named_grads(D)
        gcost = pg.wgan_gp_G_loss(G, D, z2.cuda())
        gcost.backward()
        gg = named_grads(G)
        torch.cuda.synchronize()
        out[rank] = (float(cost.detach()), float(gcost.detach()), gd, gg)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs (gpurun --gpus 2)')
def test_nccl_world2_gradients_are_the_mean_of_the_per_shard_oracle_gradients():
    from _gpu_util import O
    world, c = 2, CFG
    port = _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = {k: v for k, v in out.items()}
    pgp, pdp = make_params()
    nb = O.n_blocks_for(c['res'])
    gd_o, gg_o, costs = [], [], []
    for rank in range(world):
        z1, z2, real, mix = shard_inputs(rank)
        cost, _, _, gd = O.d_step_grads(pdp, pgp, real, z1, mix, c['depth'], c['alpha'], nb)
        gcost, gg = O.g_step_grads(pgp, pdp, z2, c['depth'], c['alpha'], nb)
        gd_o.append(gd), gg_o.append(gg), costs.append((float(cost), float(gcost)))
    mean = lambda ds: {k: sum(d[k] for d in ds) / len(ds) for k in ds[0]}
    gd_m, gg_m = mean(gd_o), mean(gg_o)
    for rank in range(world):
        cost, gcost, gd, gg = res[rank]
        # the losses stay per replica (each rank logs its own), the gradients are the averaged ones
        assert abs(cost - costs[rank][0]) < 1e-3 * abs(costs[rank][0]) and abs(gcost - costs[rank][1]) < 1e-3 * max(1e-3, abs(costs[rank][1]))
        assert set(gd) == set(gd_m) and set(gg) == set(gg_m)
        # (unscreened inputs: a LeakyReLU unit deciding differently than in the oracle costs ~1e-3 on a gradient tensor --
        # tests/test_gpu_baseline_widths.py; what is under test here is the averaging, which is exact or off by O(1))
        for k, v in gd_m.items():
            assert rel_err(gd[k], v) < 5e-3, (rank, k)
        for k, v in gg_m.items():
            assert rel_err(gg[k], v) < 5e-3, (rank, k)
    for k in res[0][2]:
        assert torch.equal(res[0][2][k], res[1][2][k]), k          # one all-reduce: every rank holds the same bits
    for k in res[0][3]:
        assert torch.equal(res[0][3][k], res[1][3][k]), k
