"""Host helpers on the path (reference utils.py:56-57)."""
import numpy as np
import torch


def random_latents(num_latents, latent_size):
    """Gaussian latents from numpy's global RNG as a CPU fp32 tensor (utils.py:56-57)."""
    return torch.from_numpy(np.random.randn(num_latents, latent_size).astype(np.float32))


def device_random_latents(num_latents, latent_size, device='cuda', seed=None):
    """Opt-in replacement for `partial(random_latents, n, latent_size)` (train.py:161-163, the `create_rlg` handed to
    DepthManager): returns a callable that draws the Gaussian latents ON THE DEVICE (torch generator), so that
    Trainer.train() (trainer.py:86,103) has no host-side draw and no host-to-device copy per step.  The default stays
    `random_latents`: the reference draws from numpy's global generator, and that order of draws is observable."""
    gen = torch.Generator(device=device)
    if seed is not None:
        gen.manual_seed(seed)

    def draw():
        return torch.randn(num_latents, latent_size, device=device, generator=gen)
    return draw
