"""Host helpers on the path (reference utils.py:56-57)."""
import numpy as np
import torch


def random_latents(num_latents, latent_size):
    """Gaussian latents from numpy's global RNG as a CPU fp32 tensor (utils.py:56-57)."""
    return torch.from_numpy(np.random.randn(num_latents, latent_size).astype(np.float32))
