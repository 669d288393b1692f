"""multiscale_variational_autoencoder_b200: the multiscale-VAE training step of
NikolasMarkou/multiscale_variational_autoencoder (mvae/multiscale_vae.py) on hand-written sm_100a kernels.

Exports the same three names as the reference package (mvae/__init__.py:7-15) plus `coord`."""

__version__ = "0.1.0"

from .vae import VAE
from . import layer_blocks
from . import coord
from .multiscale_vae import MultiscaleVAE

__all__ = ["VAE", "layer_blocks", "MultiscaleVAE", "coord"]
