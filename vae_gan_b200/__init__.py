"""vae_gan_b200 - B200-native (sm_100a) VAE-GAN training step behind the reference notebook's
nn.Module surface.  See DESIGN.md / INTEGRATION.md."""
from . import functional
from .functional import compute_dtype, config, rng
from .modules import (Decoder, Discriminator, Encoder, ResBlockDiscriminator, ResBlockVAE,
                      SpatialVAECodeProcessor, UnsupervisedGeneratorNetwork, build_vae_gan, init_weights)
from .train import VaeGanTrainer
from .data import InputPipeline, normalize_images
from ._lib import is_deterministic, set_deterministic

__all__ = ["ResBlockVAE", "Encoder", "Decoder", "ResBlockDiscriminator", "Discriminator",
           "SpatialVAECodeProcessor", "UnsupervisedGeneratorNetwork", "init_weights", "build_vae_gan",
           "VaeGanTrainer", "InputPipeline", "normalize_images", "set_deterministic", "is_deterministic", "functional", "compute_dtype", "config", "rng"]
