"""B200-native HoloGAN generator hot path (sm_100a kernels behind a C ABI, torch for plumbing).

Public surface mirrors the reference's `core.models.hologan_generator` /
`core.models.hologan_discriminator` modules; see `lightning_gan_zoo_b200.compat.install()` to
expose them under the reference's dotted paths for Hydra `_target_` strings.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "ops"]
__version__ = "0.1.0"
