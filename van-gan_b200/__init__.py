"""van-gan_b200: B200-native (sm_100a) drop-in for VAN-GAN's volumetric train-step and
sliding-window hot path.  Import as `van_gan_b200` (the hyphenated directory holds the sources;
`van_gan_b200/__init__.py` is the import shim)."""
from . import _lib  # noqa: F401  (raises on first use if libvangan_b200.so is missing)

__all__ = ["_lib"]
