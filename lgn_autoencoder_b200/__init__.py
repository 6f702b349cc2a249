"""B200-native LGAE hot path: hand-written sm_100a kernels behind the reference's ``lgn`` module API."""
__version__ = "0.1.0"
