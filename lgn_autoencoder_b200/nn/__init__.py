"""NN building blocks (API of the reference's lgn/nn/__init__.py:1-10; the classes the LGAE never instantiates --
BasicMLP, InputLinear, InputMPNN, MaskLevel -- are out of scope)."""
from .g_nn import CatMixReps, CatReps, MixReps
from .generic_levels import get_activation_fn
from .position_levels import RadialFilters, RadPolyTrig

__all__ = ["MixReps", "CatReps", "CatMixReps", "RadialFilters", "RadPolyTrig", "get_activation_fn"]
