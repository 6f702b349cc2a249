"""Containers and complex helpers (API of the reference's lgn/g_lib/__init__.py:1-22)."""
from . import cplx_lib, g_torch, rotations
from .g_tau import GTau
from .g_tensor import GScalar, GTensor, GVec, GWeight
from .g_torch import add, cat, cat_mix, div, mix, mul, sub
from .weight_dict import GWeightDict, ParameterDictNew

__all__ = ["GTau", "GTensor", "GVec", "GScalar", "GWeight", "GWeightDict", "ParameterDictNew", "cplx_lib", "g_torch", "rotations",
           "add", "sub", "mul", "div", "cat", "mix", "cat_mix"]
