"""Element-wise and structural operations on G containers (reference: lgn/g_lib/g_torch.py:18-300)."""
import torch

from . import cplx_lib
from .g_tensor import GScalar, GTensor, GVec, GWeight


def _binary(op, a, b):
    if isinstance(a, GTensor) and isinstance(b, GTensor):
        if set(a.keys()) != set(b.keys()):
            raise ValueError(f"operands have different irreps: {list(a.keys())} vs {list(b.keys())}")
        return type(a)({k: op(a[k], b[k]) for k in a.keys()}, ignore_check=True)
    if isinstance(a, GTensor):
        return type(a)({k: op(v, b) for k, v in a.items()}, ignore_check=True)
    if isinstance(b, GTensor):
        return type(b)({k: op(a, v) for k, v in b.items()}, ignore_check=True)
    return op(a, b)


def add(a, b):
    return _binary(torch.add, a, b)


def sub(a, b):
    return _binary(torch.sub, a, b)


def div(a, b):
    return _binary(torch.div, a, b)


def mul(a, b):
    """Complex product GScalar x GVec (edge features, lgn/models/lgn_cg.py:167), GScalar x GScalar, or a plain
    element-wise product with a real number / tensor."""
    if isinstance(a, GVec) and isinstance(b, GScalar):
        a, b = b, a
    if isinstance(a, GScalar) and isinstance(b, GVec):
        if set(a.keys()) != set(b.keys()):
            raise ValueError(f"operands have different irreps: {list(a.keys())} vs {list(b.keys())}")
        return GVec({k: cplx_lib.mul_zscalar_zirrep(a[k], b[k]) for k in a.keys()}, ignore_check=True)
    if isinstance(a, GScalar) and isinstance(b, GScalar):
        return GScalar({k: cplx_lib.mul_zscalar_zscalar(a[k], b[k]) for k in a.keys()}, ignore_check=True)
    if isinstance(a, GTensor) and isinstance(b, GTensor):
        raise ValueError(f"cannot multiply {type(a).__name__} with {type(b).__name__}")
    return _binary(torch.mul, a, b)


def cat(reps_list):
    """Concatenate on the channel axis; irreps are visited in the iteration order of a set union, which is
    what fixes the part order downstream (reference g_torch.py:190-214)."""
    reps_list = [r for r in reps_list if r is not None and len(r) > 0]
    cls = type(reps_list[0])
    if not all(type(r) is cls for r in reps_list):
        raise ValueError("all reps must have the same type")
    all_keys = set().union(*[set(r.keys()) for r in reps_list])
    return cls({k: torch.cat([r[k] for r in reps_list if k in r], dim=cls.cdim) for k in all_keys}, ignore_check=True)


def mix(weights, rep, key_order=None):
    """Per-irrep complex channel mixing W.x (reference g_torch.py:217-255)."""
    wkeys = set(weights.keys())
    if wkeys != set(rep.keys()):
        raise ValueError("Must have one mixing weight for each part of the rep")
    order = list(key_order) if key_order is not None else list(wkeys)
    if isinstance(rep, GVec):
        return GVec({k: cplx_lib.mix_zweight_zvec(weights[k], rep[k]) for k in order}, ignore_check=True)
    if isinstance(rep, GScalar):
        return GScalar({k: cplx_lib.mix_zweight_zscalar(weights[k], rep[k]) for k in order}, ignore_check=True)
    raise ValueError(f"cannot mix a {type(rep).__name__}")


def cat_mix(weights, reps_list):
    return mix(weights, cat(reps_list))
