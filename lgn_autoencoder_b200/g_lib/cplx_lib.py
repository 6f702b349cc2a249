"""Complex arithmetic on planar (2, ...) tensors (reference: lgn/g_lib/cplx_lib.py:7-72).  These are the
generic building blocks of the layer-level API; the fused sm_100a path never materialises them."""
import torch


def _c(x):
    return torch.complex(x[0], x[1])


def _p(z):
    return torch.stack((z.real, z.imag), 0)


def mix_zweight_zvec(weight, part, zdim=0):
    """(2,C',C) complex weight applied on the channel axis of a (2,...,C,d) part (kernel: csrc/lgae_cg.cu)."""
    from .. import layer_ops
    return layer_ops.mix(weight, part)


def mix_zweight_zscalar(weight, part, zdim=0):
    """(2,C',C) applied on the channel axis of a (2,...,C) part."""
    from .. import layer_ops
    return layer_ops.mix(weight, part.unsqueeze(-1)).squeeze(-1)


def mul_zscalar_zirrep(scalar, part, zdim=0):
    """(2,...,C) x (2,...,C,d) -> (2,...,C,d) (kernel: csrc/lgae_layers.cu)."""
    from .. import layer_ops
    return layer_ops.scalar_irrep(scalar, part)


def mul_zscalar_zscalar(s1, s2, zdim=0):
    return _p(_c(s1) * _c(s2))
