"""Learnable radial functions of the pairwise Minkowski norms (reference: lgn/nn/position_levels.py:7-310)."""
import torch
import torch.nn as nn

from ..g_lib import GScalar


class _MaskedOutRadialFn(torch.autograd.Function):
    """RadPolyTrig.forward with an all-zero edge mask (the decoder, lgn_decoder.py:335-340): every bell is replaced by 0, so
    each output is exactly its Linear's bias, broadcast over the edges.  Same values and gradients as the general kernel (the
    bias gradient is the sum over the edges; a, b, c and the Linear weights get exact zeros) without
    evaluating 2K reciprocals per edge in either direction."""

    @staticmethod
    def forward(ctx, shape, n_l, a, b, c, *wb):
        ctx.meta = (tuple(a.shape), [tuple(wb[2 * l].shape) for l in range(n_l)], n_l)
        return tuple(wb[2 * l + 1].detach().expand(*shape, -1).contiguous() for l in range(n_l))

    @staticmethod
    def backward(ctx, *gs):
        ashape, wshapes, n_l = ctx.meta
        g0 = gs[0]
        zeros = lambda shp: torch.zeros(shp, dtype=g0.dtype, device=g0.device)
        out = [None, None, zeros(ashape), zeros(ashape), zeros(ashape)]
        for l in range(n_l):
            out += [zeros(wshapes[l]), gs[l].reshape(-1, gs[l].shape[-1]).sum(0)]
        return tuple(out)


class RadPolyTrig(nn.Module):
    """phi_k(n) = b_k / (1 + (c_k n)^2 + 1e-16) + a_k for 2*num_basis_fn bells, zeroed on masked edges, followed by
    one Linear per zonal degree l.  Cartesian input: Linear(2K' -> 2C), outputs (2c, 2c+1) = (re, im) of channel c.
    Canonical (complex) input: the norms carry a leading complex axis and the same real Linear(2K' -> C) is
    applied to its two slices."""

    def __init__(self, max_zf, num_basis_fn, num_channels, mix=True, input_basis="cartesian", device=None, dtype=torch.float64):
        super().__init__()
        if device is None:
            device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        if input_basis.lower() not in ("cartesian", "canonical"):
            raise ValueError("Input basis can only be 'cartesian' or 'canonical'!")
        self.max_zf, self.num_basis_fn, self.num_channels, self.input_basis, self.mix = max_zf, num_basis_fn, num_channels, input_basis, mix
        self.basis_size = (1, 1, 1, 2 * num_basis_fn)
        self.a = nn.Parameter(torch.randn(self.basis_size).to(device=device, dtype=dtype))
        self.b = nn.Parameter(torch.randn(self.basis_size).to(device=device, dtype=dtype))
        self.c = nn.Parameter(torch.randn(self.basis_size).to(device=device, dtype=dtype))
        if mix == "cplx" or mix is True:
            out = num_channels if input_basis == "canonical" else 2 * num_channels
            self.linear = nn.ModuleList([nn.Linear(2 * num_basis_fn, out).to(device=device, dtype=dtype) for _ in range(max_zf + 1)])
            self.radial_types = (num_channels,) * max_zf
        elif mix == "real":
            self.linear = nn.ModuleList([nn.Linear(2 * num_basis_fn, num_channels).to(device=device, dtype=dtype) for _ in range(max_zf + 1)])
            self.radial_types = (num_channels,) * max_zf
        elif mix == "none" or mix is False:
            self.linear = None
            self.radial_types = (num_basis_fn,) * max_zf
        else:
            raise ValueError(f"Can only specify mix = real, cplx, or none: {mix}")
        self.device = device

    def forward_masked_out(self, shape):
        """The radial functions of ``norms`` of the given shape when EVERY edge is masked out (canonical basis, mix 'cplx')."""
        if not (self.mix == "cplx" or self.mix is True) or self.input_basis != "canonical":
            raise NotImplementedError("forward_masked_out covers the decoder's configuration (canonical basis, complex mixing)")
        wb = []
        for lin in self.linear:
            wb += [lin.weight, lin.bias]
        outs = _MaskedOutRadialFn.apply(tuple(shape), len(self.linear), self.a, self.b, self.c, *wb)
        return GScalar({(l, l): o for l, o in enumerate(outs)}, ignore_check=True)

    def forward(self, norms, edge_mask):
        if self.mix == "cplx" or self.mix is True:
            # one kernel for the bells, the mask and the Linear of every zonal degree (csrc/lgae_layers.cu)
            from .. import layer_ops
            outs = layer_ops.radial_functions(norms, edge_mask, self.a, self.b, self.c, [(lin.weight, lin.bias) for lin in self.linear],
                                              planar=self.input_basis != "canonical")
            return GScalar({(l, l): o for l, o in enumerate(outs)}, ignore_check=True)
        s = tuple(norms.shape)
        mask = (edge_mask != 0).unsqueeze(-1)
        x = norms.unsqueeze(-1)
        bells = self.b * (torch.ones_like(self.b) + (self.c * x).pow(2) + 1e-16).pow(-1) + self.a
        bells = torch.where(mask, bells, torch.zeros((), dtype=norms.dtype, device=norms.device))
        bells = bells.view(s + (1, 2 * self.num_basis_fn))
        if self.mix == "cplx" or self.mix is True:
            if self.input_basis == "canonical":
                parts = {(l, l): lin(bells).view(s + (self.num_channels,)) for l, lin in enumerate(self.linear)}
            else:
                nd = len(s)
                parts = {(l, l): lin(bells).view(s + (self.num_channels, 2)).permute(nd + 1, *range(nd + 1)) for l, lin in enumerate(self.linear)}
        elif self.mix == "real":
            r = {(l, l): lin(bells).view(s + (self.num_channels,)) for l, lin in enumerate(self.linear)}
            parts = {k: torch.stack((v, torch.zeros_like(v)), 0) for k, v in r.items()}
        else:
            b2 = bells.view(s + (self.num_basis_fn, 2)).permute(len(s) + 1, *range(len(s) + 1))
            parts = {(l, l): b2 for l in range(self.max_zf + 1)}
        return GScalar(parts, ignore_check=True)


class RadialFilters(nn.Module):
    """One RadPolyTrig per LGN level."""

    def __init__(self, max_zf, num_basis_fn, num_channels_out, num_levels, mix=True, input_basis="cartesian", device=None, dtype=torch.float64):
        super().__init__()
        if device is None:
            device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.num_levels, self.max_zf = num_levels, max_zf
        self.rad_funcs = nn.ModuleList([RadPolyTrig(max_zf[lvl], num_basis_fn, num_channels_out[lvl], mix=mix, input_basis=input_basis,
                                                    device=device, dtype=dtype) for lvl in range(num_levels)])
        self.tau = [{(l, l): rf.radial_types[l - 1] for l in range(0, mz + 1)} for rf, mz in zip(self.rad_funcs, max_zf)]
        self.num_rad_channels = self.tau[0][(1, 1)] if len(self.tau) > 0 else 0

    def forward(self, norms, base_mask, basis="cartesian"):
        return [rf(norms, base_mask) for rf in self.rad_funcs]

    def forward_masked_out(self, shape):
        return [rf.forward_masked_out(shape) for rf in self.rad_funcs]
