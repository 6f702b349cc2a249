"""One LGN message-passing level and the scalar MLP (reference: lgn/models/lgn_levels.py:9-241).

Inside LGNEncoder / LGNDecoder these modules only own the parameters: the forward and backward of a whole level
run in csrc/lgae_level.cu and csrc/lgae_mlp.cu.  Their own ``forward`` is the generic layer-level composite used
for stand-alone calls and for configurations outside the fused path."""
import torch
import torch.nn as nn

from .. import layer_ops
from ..cg_lib import CGProduct
from ..g_lib import GTau
from ..nn import CatMixReps, get_activation_fn


class LGNNodeLevel(nn.Module):
    def __init__(self, tau_in, tau_pos, maxdim, num_channels, level_gain, weight_init, device=None, dtype=torch.float64, cg_dict=None):
        super().__init__()
        if device is None:
            device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.maxdim, self.num_channels = maxdim, num_channels
        self.tau_in, self.tau_pos = tau_in, tau_pos
        self.cg_power = CGProduct(tau_in, tau_in, maxdim=maxdim, device=device, dtype=dtype, cg_dict=cg_dict)
        self.cg_aggregate = CGProduct(tau_in, tau_pos, maxdim=maxdim, aggregate=True, device=device, dtype=dtype, cg_dict=cg_dict)
        tau_sq, tau_ag = self.cg_power.tau_out, self.cg_aggregate.tau_out
        self.cat_mix = CatMixReps([tau_ag, tau_in, tau_sq], num_channels, maxdim=maxdim, weight_init=weight_init, gain=level_gain,
                                  device=device, dtype=dtype)
        self.tau_out = self.cat_mix.taus_out

    def forward(self, node_feature, edge_feature, mask=None):
        """ag = CG_aggregate(node, edge); sq = CG(node, node); out = Mix(cat([ag, node, sq])).  ``mask`` is accepted and
        ignored, as in the reference."""
        reps_ag = self.cg_aggregate(node_feature, edge_feature)
        reps_sq = self.cg_power(node_feature, node_feature)
        return self.cat_mix([reps_ag, node_feature, reps_sq])


class CGMLP(nn.Module):
    """MLP on the (0,0) part: (2,B,N,C,1) -> rows of 2C interleaved (re, im) -> Linear/activation stack -> back.
    Like the reference it pops (0,0) from its input and re-inserts it, i.e. it mutates the GVec and moves (0,0)
    to the end of the part order."""

    def __init__(self, tau, num_hidden=3, layer_width_mul=2, activation="sigmoid", device=None, dtype=torch.float64):
        super().__init__()
        if device is None:
            device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.tau = tau
        self.num_scalars = 2 * GTau(tau)[(0, 0)]
        width = layer_width_mul * self.num_scalars
        self.num_hidden, self.layer_width = num_hidden, width
        self.linear = nn.ModuleList()
        if num_hidden > 0:
            self.linear.append(nn.Linear(self.num_scalars, width))
            for _ in range(num_hidden - 1):
                self.linear.append(nn.Linear(width, width))
            self.linear.append(nn.Linear(width, self.num_scalars))
        else:
            self.linear.append(nn.Linear(self.num_scalars, self.num_scalars))
        act = get_activation_fn(activation)
        self.activations = nn.ModuleList([act for _ in range(num_hidden)])
        self.to(device=device, dtype=dtype)

    def forward(self, node_feature_in, mask=None):
        out = node_feature_in
        x = out.pop((0, 0)).squeeze(-1)
        s = x.shape
        x = x.permute(1, 2, 3, 0).contiguous().view(s[1:3] + (self.num_scalars,))
        # every layer is one launch of the tiled fp64 GEMM kernel (csrc/lgae_layers.cu); LeakyReLU is fused into it
        for lin, act in zip(self.linear, self.activations):
            if isinstance(act, nn.LeakyReLU):
                x = layer_ops.linear(x, lin.weight, lin.bias, leaky_slope=act.negative_slope)
            else:
                x = act(layer_ops.linear(x, lin.weight, lin.bias))
        x = layer_ops.linear(x, self.linear[-1].weight, self.linear[-1].bias)
        if mask is not None:
            x = torch.where(mask, x, torch.zeros((), dtype=x.dtype, device=x.device))
        out[(0, 0)] = x.view(s[1:] + (2,)).permute(3, 0, 1, 2).unsqueeze(-1)
        return out

    def scale_weights(self, scale):
        self.linear[-1].weight.data *= scale
        if self.linear[-1].bias is not None:
            self.linear[-1].bias.data *= scale
