"""Stack of LGN levels (reference: lgn/models/lgn_cg.py:8-180)."""
import torch.nn as nn

from ..cg_lib import CGModule
from .lgn_levels import CGMLP, LGNNodeLevel


class LGNCG(CGModule):
    def __init__(self, maxdim, max_zf, tau_in, tau_pos, num_cg_levels, num_channels, level_gain, weight_init, mlp=True, mlp_depth=None,
                 mlp_width=None, activation="leakyrelu", device=None, dtype=None, cg_dict=None):
        super().__init__(device=device, dtype=dtype, cg_dict=cg_dict)
        device, dtype, cg_dict = self.device, self.dtype, self.cg_dict
        self.max_zf, self.mlp = max_zf, mlp
        tau_node_in = tau_in.tau if isinstance(tau_in, CGModule) else tau_in
        self.node_levels = nn.ModuleList()
        if mlp:
            self.mlp_levels = nn.ModuleList()
        tau_node = tau_node_in
        for layer in range(num_cg_levels):
            lvl = LGNNodeLevel(tau_node, tau_pos[layer], maxdim[layer], num_channels[layer + 1], level_gain[layer], weight_init,
                               device=device, dtype=dtype, cg_dict=cg_dict)
            self.node_levels.append(lvl)
            if mlp:
                self.mlp_levels.append(CGMLP(lvl.tau_out, activation=activation, num_hidden=mlp_depth, layer_width_mul=mlp_width,
                                             device=device, dtype=dtype))
            tau_node = lvl.tau_out
        self.tau_levels_node = [tau_node_in] + [lvl.tau_out for lvl in self.node_levels]

    def forward(self, node_feature, node_mask, rad_funcs, zonal_functions):
        if len(self.node_levels) != len(rad_funcs):
            raise ValueError(f"The number of layer ({len(self.node_levels)}) and the number of available radial functions "
                             f"({len(rad_funcs)}) are not equal!")
        nodes_features = [node_feature]
        for idx, node_level in enumerate(self.node_levels):
            edge = rad_funcs[idx] * zonal_functions
            node_feature = node_level(node_feature, edge, node_mask)
            if self.mlp:
                node_feature = self.mlp_levels[idx](node_feature)
            nodes_features.append(node_feature)
        return nodes_features
