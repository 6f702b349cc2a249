"""Per-irrep complex linear layers (reference: lgn/nn/g_nn.py:7-282)."""
import torch
import torch.nn as nn

from ..cg_lib import CGModule
from ..g_lib import GTau, GVec, GWeight, GWeightDict, g_torch


class MixReps(CGModule):
    """out[(k,n)] = W_(k,n) . x[(k,n)] on the channel axis, W complex (2, C_out, C_in).

    Initialisation follows the reference (g_nn.py:69-93): re and im parts drawn from randn (or rand) and scaled
    by gain / max(2, C_out, C_in), with a further 10^-k for the diagonal irreps (k, k)."""

    def __init__(self, tau_in, tau_out, real=False, weight_init="randn", gain=1, device=None, dtype=torch.float64):
        super().__init__(device=device, dtype=dtype)
        tau_in = GTau({k: v for k, v in GTau(tau_in).items() if v})
        if isinstance(tau_out, int):
            tau_out = {k: tau_out for k in tau_in.keys()}
        self.tau_in, self.tau_out = tau_in, GTau(tau_out)
        self.real, self.weight_init = real, weight_init
        if weight_init == "randn":
            w = GWeight.randn(self.tau_in, self.tau_out, device=self.device, dtype=self.dtype)
        elif weight_init == "rand":
            w = GWeight.rand(self.tau_in, self.tau_out, device=self.device, dtype=self.dtype)
        else:
            raise NotImplementedError(f"weight_init can only be 'randn' or 'rand'; other choices are not implemented yet ({weight_init})!")
        self.weights = GWeightDict()
        for key, val in w.items():
            scale = gain / max(val.shape) / (10 ** key[0] if key[0] == key[1] else 1)
            self.weights[key] = nn.Parameter(val * scale)

    def forward(self, reps):
        if isinstance(reps, dict):
            reps = GVec(reps)
        if GTau.from_rep(reps) != self.tau_in:
            raise ValueError(f"Tau of input reps, {GTau.from_rep(reps)}, does not match initialized tau, {self.tau_in}!")
        return g_torch.mix(self.weights, reps, key_order=self.weights.keys())

    @property
    def tau(self):
        return self.tau_out


class CatReps(nn.Module):
    """Truncate to max(key) < maxdim and concatenate on the channel axis (reference g_nn.py:124-193)."""

    def __init__(self, taus_in, maxdim=None):
        super().__init__()
        self.taus_in = taus_in = [GTau(t) for t in taus_in if t]
        if maxdim is None:
            maxdim = max(t.maxdim for t in taus_in)
        self.maxdim = maxdim
        self.taus_in = [GTau({k: v for k, v in t.items() if max(k) < maxdim}) for t in taus_in]
        self.tau_out = GTau.cat(self.taus_in)
        self.all_keys = list(self.tau_out.keys())

    def forward(self, reps):
        reps = [r for r in reps if r is not None and len(r) > 0]
        reps = [r.truncate(self.maxdim) for r in reps]
        taus = [GTau.from_rep(r) for r in reps]
        if taus != self.taus_in:
            raise ValueError(f"Tau of input reps does not match predefined version! got: {taus} expected: {self.taus_in}")
        return g_torch.cat(reps)

    @property
    def tau(self):
        return self.tau_out


class CatMixReps(CGModule):
    """CatReps followed by MixReps (reference g_nn.py:196-282)."""

    def __init__(self, taus_in, tau_out, maxdim=None, real=False, weight_init="randn", gain=1, device=None, dtype=torch.float64):
        super().__init__(device=device, dtype=dtype)
        self.cat_reps = CatReps(taus_in, maxdim=maxdim)
        self.mix_reps = MixReps(self.cat_reps.tau, tau_out, real=real, weight_init=weight_init, gain=gain, device=self.device, dtype=self.dtype)
        self.taus_in = taus_in
        self.taus_out = GTau(self.mix_reps.tau)

    def forward(self, reps_in):
        return self.mix_reps(self.cat_reps(reps_in))

    @property
    def tau(self):
        return self.taus_out
