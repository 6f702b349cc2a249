"""Mixin that keeps all parameters of a model as views of one flat fp64 buffer (what the kernels read) while
preserving the reference's per-tensor ``nn.Parameter`` names and shapes."""
from collections import OrderedDict

import torch

from .. import fused


class FusedParamsMixin:
    _plan = None
    _theta = None

    def _build_plan(self, kind, **geometry):
        shapes = OrderedDict((name, tuple(p.shape)) for name, p in self.named_parameters())
        self._plan_kind, self._plan_geometry = kind, geometry
        self._plan = fused.FusedPlan(kind, shapes, **geometry)
        self._theta = None

    def _flat_params(self):
        """The flat buffer, rebuilt (and the parameters re-pointed into it) whenever a parameter no longer aliases
        it, e.g. after ``.to(device)``; ``load_state_dict`` and optimizer steps write in place and keep the aliasing."""
        params = OrderedDict(self.named_parameters())
        theta = self._theta
        ok = theta is not None
        if ok:
            base = theta.data_ptr()
            for name, (off, n, _) in self._plan.offsets.items():
                p = params[name]
                if p.data_ptr() != base + 8 * off or p.device != theta.device or not p.is_contiguous():
                    ok = False
                    break
        if not ok:
            device = next(iter(params.values())).device
            theta = self._plan.flatten({k: v.data for k, v in params.items()}, device)
            for name, view in self._plan.views(theta).items():
                params[name].data = view
            self._theta = theta
        return theta, list(params.values())

    def _theta_node(self):
        """The flat parameter buffer as one autograd tensor (``fused._FlatParamsFn``)."""
        theta, params = self._flat_params()
        return fused._FlatParamsFn.apply(self._plan, theta, *params)

    def l1_norm(self):
        """sum |p| over all parameters (lgn_encoder.py:249-250), evaluated on the flat buffer in one reduction."""
        if self._plan is None:
            return sum(p.abs().sum() for p in self.parameters())
        return _FlatNormFn.apply(self._theta_node(), 1)

    def l2_norm(self):
        """sum p^2 over all parameters (lgn_encoder.py:252-253)."""
        if self._plan is None:
            return sum(torch.pow(p, 2).sum() for p in self.parameters())
        return _FlatNormFn.apply(self._theta_node(), 2)


class _FlatNormFn(torch.autograd.Function):
    """L1 / squared-L2 norm of all parameters through the flat buffer they alias: two launches instead of two per tensor."""

    @staticmethod
    def forward(ctx, theta, order):
        ctx.order = order
        ctx.save_for_backward(theta)
        return theta.abs().sum() if order == 1 else (theta * theta).sum()

    @staticmethod
    def backward(ctx, g):
        (theta,) = ctx.saved_tensors
        return (torch.sign(theta) * g if ctx.order == 1 else 2.0 * theta * g), None
