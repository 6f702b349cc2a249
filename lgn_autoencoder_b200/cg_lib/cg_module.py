"""CGModule: nn.Module that carries a (shared) CGDict plus device/dtype (reference: lgn/cg_lib/cg_module.py:7-212).
fp64 only, like the reference (cg_module.py:62-73)."""
import torch
import torch.nn as nn

from .cg_dict import CGDict


class CGModule(nn.Module):
    def __init__(self, cg_dict=None, maxdim=None, device=None, dtype=None):
        super().__init__()
        if device is None:
            device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        if dtype is None:
            dtype = torch.float64
        if dtype != torch.float64:
            raise ValueError(f"CG Module only takes float64 (the reference and the sm_100a kernels are fp64 only), got {dtype}")
        self._device, self._dtype = torch.device(device), dtype
        self._init_cg_dict(cg_dict, maxdim)

    def _init_cg_dict(self, cg_dict, maxdim):
        if cg_dict is None and maxdim is None:
            self._cg_dict, self._maxdim = None, None
            return
        if cg_dict is None:
            cg_dict = CGDict(maxdim=maxdim, transpose=True, dtype=self._dtype, device=self._device)
        else:
            if cg_dict.dtype != self._dtype:
                raise ValueError(f"CGDict dtype {cg_dict.dtype} does not match module dtype {self._dtype}")
            if cg_dict.device != self._device:
                cg_dict.to(device=self._device)
            if maxdim is not None:
                cg_dict.update_maxdim(maxdim)
        self._cg_dict = cg_dict
        self._maxdim = maxdim if maxdim is not None else cg_dict.maxdim

    @property
    def device(self):
        return self._device

    @property
    def dtype(self):
        return self._dtype

    @property
    def maxdim(self):
        return self._maxdim

    @property
    def cg_dict(self):
        return self._cg_dict

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        probe = fn(torch.empty(0, dtype=self._dtype, device=self._device))
        if probe.dtype != torch.float64:
            raise ValueError("CG modules are fp64 only")
        self._device = probe.device
        if self._cg_dict is not None:
            self._cg_dict.to(device=self._device)
        return out
