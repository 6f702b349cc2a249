"""Learnable per-irrep complex weights.  Plays the role of the reference's ParameterDictNew
(lgn/g_lib/parameter_dict_new.py:4-18): parameters are registered under ``str((k, n))`` so the state-dict
entries read ``...weights.(0, 0)``, and ``keys()`` iterates a *set* of tuple keys, whose order fixes the part
order of every MixReps output.  Unlike the reference's class, items()/values()/repr work
under torch >= 2."""
import torch.nn as nn


class GWeightDict(nn.Module):
    def __init__(self, weights=None):
        super().__init__()
        self._order = []
        if weights is not None:
            for key, val in dict(weights).items():
                self[key] = val

    def __setitem__(self, key, param):
        if not isinstance(param, nn.Parameter):
            param = nn.Parameter(param)
        if key not in self._order:
            self._order.append(key)
        self.register_parameter(str(key), param)

    def __getitem__(self, key):
        return self._parameters[str(key)]

    def __contains__(self, key):
        return str(key) in self._parameters

    def __len__(self):
        return len(self._parameters)

    def keys(self):
        """Tuple keys in the iteration order of a set (as the reference's ``set(map(eval, ...))``)."""
        return list(set(self._order))

    def registration_keys(self):
        return list(self._order)

    def values(self):
        return [self[k] for k in self.keys()]

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def __iter__(self):
        return iter(self.keys())

    def extra_repr(self):
        return ", ".join(f"{k}: {tuple(self[k].shape)}" for k in self._order)


# name kept for drop-in imports
ParameterDictNew = GWeightDict
