"""Activation lookup (reference: lgn/nn/generic_levels.py:119-135)."""
import torch.nn as nn


def get_activation_fn(activation):
    activation = activation.lower()
    table = {"leakyrelu": lambda: nn.LeakyReLU(), "relu": lambda: nn.ReLU(), "elu": lambda: nn.ELU(), "sigmoid": lambda: nn.Sigmoid(),
             "logsigmoid": lambda: nn.LogSigmoid(), "tanh": lambda: nn.Tanh(), "atan": lambda: _Atan()}
    if activation not in table:
        raise ValueError(f"Activation function {activation} not implemented!")
    return table[activation]()


class _Atan(nn.Module):
    def forward(self, x):
        import torch
        return torch.atan(x)
