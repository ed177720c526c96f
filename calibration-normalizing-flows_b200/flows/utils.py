"""Conditioner network container -- drop-in for the reference's ``flows/utils.py:6-31``.

``MLP`` only *holds* the parameters (same attribute names, shapes, init and state_dict keys
as the reference: ``layers.{j}.weight`` [out, in], ``layers.{j}.bias``).  Inside a coupling
layer the network is never evaluated module by module: the fused CUDA kernel reads the
packed weights.  ``forward`` exists for standalone use of the container (reference
``run_experiment3D.py:40`` builds a bare MLP as its "dnn" baseline, out of this path's scope).
"""
import math

import torch
from torch import nn
import torch.nn.functional as F


class MLP(nn.Module):
    def __init__(self, dim, hidden_size=[], activation=F.relu, wscale=1.):
        super().__init__()
        self.activation = activation
        widths = [dim, *hidden_size, dim]
        self.layers = nn.ModuleList()
        for fan_in, fan_out in zip(widths, widths[1:]):
            lin = nn.Linear(fan_in, fan_out)
            # nn.Linear default init, then weight AND bias scaled (reference flows/utils.py:19-21)
            with torch.no_grad():
                lin.weight.mul_(wscale)
                lin.bias.mul_(wscale)
            self.layers.append(lin)

    @property
    def hidden_size(self):
        return [lin.out_features for lin in list(self.layers)[:-1]]

    def canonical_parameters(self):
        out = []
        for lin in self.layers:
            out += [lin.weight, lin.bias]
        return out

    def forward(self, x):
        *hidden, last = list(self.layers)
        for lin in hidden:
            x = self.activation(lin(x))
        return last(x)


class TempScaler(nn.Module):
    """Drop-in for reference ``flows/utils.py:34-48``: ``z = x / |T|`` (and its inverse), through the
    per-dimension affine kernel with ``s_k = -log|T|``.  Returns the tensor only, as the reference."""

    def __init__(self):
        super().__init__()
        self.T = nn.Parameter(torch.ones(1))

    def _s(self, x):
        return (-torch.log(torch.abs(self.T))).expand(x.shape[1])

    def forward(self, x):
        from .flows import _AffineConstFn, affine_const_apply
        if torch.is_grad_enabled():
            return _AffineConstFn.apply(x, self._s(x), None)
        return affine_const_apply(x, self._s(x), None)

    def backward(self, z):
        from .flows import affine_const_apply
        return affine_const_apply(z, self._s(z), None, inverse=True)
