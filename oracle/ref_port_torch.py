"""CPU port of the reference path in torch CPU ops.  TEST / BASELINE INFRASTRUCTURE ONLY.

The reference's arithmetic lives in torch, so the honest CPU baseline is the same
op sequence on the same library: this file restates flows/flows.py:101-126, flows/utils.py:26-31
and the train step of calibrators.py:287-295 functionally (weights passed in as a flat vector in
the framework's canonical order), keeping the reference's dense formulation -- full-width
mask*x inputs, full K-wide conditioner outputs multiplied by (1-mask), a flip per layer and one
[N,H] hidden tensor per Linear.  Used by bench.py (cpu_baseline and --impl reference) and by
tests as a second, autograd-based checker.  Never imported by the product package.
"""
import torch


def unflatten(flat, K, L, hidden, scale=True, shift=True):
    units = [K] + list(hidden) + [K]
    layers, off = [], 0
    for _ in range(L):
        lay = {}
        for name, present in (('s', scale), ('t', shift)):
            if not present:
                lay[name] = None
                continue
            net = []
            for i, o in zip(units[:-1], units[1:]):
                W = flat[off:off + o * i].view(o, i); off += o * i
                b = flat[off:off + o]; off += o
                net.append((W, b))
            lay[name] = net
        layers.append(lay)
    assert off == flat.numel()
    return layers


def _mlp(net, x):
    for W, b in net[:-1]:
        x = torch.relu(torch.nn.functional.linear(x, W, b))
    W, b = net[-1]
    return torch.nn.functional.linear(x, W, b)


def forward(layers, x):
    K = x.shape[1]
    mask = torch.zeros(1, K, dtype=x.dtype)
    mask[:, K // 2:] = 1
    cum = 0.0
    zs = []
    for lay in layers:
        xb = mask * x
        b1 = 1 - mask
        s = _mlp(lay['s'], xb) if lay['s'] is not None else x.new_zeros(x.size())
        t = _mlp(lay['t'], xb) if lay['t'] is not None else x.new_zeros(x.size())
        z = xb + b1 * (x * torch.exp(s) + t)
        cum = cum + torch.sum(b1 * s, dim=1)
        x = z.flip((1,))
        zs.append(x)
    return zs, cum


def inverse(layers, z):
    K = z.shape[1]
    mask = torch.zeros(1, K, dtype=z.dtype)
    mask[:, K // 2:] = 1
    cum = 0.0
    for lay in layers[::-1]:
        z = z.flip((1,))
        xb = mask * z
        b1 = 1 - mask
        s = _mlp(lay['s'], xb) if lay['s'] is not None else z.new_zeros(z.size())
        t = _mlp(lay['t'], xb) if lay['t'] is not None else z.new_zeros(z.size())
        z = xb + b1 * (z - t) * torch.exp(-s)
        cum = cum + torch.sum(b1 * (-s), dim=1)
    return z, cum


def nll_loss(layers, x, y):
    zs, ld = forward(layers, x)
    probs = torch.softmax(zs[-1], dim=1)
    ce = torch.log(probs.gather(1, y.view(-1, 1)) + 1e-7)
    return -torch.mean(ce.squeeze() + ld)


class TrainState:
    """flat parameter leaf + torch.optim.Adam, one step = calibrators.py:287-295."""

    def __init__(self, flat, K, L, hidden, scale=True, shift=True):
        self.flat = flat.clone().requires_grad_(True)
        self.cfg = (K, L, hidden, scale, shift)
        self.opt = torch.optim.Adam([self.flat])

    def step(self, x, y):
        layers = unflatten(self.flat, *self.cfg)
        loss = nll_loss(layers, x, y)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return float(loss)
