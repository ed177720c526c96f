"""ECE / NLL / accuracy -- drop-in for reference ``utils/metrics.py`` (:35-73, :6-15, :76-80).

numpy (or torch) in, Python floats out, like the reference; the arithmetic is one pass of the
``cnf_metrics`` CUDA kernel (histogram reduction over right-closed bins).  float32 and
float64 probability arrays are both accepted and binned in their own dtype, as NumPy does.

Sharded evaluation: the kernel produces additive sufficient statistics
(``3*bins+3`` float64 values); with ``sharded=True`` every rank of an initialised torch.distributed
group passes its own shard and the statistics are all-reduced before the metric is formed, so each
rank returns the metric of the whole set.  ``sharded`` is a collective: call it on every rank.
"""
import ctypes

import numpy as np
import torch

from .. import _lib
from .._engine import _ptr, _stream, on_device


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError('cnf_b200.utils.metrics needs a CUDA device; there is no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def _as_device(a, dev):
    if isinstance(a, torch.Tensor):
        return a.to(dev)
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def bin_edges(bins):
    width = 1. / bins                      # utils/metrics.py:57, 61: edges are i*width
    return np.array([i * width for i in range(bins + 1)], dtype=np.float64)


def _labels(target, n_cols):
    if isinstance(target, torch.Tensor):
        t = target
        if t.dim() == 2:
            t = t.argmax(dim=1)
        return t.to(torch.int64)
    t = np.asarray(target)
    if t.ndim == 2:
        t = np.argmax(t, axis=1)
    return t.astype(np.int64)


def statistics(values, target, bins=15, mode=_lib.METRICS_PROBS, log_priors=None, device=None):
    """Raw sufficient statistics as a float64 numpy vector [3*bins+3]:
    per-bin count, sum of confidences, sum of correct; then sum -log(p_y+1e-7), #correct, N.
    `values` are probabilities (mode PROBS) or float32 logits (modes LOGITS / CALIBRATED)."""
    dev = device or (values.device if isinstance(values, torch.Tensor) and values.is_cuda else _device())
    v = _as_device(values, dev)
    if v.dtype not in (torch.float32, torch.float64):
        v = v.to(torch.float64)
    if mode != _lib.METRICS_PROBS:
        v = v.to(torch.float32)
    v = v.contiguous()
    N, K = v.shape
    y = _as_device(_labels(target, K), dev).to(torch.int64).contiguous()
    acc = torch.zeros(3 * bins + 3, dtype=torch.float64, device=dev)
    edges = torch.from_numpy(bin_edges(bins)).to(dev)
    lp = None
    if log_priors is not None:
        lp = _as_device(np.asarray(log_priors, dtype=np.float64), dev)
    if v.dim() != 2:
        raise ValueError('cnf_b200.metrics: expected a [N, K] array, got shape %s' % (tuple(v.shape),))
    if y.dim() != 1 or y.shape[0] != N:
        raise ValueError('cnf_b200.metrics: %d rows but labels of shape %s' % (N, tuple(y.shape)))
    if lp is not None and lp.numel() != K:
        raise ValueError('cnf_b200.metrics: log_priors must have %d entries' % K)
    with on_device(dev):
        _lib.call('cnf_metrics', _ptr(v), ctypes.c_int32(1 if v.dtype == torch.float64 else 0), _ptr(y),
                  ctypes.c_int64(N), ctypes.c_int32(K), ctypes.c_int32(bins), ctypes.c_int32(mode), _ptr(lp),
                  _ptr(edges), _ptr(acc), _stream(dev))
    return acc


def reduce_statistics(stats):
    """Sum the sufficient statistics over the ranks of the default torch.distributed group (no-op when it
    is not initialised or has one rank).  Works on CUDA tensors (NCCL) and CPU tensors (gloo)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return stats
    t = stats if isinstance(stats, torch.Tensor) else torch.as_tensor(np.asarray(stats, dtype=np.float64))
    t = t.clone()
    dist.all_reduce(t)
    return t


def metrics_from_statistics(stats, bins):
    """{'ece', 'nll', 'accuracy', 'n'} from a statistics vector (utils/metrics.py:66-73, :15, :80)."""
    st = np.asarray(stats.cpu() if isinstance(stats, torch.Tensor) else stats, dtype=np.float64)
    n = st[3 * bins + 2]
    return {'ece': float(ece_from_statistics(st, bins)), 'nll': float(st[3 * bins] / n) if n > 0 else 0.0,
            'accuracy': float(st[3 * bins + 1] / n) if n > 0 else 0.0, 'n': int(n)}


def ece_from_statistics(stats, bins):
    stats = np.asarray(stats.cpu() if isinstance(stats, torch.Tensor) else stats, dtype=np.float64)
    cnt, sconf, sacc = stats[:bins], stats[bins:2 * bins], stats[2 * bins:3 * bins]
    n = stats[3 * bins + 2]
    total = 0.0
    for c, sc, sa in zip(cnt, sconf, sacc):
        if c > 0:                          # empty bins are skipped (utils/metrics.py:66)
            total += abs(sa / c - sc / c) * c
    return total / n if n > 0 else 0.0


def _two_columns(probs):
    """1-D (binary) inputs: column 0 = 1-p, column 1 = p (utils/metrics.py:9-10, 50-52)."""
    p = np.asarray(probs)
    if p.ndim < 2 or p.shape[1] == 1:
        p = p.reshape(-1)
        return np.stack([1. - p, p], axis=1)
    return probs


def expected_calibration_error(probs, target, bins=15, sharded=False):
    if not isinstance(probs, torch.Tensor):
        probs = _two_columns(probs)
    stats = statistics(probs, target, bins=bins)
    if sharded:
        stats = reduce_statistics(stats)
    return float(ece_from_statistics(stats, bins))


def neg_log_likelihood(probs, target, sharded=False):
    if not isinstance(probs, torch.Tensor):
        probs = _two_columns(probs)
    stats = statistics(probs, target, bins=1)
    if sharded:
        stats = reduce_statistics(stats)
    stats = stats.cpu().numpy()
    return float(stats[3] / stats[5])


def accuracy(probs, target, sharded=False):
    stats = statistics(probs, target, bins=1)
    if sharded:
        stats = reduce_statistics(stats)
    stats = stats.cpu().numpy()
    return float(stats[4] / stats[5])
