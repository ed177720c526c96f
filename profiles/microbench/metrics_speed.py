"""Timing of the metrics kernel (HBM-bound): probs fp32/fp64 and logits modes at K=10."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np, torch, cnf_b200, bench
from cnf_b200 import _lib
from cnf_b200._engine import _ptr, _stream
from cnf_b200.utils import metrics as M
dev = torch.device('cuda:0')
n = 12_500_000
x, y = bench.synth(n, 7, dev)
p32 = torch.softmax(x, dim=1).contiguous()
p64 = p32.double()
lp = torch.zeros(10, dtype=torch.float64, device=dev)
edges = torch.from_numpy(M.bin_edges(15)).to(dev)
acc = torch.zeros(48, dtype=torch.float64, device=dev)
def run(v, is64, mode):
    _lib.call('cnf_metrics', _ptr(v), ctypes.c_int32(is64), _ptr(y), ctypes.c_int64(n), ctypes.c_int32(10),
              ctypes.c_int32(15), ctypes.c_int32(mode), _ptr(lp), _ptr(edges), _ptr(acc), _stream(dev))
def timeit(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for name, v, is64, mode, bps in (('probs f32', p32, 0, 0, 48), ('probs f64', p64, 1, 0, 88), ('logits f32', x, 0, 1, 48), ('calibrated', x, 0, 2, 48)):
    ms = timeit(lambda: run(v, is64, mode))
    print('%-11s %.3f ms per 12.5M  -> %.1f G samples/s, %.0f GB/s' % (name, ms, n / ms / 1e6, n * bps / ms / 1e6))
