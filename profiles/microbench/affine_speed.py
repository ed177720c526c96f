"""Timing of the HBM-bound streaming kernels: AffineConstantLayer forward / inverse / backward,
TempScaler, calibrated-probability tail, next to a plain device copy of the same bytes."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np, torch, cnf_b200, bench
from cnf_b200 import _lib
from cnf_b200._engine import _ptr, _stream
dev = torch.device('cuda:0')
def timeit(f, reps=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for K in (10, 100, 3):
    n = 100_000_000 // K
    x = torch.randn(n, K, device=dev)
    z = torch.empty_like(x)
    gz = torch.randn(n, K, device=dev)
    gx = torch.empty_like(x)
    s = torch.randn(K, device=dev) * 0.1
    t = torch.randn(K, device=dev)
    gs = torch.zeros(K, device=dev); gt = torch.zeros(K, device=dev)
    st = _stream(dev)
    ms_copy = timeit(lambda: z.copy_(x))
    ms_f = timeit(lambda: _lib.call('cnf_affine_const', _ptr(x), _ptr(s), _ptr(t), _ptr(z), ctypes.c_int64(n), ctypes.c_int32(K), ctypes.c_int32(0), st))
    ms_i = timeit(lambda: _lib.call('cnf_affine_const', _ptr(x), _ptr(s), _ptr(t), _ptr(z), ctypes.c_int64(n), ctypes.c_int32(K), ctypes.c_int32(1), st))
    ms_b = timeit(lambda: _lib.call('cnf_affine_const_backward', _ptr(x), _ptr(gz), _ptr(s), _ptr(gx), _ptr(gs), _ptr(gt), ctypes.c_int64(n), ctypes.c_int32(K), st))
    gb = n * K * 8 / 1e6
    print('K=%3d N=%9d  copy %.3f ms (%.0f GB/s) | affine fwd %.3f ms (%.0f GB/s) inv %.3f ms (%.0f GB/s) | backward %.3f ms (%.0f GB/s of 12K B/sample)'
          % (K, n, ms_copy, gb / ms_copy, ms_f, gb / ms_f, ms_i, gb / ms_i, ms_b, n * K * 12 / 1e6 / ms_b))

# PlanarLayer / RadialLayer streaming kernels (8K+4 / 8K bytes per sample forward)
import cnf_b200
for K in (10, 100):
    n = 100_000_000 // K
    x = torch.randn(n, K, device=dev)
    pl = cnf_b200.PlanarLayer(K).to(dev)
    rl = cnf_b200.RadialLayer(K).to(dev)
    with torch.no_grad():
        ms_p = timeit(lambda: pl(x))
        ms_r = timeit(lambda: rl(x))
    xg = x.clone().requires_grad_(True)
    def fb(layer):
        z, ld = layer(xg)
        (z.sum() + ld.sum()).backward()
    ms_pb = timeit(lambda: fb(pl), reps=5)
    print('K=%3d planar fwd %.3f ms (%.0f GB/s)  radial fwd %.3f ms (%.0f GB/s)  planar fwd+bwd via autograd %.3f ms'
          % (K, ms_p, n * (8 * K + 4) / 1e6 / ms_p, ms_r, n * 8 * K / 1e6 / ms_r, ms_pb))
