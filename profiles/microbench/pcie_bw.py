"""Raw pinned-memory PCIe bandwidth of the box (one direction and both at once) and the e2e API at a few chunkings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, bench
dev = torch.device('cuda:0')
n = 64 << 20
h1 = torch.empty(n, dtype=torch.float32, pin_memory=True); h2 = torch.empty(n, dtype=torch.float32, pin_memory=True)
d1 = torch.empty(n, dtype=torch.float32, device=dev); d2 = torch.empty(n, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(f, reps=5):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
gb = n * 4 / 1e9
print('H2D alone  %.1f GB/s' % (gb / t(lambda: d1.copy_(h1, non_blocking=True))))
print('D2H alone  %.1f GB/s' % (gb / t(lambda: h2.copy_(d2, non_blocking=True))))
def both():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print('both ways  %.1f GB/s each' % (gb / t(both)))
import cnf_b200
m = bench.make_weights().to(dev); m.flow.precision = 'bf16'
N = 1_000_000
xh = torch.empty((N, 10), dtype=torch.float32, pin_memory=True); xh.copy_(bench.synth(N, 3)[0])
zh = torch.empty((N, 10), dtype=torch.float32, pin_memory=True); lh = torch.empty(N, dtype=torch.float32, pin_memory=True)
for chunk, slots in ((1 << 18, 3), (1 << 17, 4), (1 << 16, 4), (1 << 19, 2), (1 << 20, 1), (125000, 4)):
    f = lambda: m.transform_host(xh, zh, lh, device=dev, chunk=chunk, slots=slots)
    dt = t(f, 10)
    print('e2e chunk %7d slots %d: %.3f ms  %.2f G samples/s' % (chunk, slots, dt * 1e3, N / dt / 1e9))

# zero-copy: the flow kernel reads the pinned host logits and writes z / log-det straight over PCIe
import ctypes
from cnf_b200 import _lib
from cnf_b200._engine import _ptr, _stream
eng = m.engine(); eng.ensure(dev); eng.pack(tc=True)
def zc(desc, packed):
    _lib.call('cnf_flow_forward', ctypes.byref(desc), _ptr(packed), _ptr(eng.tables), _ptr(xh), _ptr(zh), _ptr(lh),
              None, ctypes.c_int64(N), _stream(dev))
    torch.cuda.current_stream().synchronize()
zref = zh.clone()
dt = t(lambda: zc(eng.desc_tc, eng.packed_tc), 10)
print('zero-copy bf16 kernel: %.3f ms  %.2f G samples/s   same result: %s' % (dt * 1e3, N / dt / 1e9, bool(torch.equal(zref, zh))))
dt = t(lambda: zc(eng.desc, eng.packed), 10)
print('zero-copy fp32 kernel: %.3f ms  %.2f G samples/s' % (dt * 1e3, N / dt / 1e9))
