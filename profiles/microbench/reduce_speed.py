"""In-situ (L2-hot) time of the partial-row reduction after a small-batch fp32 step."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, cnf_b200, bench
from cnf_b200 import _lib
from cnf_b200._engine import _ptr, _stream
dev = torch.device('cuda:0')
for n in (1000, 5000, 1 << 20):
    xt, yt = bench.synth(n, 77, dev)
    m = bench.make_weights(seed=2).to(dev)
    eng = m.engine()
    tr = cnf_b200.FusedNLLTrainer(eng, xt, yt, precision='fp32')
    tr.step()
    rows = min((n + 31) // 32, 296)
    st = _stream(dev)
    def red():
        _lib.call('cnf_grad_reduce_rows', ctypes.byref(eng.desc), _ptr(eng.partials), ctypes.c_int64(rows), _ptr(eng.gather), _ptr(eng.flat_grad), st)
    for _ in range(5): red()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): red()
    e1.record(); torch.cuda.synchronize()
    print('N=%d rows=%d: reduce (memset + kernel) %.2f us' % (n, rows, e0.elapsed_time(e1) / 200 * 1e3))
