"""bf16 training step at the C2/C3 shape, 2^21 samples: whole step and tape-writing forward alone
(CNF_B200_LIB picks an experimental build; with CNF_TCB_EXP builds only the timing means anything)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, cnf_b200, bench
dev = torch.device('cuda:0')
n = int(os.environ.get('N', 1 << 21))
xt, yt = bench.synth_dev(n, 7000, dev)
m = bench.make_model(seed=2, wmult=1.0).to(dev)
tr = cnf_b200.FusedNLLTrainer(m.engine(), xt, yt, n_total=n, precision='bf16')
def timeit(f, reps=8):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ts = timeit(tr.step)
te = timeit(tr.evaluate)
print('%-24s step %.3f ms = %.1f M samples/s   evaluation pass %.3f ms' % (
    os.path.basename(os.environ.get('CNF_B200_LIB', 'default')), ts, n / ts / 1e3, te))
