"""Split of the bf16 training step at 2^21 samples: forward without tape, whole step, evaluation pass."""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch, cnf_b200, bench
dev = torch.device('cuda:0')
n = 1 << 21
xt, yt = bench.synth_dev(n, 5000, dev)
m = bench.make_model(seed=2, wmult=1.0).to(dev)
eng = m.engine()
tr = cnf_b200.FusedNLLTrainer(eng, xt, yt, precision='bf16')
def timeit(f, reps=5):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print('forward (no tape)  %.3f ms' % timeit(lambda: eng.apply(xt, precision='bf16', repack=False)))
print('evaluation pass    %.3f ms  (forward + loss head only)' % timeit(tr.evaluate))
print('training step      %.3f ms  (forward + tape, backward, reduce, Adam, repack)' % timeit(tr.step))
