"""Quick timing of the fp32 kernels at the C2 shape (forward 1M samples, train step 1M samples)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch, cnf_b200, bench
dev = torch.device('cuda:0')
n = 1 << 20
xt, yt = bench.synth(n, 5000, dev)
m = bench.make_weights(seed=2).to(dev)
eng = m.engine()
def timeit(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
eng.ensure(dev); eng.pack(tc=True)
print('fp32 forward  ms/1Mi: %.3f' % timeit(lambda: eng.apply(xt, repack=False)))
print('fp32 inverse  ms/1Mi: %.3f' % timeit(lambda: eng.apply(xt, inverse=True, repack=False)))
print('bf16 forward  ms/1Mi: %.3f' % timeit(lambda: eng.apply(xt, precision='bf16', repack=False)))
tr = cnf_b200.FusedNLLTrainer(eng, xt, yt)
print('train step    ms/1Mi: %.3f' % timeit(tr.step, 5))
print('eval pass     ms/1Mi: %.3f' % timeit(tr.evaluate, 5))
