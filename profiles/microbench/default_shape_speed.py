"""The reference's DEFAULT conditioner (NvpCouplingLayer(dim, hidden_size=[5, 5]), flows/flows.py:69) at large N:
fp32 forward and NLL + Adam step, K = 10, L = 6 (and a few neighbours)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, cnf_b200, bench
dev = torch.device('cuda:0')
N = int(os.environ.get('N', 10_000_000))
x, y = bench.synth_dev(N, 3, dev)
def timeit(f, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for hidden in ([5, 5], [5], [16], [32], [64], [128]):
    torch.manual_seed(1)
    flow = cnf_b200.Flow([cnf_b200.NvpCouplingLayer(10, hidden) for _ in range(6)]).to(dev)
    eng = flow.engine()
    eng.ensure(dev); eng.pack()
    tf = timeit(lambda: eng.apply(x, repack=False))
    nt = min(N, 4_000_000)
    tr = cnf_b200.FusedNLLTrainer(eng, x[:nt], y[:nt])
    tt = timeit(tr.step, 3)
    macs = 2 * sum(a * b for a, b in zip([5] + hidden, hidden + [5])) * 6
    print('hidden=%-10s forward %.3f ms = %.2f G samples/s (%.1f TFLOP/s minimal, HBM %.0f GB/s)   train step %.2f ms = %.1f M samples/s'
          % (hidden, tf, N / tf / 1e6, 2 * macs * N / tf / 1e9, 84 * N / tf / 1e6, tt, nt / tt / 1e3))
