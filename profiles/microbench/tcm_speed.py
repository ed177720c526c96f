"""Conditioners with two hidden layers on the tensor cores (cnf_flow_tcm.cu) against the fp32 deep kernel:
forward + log-det at K = 10, L = 6 (and a K = 40 case), device-resident logits, CUDA events."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, cnf_b200, bench
dev = torch.device('cuda:0')
N = int(os.environ.get('N', 4_000_000))
def timeit(f, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for K, L, hidden in ((10, 6, [128, 128]), (10, 6, [64, 64]), (10, 6, [128, 64]), (40, 4, [128, 128]), (10, 6, [128, 128, 128]),
                     (10, 6, [128, 128, 128, 128])):
    x, y = bench.synth_dev(N, 3, dev, k=K)
    torch.manual_seed(1)
    flow = cnf_b200.Flow([cnf_b200.NvpCouplingLayer(K, hidden) for _ in range(L)]).to(dev)
    with torch.no_grad():
        for p in flow.parameters():
            if p.requires_grad: p.mul_(30.0 if len(hidden) == 2 else 60.0)
    eng = flow.engine()
    eng.ensure(dev); eng.pack(); eng.pack(tc=True)
    tb = timeit(lambda: eng.apply(x, precision='bf16', repack=False))
    n32 = min(N, 200_000)
    t32 = timeit(lambda: eng.apply(x[:n32], repack=False), 2)
    d0, d1 = K // 2, K - K // 2
    flop = 2 * 2 * L * (d1 * hidden[0] + sum(a * b for a, b in zip(hidden[:-1], hidden[1:])) + hidden[-1] * d0)
    print('K=%d L=%d hidden=%-11s bf16 tcgen05 %.3f ms = %.3f G samples/s (%.0f TFLOP/s minimal)   fp32 %.1f M samples/s'
          % (K, L, hidden, tb, N / tb / 1e6, flop * N / tb / 1e9, n32 / t32 / 1e3), flush=True)
