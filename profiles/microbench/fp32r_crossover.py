"""Where the register-resident fp32 kernels start to pay: forward (CNF_FP32R variants vs the generic / tile kernels)
and training (CNF_FP32R_TRAIN variants vs the split kernel) over batch sizes, CUDA events."""
import os
import sys

os.environ['CNF_LIVE_ENV'] = '1'
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import bench  # noqa: E402
import cnf_b200  # noqa: E402,F401

dev = torch.device('cuda:0')
model = bench.make_model().to(dev)
eng = model.engine()
eng.ensure(dev)
eng.pack()
xall, yall = bench.synth_dev(4 << 20, 1, dev)


def t_fwd(n, reps=20):
    x = xall[:n]
    eng.apply(x, repack=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.apply(x, repack=False)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def t_train(n, reps=10):
    x, y = xall[:n], yall[:n]
    acc = torch.zeros(4, dtype=torch.float64, device=dev)
    eng.nll_step(x, y, acc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.nll_step(x, y, acc)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


sizes = [40_000, 65536, 100_000, 131072, 200_000, 262144, 400_000, 524288, 1 << 20, 4 << 20]
fv = {'off': 'generic', '0': '128x8', '1': '128x4', '4': '128x2', '5': '128x1', '6': '64x2', '7': '64x1'}
print('forward, us per call:  N ' + ' '.join('%9s' % v for v in fv.values()))
for n in sizes:
    row = []
    for k in fv:
        os.environ['CNF_FP32R'] = k
        row.append(t_fwd(n))
    print('%22d ' % n + ' '.join('%9.1f' % v for v in row), flush=True)
tv = {'off': 'split', '0': 'auto', '2': '128x8', '4': '256x6', '5': '128x4'}
print('training, us per fused fwd+bwd pass:  N ' + ' '.join('%9s' % v for v in tv.values()))
for n in sizes:
    row = []
    for k in tv:
        os.environ['CNF_FP32R_TRAIN'] = k
        row.append(t_train(n))
    print('%37d ' % n + ' '.join('%9.1f' % v for v in row), flush=True)
