"""fp32 NLL training step at the C2 shape: register-resident kernel variants (CNF_FP32R_TRAIN=0..4) against the
32-sample-tile split kernel (CNF_FP32R_TRAIN=off): CUDA events per fused fwd+bwd pass, gradient and loss agreement."""
import os
import sys

os.environ['CNF_LIVE_ENV'] = '1'
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import bench  # noqa: E402
import cnf_b200  # noqa: E402,F401

dev = torch.device('cuda:0')
N = int(os.environ.get('N', 4 << 20))
x, y = bench.synth_dev(N, 7, dev)
model = bench.make_model(seed=2, wmult=float(os.environ.get('WMULT', 1.0))).to(dev)
eng = model.engine()
eng.ensure(dev)
eng.pack()
names = {'off': 'split kernel (32-sample tiles)', '0': 'reg 128 thr x 4 u2 (2/SM)', '1': 'reg 128 x 6 u2 (1/SM)', '2': 'reg 128 x 8 u2', '3': 'reg 256 x 8 u2', '4': 'reg 256 x 6 u2', '5': 'reg 128 x 8 u1', '6': 'reg 256 x 8 u1'}
ref = None
for v in ['off', '0', '1', '2', '3', '4', '5', '6']:
    os.environ['CNF_FP32R_TRAIN'] = v
    acc = torch.zeros(4, dtype=torch.float64, device=dev)
    eng.nll_step(x, y, acc)
    torch.cuda.synchronize()
    g = eng.flat_grad.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        eng.nll_step(x, y, acc)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    if ref is None:
        ref = (g, acc.clone())
    err = float((g - ref[0]).abs().max() / ref[0].abs().max())
    print('%-32s %.3f ms  %.1f M samples/s   grad max rel diff %.1e   loss sums rel diff %.1e  nonfinite %d' % (
        names[v], ms, N / ms / 1e3, err, float(((acc - ref[1] * (acc[0] / ref[1][0])).abs().max()) / acc[0].abs()), int(acc[3])), flush=True)
