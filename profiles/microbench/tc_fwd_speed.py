"""bf16 tcgen05 forward at the C2 shape, 10^7 samples, CUDA events (CNF_B200_LIB picks an experimental build)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import bench  # noqa: E402
import cnf_b200  # noqa: E402,F401

dev = torch.device('cuda:0')
model = bench.make_model().to(dev)
eng = model.engine()
eng.ensure(dev)
N = int(os.environ.get('N', 10_000_000))
x, _ = bench.synth_dev(N, 1, dev)
eng.pack(tc=True)
z, ld, _ = eng.apply(x, precision='bf16')
zf, ldf, _ = eng.apply(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    eng.apply(x, precision='bf16', repack=False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print('%s: %.3f ms  %.3f G samples/s   vs fp32 kernel: z %.1e ld %.1e' % (
    os.environ.get('CNF_B200_LIB', 'default lib'), ms, N / ms / 1e6,
    float((z - zf).abs().max() / zf.abs().max()), float((ld - ldf).abs().max() / max(1.0, float(ldf.abs().max())))))
