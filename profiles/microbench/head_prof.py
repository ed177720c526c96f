"""ncu target: the headline launches of bench.py in isolation -- cnf_flow_forward at the C2 shape over N logits in ONE
launch, bf16 (flow_tc_kernel) then fp32 (flow_reg10_kernel), then the fused statistics pass of both.  N from the
environment (default 10^8, the bench's launch size)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import bench  # noqa: E402
import cnf_b200  # noqa: E402,F401

dev = torch.device('cuda:0')
N = int(os.environ.get('N', 100_000_000))
model = bench.make_model().to(dev)
eng = model.engine()
eng.ensure(dev)
eng.pack(tc=True)
x, y = bench.synth_dev(N, 1000, dev)
lp = torch.log(torch.bincount(y, minlength=10).double() / N).cpu().numpy()
for prec in ('bf16', 'fp32'):
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        z, ld, _ = eng.apply(x, precision=prec, repack=False)
        e1.record()
        torch.cuda.synchronize()
        del z, ld
    print('%s forward: %.3f ms  %.3f G samples/s' % (prec, e0.elapsed_time(e1), N / e0.elapsed_time(e1) / 1e6))
    for rep in range(2):
        e0.record()
        eng.predict(x, center=False, log_priors=lp, y=y, bins=15, precision=prec, repack=False)
        e1.record()
        torch.cuda.synchronize()
    print('%s fused statistics pass: %.3f ms  %.3f G samples/s' % (prec, e0.elapsed_time(e1), N / e0.elapsed_time(e1) / 1e6))
