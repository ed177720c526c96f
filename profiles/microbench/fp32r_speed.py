"""fp32 forward at the C2 shape: register-resident kernel variants (CNF_FP32R=0..5) against the generic
flow_apply_kernel (CNF_FP32R=off), CUDA events, 10^7 samples; parity of every variant against the generic kernel."""
import os
import sys

os.environ['CNF_LIVE_ENV'] = '1'     # switches are flipped between calls below

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
names = {'off': 'generic flow_apply_kernel', '0': '128 thr x 8 (2 CTAs/SM)', '1': '128 x 4 (3/SM)', '2': '128 x 6 (3/SM)', '3': '256 x 8 (1/SM)',
         '4': '128 x 2 (3/SM)', '5': '128 x 1', '6': '64 x 2', '7': '64 x 1'}
ref = None
for v in ['off'] + [str(i) for i in range(8)]:
    os.environ['CNF_FP32R'] = v
    for inverse in (False, True):
        z, ld, _ = eng.apply(x, inverse=inverse)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            eng.apply(x, inverse=inverse, repack=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        if v == 'off':
            ref = (ref or {})
            ref[inverse] = (z.clone(), ld.clone())
            err = (0.0, 0.0)
        else:
            err = (float((z - ref[inverse][0]).abs().max() / ref[inverse][0].abs().max()),
                   float((ld - ref[inverse][1]).abs().max() / max(1.0, float(ref[inverse][1].abs().max()))))
        print('%-28s %s  %.3f ms  %.3f G samples/s  max rel diff z %.1e ld %.1e' % (
            names[v], 'inv' if inverse else 'fwd', ms, N / ms / 1e6, err[0], err[1]), flush=True)
