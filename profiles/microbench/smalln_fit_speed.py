"""Per-epoch latency of the full-batch calibrator step at realistic validation-set sizes (C2 shape)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, cnf_b200, bench
dev = torch.device('cuda:0')
for n in (1000, 5000, 10000, 50000):
    xt, yt = bench.synth_dev(n, 77, dev)
    for prec in ('fp32', 'bf16'):
        m = bench.make_model(seed=2, wmult=1.0).to(dev)
        tr = cnf_b200.FusedNLLTrainer(m.engine(), xt, yt, precision=prec)
        for _ in range(5): tr.step(); tr.evaluate()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(200): tr.step(); tr.evaluate()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 200
        t0 = time.perf_counter()
        for _ in range(200): tr.step()
        torch.cuda.synchronize()
        ds = (time.perf_counter() - t0) / 200
        print('N=%6d %s  %.1f us per epoch (step + evaluate)   %.1f us per step alone' % (n, prec, dt * 1e6, ds * 1e6))
