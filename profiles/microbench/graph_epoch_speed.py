"""Steady-state time of one full-batch epoch (optimiser step + evaluation pass) at N=5,000, C2 shape:
eager launches against the replayed CUDA graph, per precision."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, cnf_b200, bench
dev = torch.device('cuda:0')
xt, yt = bench.synth(5000, 77, dev)
for prec in ('fp32', 'bf16'):
    for mode in ('eager', 'graph'):
        m = bench.make_weights(seed=2).to(dev)
        tr = cnf_b200.FusedNLLTrainer(m.engine(), xt, yt, precision=prec)
        run = (lambda: (tr.step(), tr.evaluate())) if mode == 'eager' else tr.epoch_graph
        for _ in range(5): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(300): run()
        e1.record(); torch.cuda.synchronize()
        print('%s %s: %.1f us per epoch' % (prec, mode, e0.elapsed_time(e1) / 300 * 1e3))
