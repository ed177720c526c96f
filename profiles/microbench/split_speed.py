"""fp32 training step at the C2 shape over batch sizes: the small-batch (split) kernel against the
one-thread-per-sample kernel (CNF_SPLIT_TRAIN=1 / 0 in the environment of separate runs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, cnf_b200, bench
dev = torch.device('cuda:0')
for n in (1000, 5000, 10000, 20000, 50000, 200000, 1 << 20):
    xt, yt = bench.synth(n, 77, dev)
    m = bench.make_weights(seed=2).to(dev)
    tr = cnf_b200.FusedNLLTrainer(m.engine(), xt, yt, precision='fp32')
    for _ in range(5): tr.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50 if n <= 50000 else 5
    e0.record()
    for _ in range(reps): tr.step()
    e1.record(); torch.cuda.synchronize()
    print('N=%8d  %.1f us per step  (CNF_SPLIT_TRAIN=%s)' % (n, e0.elapsed_time(e1) / reps * 1e3, os.environ.get('CNF_SPLIT_TRAIN', 'default')))
