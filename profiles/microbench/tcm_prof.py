"""One launch of flow_tcm_kernel at hidden [128, 128] (K = 10, L = 6) for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, cnf_b200, bench
dev = torch.device('cuda:0')
N = int(os.environ.get('N', 1_000_000))
x, _ = bench.synth_dev(N, 3, dev)
m = bench.make_model(seed=7, wmult=30.0, hidden=[int(h) for h in os.environ.get('H', '128,128').split(',')]).to(dev)
e = m.engine(); e.ensure(dev); e.pack(tc=True)
for _ in range(3):
    e.apply(x, precision='bf16', repack=False)
torch.cuda.synchronize()
