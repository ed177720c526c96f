"""A few full-batch epochs at N=5,000 on the tensor-core trainer (ncu launch-list target)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, cnf_b200, bench
dev = torch.device('cuda:0')
xt, yt = bench.synth(5000, 77, dev)
m = bench.make_weights(seed=2).to(dev)
tr = cnf_b200.FusedNLLTrainer(m.engine(), xt, yt, precision=os.environ.get('CNF_PREC', 'bf16'))
for _ in range(6):
    tr.step()
tr.evaluate()
torch.cuda.synchronize()
print('ok')
