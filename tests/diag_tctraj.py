"""Loss trajectories of the fp32 and bf16 trainers from the same initialisation (GPU box)."""
import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
import flow_oracle as orc
import cnf_b200
from cnf_b200.calibrators import FusedNLLTrainer
dev = torch.device('cuda:0')
K, N = 10, 200000
x, y = orc.synth_logits(N, K, seed=11)
xt, yt = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
for prec in ('fp32', 'bf16'):
    torch.manual_seed(3)
    flow = cnf_b200.RealNvpFlow(K, layers=6, hidden_size=[128]).to(dev)
    tr = FusedNLLTrainer(flow.engine(), xt, yt, lr=1e-3, precision=prec)
    out = []
    for i in range(40):
        tr.step()
        a = tr.loss_acc.cpu().numpy()
        out.append('%.4f/%.3f/%.3f/%d' % (-a[0] / N, -a[1] / N, a[2] / N, a[3]))
    print(prec, ' '.join(out))
