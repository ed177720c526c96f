"""C4-shaped fp32 training steps (lean kernel), 20,000 samples (ncu target for flow_train_lean_kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, cnf_b200
dev = torch.device('cuda:0')
K, L, H, N = 100, 8, 512, 20000
g = torch.Generator().manual_seed(1)
x = 1.5 * torch.randn(N, K, generator=g); y = torch.randint(0, K, (N,), generator=g)
x[torch.arange(N), y] += 3.0; x -= x.mean(dim=1, keepdim=True)
torch.manual_seed(2)
m = cnf_b200.RealNvpFlow(K, layers=L, hidden_size=[H]).to(dev)
tr = cnf_b200.FusedNLLTrainer(m.engine(), x.to(dev), y.to(dev))
for _ in range(3): tr.step()
torch.cuda.synchronize()
print('ok')
