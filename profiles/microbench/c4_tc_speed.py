"""Config C4 shape (K=100, L=8, H=512), wide tensor-core kernel only, CUDA events (CNF_B200_LIB picks an experimental build)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, cnf_b200
dev = torch.device('cuda:0')
K, L, H = 100, 8, [512]
torch.manual_seed(4)
m = cnf_b200.RealNvpFlow(K, layers=L, hidden_size=H)
with torch.no_grad():
    for p in m.parameters():
        if p.requires_grad: p.mul_(60.0)
m.to(dev)
eng = m.engine()
n = int(os.environ.get('N', 2_000_000))
x = 1.5 * torch.randn(n, K, device=dev)
x = x - x.mean(dim=1, keepdim=True)
eng.ensure(dev); eng.pack(tc=True)
for _ in range(2): eng.apply(x, precision='bf16', repack=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): eng.apply(x, precision='bf16', repack=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print('%-28s %.2f ms  %.1f M samples/s' % (os.path.basename(os.environ.get('CNF_B200_LIB', 'default')), ms, n / ms / 1e3))
