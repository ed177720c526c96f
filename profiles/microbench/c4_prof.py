"""C4 forward on the wide tensor-core kernel, 400k samples x 6 launches (ncu target for flow_tcw_kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np, torch, cnf_b200
dev = torch.device('cuda:0')
K, L, H = 100, 8, [512]
torch.manual_seed(4)
m = cnf_b200.RealNvpFlow(K, layers=L, hidden_size=H)
with torch.no_grad():
    for p in m.parameters():
        if p.requires_grad: p.mul_(60.0)
m.to(dev)
eng = m.engine()
n = 400_000
g = torch.Generator().manual_seed(1)
x = (1.5 * torch.randn(n, K, generator=g))
x[torch.arange(n), torch.randint(0, K, (n,), generator=g)] += 3.0
x = (x - x.mean(dim=1, keepdim=True)).to(dev)
eng.ensure(dev); eng.pack(tc=True)
for _ in range(2): eng.apply(x, precision='bf16', repack=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(4): eng.apply(x, precision='bf16', repack=False)
e1.record(); torch.cuda.synchronize()
print('C4 bf16 ms per 400k: %.3f  -> %.1f M samples/s' % (e0.elapsed_time(e1) / 4, n / (e0.elapsed_time(e1) / 4) / 1e3))
