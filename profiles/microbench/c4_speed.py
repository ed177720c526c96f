"""Config C4 (K=100, L=8, H=512): fp32 kernel vs wide tensor-core kernel, accuracy against each other."""
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
n = 200_000
g = torch.Generator().manual_seed(1)
x = (1.5 * torch.randn(n, K, generator=g))
x[torch.arange(n), torch.randint(0, K, (n,), generator=g)] += 3.0
x = (x - x.mean(dim=1, keepdim=True)).to(dev)
eng.ensure(dev); eng.pack(tc=True)
def timeit(f, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
z32, ld32, _ = eng.apply(x, repack=False)
zbf, ldbf, _ = eng.apply(x, precision='bf16', repack=False)
print('bf16 vs fp32: z rel %.3e  logdet rel %.3e' % (float((zbf - z32).abs().max() / z32.abs().max()),
      float((ldbf - ld32).abs().max() / max(1.0, float(ld32.abs().max())))))
t32 = timeit(lambda: eng.apply(x, repack=False), 3)
tbf = timeit(lambda: eng.apply(x, precision='bf16', repack=False))
print('fp32 %.2f ms (%.1f M samples/s)   bf16-tc %.2f ms (%.1f M samples/s)' % (t32, n / t32 / 1e3, tbf, n / tbf / 1e3))
