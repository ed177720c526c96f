"""Training-step latency at CIFAR-100-sized problems (K=100) on the fp32 kernels (the only training path for
wide shapes) and, for scale, the reference op sequence on the host cores."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch, cnf_b200
dev = torch.device('cuda:0')
for (K, L, H, N) in ((100, 8, 512, 10000), (100, 4, 100, 10000), (100, 8, 512, 100000)):
    g = torch.Generator().manual_seed(1)
    x = 1.5 * torch.randn(N, K, generator=g)
    y = torch.randint(0, K, (N,), generator=g)
    x[torch.arange(N), y] += 3.0
    x -= x.mean(dim=1, keepdim=True)
    torch.manual_seed(2)
    m = cnf_b200.RealNvpFlow(K, layers=L, hidden_size=[H]).to(dev)
    tr = cnf_b200.FusedNLLTrainer(m.engine(), x.to(dev), y.to(dev))
    for _ in range(2): tr.step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): tr.step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    with torch.no_grad():
        m(x.to(dev)); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5): m(x.to(dev))
        torch.cuda.synchronize()
        df = (time.perf_counter() - t0) / 5
    print('K=%d L=%d H=%d N=%d: fp32 train step %.2f ms (%.2f M samples/s), fp32 forward incl. H2D %.2f ms' % (K, L, H, N, dt * 1e3, N / dt / 1e6, df * 1e3))
# host reference op sequence (torch CPU autograd + Adam) for the first case
import ref_port_torch as rp
K, L, H, N = 100, 8, 512, 10000
g = torch.Generator().manual_seed(1)
x = 1.5 * torch.randn(N, K, generator=g); y = torch.randint(0, K, (N,), generator=g)
n_par = L * 2 * (H * K + H + K * H + K)
st = rp.TrainState(0.001 * torch.randn(n_par, generator=g), K, L, [H])
torch.set_num_threads(os.cpu_count())
st.step(x, y)
t0 = time.perf_counter()
for _ in range(3): st.step(x, y)
print('host reference port: %.1f ms per step at N=%d (%d threads)' % ((time.perf_counter() - t0) / 3 * 1e3, N, torch.get_num_threads()))
