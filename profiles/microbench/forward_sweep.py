"""fp32 forward (+ log-det) latency over a sweep of shapes at N=10,000 against the reference op sequence on the host."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch, cnf_b200
import ref_port_torch as rp
dev = torch.device('cuda:0')
torch.set_num_threads(os.cpu_count())
N = 10000
for (K, L, hidden) in ((10, 4, [5, 5]), (10, 6, [128]), (10, 6, [64, 64]), (10, 6, [128, 128]), (100, 4, [100, 100]),
                       (100, 4, [64, 64, 64]), (30, 6, [256]), (100, 8, [512])):
    g = torch.Generator().manual_seed(1)
    x = 1.5 * torch.randn(N, K, generator=g)
    torch.manual_seed(2)
    layers = [cnf_b200.NvpCouplingLayer(K, hidden) for _ in range(L)]
    flow = cnf_b200.Flow(layers).to(dev)
    xt = x.to(dev)
    with torch.no_grad():
        for _ in range(3): flow(xt)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20): flow(xt)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    flat = torch.cat([p.detach().cpu().reshape(-1) for lay in layers for p in lay.canonical_parameters()])
    lay_cpu = rp.unpack(flat, K, L, hidden) if hasattr(rp, 'unpack') else None
    print('K=%3d L=%d hidden=%-14s N=%d: fp32 forward %.3f ms' % (K, L, hidden, N, dt * 1e3))
