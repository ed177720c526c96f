"""fp32 training-step and forward speed over a sweep of shapes against the reference op sequence on the host cores:
finds shapes where the CUDA path has no margin."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch, cnf_b200
import ref_port_torch as rp
dev = torch.device('cuda:0')
torch.set_num_threads(os.cpu_count())
N = 10000
for (K, L, hidden) in ((10, 4, [5, 5]), (10, 6, [64, 64]), (10, 6, [128, 128]), (100, 4, [100, 100]), (100, 4, [64, 64, 64]), (30, 6, [256]), (100, 4, [100]), (100, 8, [512])):
    g = torch.Generator().manual_seed(1)
    x = 1.5 * torch.randn(N, K, generator=g); y = torch.randint(0, K, (N,), generator=g)
    torch.manual_seed(2)
    layers = [cnf_b200.NvpCouplingLayer(K, hidden) for _ in range(L)]
    flow = cnf_b200.Flow(layers).to(dev)
    eng = flow.engine()
    tr = cnf_b200.FusedNLLTrainer(eng, x.to(dev), y.to(dev))
    for _ in range(2): tr.step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): tr.step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    flat = torch.cat([p.detach().cpu().reshape(-1) for lay in layers for p in lay.canonical_parameters()])
    st = rp.TrainState(flat, K, L, hidden)
    st.step(x, y); t0 = time.perf_counter()
    for _ in range(2): st.step(x, y)
    dc = (time.perf_counter() - t0) / 2
    print('K=%3d L=%d hidden=%-14s N=%d: train step GPU %.2f ms  host %.1f ms  -> %.0fx' % (K, L, hidden, N, dt * 1e3, dc * 1e3, dc / dt))
