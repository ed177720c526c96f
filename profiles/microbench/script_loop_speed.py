"""The script loop of run_experiment3D.py:98-135 through the drop-in Flow (autograd + torch Adam), K = 3, 10 couplings of
hidden [5, 5], N = 1,500: time per step with the register-resident backward kernel and with the tile kernel."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ['CNF_LIVE_ENV'] = '1'
import numpy as np, torch, cnf_b200
dev = torch.device('cuda:0')
rs = np.random.RandomState(0)
N, K = 1500, 3
y = rs.randint(0, K, size=N)
x = (1.5 * rs.randn(N, K)).astype(np.float32)
x[np.arange(N), y] += 3.0 * (rs.rand(N) < 0.8)
xt, yt = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
for sw in (None, 'off'):
    if sw: os.environ['CNF_FP32R_TRAIN'] = sw
    torch.manual_seed(1)
    flow = cnf_b200.Flow([cnf_b200.NvpCouplingLayer(K, [5, 5]) for _ in range(10)]).to(dev)
    opt = torch.optim.Adam(flow.parameters(), lr=1e-4)
    ce = torch.nn.CrossEntropyLoss()
    def step():
        opt.zero_grad()
        zs, ld = flow(xt)
        loss = ce(zs[-1], yt) - ld.mean()
        loss.backward()
        opt.step()
    for _ in range(30): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(300): step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 300
    print('backward on the %s kernel: %.1f us per script step (forward + CE + backward + torch Adam)' % ('tile' if sw else 'register', dt * 1e6), flush=True)
