"""A few full-batch Adam steps at the reference's notebook shape (K = 3, N = 1,500, 5 x [3, 3]) for the ncu launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch, cnf_b200
dev = torch.device('cuda:0')
rs = np.random.RandomState(0)
N, K = 1500, 3
y = rs.randint(0, K, size=N)
x = (1.5 * rs.randn(N, K)).astype(np.float32)
x[np.arange(N), y] += 3.0 * (rs.rand(N) < 0.8)
xt, yt = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
for scale in (True, False):
    torch.manual_seed(1)
    flow = cnf_b200.Flow([cnf_b200.NvpCouplingLayer(K, [3, 3], scale=scale) for _ in range(5)]).to(dev)
    tr = cnf_b200.FusedNLLTrainer(flow.engine(), xt, yt)
    h = tr.fit_loop(6, N, None)
    torch.cuda.synchronize()
