"""ncu launch-list target: full-batch Adam steps at the notebook sizes (N=1500, K=3; RealNVP 5 x [3,3], NICE 5 x [3,3],
10 x NvpCouplingLayer(3, [5,5]))."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch, cnf_b200
dev = torch.device('cuda:0')
rs = np.random.RandomState(0)
N, K = 1500, 3
y = rs.randint(0, K, size=N)
x = (1.5 * rs.randn(N, K)).astype(np.float32)
xt, yt = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
which = os.environ.get('WHICH', 'nvp')
if which == 'nvp':
    flow = cnf_b200.RealNvpFlow(K, layers=5, hidden_size=[3, 3]).to(dev)
elif which == 'nice':
    flow = cnf_b200.NiceFlow(K, layers=5, hidden_size=[3, 3]).to(dev)
else:
    flow = cnf_b200.Flow([cnf_b200.NvpCouplingLayer(3, [5, 5]) for _ in range(10)]).to(dev)
tr = cnf_b200.FusedNLLTrainer(flow.engine(), xt, yt)
for _ in range(8):
    tr.step()
torch.cuda.synchronize()
print('done')
