"""Diagnostic: the notebook fit (N=1500, K=3, 5 couplings, hidden [3,3]) step by step on the GPU; saves the weights
before the first non-finite loss for an offline comparison with the float64 oracle."""
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
x = x - x.mean(axis=1, keepdims=True)
for seed in (0, 1, 2):
    torch.manual_seed(seed)
    flow = cnf_b200.RealNvpFlow(K, layers=5, hidden_size=[3, 3]).to(dev)
    eng = flow.engine()
    xt, yt = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
    tr = cnf_b200.FusedNLLTrainer(eng, xt, yt)
    flat0 = eng.flat.detach().cpu().numpy().copy()
    prev = None
    for ep in range(5000):
        snap = (eng.flat.detach().cpu().numpy().copy(), eng.adam_m.cpu().numpy().copy() if eng.adam_m is not None else None,
                eng.adam_v.cpu().numpy().copy() if eng.adam_v is not None else None) if ep % 50 == 0 or (prev is not None and prev > 5) else None
        tr.step()
        loss = -float(tr.loss_acc[0]) / N
        if snap is not None:
            last_snap, last_ep = snap, ep
        if not np.isfinite(loss) or float(tr.loss_acc[3]) != 0:
            print('seed %d: non-finite loss at epoch %d (prev loss %s), bad count %s' % (seed, ep, prev, float(tr.loss_acc[3])))
            np.savez(os.path.join(ROOT, 'gpurun_out', 'nb_nan_seed%d.npz' % seed), flat0=flat0, flat=last_snap[0], snap_epoch=last_ep,
                     grad=eng.flat_grad.cpu().numpy(), x=x, y=y)
            break
        prev = loss
        if ep % 500 == 0: print('seed %d epoch %d loss %.5f' % (seed, ep, loss))
    else:
        print('seed %d: finite through 5000 epochs, final loss %.5f' % (seed, prev))
