"""The two fits whose wall-clock prints are the reference's only performance evidence:
  * TorchFlowCalibrator(RealNvpFlow, N=1500, K=3, layers=5, hidden_size=[3,3], epochs=5000): 188.3 s on the author's CPU
    (notebooks/simulated-predictions-flows.ipynb:165, 214); NiceFlow: 160.7 s
  * Flow([NvpCouplingLayer(3, [5,5])] x 10), N=1500, full-batch Adam steps: 22 ms/step (logits-to-categorical.ipynb:647)
through the drop-in API, host numpy in / out."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch, cnf_b200
dev = torch.device('cuda:0')
rs = np.random.RandomState(0)
N, K = 1500, 3
y = rs.randint(0, K, size=N)
x = (1.5 * rs.randn(N, K)).astype(np.float32)
x[np.arange(N), y] += 3.0 * (rs.rand(N) < 0.8)
t = np.eye(K, dtype=np.float32)[y]
epochs = int(os.environ.get('EPOCHS', 5000))
for name, fac in (('RealNvpFlow', cnf_b200.RealNvpFlow), ('NiceFlow', cnf_b200.NiceFlow)):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        cal = cnf_b200.TorchFlowCalibrator(fac, x, t, layers=5, hidden_size=[3, 3], epochs=epochs, dev=dev)
        p = cal.predict(x)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('%-12s fit %d epochs + predict: %.3f s  (%.1f us per epoch)  final loss %.4f' % (name, epochs, dt, dt / epochs * 1e6, cal.history['loss'][-1]))
flow = cnf_b200.Flow([cnf_b200.NvpCouplingLayer(3, [5, 5]) for _ in range(10)]).to(dev)
xt, yt = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
tr = cnf_b200.FusedNLLTrainer(flow.engine(), xt, yt, eps=0.0, gamma=1.0, lr=1e-4)
for _ in range(20): tr.step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(500): tr.step()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 500
print('Flow([NvpCouplingLayer(3,[5,5])]x10) N=1500: %.1f us per full-batch Adam step' % (dt * 1e6))
