"""Calibrator fit at a CIFAR-10-sized validation set (K=10, N=5,000, RealNVP 6 couplings, hidden 128), 1000
full-batch epochs through the drop-in API: per-epoch wall time by precision and with / without graph capture."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch, cnf_b200
dev = torch.device('cuda:0')
rs = np.random.RandomState(3)
N, K = 5000, 10
y = rs.randint(0, K, size=N)
x = (1.5 * rs.randn(N, K)).astype(np.float32)
x[np.arange(N), y] += 3.0 * (rs.rand(N) < 0.8)
t = np.eye(K, dtype=np.float32)[y]
cnf_b200.TorchFlowCalibrator(cnf_b200.RealNvpFlow, x, t, layers=6, hidden_size=[128], epochs=3, dev=dev)   # warm the context
for prec in ('fp32', 'bf16'):
    for graph in (False, True):
        torch.manual_seed(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cal = cnf_b200.TorchFlowCalibrator(cnf_b200.RealNvpFlow, x, t, layers=6, hidden_size=[128], epochs=1000, dev=dev,
                                           precision=prec, cuda_graph=graph)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print('%s cuda_graph=%-5s fit 1000 epochs: %7.1f ms (%.1f us per epoch)  final loss %.5f'
              % (prec, graph, dt * 1e3, dt * 1e3, float(cal.history['loss'][-1])))
