"""C1-sized calibrator fit (K=3, N=10,000, NICE, 4 couplings, hidden 32), 1000 full-batch epochs (the reference's
default): captured CUDA-graph epochs vs eager launches."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch, cnf_b200
dev = torch.device('cuda:0')
rs = np.random.RandomState(3)
y = rs.randint(0, 3, size=10_000)
x = (1.5 * rs.randn(10_000, 3)).astype(np.float32)
x[np.arange(10_000), y] += 3.0 * (rs.rand(10_000) < 0.8)
t = np.eye(3, dtype=np.float32)[y]
for graph in (True, False, True, False):
    torch.manual_seed(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cal = cnf_b200.TorchFlowCalibrator(cnf_b200.NiceFlow, x, t, layers=4, hidden_size=[32], epochs=1000, dev=dev, cuda_graph=graph)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('cuda_graph=%s  fit 1000 epochs: %.1f ms (%.1f us per epoch)  final loss %.5f' % (graph, dt * 1e3, dt * 1e3, float(cal.history['loss'][-1])))
