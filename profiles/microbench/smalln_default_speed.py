"""Per-step latency at calibration-set sizes for the reference's default conditioner (K = 10, 6 couplings,
hidden_size=[5, 5]): full-batch Adam step and evaluation pass, fp32; CNF_FP32R_TRAIN=off / CNF_FP32R=off give the generic
kernels these shapes used before."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, cnf_b200, bench
dev = torch.device('cuda:0')
for n in (1500, 5000, 10000, 50000):
    xt, yt = bench.synth_dev(n, 77, dev)
    m = bench.make_model(seed=2, wmult=1.0, hidden=[5, 5]).to(dev)
    tr = cnf_b200.FusedNLLTrainer(m.engine(), xt, yt)
    for _ in range(5): tr.step(); tr.evaluate()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200): tr.step()
    torch.cuda.synchronize(); ds = (time.perf_counter() - t0) / 200
    t0 = time.perf_counter()
    for _ in range(200): tr.evaluate()
    torch.cuda.synchronize(); de = (time.perf_counter() - t0) / 200
    print('N=%6d  step %.1f us   evaluation pass %.1f us' % (n, ds * 1e6, de * 1e6))
