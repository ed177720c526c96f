"""The reference's notebook shapes (K = 3, N = 1,500) per full-batch Adam step / evaluation pass, device-resident:
register-resident training kernel (default) against the generic tile kernel (CNF_FP32R_TRAIN=off)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ['CNF_LIVE_ENV'] = '1'
import numpy as np, torch, cnf_b200
dev = torch.device('cuda:0')
rs = np.random.RandomState(0)
def data(N, K):
    y = rs.randint(0, K, size=N)
    x = (1.5 * rs.randn(N, K)).astype(np.float32)
    x[np.arange(N), y] += 3.0 * (rs.rand(N) < 0.8)
    return torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
def timeit(f, reps=300):
    for _ in range(20): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e6
CASES = ((3, 5, [3, 3], True, 1500), (3, 5, [3, 3], False, 1500), (3, 10, [5, 5], True, 1500), (3, 4, [32], False, 10000),
         (5, 6, [5, 5], True, 5000))
if os.environ.get('SINGLE'):      # single-hidden-layer shapes at calibration-set sizes: where does the register kernel win?
    CASES = tuple((K, L, [H], sc, N) for (K, L, sc) in ((3, 4, False), (10, 6, True)) for H in (16, 32, 64, 128) for N in (5000, 20000))
for K, L, hidden, scale, N in CASES:
    x, y = data(N, K)
    out = []
    for sw in ('0' if len(hidden) == 1 else None, 'off'):
        if sw is None: os.environ.pop('CNF_FP32R_TRAIN', None)
        else: os.environ['CNF_FP32R_TRAIN'] = sw
        torch.manual_seed(1)
        flow = cnf_b200.Flow([cnf_b200.NvpCouplingLayer(K, hidden, scale=scale) for _ in range(L)]).to(dev)
        tr = cnf_b200.FusedNLLTrainer(flow.engine(), x, y)
        out.append((timeit(tr.step), timeit(tr.evaluate)))
    print('K=%d L=%d hidden=%-8s scale=%-5s N=%-6d step / evaluation us: register kernel %.1f / %.1f   tile kernel %.1f / %.1f'
          % (K, L, hidden, scale, N, out[0][0], out[0][1], out[1][0], out[1][1]), flush=True)
