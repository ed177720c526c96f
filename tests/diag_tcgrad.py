"""Per-tensor error of the tensor-core training gradient (bf16) against the float64 oracle.
Diagnostic (test infrastructure, uses the oracle as the checker).  Run on the GPU box from the repo root:
    python tests/diag_tcgrad.py [N]"""
import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
import flow_oracle as orc
from conftest import load_golden, oracle_params_from_golden
from helpers import build_flow_from_golden
dev = torch.device('cuda:0')
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
g = load_golden('flow_c2_nvp_k10')
flow = build_flow_from_golden(g, dev, precision='bf16')
eng = flow.engine()
p = oracle_params_from_golden(g, np.float64)
xn, yn = orc.synth_logits(N, 10, seed=5)
x = torch.from_numpy(xn).to(dev); y = torch.from_numpy(yn).to(dev)
eng.ensure(dev)
for tag, eps, gamma in [('cal', 1e-7, 1.0), ('nodet', 0.0, 0.0)]:
    _, _, _, gref, _ = orc.train_step_grads(p, xn.astype(np.float64), yn, eps=eps, gamma=gamma)
    ref = orc.flatten(gref)
    res = {}
    for prec in ('bf16', 'fp32'):
        eng.pack(tc=(prec == 'bf16'))
        acc = torch.zeros(4, dtype=torch.float64, device=dev)
        eng.nll_step(x, y, acc, eps=eps, gamma=gamma, precision=prec)
        res[prec] = eng.flat_grad.cpu().numpy().copy()
    K, H, L = 10, 128, 6
    off = 0
    print(tag, 'N', N, 'overall bf16 %.2e fp32 %.2e' % (np.abs(res['bf16'] - ref).max() / np.abs(ref).max(), np.abs(res['fp32'] - ref).max() / np.abs(ref).max()))
    worst = {}
    for l in range(L):
        for net in 'st':
            for nm, n in [('W0', H * K), ('b0', H), ('W1', K * H), ('b1', K)]:
                b = ref[off:off + n]
                e = np.abs(res['bf16'][off:off + n] - b).max() / max(np.abs(b).max(), 1e-30)
                e32 = np.abs(res['fp32'][off:off + n] - b).max() / max(np.abs(b).max(), 1e-30)
                k = net + nm
                worst[k] = max(worst.get(k, (0, 0)), (e, e32))
                off += n
    for k, v in worst.items():
        print('   %s worst-over-layers rel-to-own-max: bf16 %.4f   fp32 %.2e' % (k, v[0], v[1]))
