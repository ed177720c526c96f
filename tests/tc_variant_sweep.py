"""Debug helper (not a test): accuracy and speed of the tensor-core kernel's epilogue variants
(CNF_TC_EPI=0 round-to-nearest F2FP, 1 truncate+compensate) against the float64 golden outputs.
    timeout -s KILL 120 python tests/tc_variant_sweep.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import load_golden  # noqa: E402
from helpers import build_flow_from_golden, rel_err  # noqa: E402
import flow_oracle as orc  # noqa: E402
from conftest import oracle_params_from_golden  # noqa: E402

dev = torch.device('cuda:0')
variants = [int(v) for v in sys.argv[1:]] or [0, 1]
for name in ('c2_nvp_k10', 'c1_nice_k3', 'nvp_k2'):
    g = load_golden('flow_' + name)
    flow = build_flow_from_golden(g, dev)
    eng = flow.engine()
    K = int(g['K'])
    xbig, _ = orc.synth_logits(200_000, K, seed=5)
    p = oracle_params_from_golden(g, np.float64)
    zo, ldo = orc.flow_forward(p, xbig.astype(np.float64))
    x = torch.from_numpy(xbig).to(dev)
    xl = torch.from_numpy(orc.synth_logits(1_000_000, K, seed=6)[0]).to(dev)
    z32, ld32, _ = eng.apply(x)
    print(name, 'fp32 path err', rel_err(z32.cpu().numpy(), zo[-1]), 'tc_bytes', eng.tc_bytes, flush=True)
    for v in variants:
        os.environ['CNF_TC_EPI'] = str(v)
        z, ld, _ = eng.apply(x, precision='bf16')
        torch.cuda.synchronize()
        zz, ll = z.cpu().numpy(), ld.cpu().numpy()
        pz = orc.softmax(zz.astype(np.float64)); po = orc.softmax(zo[-1])
        print('  epi %d: z err %.3e  logdet err %.3e  prob abs err %.3e  mean signed logdet err %.2e  finite %s' % (
            v, rel_err(zz, zo[-1]), np.max(np.abs(ll - ldo)) / max(1.0, np.max(np.abs(ldo))),
            np.max(np.abs(pz - po)), float(np.mean(ll - ldo)), bool(np.isfinite(zz).all())), flush=True)
        xr, ldr, _ = eng.apply(z, inverse=True, precision='bf16')
        torch.cuda.synchronize()
        print('         round trip err %.3e' % rel_err(xr.cpu().numpy(), xbig), flush=True)
        for _ in range(3):
            eng.apply(xl, precision='bf16', repack=False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            eng.apply(xl, precision='bf16', repack=False)
        e1.record()
        torch.cuda.synchronize()
        print('         %.3f ms per 1M samples -> %.2f G samples/s' % (e0.elapsed_time(e1) / 20, 20e-3 / e0.elapsed_time(e1) * 1e3 / 1e3 * 1.0), flush=True)
