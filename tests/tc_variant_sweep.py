"""Debug helper (not a test): run the tensor-core kernel under each descriptor/packing variant and
print its error against the reference's float64 golden outputs.  Usage on the GPU box:
    timeout -s KILL 120 python tests/tc_variant_sweep.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import load_golden  # noqa: E402
from helpers import build_flow_from_golden, rel_err  # noqa: E402

dev = torch.device('cuda:0')
variants = [int(v) for v in sys.argv[1:]] or [0, 1, 2, 3]
for name in ('c2_nvp_k10', 'c1_nice_k3'):
    g = load_golden('flow_' + name)
    flow = build_flow_from_golden(g, dev)
    eng = flow.engine()
    x = torch.from_numpy(g['x']).to(dev)
    z32, ld32, _ = eng.apply(x)
    print(name, 'fp32 path err', rel_err(z32.cpu().numpy(), g['z64']), 'tc_bytes', eng.tc_bytes, flush=True)
    for v in variants:
        os.environ['CNF_TC_VARIANT'] = str(v)
        z, ld, _ = eng.apply(x, precision='bf16')
        torch.cuda.synchronize()
        zz, ll = z.cpu().numpy(), ld.cpu().numpy()
        print('  variant %d: z err %.3e  logdet err %.3e  finite %s' % (
            v, rel_err(zz, g['z64']), np.max(np.abs(ll - g['logdet64'])) / max(1.0, np.max(np.abs(g['logdet64']))),
            bool(np.isfinite(zz).all())), flush=True)
        xr, ldr, _ = eng.apply(z, inverse=True, precision='bf16')
        torch.cuda.synchronize()
        print('             round trip err %.3e' % rel_err(xr.cpu().numpy(), g['x']), flush=True)
