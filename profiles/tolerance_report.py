"""bf16 tensor-core path vs fp32 path vs the float64 oracle at the BASELINE shapes (run on the GPU box):
    python profiles/tolerance_report.py > profiles/r01_tolerance_report.txt
Errors are max|a-b| / max|b| for z, max|a-b| / max(1, max|b|) for log-det, max abs for softmax(z)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np, torch, cnf_b200
import flow_oracle as orc

dev = torch.device('cuda:0')
print('%-34s %-6s %10s %10s %10s %10s' % ('config', 'path', 'z', 'logdet', 'probs', 'roundtrip'))
for name, K, L, H, scale, wmul, N in (('C1 NICE K=3 L=4 H=32', 3, 4, [32], False, 300.0, 10000),
                                      ('C2 RealNVP K=10 L=6 H=128', 10, 6, [128], True, 300.0, 20000),
                                      ('C2 at reference init (x1)', 10, 6, [128], True, 1.0, 20000),
                                      ('C4 RealNVP K=100 L=8 H=512', 100, 8, [512], True, 60.0, 2000)):
    torch.manual_seed(K)
    flow = cnf_b200.CouplingStack(K, layers=L, hidden_size=H, scale=scale)
    with torch.no_grad():
        for p in flow.parameters():
            if p.requires_grad:
                p.mul_(wmul)
    flat = np.concatenate([p.detach().numpy().reshape(-1) for lay in flow.layers for p in lay.canonical_parameters()])
    params = orc.unflatten(flat.astype(np.float64), orc.init_params(K, L, H, scale, True))
    x, _ = orc.synth_logits(N, K, seed=K)
    zo, ldo = orc.flow_forward(params, x.astype(np.float64))
    po = orc.softmax(zo[-1])
    flow.to(dev)
    eng = flow.engine()
    xt = torch.from_numpy(x).to(dev)
    for path in ('fp32', 'bf16'):
        z, ld, _ = eng.apply(xt, precision=path)
        xr, _, _ = eng.apply(z, inverse=True, precision=path)
        zz = z.cpu().numpy().astype(np.float64)
        ez = np.max(np.abs(zz - zo[-1])) / np.max(np.abs(zo[-1]))
        el = np.max(np.abs(ld.cpu().numpy() - ldo)) / max(1.0, np.max(np.abs(ldo)))
        ep = np.max(np.abs(orc.softmax(zz) - po))
        er = float((xr - xt).abs().max() / xt.abs().max())
        print('%-34s %-6s %10.2e %10.2e %10.2e %10.2e' % (name, path, ez, el, ep, er))
print('\nstated tolerances: fp32 1e-5 (z, logdet), 5e-5 round trip; bf16 1e-2 (z, logdet, probs absolute)')
