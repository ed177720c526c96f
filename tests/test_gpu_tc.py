"""GPU tests of the bf16 tcgen05 tensor-core path (precision='bf16').

Stated bf16 tolerance (BASELINE.json north_star): against the float64 oracle,
  |z - z64|           <= 1e-2 * max|z64|
  |logdet - ld64|     <= 1e-2 * max(1, max|ld64|)
  |softmax(z) - p64|  <= 1e-2 absolute
(measured on B200: ~6e-4, ~2.5e-3, ~1e-3 at config C2 with trained-like weights)."""
import numpy as np
import pytest

import flow_oracle as orc
from conftest import load_golden, oracle_params_from_golden
from helpers import build_flow_from_golden, rel_err

pytestmark = pytest.mark.gpu
RTOL, ATOL_P = 1e-2, 1e-2
TC_CASES = ['c2_nvp_k10', 'c2_nvp_k10_init', 'c1_nice_k3', 'nvp_k2', 'nvp_k7_randflip']


@pytest.mark.parametrize('name', TC_CASES)
def test_tc_forward_inverse_vs_float64_oracle(name, cuda_device):
    import torch
    g = load_golden('flow_' + name)
    flow = build_flow_from_golden(g, cuda_device, precision='bf16')
    eng = flow.engine()
    assert eng.tc_bytes > 0, 'tensor-core path should cover this shape'
    p = oracle_params_from_golden(g, np.float64)
    K = int(g['K'])
    for N in (1, 127, 128, 129, 5000, 33333):
        x, _ = orc.synth_logits(N, K, seed=N)
        zo, ldo = orc.flow_forward(p, x.astype(np.float64))
        with torch.no_grad():
            zs, ld = flow(torch.from_numpy(x).to(cuda_device))
            z = zs[-1]
            xr, ldr = flow.backward(z)
        zz = z.cpu().numpy()
        assert np.isfinite(zz).all()
        assert rel_err(zz, zo[-1]) < RTOL
        assert np.max(np.abs(ld.cpu().numpy().reshape(-1) - ldo)) < RTOL * max(1.0, np.max(np.abs(ldo)))
        assert np.max(np.abs(orc.softmax(zz.astype(np.float64)) - orc.softmax(zo[-1]))) < ATOL_P
        # the inverse of the same (bf16) map: round trip and log-det antisymmetry
        assert rel_err(xr[-1].cpu().numpy(), x) < RTOL
        assert float((ld + ldr).abs().max()) < RTOL * max(1.0, np.max(np.abs(ldo)))
        if not int(g['scale']):
            assert float(ld.abs().max()) == 0.0


def test_tc_path_rejects_unsupported_shapes_loudly(cuda_device):
    import torch
    import cnf_b200
    flow = cnf_b200.Flow([cnf_b200.NvpCouplingLayer(10, [5, 5]) for _ in range(2)], precision='bf16').to(cuda_device)
    assert flow.engine().tc_bytes == 0
    with torch.no_grad(), pytest.raises(NotImplementedError, match='tensor-core'):
        flow(torch.zeros(4, 10, device=cuda_device))


def test_tc_matches_fp32_path_statistically_at_c2_size(cuda_device):
    """Full C2 size: the bf16 path's calibrated metrics agree with the fp32 path's."""
    import torch
    import cnf_b200
    from cnf_b200 import _lib
    from cnf_b200.utils import metrics as M
    g = load_golden('flow_c2_nvp_k10')
    flow = build_flow_from_golden(g, cuda_device)
    x, y = orc.synth_logits(1_000_000, 10, seed=4)
    xt = torch.from_numpy(x).to(cuda_device)
    lp = orc.log_priors(orc.onehot_encode(y))
    eng = flow.engine()
    z32, ld32, _ = eng.apply(xt)
    zbf, ldbf, _ = eng.apply(xt, precision='bf16')
    assert float((zbf - z32).abs().max() / z32.abs().max()) < RTOL
    assert float((ldbf - ld32).abs().max()) < RTOL * max(1.0, float(ld32.abs().max()))
    s32 = M.statistics(z32, y, bins=15, mode=_lib.METRICS_CALIBRATED, log_priors=lp).cpu().numpy()
    sbf = M.statistics(zbf, y, bins=15, mode=_lib.METRICS_CALIBRATED, log_priors=lp).cpu().numpy()
    assert abs(M.ece_from_statistics(s32, 15) - M.ece_from_statistics(sbf, 15)) < 1e-3
    assert abs(s32[45] / s32[47] - sbf[45] / sbf[47]) < 1e-3          # NLL
    assert abs(s32[46] - sbf[46]) / s32[47] < 1e-3                    # accuracy


@pytest.mark.parametrize('K,L,hidden,scale,shift,wmul', [(100, 8, [512], True, True, 60.0), (33, 3, [200], False, True, 100.0),
                                                         (40, 2, [24], True, True, 200.0), (126, 2, [130], True, True, 60.0),
                                                         (20, 4, [300], True, False, 100.0)])
def test_tc_wide_kernel_vs_float64_oracle(K, L, hidden, scale, shift, wmul, cuda_device):
    """Shapes beyond the resident-weight kernel run on the streamed-weight kernel (cnf_flow_tcw.cu)."""
    import torch
    import cnf_b200
    torch.manual_seed(K + L)
    flow = cnf_b200.Flow([cnf_b200.NvpCouplingLayer(K, hidden, scale=scale, shift=shift) for _ in range(L)],
                         precision='bf16')
    with torch.no_grad():
        for p in flow.parameters():
            if p.requires_grad:
                p.mul_(wmul)
    like = orc.init_params(K, L, hidden, scale, shift)
    flat = np.concatenate([p.detach().numpy().reshape(-1) for lay in flow.layers for p in lay.canonical_parameters()])
    params = orc.unflatten(flat.astype(np.float64), like)
    flow.to(cuda_device)
    assert flow.engine().tc_bytes > 0
    for N in (1, 300, 128 * 148 * 2 + 77):
        x, _ = orc.synth_logits(N, K, seed=3 + N)
        with torch.no_grad():
            zs, ld = flow(torch.from_numpy(x).to(cuda_device))
            xr, ldr = flow.backward(zs[-1])
        zo, ldo = orc.flow_forward(params, x.astype(np.float64))
        zz = zs[-1].cpu().numpy()
        assert np.isfinite(zz).all()
        assert rel_err(zz, zo[-1]) < RTOL
        assert np.max(np.abs(ld.cpu().numpy().reshape(-1) - ldo)) < RTOL * max(1.0, np.max(np.abs(ldo)))
        assert rel_err(xr[-1].cpu().numpy(), x) < RTOL
