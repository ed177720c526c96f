"""GPU tests of the fused predict / metrics pass (cnf_flow_predict): row-mean centring
prologue + flow + Calibrator.predict tail + ECE / NLL / accuracy statistics in ONE launch, against
  * the separate kernels (bit-exact: same arithmetic, z just never leaves the SM),
  * the reference-generated golden fixtures (tests/golden/calibrator_*.npz `pred`) and the oracle's metrics.
Tolerances: fp32 path 2e-5 absolute on probabilities (the golden test's own bound), integer statistics
bit-exact; bf16 tensor-core path 1e-2 absolute on probabilities (the stated bf16 tolerance)."""
import numpy as np
import pytest

import flow_oracle as orc
from conftest import load_golden, oracle_params_from_golden
from helpers import build_flow_from_golden

pytestmark = pytest.mark.gpu


def _raw_logits(n, K, seed):
    x, y = orc.synth_logits(n, K, seed=seed)
    rs = np.random.RandomState(seed)
    x = (x + rs.randn(n, 1).astype(np.float32) * 2.0).astype(np.float32)     # un-centre: a per-row offset
    return x, y


@pytest.mark.parametrize('name,precision', [('c2_nvp_k10', 'fp32'), ('c2_nvp_k10', 'bf16'), ('c1_nice_k3', 'fp32'),
                                            ('nvp_k5_oddL', 'fp32'), ('nvp_k7_randflip', 'fp32'),
                                            ('nvp_k40_wide', 'fp32'), ('nvp_k7_randflip', 'bf16')])
@pytest.mark.parametrize('N', [1, 127, 1000, 33333, 262145])
def test_fused_pass_equals_separate_kernels(name, precision, N, cuda_device):
    import torch
    from cnf_b200 import _lib
    from cnf_b200.utils import metrics as M
    g = load_golden('flow_' + name)
    K = int(g['K'])
    flow = build_flow_from_golden(g, cuda_device, precision=precision)
    eng = flow.engine()
    if precision == 'bf16' and eng.tc_bytes == 0:
        pytest.skip('shape not on the tensor-core path')
    x, y = _raw_logits(N, K, seed=5 + N)
    lp = orc.log_priors(orc.onehot_encode(np.concatenate([y, np.arange(K)])))
    xt, yt = torch.from_numpy(x).to(cuda_device), torch.from_numpy(y).to(cuda_device)
    res = eng.predict(xt, center=True, log_priors=lp, y=yt, bins=15, want_z=True, want_probs=True,
                      precision=precision)
    # separate passes: numpy centring (calibrators.py:42), flow kernel, calibrated-probability kernel, metrics kernel
    xc = x - x.mean(axis=1, keepdims=True)
    assert xc.dtype == np.float32
    z2, ld2, _ = eng.apply(torch.from_numpy(xc).to(cuda_device), precision=precision)
    if N > 32768 or precision == 'bf16':
        # the same kernel serves both calls: centring prologue + flow must reproduce numpy centring + flow bit for bit
        assert torch.equal(res['z'], z2) and torch.equal(res['logdet'], ld2)
    else:
        # fp32 batches of <= 32,768 rows go to the 32-sample-tile kernel in apply(): same arithmetic, another summation order
        assert float((res['z'] - z2).abs().max()) <= 1e-5 * float(z2.abs().max())
        assert float((res['logdet'] - ld2).abs().max()) <= 1e-5 * max(1.0, float(ld2.abs().max()))
    z2 = res['z']
    st2 = M.statistics(z2, yt, bins=15, mode=_lib.METRICS_CALIBRATED, log_priors=lp, device=cuda_device)
    st = res['stats'].cpu().numpy()
    st2 = st2.cpu().numpy()
    assert np.array_equal(st[:15], st2[:15]) and np.array_equal(st[30:45], st2[30:45])      # counts, correct
    assert st[46] == st2[46] and st[47] == N
    # (the streaming metrics kernel reads rows of 8k words lane-rotated, which reorders the float32 softmax sum)
    tol = 1e-12 if (K * 4) % 32 else 1e-6
    assert np.allclose(st[15:30], st2[15:30], rtol=tol, atol=1e-12)
    assert np.isclose(st[45], st2[45], rtol=tol)
    # probabilities: the same formula as the oracle's predict tail on the same z
    pc = orc.calibrated_probs(z2.cpu().numpy(), lp) if hasattr(orc, 'calibrated_probs') else None
    probs = res['probs'].cpu().numpy()
    assert probs.dtype == np.float64 and np.allclose(probs.sum(axis=1), 1.0, atol=1e-12)
    if pc is not None:
        assert np.max(np.abs(probs - pc)) < 2e-6
    # statistics-only call: nothing but the 48 doubles comes back
    res3 = eng.predict(xt, center=True, log_priors=lp, y=yt, bins=15, precision=precision)
    assert set(res3) == {'stats'}
    st3 = res3['stats'].cpu().numpy()
    assert np.array_equal(st3[:15], st[:15]) and np.allclose(st3, st, rtol=1e-12, atol=1e-12)


def test_fused_logits_mode_and_uncentred(cuda_device):
    """log_priors=None: statistics of softmax(z) (predict_post, calibrators.py:350-353); center=False leaves x as is."""
    import torch
    from cnf_b200 import _lib
    from cnf_b200.utils import metrics as M
    g = load_golden('flow_c2_nvp_k10')
    flow = build_flow_from_golden(g, cuda_device)
    eng = flow.engine()
    x, y = orc.synth_logits(5000, 10, seed=3)
    xt, yt = torch.from_numpy(x).to(cuda_device), torch.from_numpy(y).to(cuda_device)
    res = eng.predict(xt, center=False, y=yt, bins=10, want_z=True)
    z2, ld2, _ = eng.apply(xt)       # (N <= 32768: the 32-sample-tile kernel, another summation order)
    assert float((res['z'] - z2).abs().max()) <= 1e-5 * float(z2.abs().max())
    assert float((res['logdet'] - ld2).abs().max()) <= 1e-5 * max(1.0, float(ld2.abs().max()))
    z2 = res['z']
    st2 = M.statistics(z2, yt, bins=10, mode=_lib.METRICS_LOGITS, device=cuda_device).cpu().numpy()
    st = res['stats'].cpu().numpy()
    assert np.array_equal(st[:10], st2[:10]) and np.array_equal(st[20:30], st2[20:30])
    assert np.allclose(st, st2, rtol=1e-6)


@pytest.mark.parametrize('name', ['cal_nice_k3', 'cal_nvp_k10'])
@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_calibrator_predict_and_evaluate_vs_reference_golden(name, precision, cuda_device):
    """TorchFlowCalibrator.predict / .evaluate through the fused pass against the reference's own `pred`
    (tests/golden/calibrator_*.npz, written by the unmodified reference) and the oracle's metrics on it."""
    import torch
    import cnf_b200
    g = load_golden('calibrator_' + name)
    hidden = [int(h) for h in g['hidden']]

    class Factory(cnf_b200.CouplingStack):
        def __init__(self, dim, **kw):
            super().__init__(dim, layers=int(g['layers']), hidden_size=hidden, scale=bool(g['scale']), **{
                k: v for k, v in kw.items() if k not in ('layers', 'hidden_size', 'scale')})
            flat = torch.from_numpy(g['flat_end'].astype(np.float32))      # the reference's trained weights
            off = 0
            with torch.no_grad():
                for lay in self.layers:
                    for p in lay.canonical_parameters():
                        p.copy_(flat[off:off + p.numel()].view(p.shape))
                        off += p.numel()

    cal = cnf_b200.TorchFlowCalibrator(Factory, g['x'], g['y'], epochs=0, dev=cuda_device)
    cal.force_fused = True        # (fp32 sets of <= 32,768 rows default to the tile kernel + tail kernel: quicker there)
    if precision == 'bf16':
        if cal.flow.engine().tc_bytes == 0:
            pytest.skip('shape not on the tensor-core path')
        cal.precision = 'bf16'
    calls = []
    orig = cal.flow.engine().predict
    cal.flow.engine().predict = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    tol = 2e-5 if precision == 'fp32' else 1e-2
    pred = cal.predict(g['x_test'])
    assert calls, 'predict must take the fused pass'
    assert pred.dtype == np.float64 and np.max(np.abs(pred - g['pred'])) < tol
    rs = np.random.RandomState(0)
    yt = np.where(rs.rand(len(pred)) < 0.7, g['pred'].argmax(1), rs.randint(0, pred.shape[1], len(pred)))
    m = cal.evaluate(g['x_test'], yt, bins=15)
    ref_ece = orc.expected_calibration_error(g['pred'], yt, 15)
    ref_nll = orc.neg_log_likelihood(g['pred'], yt)
    if precision == 'fp32':
        assert abs(m['ece'] - ref_ece) < 1e-5 and abs(m['nll'] - ref_nll) < 1e-5
        assert m['accuracy'] == orc.accuracy(g['pred'], yt) and m['n'] == len(yt)
    else:
        assert abs(m['ece'] - ref_ece) < 2e-2 and abs(m['nll'] - ref_nll) < 2e-2
    # one-hot targets and float64 logits (the host-centring route) give the same numbers
    oh = np.eye(pred.shape[1], dtype=np.int32)[yt]
    m2 = cal.evaluate(g['x_test'].astype(np.float64), oh, bins=15)
    assert abs(m2['ece'] - m['ece']) < (1e-6 if precision == 'fp32' else 2e-2)


def test_boundary_validation_raises_like_the_reference(cuda_device):
    """Wrong shapes / dtypes never reach the kernels (the C ABI reads N*K floats blindly)."""
    import torch
    import cnf_b200
    g = load_golden('flow_c2_nvp_k10')
    flow = build_flow_from_golden(g, cuda_device)
    eng = flow.engine()
    x, y = orc.synth_logits(64, 10, seed=1)
    xt, yt = torch.from_numpy(x).to(cuda_device), torch.from_numpy(y).to(cuda_device)
    with pytest.raises(ValueError):
        eng.apply(xt[:, :9])
    with pytest.raises(ValueError):
        eng.apply(xt.view(-1))
    z64, _, _ = eng.apply(xt.double())                      # float64 is cast, as apply() always did
    z32, _, _ = eng.apply(xt)
    assert torch.equal(z64, z32)
    zt, _, _ = eng.apply(xt.t().contiguous().t())           # non-contiguous view
    assert torch.equal(zt, z32)
    with pytest.raises(ValueError):
        cnf_b200.FusedNLLTrainer(eng, xt.double(), yt)      # the trainer keeps x resident: no silent copy
    with pytest.raises(ValueError):
        cnf_b200.FusedNLLTrainer(eng, xt, yt[:-1])
    with pytest.raises(ValueError):
        cnf_b200.FusedNLLTrainer(eng, xt, yt.float())
    tr = cnf_b200.FusedNLLTrainer(eng, xt, yt.to(torch.int32))     # integer labels are widened
    tr.step()
    assert torch.isfinite(tr.loss_acc).all()
    with pytest.raises(ValueError):
        eng.predict(xt, y=yt[:-1])
    with pytest.raises(ValueError):
        eng.predict(xt, log_priors=np.zeros(9), want_probs=True)
    lay = cnf_b200.AffineConstantLayer(7).to(cuda_device)
    with pytest.raises(ValueError):
        lay(xt)                                             # 10 columns into a 7-dim layer


def test_flow_on_a_non_current_device(cuda_device):
    """A flow on cuda:1 while the current device is cuda:0 (the reference's dev= argument): every call into
    the library runs under a device guard."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    g = load_golden('flow_c2_nvp_k10')
    dev1 = torch.device('cuda:1')
    torch.cuda.set_device(0)
    flow0 = build_flow_from_golden(g, cuda_device)
    flow1 = build_flow_from_golden(g, dev1)
    x, y = orc.synth_logits(3000, 10, seed=2)
    z0, ld0 = flow0(torch.from_numpy(x).to(cuda_device))
    z1, ld1 = flow1(torch.from_numpy(x).to(dev1))
    assert torch.cuda.current_device() == 0
    assert torch.equal(z0[-1].cpu(), z1[-1].cpu()) and torch.equal(ld0.cpu(), ld1.cpu())
    import cnf_b200
    tr = cnf_b200.FusedNLLTrainer(flow1.engine(), torch.from_numpy(x).to(dev1), torch.from_numpy(y).to(dev1))
    tr.step()
    assert torch.isfinite(tr.loss_acc).all()
    p = torch.softmax(z1[-1], dim=1)
    assert abs(cnf_b200.expected_calibration_error(p, torch.from_numpy(y).to(dev1)) -
               cnf_b200.expected_calibration_error(p.cpu().numpy(), y)) < 1e-6


def test_lazy_intermediates_refuse_stale_weights(cuda_device):
    """zs[i < L-1] are materialised lazily: reading them after the weights changed must raise, not return
    the outputs of the new weights; zs[-1] stays valid; a fresh pass works again."""
    import torch
    import cnf_b200
    g = load_golden('flow_c2_nvp_k10')
    flow = build_flow_from_golden(g, cuda_device)
    x = torch.from_numpy(g['x']).to(cuda_device)
    y = torch.from_numpy(g['y']).to(cuda_device)
    with torch.no_grad():
        zs, ld = flow(x)
        z0 = zs[0].clone()                         # read before any update: fine
        assert np.max(np.abs(z0.cpu().numpy() - g['zs'][0])) <= 1e-5 * np.max(np.abs(g['zs'][0]))
        zs2, _ = flow(x)
    tr = cnf_b200.FusedNLLTrainer(flow.engine(), x, y)
    tr.step()                                      # fused Adam through the C ABI: torch's version counter does not move
    with pytest.raises(RuntimeError):
        zs2[1]
    assert torch.equal(zs2[-1], zs[-1])
    with torch.no_grad():
        zs3, _ = flow(x)
        zs4, _ = flow(x)
        first = zs3[1].clone()                     # materialises (and caches) the whole segment with the weights of the pass
        for p in flow.parameters():
            if p.requires_grad:
                p.mul_(1.01)                       # a torch-side in-place update is seen too
        assert torch.equal(zs3[2], zs3[2]) and torch.equal(zs3[1], first)    # cached: still the old weights' outputs
        with pytest.raises(RuntimeError):
            zs4[2]                                 # never materialised before the update
