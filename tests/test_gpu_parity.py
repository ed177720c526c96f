"""GPU parity tests: the CUDA path (through the drop-in modules -> C ABI) against the golden
fixtures generated from the reference, and against the numpy oracle on seeded inputs.

Tolerances (stated, per BASELINE.json north_star): fp32 path 1e-5 relative, measured as
max|a-b| / max|b| (per-element relative error is meaningless where log_det ~ 0)."""
import numpy as np
import pytest

import flow_oracle as orc
from conftest import golden_flow_names, load_golden, oracle_params_from_golden
from helpers import build_flow_from_golden, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def ld_err(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - b)) / max(1.0, np.max(np.abs(b))))


@pytest.mark.parametrize('name', golden_flow_names())
def test_forward_inverse_vs_reference_golden(name, cuda_device):
    import torch
    g = load_golden('flow_' + name)
    flow = build_flow_from_golden(g, cuda_device)
    x = torch.from_numpy(g['x']).to(cuda_device)
    with torch.no_grad():
        zs, ld = flow(x)
        assert len(zs) == int(g['L'])
        assert rel_err(zs[-1].cpu().numpy(), g['z64']) < TOL
        assert ld_err(ld.cpu().numpy(), g['logdet64']) < TOL
        for l in range(int(g['L'])):
            assert rel_err(zs[l].cpu().numpy(), g['zs'][l]) < TOL
        xs, ldi = flow.backward(torch.from_numpy(g['zs'][-1]).to(cuda_device))
        for l in range(int(g['L'])):
            assert rel_err(xs[l].cpu().numpy(), g['xs'][l]) < 5 * TOL
        assert ld_err(ldi.cpu().numpy(), g['logdet_inv'].astype(np.float64)) < TOL
        # round trip and log-det antisymmetry
        xr, ldr = flow.backward(zs[-1])
        assert rel_err(xr[-1].cpu().numpy(), g['x']) < 5 * TOL
        assert ld_err((ld + ldr).cpu().numpy(), np.zeros(1)) < 5 * TOL
        if not int(g['scale']):
            assert float(ld.abs().max()) == 0.0          # NICE: log-det exactly zero
        # layer-level API, and N == 1 gives a 0-d log-det (flows/flows.py:109)
        z1, ld1 = flow.layers[0](x)
        assert rel_err(z1.cpu().numpy(), g['zs'][0]) < TOL
        _, ld_one = flow(x[:1])
        assert ld_one.dim() == 0
        xb, _ = flow.layers[0].backward(z1)
        assert rel_err(xb.cpu().numpy(), g['x']) < 5 * TOL


def test_odd_k_middle_dim_untouched_and_odd_l_reversal(cuda_device):
    import torch
    g = load_golden('flow_nvp_k5_oddL')
    flow = build_flow_from_golden(g, cuda_device)
    x = torch.from_numpy(g['x']).to(cuda_device)
    with torch.no_grad():
        zs, _ = flow(x)
    z = zs[-1].cpu().numpy()
    # K=5: physical slot 2 is never transformed; L=3 is odd so the output order is reversed
    assert np.array_equal(z[:, 2], g['x'][:, 2])


@pytest.mark.parametrize('N', [0, 1, 127, 128, 129, 1000, 70001])
def test_ragged_sizes_vs_oracle(N, cuda_device):
    import torch
    g = load_golden('flow_c2_nvp_k10')
    flow = build_flow_from_golden(g, cuda_device)
    p = oracle_params_from_golden(g, np.float64)
    x, _ = orc.synth_logits(max(N, 1), 10, seed=N + 5)
    x = x[:N]
    with torch.no_grad():
        zs, ld = flow(torch.from_numpy(x).to(cuda_device))
        z = zs[-1].cpu().numpy()
    assert z.shape == (N, 10)
    if N == 0:
        return
    zo, ldo = orc.flow_forward(p, x.astype(np.float64))
    assert rel_err(z, zo[-1]) < TOL
    assert ld_err(ld.cpu().numpy().reshape(-1), ldo) < TOL


@pytest.mark.parametrize('K,L,hidden,scale,shift', [(100, 8, [512], True, True), (100, 2, [64, 64], True, True),
                                                    (33, 3, [200], False, True), (10, 4, [300, 40, 17], True, True)])
def test_wide_shapes_vs_oracle(K, L, hidden, scale, shift, cuda_device):
    import torch
    import cnf_b200
    torch.manual_seed(K + L)
    flow = cnf_b200.Flow([cnf_b200.NvpCouplingLayer(K, hidden, scale=scale, shift=shift) for _ in range(L)])
    with torch.no_grad():
        for p in flow.parameters():
            if p.requires_grad:
                p.mul_(100.0)
    like = orc.init_params(K, L, hidden, scale, shift)
    flat = np.concatenate([p.detach().numpy().reshape(-1) for lay in flow.layers for p in lay.canonical_parameters()])
    params = orc.unflatten(flat.astype(np.float64), like)
    x, _ = orc.synth_logits(300, K, seed=3)
    flow.to(cuda_device)
    with torch.no_grad():
        zs, ld = flow(torch.from_numpy(x).to(cuda_device))
        xs, ldi = flow.backward(zs[-1])
    zo, ldo = orc.flow_forward(params, x.astype(np.float64))
    assert rel_err(zs[-1].cpu().numpy(), zo[-1]) < TOL
    assert ld_err(ld.cpu().numpy(), ldo) < TOL
    assert rel_err(xs[-1].cpu().numpy(), x) < 1e-4


@pytest.mark.parametrize('name', golden_flow_names())
@pytest.mark.parametrize('tag,eps,gamma', [('cal', 1e-7, 1.0), ('script', 0.0, 1.0), ('script_nodet', 0.0, 0.0)])
def test_fused_train_step_gradients_vs_reference_autograd(name, tag, eps, gamma, cuda_device):
    import torch
    g = load_golden('flow_' + name)
    flow = build_flow_from_golden(g, cuda_device)
    eng = flow.engine()
    x = torch.from_numpy(g['x']).to(cuda_device)
    y = torch.from_numpy(g['y']).to(cuda_device)
    eng.ensure(cuda_device)
    eng.pack()
    acc = torch.zeros(4, dtype=torch.float64, device=cuda_device)
    eng.nll_step(x, y, acc, eps=eps, gamma=gamma)
    N = x.shape[0]
    loss = -float(acc[0]) / N
    assert abs(loss - float(g['loss_' + tag])) < 1e-5 * max(1.0, abs(float(g['loss_' + tag])))
    grad = eng.flat_grad.cpu().numpy()
    ref = g['grad_' + tag]
    assert rel_err(grad, ref) < 2e-4
    assert np.all(grad[ref == 0] == 0)                 # dead weights: exactly zero
    assert float(acc[3]) == 0.0


@pytest.mark.parametrize('name', ['c1_nice_k3', 'nvp_k5_oddL', 'c2_nvp_k10', 'nvp_k7_randflip', 'scaleonly_k4_h888',
                                  'nvp_k10_nohidden'])
def test_autograd_through_drop_in_modules(name, cuda_device):
    """User code doing loss.backward() on the drop-in Flow (run_experiment3D.py:102-107,133-135)."""
    import torch
    g = load_golden('flow_' + name)
    flow = build_flow_from_golden(g, cuda_device)
    x = torch.from_numpy(g['x']).to(cuda_device).requires_grad_(True)
    y = torch.from_numpy(g['y']).to(cuda_device)
    zs, ld = flow(x)
    loss = torch.nn.CrossEntropyLoss()(zs[-1], y) - torch.mean(ld)
    loss.backward()
    assert abs(float(loss) - float(g['loss_script'])) < 1e-5 * max(1.0, abs(float(g['loss_script'])))
    grad = np.concatenate([(p.grad if p.grad is not None else torch.zeros_like(p)).cpu().numpy().reshape(-1)
                           for lay in flow.layers for p in lay.canonical_parameters()])
    assert rel_err(grad, g['grad_script']) < 2e-4
    assert rel_err(x.grad.cpu().numpy(), g['gx_script']) < 2e-4
    # a torch optimiser stepping the aliased parameters is seen by the next forward
    opt = torch.optim.SGD([p for p in flow.parameters() if p.requires_grad], lr=1e-2, weight_decay=1e-2)
    opt.step()
    opt.zero_grad()
    zs, ld = flow(x.detach())
    loss = torch.nn.CrossEntropyLoss()(zs[-1], y) - torch.mean(ld)
    loss.backward()
    opt.step()
    flat = np.concatenate([p.detach().cpu().numpy().reshape(-1) for lay in flow.layers
                           for p in lay.canonical_parameters()])
    assert rel_err(flat - g['flat'], g['sgd2_flat'] - g['flat']) < 2e-3


@pytest.mark.parametrize('name', golden_flow_names())
def test_fused_adam_and_sgd_vs_torch_optim(name, cuda_device):
    import torch
    from cnf_b200 import FusedNLLTrainer
    g = load_golden('flow_' + name)
    flow = build_flow_from_golden(g, cuda_device)
    x = torch.from_numpy(g['x']).to(cuda_device)
    y = torch.from_numpy(g['y']).to(cuda_device)
    tr = FusedNLLTrainer(flow.engine(), x, y)
    losses = []
    for _ in range(3):
        tr.step()
        losses.append(-float(tr.loss_acc[0]) / x.shape[0])
    assert np.allclose(losses, g['adam3_losses'], rtol=1e-4, atol=1e-5)
    flat = flow.engine().flat.cpu().numpy()
    disp, disp_ref = flat - g['flat'], g['adam3_flat'] - g['flat']
    # Adam normalises every entry to ~lr per step, so an entry whose gradient is rounding noise can land +-lr either
    # way: all entries within 5 % of the largest displacement, and the entries with a gradient well above the noise
    # (> 1e-3 of the largest) -- where the update is a smooth function of the gradient -- within 5e-3 of it
    scale = np.max(np.abs(disp_ref))
    assert np.max(np.abs(disp - disp_ref)) < 0.05 * scale
    solid = np.abs(g['grad_cal']) > 1e-3 * np.max(np.abs(g['grad_cal']))
    assert solid.any() and np.max(np.abs(disp - disp_ref)[solid]) < 5e-3 * scale, np.max(np.abs(disp - disp_ref)[solid]) / scale
    assert np.all(disp[g['grad_cal'] == 0] == 0)
    # the nn.Parameters alias the flat buffer: state_dict sees the update
    p0 = flow.layers[0].canonical_parameters()[0]
    assert np.array_equal(p0.detach().cpu().numpy().reshape(-1), flat[:p0.numel()])

    flow = build_flow_from_golden(g, cuda_device)
    tr = FusedNLLTrainer(flow.engine(), x, y, eps=0.0, gamma=1.0, lr=1e-2, weight_decay=1e-2, optim='sgd')
    for _ in range(2):
        tr.step()
    flat = flow.engine().flat.cpu().numpy()
    assert rel_err(flat - g['flat'], g['sgd2_flat'] - g['flat']) < 2e-3


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
@pytest.mark.parametrize('name', ['c2_nvp_k10', 'c1_nice_k3', 'nvp_k7_randflip'])
def test_one_launch_optimiser_tail_is_bitwise_the_three_launch_tail(name, precision, cuda_device):
    """cnf_reduce_adam_pack_rows / cnf_reduce_adam_pack_tc (reduce + Adam + repack in one launch) against
    cnf_grad_reduce_* + cnf_adam_step + cnf_pack_weights* on the same partial rows."""
    import torch
    from cnf_b200 import FusedNLLTrainer
    g = load_golden('flow_' + name)
    x = torch.from_numpy(g['x']).to(cuda_device)
    y = torch.from_numpy(g['y']).to(cuda_device)
    state = []
    for fused in (True, False):
        flow = build_flow_from_golden(g, cuda_device)
        e = flow.engine()
        tr = FusedNLLTrainer(e, x, y, precision=precision)
        if precision == 'bf16':
            assert e.tc_tail_maps() is not None
            if not fused:
                e._scatter_tc_host, e._scatter_tc_dev = False, None      # step() falls back to the separate launches
        else:
            assert e.gather_one_to_one
            if not fused:
                e.gather_one_to_one = False
        for _ in range(3):
            tr.step()
        blob = e.packed_tc if precision == 'bf16' else e.packed
        state.append([t.detach().cpu().numpy().copy() for t in (e.flat, e.flat_grad, e.adam_m, e.adam_v, blob)])
        assert e.adam_t == 3
    for what, a, b in zip(('flat', 'flat_grad', 'adam_m', 'adam_v', 'packed'), *state):
        assert np.array_equal(a, b), (what, int(np.sum(a != b)), float(np.max(np.abs(a.astype(np.float64) - b))))


def test_metrics_vs_reference_golden(cuda_device):
    import cnf_b200
    from cnf_b200.utils import metrics as M
    g = load_golden('metrics')
    names = sorted(k[:-len('_probs')] for k in g if k.endswith('_probs'))
    for n in names:
        p, y = g[n + '_probs'], g[n + '_y']
        assert abs(cnf_b200.expected_calibration_error(p, y, bins=15) - float(g[n + '_ece15'])) < 1e-6
        assert abs(cnf_b200.expected_calibration_error(p, y, bins=10) - float(g[n + '_ece10'])) < 1e-6
        oh = np.zeros(p.shape, dtype=np.int32)
        oh[np.arange(len(y)), y] = 1
        assert abs(cnf_b200.neg_log_likelihood(p, oh) - float(g[n + '_nll'])) < 1e-6 * max(1, abs(float(g[n + '_nll'])))
        assert cnf_b200.accuracy(p, oh) == float(g[n + '_acc'])
        # integer statistics are bit-exact against the oracle's right-closed binning
        cnt, sconf, sacc = orc.ece_bins(p, y, 15)
        st = M.statistics(p, y, bins=15).cpu().numpy()
        assert np.array_equal(st[:15], cnt.astype(np.float64))
        assert np.array_equal(st[30:45], sacc)
        assert np.allclose(st[15:30], sconf, rtol=1e-12 if p.dtype == np.float64 else 1e-6)


def test_metrics_logits_and_calibrated_modes_vs_oracle(cuda_device):
    import torch
    from scipy.special import softmax
    from cnf_b200 import _lib
    from cnf_b200.utils import metrics as M
    z, y = orc.synth_logits(20000, 10, seed=11)
    lp = orc.log_priors(orc.onehot_encode(y))
    st = M.statistics(z, y, bins=15, mode=_lib.METRICS_LOGITS).cpu().numpy()
    p32 = softmax(z, axis=1)
    assert p32.dtype == np.float32
    assert abs(M.ece_from_statistics(st, 15) - orc.expected_calibration_error(p32, y, 15)) < 1e-6
    assert abs(st[45] / st[47] - orc.neg_log_likelihood(p32, y)) < 1e-6
    st = M.statistics(z, y, bins=15, mode=_lib.METRICS_CALIBRATED, log_priors=lp).cpu().numpy()
    pc = softmax(np.log(p32 + 1e-7) - lp, axis=1)          # calibrators.py:44 on float32 probs
    assert pc.dtype == np.float64
    assert abs(M.ece_from_statistics(st, 15) - orc.expected_calibration_error(pc, y, 15)) < 1e-6
    assert abs(st[45] / st[47] - orc.neg_log_likelihood(pc, y)) < 1e-6
    assert st[46] / st[47] == orc.accuracy(pc, y)


@pytest.mark.parametrize('name', ['cal_nice_k3', 'cal_nvp_k10'])
def test_calibrator_drop_in_vs_reference(name, cuda_device):
    import torch
    import cnf_b200
    g = load_golden('calibrator_' + name)
    K = int(g['K'])
    hidden = [int(h) for h in g['hidden']]

    class Factory(cnf_b200.CouplingStack):
        def __init__(self, dim, **kw):
            super().__init__(dim, layers=int(g['layers']), hidden_size=hidden, scale=bool(g['scale']), **{
                k: v for k, v in kw.items() if k not in ('layers', 'hidden_size', 'scale')})
            flat = torch.from_numpy(g['flat0'].astype(np.float32))
            off = 0
            with torch.no_grad():
                for lay in self.layers:
                    for p in lay.canonical_parameters():
                        p.copy_(flat[off:off + p.numel()].view(p.shape))
                        off += p.numel()

    cal = cnf_b200.TorchFlowCalibrator(Factory, g['x'], g['y'], epochs=int(g['epochs']), dev=cuda_device)
    assert np.allclose(cal.log_priors, g['log_priors'])
    hist = {k: np.array([float(v) for v in cal.history[k]]) for k in ('loss', 'ce', 'log_det')}
    assert all(v.dim() == 0 for v in cal.history['loss'])
    assert np.allclose(hist['loss'], g['hist_loss'], rtol=2e-4, atol=1e-5)
    assert np.allclose(hist['ce'], g['hist_ce'], rtol=2e-4, atol=1e-5)
    assert np.allclose(hist['log_det'], g['hist_log_det'], rtol=2e-3, atol=1e-5)
    pred = cal.predict(g['x_test'])
    assert pred.dtype == np.float64
    assert np.max(np.abs(pred - g['pred'])) < 2e-5
    assert np.max(np.abs(cal(g['x_test']) - g['pred'])) < 2e-5
    pl = cal.predict_logits(orc.center(g['x_test']))
    assert rel_err(pl, g['pred_logits']) < 1e-4
    pp = cal.predict_post(orc.center(g['x_test']))
    assert np.allclose(pp.sum(axis=1), 1.0, atol=1e-5)
    # mini-batch mode runs and keeps finite history
    cal2 = cnf_b200.TorchFlowCalibrator(Factory, g['x'], g['y'], epochs=2, batch_size=64, dev=cuda_device)
    assert np.isfinite([float(v) for v in cal2.history['loss']]).all()


def test_calibrator_minibatch_fit_vs_oracle_with_pinned_permutation(cuda_device):
    """The mini-batch branch of TorchFlowCalibrator.fit (calibrators.py:274-317: shuffled batches, one Adam step per
    batch, history from the LAST evaluation batch only) against the float64 oracle driven with the same batch
    order (the shuffle is pinned through the _epoch_permutation hook)."""
    import torch
    import cnf_b200
    g = load_golden('calibrator_cal_nvp_k10')
    K, hidden, layers = int(g['K']), [int(h) for h in g['hidden']], int(g['layers'])
    N, bs, epochs = g['x'].shape[0], 64, 3
    perms = [np.random.default_rng(40 + i).permutation(N) for i in range(2 * epochs)]

    class Factory(cnf_b200.CouplingStack):
        def __init__(self, dim, **kw):
            super().__init__(dim, layers=layers, hidden_size=hidden, scale=bool(g['scale']), **{
                k: v for k, v in kw.items() if k not in ('layers', 'hidden_size', 'scale')})
            flat = torch.from_numpy(g['flat0'].astype(np.float32))
            off = 0
            with torch.no_grad():
                for lay in self.layers:
                    for p in lay.canonical_parameters():
                        p.copy_(flat[off:off + p.numel()].view(p.shape))
                        off += p.numel()

    class Pinned(cnf_b200.TorchFlowCalibrator):
        calls = 0

        def _epoch_permutation(self, n_local, gen):
            p = perms[Pinned.calls]
            Pinned.calls += 1
            return torch.from_numpy(p).to(self.dev)

    cal = Pinned(Factory, g['x'], g['y'], epochs=epochs, batch_size=bs, dev=cuda_device)
    assert Pinned.calls == 2 * epochs
    hist = {k: np.array([float(v) for v in cal.history[k]]) for k in ('loss', 'ce', 'log_det')}
    # oracle: same centring, same batches, Adam with torch defaults
    x = orc.center(g['x']).astype(np.float64)
    y = np.argmax(orc.onehot_encode(g['y']), axis=1) if g['y'].ndim == 1 else np.argmax(g['y'], axis=1)
    like = orc.init_params(K, layers, hidden, bool(g['scale']), True)
    flat = g['flat0'].astype(np.float64)
    m, v, t = np.zeros_like(flat), np.zeros_like(flat), 0
    ref = {'loss': [], 'ce': [], 'log_det': []}
    for e in range(epochs):
        p = perms[2 * e]
        for s in range(0, N, bs):
            idx = p[s:s + bs]
            _, _, _, grads, _ = orc.train_step_grads(orc.unflatten(flat, like), x[idx], y[idx])
            t += 1
            flat, m, v = orc.adam_step(flat, orc.flatten(grads), m, v, t)
        p = perms[2 * e + 1]
        idx = p[(N - 1) // bs * bs:]
        zs, ld = orc.flow_forward(orc.unflatten(flat, like), x[idx])
        loss, ce, ldm, _, _ = orc.nll_head(zs[-1], ld, y[idx])
        # reference quirk: the last batch's mean times its length over N (calibrators.py:309-317)
        ref['loss'].append(loss * len(idx) / N)
        ref['ce'].append(ce * len(idx) / N)
        ref['log_det'].append(ldm * len(idx) / N)
    for k in ref:
        assert np.allclose(hist[k], np.array(ref[k]), rtol=2e-4, atol=1e-6), (k, hist[k], ref[k])
    got = np.concatenate([p.detach().cpu().numpy().reshape(-1) for lay in cal.flow.layers for p in lay.canonical_parameters()])
    disp = flat - g['flat0']
    big = np.abs(disp) > 0.1 * np.max(np.abs(disp))
    assert rel_err((got - g['flat0'])[big], disp[big]) < 2e-2


@pytest.mark.parametrize('golden,N', [('flow_c2_nvp_k10', 65_536), ('flow_c2_nvp_k10', 70_003), ('flow_c2_nvp_k10_init', 131_075)])
@pytest.mark.parametrize('eps,gamma', [(1e-7, 1.0), (0.0, 1.0), (0.0, 0.0)])
def test_register_resident_training_kernel_vs_oracle(golden, N, eps, gamma, cuda_device, monkeypatch):
    """Large batches (>= 262,144 samples) at K = 10 train on train_reg10_kernel (tapeless backward, butterfly-reduced
    weight gradients): loss and gradient against the float64 oracle (fp32 tolerances: loss 1e-5, gradient 2e-4 of its
    maximum), against the 32-sample-tile kernel on the same data, bitwise repeatable, and evaluation-only calls."""
    import torch
    monkeypatch.setenv('CNF_FP32R_TRAIN', '0')      # by default the kernel takes over from 262,144 samples; forced here
    g = load_golden(golden)
    flow = build_flow_from_golden(g, cuda_device)
    eng = flow.engine()
    eng.ensure(cuda_device)
    eng.pack()
    p = oracle_params_from_golden(g, np.float64)
    x, y = orc.synth_logits(N, 10, seed=N % 1000)
    xt, yt = torch.from_numpy(x).to(cuda_device), torch.from_numpy(y).to(cuda_device)
    acc = torch.zeros(4, dtype=torch.float64, device=cuda_device)
    eng.nll_step(xt, yt, acc, eps=eps, gamma=gamma)
    grad = eng.flat_grad.cpu().numpy().copy()
    loss, ce, ldm, grads, _ = orc.train_step_grads(p, x.astype(np.float64), y, eps=eps, gamma=gamma)
    ref = orc.flatten(grads)
    assert abs(-float(acc[0]) / N - loss) < 1e-5 * max(1.0, abs(loss))
    assert abs(-float(acc[1]) / N - ce) < 1e-5 * max(1.0, abs(ce)) and float(acc[3]) == 0.0
    assert rel_err(grad, ref) < 2e-4, rel_err(grad, ref)
    assert np.all(grad[ref == 0] == 0)
    # one row of the partial buffer per warp, tiles in a fixed order on each: the sums do not depend on scheduling
    acc2 = torch.zeros(4, dtype=torch.float64, device=cuda_device)
    eng.nll_step(xt, yt, acc2, eps=eps, gamma=gamma)
    assert np.array_equal(eng.flat_grad.cpu().numpy(), grad)
    # evaluation only: same sums, no gradient buffer touched
    acc3 = torch.zeros(4, dtype=torch.float64, device=cuda_device)
    eng.nll_step(xt, yt, acc3, eps=eps, gamma=gamma, with_grad=False)
    assert np.allclose(acc3.cpu().numpy(), acc.cpu().numpy(), rtol=1e-12)
    # the tile kernel on the same batch
    monkeypatch.setenv('CNF_FP32R_TRAIN', 'off')
    acc4 = torch.zeros(4, dtype=torch.float64, device=cuda_device)
    eng.nll_step(xt, yt, acc4, eps=eps, gamma=gamma)
    assert rel_err(eng.flat_grad.cpu().numpy(), grad) < 1e-4      # (each within 2e-4 of the oracle)
    assert np.allclose(acc4.cpu().numpy()[:3], acc.cpu().numpy()[:3], rtol=1e-6)


def test_full_size_properties_c2(cuda_device):
    """BASELINE config C2 size (K=10, N=1M, L=6, H=128): size-independent properties."""
    import torch
    g = load_golden('flow_c2_nvp_k10')
    flow = build_flow_from_golden(g, cuda_device)
    N = 1_000_000
    x, _ = orc.synth_logits(N, 10, seed=99)
    xt = torch.from_numpy(x).to(cuda_device)
    with torch.no_grad():
        zs, ld = flow(xt)
        xr, ldr = flow.backward(zs[-1])
        err = float((xr[-1] - xt).abs().max() / xt.abs().max())
        assert err < 5e-5
        assert float((ld + ldr).abs().max()) < 5e-4
        assert torch.isfinite(zs[-1]).all()
        # a slice equals the same rows computed alone (tiles are independent): bitwise on the same kernel ...
        z_mid, ld_mid = flow(xt[123456:123456 + 300_001])
        assert torch.equal(z_mid[-1], zs[-1][123456:123456 + 300_001])
        assert torch.equal(ld_mid, ld[123456:123456 + 300_001])
        # ... and to fp32 rounding on the kernels that serve smaller batches (generic thread-per-sample below
        # 262,144 rows, 32-sample tiles up to 32,768: other summation orders)
        for n_small in (70_001, 40_001, 777):
            z_small, ld_small = flow(xt[123456:123456 + n_small])
            assert float((z_small[-1] - zs[-1][123456:123456 + n_small]).abs().max()) < 2e-6 * float(zs[-1].abs().max())
            assert float((ld_small - ld[123456:123456 + n_small]).abs().max()) < 2e-6 * max(1.0, float(ld.abs().max()))
    # spot-check against the oracle on a strided subset
    p = oracle_params_from_golden(g, np.float64)
    idx = np.arange(0, N, 997)
    tidx = torch.from_numpy(idx).to(cuda_device)
    zo, ldo = orc.flow_forward(p, x[idx].astype(np.float64))
    assert rel_err(zs[-1][tidx].cpu().numpy(), zo[-1]) < TOL
    assert ld_err(ld[tidx].cpu().numpy(), ldo) < TOL


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_host_buffer_api_matches_device_api(precision, cuda_device):
    """cnf_flow_apply_host (chunked H2D -> kernel -> D2H on internal streams) == device-resident call."""
    import torch
    import cnf_b200
    torch.manual_seed(0)
    model = cnf_b200.RealNvpFlow(10, layers=6, hidden_size=[128], precision=precision)
    with torch.no_grad():
        for p in model.parameters():
            if p.requires_grad:
                p.mul_(300.0)
    model.to(cuda_device)
    N = 300_007                                      # ragged: not a multiple of the chunk or the tile
    x, _ = orc.synth_logits(N, 10, seed=8)
    xh = torch.from_numpy(x).pin_memory()
    zh, lh = model.transform_host(xh, device=cuda_device, chunk=65536)
    with torch.no_grad():
        z, ld = model(torch.from_numpy(x).to(cuda_device))
    if precision == 'bf16':
        assert torch.equal(zh, z.cpu()) and torch.equal(lh, ld.cpu())
    else:
        # fp32: the 65,536-row chunks take the generic kernel, the 300,007-row device-resident call the
        # register-resident one (another summation order)
        assert float((zh - z.cpu()).abs().max()) < 2e-6 * float(z.abs().max())
        assert float((lh - ld.cpu()).abs().max()) < 2e-6 * max(1.0, float(ld.abs().max()))
    # pageable numpy input and inverse direction
    zi, li = model.engine().apply_host(zh.numpy(), inverse=True, precision=precision, device=cuda_device)
    torch.cuda.synchronize()
    assert rel_err(zi.numpy(), x) < (1e-2 if precision == 'bf16' else 5e-5)


@pytest.mark.parametrize('K,hidden', [(10, [128, 128]), (20, [64, 96, 32]), (100, [512])])
def test_host_buffer_api_on_the_streamed_weight_tensor_core_kernels(K, hidden, cuda_device):
    """cnf_flow_apply_host with the bf16 kernels that stream their weights (cnf_flow_tcw.cu, cnf_flow_tcm.cu): the
    zero-copy launch on pinned host pointers (calls of <= 2^21 rows) and the chunked copy-engine pipeline give
    exactly what the device-resident call gives, ragged sizes included."""
    import torch
    import cnf_b200
    from helpers import set_dense_weights
    flow = cnf_b200.Flow([cnf_b200.NvpCouplingLayer(K, hidden) for _ in range(3)], precision='bf16')
    set_dense_weights(flow, seed=K)
    flow.to(cuda_device)
    eng = flow.engine()
    assert eng.tc_bytes > 0
    for N, chunk in ((70_003, None), (70_003, 16_384)):
        x, _ = orc.synth_logits(N, K, seed=9)
        xh = torch.from_numpy(x).pin_memory()
        zh, lh = eng.apply_host(xh, precision='bf16', device=cuda_device, chunk=chunk)
        z, ld, _ = eng.apply(torch.from_numpy(x).to(cuda_device), precision='bf16')
        torch.cuda.synchronize()
        assert torch.equal(zh, z.cpu()) and torch.equal(lh, ld.cpu())
        xi, li = eng.apply_host(zh, inverse=True, precision='bf16', device=cuda_device, chunk=chunk)
        torch.cuda.synchronize()
        assert rel_err(xi.numpy(), x) < 1e-2


def test_affine_constant_layer_and_tempscaler_vs_reference(cuda_device):
    """AffineConstantLayer between coupling layers (loaded through the reference's
    own state_dict keys), forward / inverse / autograd; TempScaler forward and dT."""
    import torch
    import cnf_b200
    g = load_golden('affine')
    K = int(g['K'])
    flow = cnf_b200.Flow([cnf_b200.NvpCouplingLayer(K, [8]), cnf_b200.AffineConstantLayer(K),
                          cnf_b200.NvpCouplingLayer(K, [8])])
    sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('sd_')}
    assert set(sd) == set(flow.state_dict().keys())
    flow.load_state_dict(sd)
    flow.to(cuda_device)
    x = torch.from_numpy(g['x']).to(cuda_device).requires_grad_(True)
    y = torch.from_numpy(g['y']).to(cuda_device)
    zs, ld = flow(x)
    assert rel_err(zs[-1].detach().cpu().numpy(), g['z']) < 1e-5
    assert rel_err(zs[1].detach().cpu().numpy(), g['z_mid']) < 1e-5
    assert np.max(np.abs(ld.detach().cpu().numpy() - g['logdet'])) < 1e-5 * max(1.0, np.max(np.abs(g['logdet'])))
    loss = torch.nn.CrossEntropyLoss()(zs[-1], y) - torch.mean(ld)
    loss.backward()
    assert abs(float(loss.detach()) - float(g['loss'])) < 1e-5
    assert rel_err(x.grad.cpu().numpy(), g['gx']) < 2e-4
    aff = flow.layers[1]
    assert rel_err(aff.s.grad.cpu().numpy(), g['g_s']) < 2e-4
    assert rel_err(aff.t.grad.cpu().numpy(), g['g_t']) < 2e-4
    assert rel_err(flow.layers[0].s.layers[0].weight.grad.cpu().numpy(), g['g_w']) < 2e-4
    with torch.no_grad():
        xs, ldi = flow.backward(zs[-1].detach())
    assert rel_err(xs[-1].cpu().numpy(), g['x_rec']) < 5e-5
    assert np.max(np.abs(ldi.cpu().numpy() - g['logdet_inv'])) < 1e-5 * max(1.0, np.max(np.abs(g['logdet_inv'])))
    ts = cnf_b200.TempScaler().to(cuda_device)
    with torch.no_grad():
        ts.T.fill_(-1.7)
    xt = torch.from_numpy(g['x']).to(cuda_device)
    zt = ts(xt)
    assert rel_err(zt.detach().cpu().numpy(), g['temp_z']) < 1e-6
    (zt ** 2).sum().backward()
    assert rel_err(ts.T.grad.cpu().numpy(), g['temp_gT']) < 1e-4
    with torch.no_grad():
        assert rel_err(ts.backward(zt.detach()).cpu().numpy(), g['x']) < 1e-6


@pytest.mark.parametrize('K', [3, 10, 40])
def test_planar_and_radial_layers_vs_reference(K, cuda_device):
    """PlanarLayer / RadialLayer drop-ins (flows/flows.py:129-193): outputs and autograd gradients against
    the reference's own run (tests/golden/planar_radial.npz); fp32 tolerance 1e-5 / 2e-4 on gradients."""
    import torch
    import cnf_b200
    g = load_golden('planar_radial')
    t = 'planar_k%d_' % K
    lay = cnf_b200.PlanarLayer(K).to(cuda_device)
    with torch.no_grad():
        lay.w.copy_(torch.from_numpy(g[t + 'w'])); lay.u.copy_(torch.from_numpy(g[t + 'u'])); lay.b.copy_(torch.from_numpy(g[t + 'b']))
    x = torch.from_numpy(g[t + 'x']).to(cuda_device).requires_grad_(True)
    z, ld = lay(x)
    assert lay.invertible is False and ld.shape == (x.shape[0],)
    ((z * torch.from_numpy(g[t + 'cz']).to(cuda_device)).sum() + (ld * torch.from_numpy(g[t + 'cl']).to(cuda_device)).sum()).backward()
    assert rel_err(z.detach().cpu().numpy(), g[t + 'z']) < TOL
    assert np.max(np.abs(ld.detach().cpu().numpy() - g[t + 'ld'])) < TOL * max(1.0, np.max(np.abs(g[t + 'ld'])))
    assert rel_err(x.grad.cpu().numpy(), g[t + 'gx']) < 2e-4
    assert rel_err(lay.w.grad.cpu().numpy(), g[t + 'gw']) < 2e-4
    assert rel_err(lay.u.grad.cpu().numpy(), g[t + 'gu']) < 2e-4
    assert rel_err(lay.b.grad.cpu().numpy(), g[t + 'gb']) < 2e-4
    t = 'radial_k%d_' % K
    lay = cnf_b200.RadialLayer(K).to(cuda_device)
    with torch.no_grad():
        lay.z0.copy_(torch.from_numpy(g[t + 'z0'])); lay.a.copy_(torch.from_numpy(g[t + 'a'])); lay.b.copy_(torch.from_numpy(g[t + 'b']))
    x = torch.from_numpy(g[t + 'x']).to(cuda_device).requires_grad_(True)
    z, ld = lay(x)
    assert ld.dim() == 0 and float(ld) == 0.0                # the reference's constant log(1.0)
    (z * torch.from_numpy(g[t + 'cz']).to(cuda_device)).sum().backward()
    assert rel_err(z.detach().cpu().numpy(), g[t + 'z']) < TOL
    assert rel_err(x.grad.cpu().numpy(), g[t + 'gx']) < 2e-4
    assert rel_err(lay.z0.grad.cpu().numpy(), g[t + 'gz0']) < 2e-4
    assert rel_err(lay.a.grad.cpu().numpy(), g[t + 'ga']) < 2e-4
    assert rel_err(lay.b.grad.cpu().numpy(), g[t + 'gb']) < 2e-4


def test_planar_radial_inside_a_flow_and_ragged_sizes(cuda_device):
    """Planar -> coupling -> radial in one Flow (state_dict keys as the reference's), Flow.backward raises,
    and the streaming kernels agree with the oracle on ragged / multi-tile batch sizes."""
    import torch
    import cnf_b200
    g = load_golden('planar_radial')
    K = int(g['flow_K'])
    flow = cnf_b200.Flow([cnf_b200.PlanarLayer(K), cnf_b200.NvpCouplingLayer(K, hidden_size=[8]), cnf_b200.RadialLayer(K)])
    sd = {k[len('flow_sd_'):]: torch.from_numpy(np.asarray(g[k])) for k in g if k.startswith('flow_sd_')}
    assert set(flow.state_dict().keys()) == set(sd.keys())
    flow.load_state_dict(sd)
    flow.to(cuda_device)
    with torch.no_grad():
        zs, ld = flow(torch.from_numpy(g['flow_x']).to(cuda_device))
    assert rel_err(zs[-1].cpu().numpy(), g['flow_z']) < TOL
    assert np.max(np.abs(ld.cpu().numpy() - g['flow_ld'])) < TOL * max(1.0, np.max(np.abs(g['flow_ld'])))
    with pytest.raises(ValueError, match='not tractable'):
        flow.backward(zs[-1])
    rng = np.random.default_rng(5)
    for N, K in ((1, 5), (31, 10), (33, 10), (4097, 100), (100003, 10)):
        x, _ = orc.synth_logits(N, K, seed=N)
        w, u, b = rng.random(K).astype(np.float32), rng.random(K).astype(np.float32), rng.random(1).astype(np.float32)
        pl = cnf_b200.PlanarLayer(K).to(cuda_device)
        with torch.no_grad():
            pl.w.copy_(torch.from_numpy(w)); pl.u.copy_(torch.from_numpy(u)); pl.b.copy_(torch.from_numpy(b))
            z, l2 = pl(torch.from_numpy(x).to(cuda_device))
        zo, lo, _ = orc.planar_forward(w.astype(np.float64), u.astype(np.float64), b, x.astype(np.float64))
        assert rel_err(z.cpu().numpy(), zo) < TOL
        assert np.max(np.abs(l2.cpu().numpy().reshape(-1) - lo)) < TOL * max(1.0, np.max(np.abs(lo)))
        rl = cnf_b200.RadialLayer(K).to(cuda_device)
        with torch.no_grad():
            z, _ = rl(torch.from_numpy(x).to(cuda_device))
        zo, _ = orc.radial_forward(rl.z0.detach().cpu().numpy().astype(np.float64), rl.a.detach().cpu().numpy(),
                                   rl.b.detach().cpu().numpy(), x.astype(np.float64))
        assert rel_err(z.cpu().numpy(), zo) < TOL


def test_calibrator_fit_with_captured_cuda_graph_epochs_matches_eager(cuda_device):
    """cuda_graph=True (one graph launch per full-batch epoch, Adam step count on the device) reproduces the
    eager fit: same history and same predictions."""
    import torch
    import cnf_b200
    g = load_golden('calibrator_cal_nvp_k10')
    hidden = [int(h) for h in g['hidden']]
    out = {}
    for graph in (False, True):
        torch.manual_seed(5)
        cal = cnf_b200.TorchFlowCalibrator(cnf_b200.RealNvpFlow, g['x'], g['y'], layers=int(g['layers']),
                                           hidden_size=hidden, epochs=12, dev=cuda_device, cuda_graph=graph)
        out[graph] = (np.array([float(v) for v in cal.history['loss']]), cal.predict(g['x_test']))
    assert np.isfinite(out[True][0]).all()
    assert np.allclose(out[True][0], out[False][0], rtol=1e-6, atol=1e-7)
    assert np.max(np.abs(out[True][1] - out[False][1])) < 1e-6


def test_empty_batches_on_every_path(cuda_device):
    """N = 0: every entry point returns empty outputs / zero statistics instead of launching."""
    import torch
    import cnf_b200
    g = load_golden('flow_c2_nvp_k10')
    x0 = torch.zeros((0, 10), device=cuda_device)
    y0 = torch.zeros(0, dtype=torch.int64, device=cuda_device)
    for prec in ('fp32', 'bf16'):
        flow = build_flow_from_golden(g, cuda_device, precision=prec)
        with torch.no_grad():
            zs, ld = flow(x0)
            xs, ldi = flow.backward(zs[-1])
        assert zs[-1].shape == (0, 10) and ld.numel() == 0 and xs[-1].shape == (0, 10)
        eng = flow.engine()
        eng.pack(tc=True)
        acc = torch.zeros(4, dtype=torch.float64, device=cuda_device)
        eng.nll_step(x0, y0, acc, precision=prec)
        assert float(acc.abs().sum()) == 0.0
        assert float(eng.flat_grad.abs().sum()) == 0.0
    for lay in (cnf_b200.PlanarLayer(10).to(cuda_device), cnf_b200.RadialLayer(10).to(cuda_device),
                cnf_b200.AffineConstantLayer(10).to(cuda_device)):
        with torch.no_grad():
            z, _ = lay(x0)
        assert z.shape == (0, 10)


def test_misaligned_views_take_the_scalar_paths(cuda_device):
    """A contiguous row-slice x[1:] starts 4*K bytes into its storage (8-byte aligned for K=10): the 128-bit /
    TMA fast paths must step aside and the results must not change."""
    import torch
    import cnf_b200
    from cnf_b200.utils import metrics as M
    g = load_golden('flow_c2_nvp_k10')
    x, y = orc.synth_logits(3001, 10, seed=4)
    big = torch.from_numpy(x).to(cuda_device)
    view = big[1:]
    assert view.is_contiguous() and view.data_ptr() % 16 != 0
    copy = view.clone()
    assert copy.data_ptr() % 16 == 0
    yv = torch.from_numpy(y).to(cuda_device)[1:]
    for prec in ('fp32', 'bf16'):
        flow = build_flow_from_golden(g, cuda_device, precision=prec)
        with torch.no_grad():
            za, la = flow(view)
            zb, lb = flow(copy)
        assert torch.equal(za[-1], zb[-1]) and torch.equal(la, lb)
        eng = flow.engine()
        eng.pack(tc=True)
        acc_a = torch.zeros(4, dtype=torch.float64, device=cuda_device)
        eng.nll_step(view, yv.contiguous(), acc_a, precision=prec)
        ga = eng.flat_grad.clone()
        acc_b = torch.zeros(4, dtype=torch.float64, device=cuda_device)
        eng.nll_step(copy, yv.contiguous(), acc_b, precision=prec)
        assert torch.allclose(acc_a, acc_b, rtol=1e-12) and torch.allclose(ga, eng.flat_grad, rtol=1e-5, atol=1e-9)
    for lay in (cnf_b200.PlanarLayer(10).to(cuda_device), cnf_b200.RadialLayer(10).to(cuda_device),
                cnf_b200.AffineConstantLayer(10).to(cuda_device)):
        with torch.no_grad():
            za, _ = lay(view)
            zb, _ = lay(copy)
        assert torch.equal(za, zb)
    pv = torch.softmax(view, dim=1)
    big_p = torch.cat([pv[:1], pv])            # make a misaligned probability view
    sa = M.statistics(big_p[1:], yv, bins=15)
    sb = M.statistics(pv.clone(), yv, bins=15)
    assert torch.allclose(sa, sb, rtol=1e-12, atol=0)


@pytest.mark.parametrize('K,L,H,scale,shift,N', [(100, 2, 512, True, True, 300), (100, 3, 100, False, True, 517),
                                                 (33, 2, 700, True, False, 129), (4, 3, 888, True, True, 1000)])
def test_lean_kernels_for_wide_shapes_vs_oracle(K, L, H, scale, shift, N, cuda_device):
    """Shapes whose weights / tape / hidden activations do not fit the regular fp32 plan run on the lean
    kernels (staged 16-unit chunks, no tape): forward, inverse, the fused NLL step and autograd through the
    drop-in Flow, all against the float64 oracle."""
    import torch
    import cnf_b200
    torch.manual_seed(K + H)
    layers = [cnf_b200.NvpCouplingLayer(K, [H], scale=scale, shift=shift) for _ in range(L)]
    flow = cnf_b200.Flow(layers)
    with torch.no_grad():
        for p in flow.parameters():
            if p.requires_grad:
                p.mul_(100.0)
    like = orc.init_params(K, L, [H], scale, shift)
    flat = np.concatenate([p.detach().numpy().reshape(-1) for lay in layers for p in lay.canonical_parameters()])
    params = orc.unflatten(flat.astype(np.float64), like)
    x, y = orc.synth_logits(N, K, seed=K)
    flow.to(cuda_device)
    xt, yt = torch.from_numpy(x).to(cuda_device), torch.from_numpy(y).to(cuda_device)
    zo, ldo = orc.flow_forward(params, x.astype(np.float64))
    with torch.no_grad():
        zs, ld = flow(xt)
        xs, _ = flow.backward(zs[-1])
    assert rel_err(zs[-1].cpu().numpy(), zo[-1]) < TOL
    assert ld_err(ld.cpu().numpy().reshape(-1), ldo) < TOL
    assert rel_err(zs[0].cpu().numpy(), zo[0]) < TOL            # intermediates (second launch)
    assert rel_err(xs[-1].cpu().numpy(), x) < 1e-4
    eng = flow.engine()
    eng.pack()
    acc = torch.zeros(4, dtype=torch.float64, device=cuda_device)
    eng.nll_step(xt, yt, acc)
    loss, _, _, grads, _ = orc.train_step_grads(params, x.astype(np.float64), y)
    go = orc.flatten(grads)
    assert abs(-float(acc[0]) / N - loss) < 1e-5 * max(1.0, abs(loss))
    assert rel_err(eng.flat_grad.cpu().numpy(), go) < 5e-4
    # autograd through the drop-in modules (external upstream gradients, g_x returned)
    xg = xt.clone().requires_grad_(True)
    zs2, ld2 = flow(xg)
    (zs2[-1].square().sum() * 1e-3 + ld2.sum()).backward()
    gz = 2e-3 * zo[-1]
    _, gx_o = orc.flow_backward(params, x.astype(np.float64), gz, np.ones(N))
    assert rel_err(xg.grad.cpu().numpy(), gx_o) < 5e-4


@pytest.mark.parametrize('K,L,H,scale,shift,N', [(10, 6, 128, True, True, 5000),     # both nets side by side (512 threads)
                                                 (10, 6, 128, True, True, 20001),    # more tiles than SMs: nets in turn
                                                 (10, 3, 256, True, True, 333),      # 16 chunks
                                                 (3, 4, 32, False, True, 1000),      # NICE (t-net only), 2 warps
                                                 (7, 2, 20, True, False, 65),        # scale only, padded hidden layer
                                                 (40, 2, 100, True, True, 2500),     # d0, d1 > number of warps
                                                 (10, 3, [128, 128], True, True, 1500),   # deeper nets: activations
                                                 (10, 2, [128, 64], True, True, 333),     #   exchanged through shared
                                                 (10, 2, [48, 128], False, True, 97),     #   memory (flow_train_deep_kernel)
                                                 (100, 2, [64, 64, 64], True, True, 200),
                                                 (100, 2, [100, 100], True, True, 150),
                                                 (30, 2, 256, True, True, 200),           # one hidden layer, too big for the split plan
                                                 (100, 2, 100, True, True, 150),
                                                 (6, 2, [256, 256], True, False, 64)])
def test_small_batch_training_kernel_vs_oracle(K, L, H, scale, shift, N, cuda_device, monkeypatch):
    """The 32-sample-tile training kernel (hidden layer split over the warps of a CTA) against the float64
    oracle: NLL head with its loss sums, external upstream gradients with g_x, repeated steps on the same
    partial buffer, and equality with the one-thread-per-sample kernel and with the all-rows C-ABI calls."""
    import ctypes
    import torch
    import cnf_b200
    from cnf_b200 import _lib
    from cnf_b200._engine import _ptr, _stream
    hidden = H if isinstance(H, list) else [H]
    torch.manual_seed(K * 1000 + sum(hidden))
    layers = [cnf_b200.NvpCouplingLayer(K, hidden, scale=scale, shift=shift) for _ in range(L)]
    flow = cnf_b200.Flow(layers)
    with torch.no_grad():
        for p in flow.parameters():
            if p.requires_grad:
                p.mul_(100.0 if len(hidden) == 1 else 30.0)
    like = orc.init_params(K, L, hidden, scale, shift)
    flat = np.concatenate([p.detach().numpy().reshape(-1) for lay in layers for p in lay.canonical_parameters()])
    params = orc.unflatten(flat.astype(np.float64), like)
    x, y = orc.synth_logits(N, K, seed=K + 1)
    flow.to(cuda_device)
    xt, yt = torch.from_numpy(x).to(cuda_device), torch.from_numpy(y).to(cuda_device)
    eng = flow.engine()
    eng.ensure(cuda_device)
    eng.pack()
    loss, _, _, grads, _ = orc.train_step_grads(params, x.astype(np.float64), y)
    go = orc.flatten(grads)
    for _ in range(2):                       # the second step reuses the partial rows of the first
        acc = torch.zeros(4, dtype=torch.float64, device=cuda_device)
        eng.nll_step(xt, yt, acc)
        assert abs(-float(acc[0]) / N - loss) < 1e-5 * max(1.0, abs(loss))
        assert rel_err(eng.flat_grad.cpu().numpy(), go) < 2e-4
    g_split = eng.flat_grad.clone()
    # every entry of a partial row has one writer and the rows are summed in a fixed order: the gradient is
    # bitwise reproducible run to run (a shared-memory race would show up here)
    for _ in range(8):
        eng.nll_step(xt, yt, torch.zeros(4, dtype=torch.float64, device=cuda_device))
        assert torch.equal(eng.flat_grad, g_split)
    # the original entry points (every partial row cleared and reduced) give the same gradient
    acc2 = torch.zeros(4, dtype=torch.float64, device=cuda_device)
    st = _stream(cuda_device)
    _lib.call('cnf_nll_train_step', ctypes.byref(eng.desc), _ptr(eng.packed), _ptr(eng.tables), _ptr(xt), _ptr(yt),
              ctypes.c_int64(N), ctypes.c_float(1e-7), ctypes.c_float(1.0), ctypes.c_float(1.0 / N),
              _ptr(eng.partials), _ptr(acc2), st)
    g_all = torch.empty_like(eng.flat_grad)
    _lib.call('cnf_grad_reduce', ctypes.byref(eng.desc), _ptr(eng.partials), _ptr(eng.gather), _ptr(g_all), st)
    # (same partial rows; the row sums associate differently when the row count differs)
    assert rel_err(g_all.cpu().numpy(), g_split.cpu().numpy()) < 1e-6
    assert torch.allclose(acc2, acc, rtol=1e-12, atol=0)
    # one-thread-per-sample kernel on the same inputs
    monkeypatch.setenv('CNF_SPLIT_TRAIN', '0')
    monkeypatch.setenv('CNF_DEEP_TRAIN', '0')
    acc3 = torch.zeros(4, dtype=torch.float64, device=cuda_device)
    eng.nll_step(xt, yt, acc3)
    monkeypatch.delenv('CNF_SPLIT_TRAIN')
    monkeypatch.delenv('CNF_DEEP_TRAIN')
    assert rel_err(eng.flat_grad.cpu().numpy(), g_split.cpu().numpy()) < 1e-5
    assert abs(float(acc3[0]) - float(acc[0])) < 1e-6 * abs(float(acc[0]))
    # external upstream gradients through the drop-in modules
    zo, _ = orc.flow_forward(params, x.astype(np.float64))
    xg = xt.clone().requires_grad_(True)
    zs2, ld2 = flow(xg)
    (zs2[-1].square().sum() * 1e-3 + ld2.sum()).backward()
    _, gx_o = orc.flow_backward(params, x.astype(np.float64), 2e-3 * zo[-1], np.ones(N))
    assert rel_err(xg.grad.cpu().numpy(), gx_o) < 2e-4


@pytest.mark.parametrize('name', ['nvp_k5_oddL', 'scaleonly_k4_h888', 'c2_nvp_k10', 'c1_nice_k3', 'nvp_k7_randflip'])
def test_tile_kernels_forced_on_the_reference_goldens(name, cuda_device, monkeypatch):
    """The 32-sample-tile kernels for deeper conditioners (flow_apply_deep_kernel, flow_train_deep_kernel) are
    picked by default only for wide nets; forced here onto the reference's golden cases (hidden [5,5], [8,8,8],
    random flips, NICE), they must reproduce the reference's outputs, inverse and autograd gradients."""
    monkeypatch.setenv('CNF_DEEP_APPLY', '1')
    monkeypatch.setenv('CNF_DEEP_TRAIN', '1')
    monkeypatch.setenv('CNF_SPLIT_TRAIN', '0')
    test_forward_inverse_vs_reference_golden(name, cuda_device)
    for tag, eps, gamma in (('cal', 1e-7, 1.0), ('script', 0.0, 1.0), ('script_nodet', 0.0, 0.0)):
        test_fused_train_step_gradients_vs_reference_autograd(name, tag, eps, gamma, cuda_device)


@pytest.mark.parametrize('L,hidden,wmul', [(6, [5, 5], 300.0), (5, [5, 5], 300.0), (4, [3, 20], 200.0), (6, [5, 64], 100.0)])
@pytest.mark.parametrize('N', [1_500, 5_003, 70_001, 130_003, 300_001])
def test_register_kernel_with_a_small_first_hidden_layer_vs_oracle(L, hidden, wmul, N, cuda_device, monkeypatch):
    """K = 10 flows whose conditioners have two hidden layers with at most five units in the first -- the reference's
    DEFAULT NvpCouplingLayer(dim, hidden_size=[5, 5]) (flows/flows.py:69) -- run forward / inverse / fused predict on
    flow_reg10_kernel<..., M2> from 1,024 samples (1 / 2 / 4 samples per thread by batch size): against the float64
    oracle, against the generic kernel, and through the fused statistics pass."""
    import torch
    import cnf_b200
    monkeypatch.setenv('CNF_LIVE_ENV', '1')
    torch.manual_seed(L + hidden[1])
    flow = cnf_b200.Flow([cnf_b200.NvpCouplingLayer(10, hidden) for _ in range(L)])
    with torch.no_grad():
        for p in flow.parameters():
            if p.requires_grad:
                p.mul_(wmul)
    flat = np.concatenate([p.detach().numpy().reshape(-1) for lay in flow.layers for p in lay.canonical_parameters()])
    params = orc.unflatten(flat.astype(np.float64), orc.init_params(10, L, hidden, True, True))
    flow.to(cuda_device)
    x, y = orc.synth_logits(N, 10, seed=N % 1000)
    xt = torch.from_numpy(x).to(cuda_device)
    with torch.no_grad():
        zs, ld = flow(xt)
        xr, ldr = flow.backward(zs[-1])
        monkeypatch.setenv('CNF_FP32R', 'off')               # the same call on the generic kernel
        zs_g, ld_g = flow(xt)
        z_g, ldg = zs_g[-1].clone(), ld_g.clone()
        monkeypatch.delenv('CNF_FP32R')
    z = zs[-1]
    assert float((z - z_g).abs().max()) < 2e-6 * float(z_g.abs().max())
    assert float((ld - ldg).abs().max()) < 2e-6 * max(1.0, float(ldg.abs().max()))
    idx = np.unique(np.concatenate([np.arange(0, N, 499), np.arange(max(0, N - 300), N)]))   # strided subset + the ragged tail
    zo, ldo = orc.flow_forward(params, x[idx].astype(np.float64))
    tidx = torch.from_numpy(idx).to(cuda_device)
    assert rel_err(z[tidx].cpu().numpy(), zo[-1]) < TOL
    assert ld_err(ld[tidx].cpu().numpy(), ldo) < TOL
    assert rel_err(xr[-1][tidx].cpu().numpy(), x[idx]) < 5e-5
    assert ld_err(ldr[tidx].cpu().numpy(), -ldo) < 5e-5
    # fused statistics pass (register kernel with the tail) == statistics of the plain forward's logits
    from cnf_b200 import _lib
    from cnf_b200.utils import metrics as M
    yt = torch.from_numpy(y).to(cuda_device)
    st_fused = flow.engine().predict(xt, y=yt, bins=15)['stats'].cpu().numpy()
    st_plain = M.statistics(z, yt, bins=15, mode=_lib.METRICS_LOGITS).cpu().numpy()
    assert np.array_equal(st_fused[[-2, -1]], st_plain[[-2, -1]])            # correct count, N
    assert np.array_equal(st_fused[:15], st_plain[:15])                      # bin counts
    assert np.allclose(st_fused, st_plain, rtol=1e-9, atol=1e-6)


@pytest.mark.parametrize('L,hidden,wmul', [(6, [5, 5], 300.0), (5, [5, 5], 1.0), (4, [3, 20], 200.0)])
@pytest.mark.parametrize('N,eps,gamma', [(1_500, 1e-7, 1.0), (5_003, 0.0, 1.0), (40_001, 1e-7, 1.0), (65_536, 1e-7, 1.0), (70_003, 0.0, 1.0), (131_075, 0.0, 0.0)])
def test_register_training_kernel_with_a_small_first_hidden_layer_vs_oracle(L, hidden, wmul, N, eps, gamma, cuda_device,
                                                                             monkeypatch):
    """train_reg10_kernel<..., M2>: the NLL step of K = 10 flows whose conditioners have two hidden layers with at most
    five units in the first (the reference's default hidden_size=[5, 5]): loss and gradient against the float64
    oracle, bitwise repeatable, and against the generic one-thread-per-sample kernel on the same batch."""
    import torch
    import cnf_b200
    monkeypatch.setenv('CNF_LIVE_ENV', '1')
    monkeypatch.setenv('CNF_FP32R_TRAIN', '0')      # by default the kernel takes over from 160,000 samples; forced here
    torch.manual_seed(L + hidden[1])
    flow = cnf_b200.Flow([cnf_b200.NvpCouplingLayer(10, hidden) for _ in range(L)])
    with torch.no_grad():
        for p in flow.parameters():
            if p.requires_grad:
                p.mul_(wmul)
    flat = np.concatenate([p.detach().numpy().reshape(-1) for lay in flow.layers for p in lay.canonical_parameters()])
    params = orc.unflatten(flat.astype(np.float64), orc.init_params(10, L, hidden, True, True))
    flow.to(cuda_device)
    eng = flow.engine()
    eng.ensure(cuda_device)
    eng.pack()
    x, y = orc.synth_logits(N, 10, seed=N % 1000)
    xt, yt = torch.from_numpy(x).to(cuda_device), torch.from_numpy(y).to(cuda_device)
    acc = torch.zeros(4, dtype=torch.float64, device=cuda_device)
    eng.nll_step(xt, yt, acc, eps=eps, gamma=gamma)
    grad = eng.flat_grad.cpu().numpy().copy()
    loss, ce, ldm, grads, _ = orc.train_step_grads(params, x.astype(np.float64), y, eps=eps, gamma=gamma)
    ref = orc.flatten(grads)
    assert abs(-float(acc[0]) / N - loss) < 1e-5 * max(1.0, abs(loss))
    assert abs(-float(acc[1]) / N - ce) < 1e-5 * max(1.0, abs(ce)) and float(acc[3]) == 0.0
    assert rel_err(grad, ref) < 2e-4, rel_err(grad, ref)
    assert np.all(grad[ref == 0] == 0)
    acc2 = torch.zeros(4, dtype=torch.float64, device=cuda_device)
    eng.nll_step(xt, yt, acc2, eps=eps, gamma=gamma)
    assert np.array_equal(eng.flat_grad.cpu().numpy(), grad)
    monkeypatch.setenv('CNF_FP32R_TRAIN', 'off')
    acc4 = torch.zeros(4, dtype=torch.float64, device=cuda_device)
    eng.nll_step(xt, yt, acc4, eps=eps, gamma=gamma)
    assert rel_err(eng.flat_grad.cpu().numpy(), grad) < 2e-4
    assert np.allclose(acc4.cpu().numpy()[:3], acc.cpu().numpy()[:3], rtol=1e-6)


@pytest.mark.parametrize('K,L,hidden,scale,shift', [(3, 5, [3, 3], True, True), (3, 5, [3, 3], False, True), (3, 10, [5, 5], True, True),
                                                    (2, 4, [4, 7], True, True), (4, 3, [5, 5], True, False), (5, 6, [2, 20], True, True),
                                                    (7, 4, [5, 5], True, True), (8, 3, [3, 3], False, True), (9, 5, [5, 9], True, True),
                                                    (10, 4, [5, 5], False, True), (3, 4, [32], False, True), (6, 3, [40], True, True),
                                                    (10, 3, [40], False, True), (10, 2, [24], True, False)])
@pytest.mark.parametrize('N,eps,gamma', [(1_500, 1e-7, 1.0), (5_003, 0.0, 1.0), (70_001, 0.0, 0.0)])
def test_register_training_kernel_other_class_counts_vs_oracle(K, L, hidden, scale, shift, N, eps, gamma, cuda_device, monkeypatch):
    """train_reg10_kernel<..., KK> for K = 2 .. 9 (cnf_flow_fp32rk.cu) and for flows without a scale or shift net: the
    reference's notebook / script shapes (K = 3, hidden_size [3, 3] / [5, 5], NiceFlow and RealNvpFlow,
    notebooks/simulated-predictions-flows.ipynb:214, 224; run_experiment3D.py:33-38).  Loss and gradient against the
    float64 oracle (odd K: the middle logit is a conditioning dim of every layer), bitwise repeatable, dead
    weight entries exactly zero, and against the generic tile kernels on the same batch."""
    import torch
    import cnf_b200
    monkeypatch.setenv('CNF_LIVE_ENV', '1')
    monkeypatch.setenv('CNF_FP32R_TRAIN', '0')      # single-hidden-layer shapes: forced on below 160,000 samples
    torch.manual_seed(K + L + hidden[-1])
    flow = cnf_b200.Flow([cnf_b200.NvpCouplingLayer(K, hidden, scale=scale, shift=shift) for _ in range(L)])
    with torch.no_grad():
        for p in flow.parameters():
            if p.requires_grad:
                p.mul_(250.0)
    flat = np.concatenate([p.detach().numpy().reshape(-1) for lay in flow.layers for p in lay.canonical_parameters()])
    params = orc.unflatten(flat.astype(np.float64), orc.init_params(K, L, hidden, scale, shift))
    flow.to(cuda_device)
    eng = flow.engine()
    eng.ensure(cuda_device)
    eng.pack()
    x, y = orc.synth_logits(N, K, seed=N % 1000)
    xt, yt = torch.from_numpy(x).to(cuda_device), torch.from_numpy(y).to(cuda_device)
    acc = torch.zeros(4, dtype=torch.float64, device=cuda_device)
    eng.nll_step(xt, yt, acc, eps=eps, gamma=gamma)
    grad = eng.flat_grad.cpu().numpy().copy()
    loss, ce, ldm, grads, _ = orc.train_step_grads(params, x.astype(np.float64), y, eps=eps, gamma=gamma)
    ref = orc.flatten(grads)
    assert abs(-float(acc[0]) / N - loss) < 1e-5 * max(1.0, abs(loss))
    assert abs(-float(acc[1]) / N - ce) < 1e-5 * max(1.0, abs(ce)) and float(acc[3]) == 0.0
    assert abs(float(acc[2]) / N - ldm) < 1e-5 * max(1.0, abs(ldm))
    assert np.max(np.abs(ref)) > 0
    assert rel_err(grad, ref) < 2e-4, rel_err(grad, ref)
    assert np.all(grad[ref == 0] == 0)
    acc2 = torch.zeros(4, dtype=torch.float64, device=cuda_device)
    eng.nll_step(xt, yt, acc2, eps=eps, gamma=gamma)
    assert np.array_equal(eng.flat_grad.cpu().numpy(), grad)
    # forward-only (evaluation) pass of the same kernel
    acc3 = torch.zeros(4, dtype=torch.float64, device=cuda_device)
    eng.nll_step(xt, yt, acc3, eps=eps, gamma=gamma, with_grad=False)
    assert np.allclose(acc3.cpu().numpy()[:3], acc.cpu().numpy()[:3], rtol=1e-9)
    monkeypatch.setenv('CNF_FP32R_TRAIN', 'off')
    acc4 = torch.zeros(4, dtype=torch.float64, device=cuda_device)
    eng.nll_step(xt, yt, acc4, eps=eps, gamma=gamma)
    assert rel_err(eng.flat_grad.cpu().numpy(), grad) < 2e-4
    assert np.allclose(acc4.cpu().numpy()[:3], acc.cpu().numpy()[:3], rtol=1e-6)


@pytest.mark.parametrize('K,L,hidden,scale', [(3, 5, [3, 3], True), (3, 4, [32], False), (10, 6, [128], True), (10, 6, [5, 5], True)])
def test_fit_loop_enqueued_from_c_is_bitwise_the_stepwise_loop(K, L, hidden, scale, cuda_device):
    """cnf_fit_full_batch (the epoch loop of TorchFlowCalibrator.fit, calibrators.py:283-317, enqueued by one library
    call) against the same loop driven step by step from Python: the same history, bitwise the same parameters and Adam
    state."""
    import torch
    import cnf_b200
    N, epochs = 1500, 25
    x, y = orc.synth_logits(N, K, seed=K)
    xt, yt = torch.from_numpy(x).to(cuda_device), torch.from_numpy(y).to(cuda_device)
    out = []
    for fused in (True, False):
        torch.manual_seed(3)
        flow = cnf_b200.Flow([cnf_b200.NvpCouplingLayer(K, hidden, scale=scale) for _ in range(L)]).to(cuda_device)
        tr = cnf_b200.FusedNLLTrainer(flow.engine(), xt, yt, lr=1e-3)
        if fused:
            hist = tr.fit_loop(epochs, N, None)
        else:
            hist = torch.zeros((epochs, 4), dtype=torch.float64, device=cuda_device)
            for e in range(epochs):
                tr.step(acc=hist[e - 1] if e > 0 else None)
            tr.evaluate(out=hist[epochs - 1])
        eng = flow.engine()
        out.append((hist.cpu().numpy(), eng.flat.cpu().numpy().copy(), eng.adam_m.cpu().numpy().copy(), eng.adam_t))
    assert np.isfinite(out[0][0]).all() and out[0][0][-1, 0] != out[0][0][0, 0]
    # parameters and Adam state bitwise; the history rows are float64 atomic adds of per-CTA sums, whose order is free
    assert np.allclose(out[0][0], out[1][0], rtol=1e-12, atol=1e-9)
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])
    assert out[0][3] == out[1][3] == epochs
    # a second call continues the same optimiser (bias correction from step 26 on)
    h2 = tr.fit_loop(3, N, None)
    assert np.isfinite(h2.cpu().numpy()).all() and flow.engine().adam_t == epochs + 3


@pytest.mark.parametrize('K,L,hidden,scale,shift', [(3, 10, [5, 5], True, True), (3, 5, [3, 3], False, True), (10, 6, [5, 5], True, True),
                                                    (7, 4, [4, 9], True, True)])
def test_autograd_backward_on_the_register_kernel_vs_oracle(K, L, hidden, scale, shift, cuda_device):
    """The script loop (run_experiment3D.py:98-135: loss = CE(zs[-1], y) - det * mean(log_det); loss.backward()) through
    the drop-in Flow at the script's own shapes: the autograd backward (external head: upstream dL/dz, dL/dlog_det) runs
    on the register-resident kernel when the input needs no gradient.  Parameter gradients against the float64 oracle
    and against the tile kernel (CNF_FP32R_TRAIN=off)."""
    import os
    import torch
    import cnf_b200
    N = 1500
    torch.manual_seed(K + L)
    flow = cnf_b200.Flow([cnf_b200.NvpCouplingLayer(K, hidden, scale=scale, shift=shift) for _ in range(L)])
    with torch.no_grad():
        for p in flow.parameters():
            if p.requires_grad:
                p.mul_(250.0)
    flat = np.concatenate([p.detach().numpy().reshape(-1) for lay in flow.layers for p in lay.canonical_parameters()])
    params = orc.unflatten(flat.astype(np.float64), orc.init_params(K, L, hidden, scale, shift))
    flow.to(cuda_device)
    x, y = orc.synth_logits(N, K, seed=11)
    xt, yt = torch.from_numpy(x).to(cuda_device), torch.from_numpy(y).to(cuda_device)

    def grads_now():
        for p in flow.parameters():
            p.grad = None
        zs, ld = flow(xt)
        loss = torch.nn.functional.cross_entropy(zs[-1], yt) - ld.mean()
        loss.backward()
        return float(loss), np.concatenate([p.grad.detach().cpu().numpy().reshape(-1) for lay in flow.layers
                                            for p in lay.canonical_parameters()])
    loss, g = grads_now()
    lo, _, _, gro, _ = orc.train_step_grads(params, x.astype(np.float64), y, eps=0.0, gamma=1.0)
    ref = orc.flatten(gro)
    assert abs(loss - lo) < 1e-5 * max(1.0, abs(lo))
    assert rel_err(g, ref) < 2e-4, rel_err(g, ref)
    assert np.all(g[ref == 0] == 0)
    os.environ['CNF_LIVE_ENV'] = '1'
    os.environ['CNF_FP32R_TRAIN'] = 'off'
    try:
        _, g_tile = grads_now()
    finally:
        os.environ.pop('CNF_FP32R_TRAIN', None)
        os.environ.pop('CNF_LIVE_ENV', None)
    assert rel_err(g_tile, g) < 2e-4
