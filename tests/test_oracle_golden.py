"""The numpy oracle against the reference's own outputs (tests/golden, made by
oracle/make_golden.py from the live reference).  CPU only."""
import numpy as np
import pytest

import flow_oracle as orc
from conftest import golden_flow_names, load_golden, oracle_params_from_golden
from helpers import rel_err


def rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30)


@pytest.mark.parametrize('name', golden_flow_names())
def test_forward_inverse_match_reference(name):
    g = load_golden('flow_' + name)
    p = oracle_params_from_golden(g, np.float64)
    zs, ld = orc.flow_forward(p, g['x'].astype(np.float64))
    # fp64 oracle vs fp32 reference: reference's own fp32 noise floor is 2e-7..3e-6
    for l in range(int(g['L'])):
        assert rel(zs[l], g['zs'][l]) < 5e-6
    assert np.max(np.abs(ld - g['logdet'])) < 5e-6 * max(1.0, np.max(np.abs(g['logdet'])))
    assert rel(zs[-1], g['z64']) < 1e-12
    assert np.max(np.abs(ld - g['logdet64'])) < 1e-12 * max(1.0, np.max(np.abs(g['logdet64'])))
    xs, ldi = orc.flow_inverse(p, g['zs'][-1].astype(np.float64))
    for l in range(int(g['L'])):
        assert rel(xs[l], g['xs'][l]) < 2e-5
    assert np.max(np.abs(ldi - g['logdet_inv'])) < 5e-6 * max(1.0, np.max(np.abs(g['logdet_inv'])))


@pytest.mark.parametrize('name', golden_flow_names())
def test_fp32_oracle_matches_reference_fp32(name):
    g = load_golden('flow_' + name)
    p = oracle_params_from_golden(g, np.float32)
    zs, ld = orc.flow_forward(p, g['x'])
    assert zs[-1].dtype == np.float32
    assert rel(zs[-1], g['zs'][-1]) < 5e-6


@pytest.mark.parametrize('name', golden_flow_names())
@pytest.mark.parametrize('tag,eps,gamma', [('cal', 1e-7, 1.0), ('script', 0.0, 1.0), ('script_nodet', 0.0, 0.0)])
def test_analytic_gradients_match_reference_autograd(name, tag, eps, gamma):
    g = load_golden('flow_' + name)
    p = oracle_params_from_golden(g, np.float64)
    loss, ce, ldm, grads, gx = orc.train_step_grads(p, g['x'].astype(np.float64), g['y'], eps, gamma)
    assert abs(loss - g['loss_' + tag]) < 2e-6 * max(1.0, abs(g['loss_' + tag]))
    flat_g = orc.flatten(grads)
    ref = g['grad_' + tag]
    assert flat_g.shape == ref.shape
    assert rel(flat_g, ref) < 2e-4
    assert rel(gx, g['gx_' + tag]) < 2e-4
    # dead first-layer columns / last-layer rows get exactly zero
    assert np.all(flat_g[ref == 0] == 0) or np.max(np.abs(flat_g[ref == 0])) < 1e-12


@pytest.mark.parametrize('name', golden_flow_names())
def test_adam_and_sgd_trajectories(name):
    g = load_golden('flow_' + name)
    like = oracle_params_from_golden(g, np.float64)
    x, y = g['x'].astype(np.float64), g['y']
    flat = g['flat'].astype(np.float64)
    m = np.zeros_like(flat)
    v = np.zeros_like(flat)
    losses = []
    for t in range(1, 4):
        p = orc.unflatten(flat, like)
        loss, _, _, grads, _ = orc.train_step_grads(p, x, y)
        losses.append(loss)
        flat, m, v = orc.adam_step(flat, orc.flatten(grads), m, v, t)
    assert np.allclose(losses, g['adam3_losses'], rtol=1e-4, atol=1e-5)
    # after 3 Adam steps every live weight moved by ~3e-3; compare the displacement
    disp_ref = g['adam3_flat'] - g['flat']
    disp = flat - g['flat']
    assert np.max(np.abs(disp - disp_ref)) < 0.05 * np.max(np.abs(disp_ref))
    dead = g['grad_cal'] == 0
    assert np.all(disp[dead] == 0)

    flat = g['flat'].astype(np.float64)
    for _ in range(2):
        p = orc.unflatten(flat, like)
        _, _, _, grads, _ = orc.train_step_grads(p, x, y, eps=0.0, gamma=1.0)
        flat = orc.sgd_step(flat, orc.flatten(grads), lr=1e-2, wd=1e-2)
    assert rel(flat - g['flat'], g['sgd2_flat'] - g['flat']) < 1e-3


def test_metrics_match_reference():
    g = load_golden('metrics')
    names = sorted(k[:-len('_probs')] for k in g if k.endswith('_probs'))
    assert names
    for n in names:
        p, y = g[n + '_probs'], g[n + '_y']
        assert abs(orc.expected_calibration_error(p, y, 15) - g[n + '_ece15']) < 1e-6
        assert abs(orc.expected_calibration_error(p, y, 10) - g[n + '_ece10']) < 1e-6
        oh = np.zeros(p.shape, dtype=np.int32)
        oh[np.arange(len(y)), y] = 1
        assert abs(orc.neg_log_likelihood(p, oh) - g[n + '_nll']) < 1e-6 * max(1, abs(g[n + '_nll']))
        assert orc.accuracy(p, oh) == g[n + '_acc']


@pytest.mark.parametrize('name', ['cal_nice_k3', 'cal_nvp_k10'])
def test_calibrator_history_and_predict(name):
    g = load_golden('calibrator_' + name)
    K = int(g['K'])
    x = orc.center(g['x'])
    y = g['y']
    like = orc.init_params(K, int(g['layers']), [int(h) for h in g['hidden']], bool(g['scale']), True)
    flat = g['flat0'].astype(np.float64)
    m = np.zeros_like(flat)
    v = np.zeros_like(flat)
    x32 = x.astype(np.float32).astype(np.float64)     # calibrators.py:248 casts to float
    hist = []
    for t in range(1, int(g['epochs']) + 1):
        p = orc.unflatten(flat, like)
        _, _, _, grads, _ = orc.train_step_grads(p, x32, y)
        flat, m, v = orc.adam_step(flat, orc.flatten(grads), m, v, t)
        p = orc.unflatten(flat, like)
        zs, ld = orc.flow_forward(p, x32)
        loss, ce, ldm, _, _ = orc.nll_head(zs[-1], ld, y)
        hist.append((loss, ce, ldm))
    hist = np.array(hist)
    assert np.allclose(hist[:, 0], g['hist_loss'], rtol=2e-4, atol=1e-5)
    assert np.allclose(hist[:, 1], g['hist_ce'], rtol=2e-4, atol=1e-5)
    assert np.allclose(hist[:, 2], g['hist_log_det'], rtol=2e-3, atol=1e-5)
    lp = orc.log_priors(orc.onehot_encode(y))
    assert np.allclose(lp, g['log_priors'])
    p_end = orc.unflatten(g['flat_end'].astype(np.float64), like)
    xt = orc.center(g['x_test']).astype(np.float32).astype(np.float64)
    zs, _ = orc.flow_forward(p_end, xt)
    assert rel(zs[-1], g['pred_logits']) < 5e-6
    pred = orc.calibrated_probs(zs[-1].astype(np.float32), lp)
    assert np.max(np.abs(pred - g['pred'])) < 1e-5


def test_pi_maps_even_odd():
    pi = orc.pi_maps(5, 3)
    assert list(pi[0]) == [0, 1, 2, 3, 4]
    assert list(pi[1]) == [4, 3, 2, 1, 0]
    assert list(pi[2]) == [0, 1, 2, 3, 4]


def _c_oracle():
    import ctypes
    import os
    import subprocess
    from conftest import ROOT
    so = os.path.join(ROOT, 'oracle', '_ref', 'libcnf_oracle.so')
    if not os.path.exists(so):
        subprocess.check_call(['make', '-C', os.path.join(ROOT, 'oracle')])
    return ctypes.CDLL(so)


@pytest.mark.parametrize('name', golden_flow_names())
def test_c_oracle_matches_reference(name):
    """The plain-C restatement (oracle/cnf_oracle.c) against the reference's float64 outputs."""
    import ctypes
    lib = _c_oracle()
    g = load_golden('flow_' + name)
    K, L, N = int(g['K']), int(g['L']), g['x'].shape[0]
    H = np.ascontiguousarray(g['hidden'], dtype=np.int32)
    perm = np.ascontiguousarray(g['perms'], dtype=np.int32) if int(g['random_flip']) else None
    flat = np.ascontiguousarray(g['flat'], dtype=np.float64)
    x = np.ascontiguousarray(g['x'], dtype=np.float64)
    z, ld, zs = np.empty((N, K)), np.empty(N), np.empty((L, N, K))
    vp = ctypes.c_void_p

    def ptr(a):
        return a.ctypes.data_as(vp) if a is not None else None

    rc = lib.cnf_oracle_forward(K, L, len(H), ptr(H), int(g['scale']), int(g['shift']), ptr(perm), ptr(flat), ptr(x),
                                ptr(z), ptr(ld), ptr(zs), ctypes.c_long(N))
    assert rc == 0
    assert rel(z, g['z64']) < 1e-12
    assert np.max(np.abs(ld - g['logdet64'])) < 1e-12 * max(1.0, np.max(np.abs(g['logdet64'])))
    for l in range(L):
        assert rel(zs[l], g['zs'][l]) < 5e-6
    xr, ldi, xs = np.empty((N, K)), np.empty(N), np.empty((L, N, K))
    zin = np.ascontiguousarray(g['zs'][-1], dtype=np.float64)
    rc = lib.cnf_oracle_inverse(K, L, len(H), ptr(H), int(g['scale']), int(g['shift']), ptr(perm), ptr(flat),
                                ptr(zin), ptr(xr), ptr(ldi), ptr(xs), ctypes.c_long(N))
    assert rc == 0
    for l in range(L):
        assert rel(xs[l], g['xs'][l]) < 2e-5
    assert np.max(np.abs(ldi - g['logdet_inv'])) < 5e-6 * max(1.0, np.max(np.abs(g['logdet_inv'])))


def test_c_oracle_metrics_match_reference():
    import ctypes
    lib = _c_oracle()
    g = load_golden('metrics')
    for n in ('f64_k10', 'f64_k3', 'edges_f64'):
        p = np.ascontiguousarray(g[n + '_probs'], dtype=np.float64)
        y = np.ascontiguousarray(g[n + '_y'], dtype=np.int64)
        st = np.empty(48)
        lib.cnf_oracle_metrics(p.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p),
                               ctypes.c_long(p.shape[0]), p.shape[1], 15, st.ctypes.data_as(ctypes.c_void_p))
        ece = sum(abs(st[30 + i] / st[i] - st[15 + i] / st[i]) * st[i] for i in range(15) if st[i] > 0) / st[47]
        assert abs(ece - g[n + '_ece15']) < 1e-6   # reference averages float32 hit flags
        assert abs(st[45] / st[47] - g[n + '_nll']) < 1e-9
        assert st[46] / st[47] == g[n + '_acc']


def test_planar_and_radial_oracle_vs_reference_golden():
    """oracle planar_forward / planar_backward / radial_forward against the reference's PlanarLayer /
    RadialLayer outputs and autograd gradients (tests/golden/planar_radial.npz)."""
    g = load_golden('planar_radial')
    for K in (3, 10, 40):
        t = 'planar_k%d_' % K
        x = g[t + 'x'].astype(np.float64)
        z, ld, u_hat = orc.planar_forward(g[t + 'w'].astype(np.float64), g[t + 'u'].astype(np.float64), g[t + 'b'], x)
        assert rel_err(z, g[t + 'z']) < 1e-5
        assert np.max(np.abs(ld - g[t + 'ld'])) < 1e-5 * max(1.0, np.max(np.abs(g[t + 'ld'])))
        gx, gw_direct, gu_hat, gb = orc.planar_backward(g[t + 'w'], u_hat, g[t + 'b'], x, g[t + 'cz'], g[t + 'cl'])
        assert rel_err(gx, g[t + 'gx']) < 1e-4
        assert abs(gb - float(g[t + 'gb'][0])) < 1e-4 * max(1.0, abs(float(g[t + 'gb'][0])))
        t = 'radial_k%d_' % K
        z, ld = orc.radial_forward(g[t + 'z0'].astype(np.float64), g[t + 'a'], g[t + 'b'], g[t + 'x'].astype(np.float64))
        assert rel_err(z, g[t + 'z']) < 1e-5
        assert ld == 0.0 and float(g[t + 'ld']) == 0.0
