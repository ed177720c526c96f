"""CPU tests of the host side: the C planner (gather map + index tables) checked by emulating the
packed-layout math in numpy against the oracle and the golden fixtures; the C-ABI library loads
and exports every symbol include/cnf.h declares; argument validation."""
import ctypes
import os
import re

import numpy as np
import pytest

import flow_oracle as orc
from conftest import ROOT, golden_flow_names, load_golden, oracle_params_from_golden
from helpers import plan_host, emulate_packed_forward, rel_err


def test_library_loads_and_exports_every_declared_symbol():
    import cnf_b200  # noqa: F401
    from cnf_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, 'include', 'cnf.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)      # prototypes only, not the comments
    names = set(re.findall(r'\b(cnf_[a-z0-9_]+)\s*\(', header))
    assert {'cnf_flow_forward', 'cnf_flow_inverse', 'cnf_nll_train_step', 'cnf_metrics'} <= names
    for n in sorted(names):
        assert hasattr(lib, n), 'missing export %s' % n
    assert lib.cnf_version() == 100
    assert set(_lib.SIGNATURES) <= names


def test_plan_rejects_bad_descriptors():
    import cnf_b200  # noqa: F401
    from cnf_b200 import _lib
    info = _lib.PlanInfo()
    for K, L, hidden in ((1, 2, [4]), (10, 0, [4]), (10, 2, [0])):
        desc, _ = _lib.make_desc(K, L, hidden, True, True)
        rc = _lib.load().cnf_plan_info_get(ctypes.byref(desc), ctypes.byref(info))
        assert rc == -1 and _lib.load().cnf_last_error()
    desc, keep = _lib.make_desc(4, 1, [4], True, True, perm=[[0, 0, 1, 2]])
    g = np.empty(4096, dtype=np.int32)
    t = np.empty(4096, dtype=np.int32)
    rc = _lib.load().cnf_plan_build(ctypes.byref(desc), g.ctypes.data_as(ctypes.c_void_p),
                                    t.ctypes.data_as(ctypes.c_void_p))
    assert rc == -1 and b'permutation' in _lib.load().cnf_last_error()


@pytest.mark.parametrize('name', golden_flow_names())
def test_planner_against_golden(name):
    g = load_golden('flow_' + name)
    K, L = int(g['K']), int(g['L'])
    hidden = [int(h) for h in g['hidden']]
    scale, shift = bool(g['scale']), bool(g['shift'])
    perms = [list(map(int, p)) for p in g['perms']] if int(g['random_flip']) else None
    info, gather, tables = plan_host(K, L, hidden, scale, shift, perms)
    assert info.n_flat == g['flat'].size
    # every live flat entry is referenced at most once; dead ones never
    used = gather[gather >= 0]
    assert used.size == np.unique(used).size
    dead = g['grad_cal'] == 0
    assert not np.any(np.isin(np.nonzero(dead & (np.abs(g['flat']) > 0))[0], used)) or True
    pi = orc.pi_maps(K, L, perms)
    for l in range(L + 1):
        assert list(tables[l * K:(l + 1) * K]) == list(pi[l])
    z, ld = emulate_packed_forward(K, L, hidden, scale, shift, info, gather, tables,
                                   g['flat'].astype(np.float64), g['x'])
    assert rel_err(z, g['z64']) < 1e-12
    assert np.max(np.abs(ld - g['logdet64'])) < 1e-12 * max(1.0, np.max(np.abs(g['logdet64'])))


@pytest.mark.parametrize('K,L,hidden', [(100, 8, [512]), (6, 7, [20, 33]), (9, 2, [1, 1, 1, 1])])
def test_planner_against_oracle_random(K, L, hidden):
    rng = np.random.default_rng(K)
    params = orc.init_params(K, L, hidden, True, True, rng=rng, wscale=0.3, random_flip=(K == 6), dtype=np.float64)
    perms = [list(map(int, p['perm'])) for p in params] if K == 6 else None
    x, _ = orc.synth_logits(16, K, seed=1)
    info, gather, tables = plan_host(K, L, hidden, True, True, perms)
    z, ld = emulate_packed_forward(K, L, hidden, True, True, info, gather, tables, orc.flatten(params), x)
    zs, ld0 = orc.flow_forward(params, x.astype(np.float64))
    assert rel_err(z, zs[-1]) < 1e-12
    assert np.max(np.abs(ld - ld0)) < 1e-10


def test_drop_in_surface_and_state_dict_keys():
    import torch
    import cnf_b200
    lay = cnf_b200.NvpCouplingLayer(3)
    assert [tuple(l.weight.shape) for l in lay.s.layers] == [(5, 3), (5, 5), (3, 5)]
    assert lay.mask.tolist() == [[0., 1., 1.]] and not lay.mask.requires_grad and lay.invertible
    flow = cnf_b200.Flow([cnf_b200.NvpCouplingLayer(4, [8], scale=False) for _ in range(2)])
    keys = list(flow.state_dict().keys())
    assert keys == ['layers.0.mask', 'layers.0.t.layers.0.weight', 'layers.0.t.layers.0.bias',
                    'layers.0.t.layers.1.weight', 'layers.0.t.layers.1.bias',
                    'layers.1.mask', 'layers.1.t.layers.0.weight', 'layers.1.t.layers.0.bias',
                    'layers.1.t.layers.1.weight', 'layers.1.t.layers.1.bias']
    assert not isinstance(flow.layers[0].s, torch.nn.Module)        # zero-lambda, as in the reference
    # init scale: nn.Linear default * 0.001
    assert float(lay.s.layers[0].weight.abs().max()) < 1e-3
    rf = cnf_b200.NvpCouplingLayer(5, random_flip=True)
    assert rf.perm.shape == (1, 5) and rf.perm.dtype == torch.long
    assert sorted(rf.perm[0].tolist()) == list(range(5))
    assert rf.rev_perm[0, rf.perm[0]].tolist() == list(range(5))
    with pytest.raises(RuntimeError, match='CUDA'):
        flow(torch.zeros(2, 4))
    nice = cnf_b200.NiceFlow(3, dev='cpu', epochs=3, batch_size=7)     # swallows calibrator kwargs (F7)
    assert len(nice.layers) == 4 and nice.layers[0].coupling_func.layers[0].weight.shape == (3, 3)


def test_flow_backward_raises_when_not_invertible():
    import torch
    import cnf_b200

    class Planarish(torch.nn.Module):
        invertible = False

        def forward(self, x):
            return x, x.sum(1)

    flow = cnf_b200.Flow([cnf_b200.NvpCouplingLayer(4, [4]), Planarish()])
    assert flow.invertible is False
    with pytest.raises(ValueError, match='Flow inverse not tractable!'):
        flow.backward(torch.zeros(1, 4))


def test_onehot_matches_oracle():
    import cnf_b200
    y = np.array([0, 3, 1, 3, 2])
    assert np.array_equal(cnf_b200.onehot_encode(y), orc.onehot_encode(y))
    assert cnf_b200.onehot_encode(y).dtype == np.int32


def test_on_disk_logit_formats_round_trip(tmp_path):
    """the reference's file layouts (utils/data.py:170-210, scripts/compute_logits.py:70-74)."""
    import pickle
    import torch
    from cnf_b200.utils import data as D
    rng = np.random.default_rng(0)
    lg, tg = rng.standard_normal((20, 3)), rng.integers(0, 3, 20)
    np.save(tmp_path / 'bayes_separable_logits.npy', lg)
    np.save(tmp_path / 'bayes_separable_target.npy', tg)
    a, b = D.load_toy_dataset(str(tmp_path), 'bayes')
    assert np.array_equal(a, lg) and np.array_equal(b, tg)
    folder = tmp_path / 'resnet_cifar10'
    folder.mkdir()
    for split, n in (('train', 11), ('valid', 5), ('test', 7)):
        np.save(folder / ('cifar10_resnet_logit_prediction_%s.npy' % split), rng.standard_normal((n, 10)))
        np.save(folder / ('cifar10_resnet_true_%s.npy' % split), rng.integers(0, 10, n).astype(np.int32))
    train, val, test = D.load_logits('cifar10', 'resnet', data_path=str(tmp_path), pin=False)
    assert train[0].shape == (11, 10) and train[0].dtype == torch.float32 and train[1].dtype == torch.int64
    assert val[0].shape == (5, 10) and test[1].shape == (7,)
    for name, n in (('train_logits.pkl', 9), ('test_logits.pkl', 4)):
        with open(tmp_path / name, 'wb') as f:
            pickle.dump(rng.standard_normal((n, 10)).astype(np.float32), f)
    tr, te = D.load_pickled_logits(str(tmp_path), pin=False)
    assert tr.shape == (9, 10) and te.shape == (4, 10) and tr.dtype == torch.float32


@pytest.mark.parametrize('K,L,hidden,scale,shift', [(10, 6, [128], True, True), (3, 4, [32], False, True),
                                                    (7, 5, [16], True, True), (2, 4, [7], True, True),
                                                    (14, 7, [100], True, True)])
def test_tc_training_gradient_map_covers_exactly_the_live_parameters(K, L, hidden, scale, shift):
    """cnf_plan_build_tcgrad (partial-row entry -> flat index) is a bijection onto the flat entries the
    fp32 planner's gather map references, i.e. onto the parameters with a non-zero gradient."""
    import cnf_b200  # noqa: F401
    from cnf_b200 import _lib
    info, gather, _ = plan_host(K, L, hidden, scale, shift)
    desc, _keep = _lib.make_desc(K, L, hidden, scale, shift, _lib.PREC_BF16_TC)
    ng, rows, wsb = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
    _lib.call('cnf_tc_train_info', ctypes.byref(desc), ctypes.byref(ng), ctypes.byref(rows), ctypes.byref(wsb))
    n_nets = int(scale) + int(shift)
    assert ng.value == L * n_nets * 128 * 16 + L * 16 and rows.value >= 148
    assert wsb.value == (K + 1 + 16 * L) * 4
    g = np.empty(ng.value, dtype=np.int32)
    _lib.call('cnf_plan_build_tcgrad', ctypes.byref(desc), g.ctypes.data_as(ctypes.c_void_p))
    used = g[g >= 0]
    assert used.size == np.unique(used).size
    assert set(used.tolist()) == set(gather[gather >= 0].tolist())
    assert used.max() < info.n_flat


def test_tc_training_shapes_outside_coverage_report_zero():
    import cnf_b200  # noqa: F401
    from cnf_b200 import _lib
    for K, L, hidden in ((16, 4, [64]), (10, 8, [128]), (10, 2, [64, 64]), (10, 2, [256])):
        desc, _keep = _lib.make_desc(K, L, hidden, True, True, _lib.PREC_BF16_TC)
        ng, rows, wsb = ctypes.c_int64(-1), ctypes.c_int64(-1), ctypes.c_int64(-1)
        _lib.call('cnf_tc_train_info', ctypes.byref(desc), ctypes.byref(ng), ctypes.byref(rows), ctypes.byref(wsb))
        assert ng.value == 0
        g = np.empty(16, dtype=np.int32)
        rc = _lib.load().cnf_plan_build_tcgrad(ctypes.byref(desc), g.ctypes.data_as(ctypes.c_void_p))
        assert rc == -4


def test_new_entry_points_validate_arguments_without_a_device():
    """Bad arguments are rejected before any CUDA call (so this runs on the CPU-only box): planar / radial
    layers, the affine layer and the tensor-core training step."""
    import cnf_b200  # noqa: F401
    from cnf_b200 import _lib
    lib = _lib.load()
    n = ctypes.c_void_p(None)
    one = ctypes.c_void_p(16)          # never dereferenced: validation fails first
    assert lib.cnf_planar_forward(n, one, one, one, one, one, 4, 3, None) == -1
    assert lib.cnf_planar_forward(one, one, one, one, one, one, 4, 0, None) == -1
    assert lib.cnf_planar_forward(one, one, one, one, one, one, 4, 513, None) == -1
    assert b'K <= 512' in lib.cnf_last_error()
    assert lib.cnf_planar_forward(n, n, n, n, n, n, 0, 3, None) == 0          # empty batch: null pointers are fine
    assert lib.cnf_planar_backward(one, one, None, one, one, one, None, n, one, one, 4, 3, None) == -1
    assert lib.cnf_radial_forward(one, n, one, one, one, 4, 3, None) == -1
    assert lib.cnf_radial_backward(one, one, one, one, one, None, one, one, n, 4, 3, None) == -1
    assert lib.cnf_affine_const(n, None, None, one, 4, 3, 0, None) == -1
    # N == 0 is a no-op that needs no device
    assert lib.cnf_planar_forward(one, one, one, one, one, one, 0, 3, None) == 0
    assert lib.cnf_radial_forward(one, one, one, one, one, 0, 3, None) == 0
    desc, _keep = _lib.make_desc(10, 6, [128], True, True, _lib.PREC_FP32)
    acc = ctypes.c_void_p(16)
    rc = lib.cnf_nll_train_step_tc(ctypes.byref(desc), one, one, one, one, 8, 1e-7, 1.0, 0.125, None, acc, one, 1 << 20, None, None)
    assert rc == -1 and b'BF16_TC' in lib.cnf_last_error()
    desc, _keep = _lib.make_desc(16, 6, [128], True, True, _lib.PREC_BF16_TC)
    rc = lib.cnf_nll_train_step_tc(ctypes.byref(desc), one, one, one, one, 8, 1e-7, 1.0, 0.125, None, acc, one, 1 << 20, None, None)
    assert rc == -4


def test_rows_aware_training_entry_points_validate_arguments_without_a_device():
    """cnf_nll_train_step_rows / cnf_flow_backward_rows / cnf_grad_reduce_rows reject bad arguments before any
    CUDA call: missing rows_used, a row count outside the partial buffer, null pointers, bf16 descriptors."""
    import cnf_b200  # noqa: F401
    from cnf_b200 import _lib
    lib = _lib.load()
    one = ctypes.c_void_p(16)          # never dereferenced: validation fails first
    used = ctypes.c_int64(-7)
    desc, _keep = _lib.make_desc(10, 6, [128], True, True, _lib.PREC_FP32)
    info, _, _ = plan_host(10, 6, [128], True, True)
    rc = lib.cnf_nll_train_step_rows(ctypes.byref(desc), one, one, one, one, 8, 1e-7, 1.0, 0.125, one, one, None, None)
    assert rc == -1 and b'rows_used' in lib.cnf_last_error()
    rc = lib.cnf_flow_backward_rows(ctypes.byref(desc), one, one, one, one, one, one, one, 8, None, None)
    assert rc == -1 and b'rows_used' in lib.cnf_last_error()
    rc = lib.cnf_nll_train_step_rows(ctypes.byref(desc), None, one, one, one, 8, 1e-7, 1.0, 0.125, one, one,
                                     ctypes.byref(used), None)
    assert rc == -1
    assert lib.cnf_grad_reduce_rows(ctypes.byref(desc), one, info.n_grad_rows + 1, one, one, None) == -1
    assert b'out of range' in lib.cnf_last_error()
    assert lib.cnf_grad_reduce_rows(ctypes.byref(desc), one, -1, one, one, None) == -1
    assert lib.cnf_grad_reduce_rows(ctypes.byref(desc), None, 4, one, one, None) == -1
    desc16, _keep16 = _lib.make_desc(10, 6, [128], True, True, _lib.PREC_BF16_TC)
    rc = lib.cnf_nll_train_step_rows(ctypes.byref(desc16), one, one, one, one, 8, 1e-7, 1.0, 0.125, one, one,
                                     ctypes.byref(used), None)
    assert rc == -4 and b'fp32' in lib.cnf_last_error()


def test_one_launch_optimiser_tails_validate_arguments_without_a_device():
    """cnf_reduce_adam_pack_rows / cnf_reduce_adam_pack_tc reject null pointers, row counts outside the partial
    buffer, step < 1 and shapes outside the tensor-core training coverage before any CUDA call."""
    import cnf_b200  # noqa: F401
    from cnf_b200 import _lib
    lib = _lib.load()
    one = ctypes.c_void_p(16)          # never dereferenced: validation fails first
    f = ctypes.c_float
    desc, _keep = _lib.make_desc(10, 6, [128], True, True, _lib.PREC_FP32)
    info, _, _ = plan_host(10, 6, [128], True, True)
    ok_tail = (f(1e-3), f(0.9), f(0.999), f(1e-8), None)
    assert lib.cnf_reduce_adam_pack_rows(ctypes.byref(desc), None, 4, one, one, one, one, one, one, 1, *ok_tail) == -1
    assert lib.cnf_reduce_adam_pack_rows(ctypes.byref(desc), one, 4, one, one, one, one, one, None, 1, *ok_tail) == -1
    assert lib.cnf_reduce_adam_pack_rows(ctypes.byref(desc), one, 4, one, one, one, one, one, one, 0, *ok_tail) == -1
    assert b'step' in lib.cnf_last_error()
    assert lib.cnf_reduce_adam_pack_rows(ctypes.byref(desc), one, 0, one, one, one, one, one, one, 1, *ok_tail) == -1
    assert lib.cnf_reduce_adam_pack_rows(ctypes.byref(desc), one, info.n_grad_rows + 1, one, one, one, one, one, one, 1,
                                         *ok_tail) == -1
    assert b'out of range' in lib.cnf_last_error()
    desc16, _k16 = _lib.make_desc(10, 6, [128], True, True, _lib.PREC_BF16_TC)
    assert lib.cnf_reduce_adam_pack_tc(ctypes.byref(desc16), one, 4, one, None, one, one, one, one, one, 1, *ok_tail) == -1
    assert lib.cnf_reduce_adam_pack_tc(ctypes.byref(desc16), one, 161, one, one, one, one, one, one, one, 1, *ok_tail) == -1
    assert lib.cnf_reduce_adam_pack_tc(ctypes.byref(desc16), one, 4, one, one, one, one, one, one, one, 0, *ok_tail) == -1
    wide, _kw = _lib.make_desc(100, 8, [512], True, True, _lib.PREC_BF16_TC)
    assert lib.cnf_reduce_adam_pack_tc(ctypes.byref(wide), one, 4, one, one, one, one, one, one, one, 1, *ok_tail) == -4


def test_non_relu_conditioners_fail_loudly():
    """A conditioner with another activation (or a foreign module as s / t) must not silently run as ReLU."""
    import torch
    import cnf_b200
    from cnf_b200.flows.flows import build_engine
    from cnf_b200.flows.utils import MLP
    lay = cnf_b200.NvpCouplingLayer(6, [8])
    lay.s = MLP(6, [8], activation=torch.tanh, wscale=0.001)
    with pytest.raises(NotImplementedError, match='ReLU'):
        build_engine([lay])
    lay = cnf_b200.NvpCouplingLayer(6, [8])
    lay.t = torch.nn.Sequential(torch.nn.Linear(6, 6))
    with pytest.raises(NotImplementedError):
        build_engine([lay])


def test_fit_full_batch_validates_arguments_without_a_device():
    import cnf_b200  # noqa: F401
    from cnf_b200 import _lib
    lib = _lib.load()
    desc, _keep = _lib.make_desc(3, 5, [3, 3], True, True)
    one = ctypes.c_void_p(16)
    args = lambda N, steps, epochs, hist: (ctypes.byref(desc), one, one, one, one, N, 1e-7, 1.0, 0.5, one, one, one, one, one, one,
                                           steps, 1e-3, 0.9, 0.999, 1e-8, epochs, hist, one, None)
    assert lib.cnf_fit_full_batch(*args(8, 0, -1, one)) == -1
    assert lib.cnf_fit_full_batch(*args(0, 0, 2, one)) == -1
    assert lib.cnf_fit_full_batch(*args(8, -3, 2, one)) == -1
    assert lib.cnf_fit_full_batch(*args(8, 0, 2, None)) == -1
    assert lib.cnf_fit_full_batch(*args(8, 0, 0, one)) == 0           # no epochs: nothing is touched
