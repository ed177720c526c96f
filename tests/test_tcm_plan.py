"""CPU test of the deep (two to four hidden layers) tensor-core plan (cnf_flow_tcm.cu): the gather map of the bf16 phase images
(B1 | Bm | B2 in the no-swizzle K-major UMMA layout) and of the fp32 bias section, decoded in numpy with the layout
formulas of the kernel's descriptors and evaluated against the float64 oracle (flows/flows.py:101-112,
flows/utils.py:26-31).  No GPU needed: the planner is pure index logic."""
import ctypes

import numpy as np
import pytest

import flow_oracle as orc
from helpers import plan_host, rel_err


def _rup(x, m):
    return (x + m - 1) // m * m


@pytest.mark.parametrize('K,L,hidden,scale,shift,rflip', [(10, 6, [128, 128], True, True, False),
                                                          (3, 4, [32, 20], False, True, False),
                                                          (7, 3, [100, 64], True, True, True),
                                                          (65, 2, [128, 17], True, False, False),
                                                          (10, 3, [64, 128, 32], True, True, False),
                                                          (5, 2, [16, 48, 100, 128], True, True, True)])
def test_tcm_plan_images_against_oracle(K, L, hidden, scale, shift, rflip):
    import cnf_b200  # noqa: F401
    from cnf_b200 import _lib
    rng = np.random.default_rng(K + L)
    params = orc.init_params(K, L, hidden, scale, shift, rng=rng, wscale=0.3, random_flip=rflip, dtype=np.float64)
    perms = [list(map(int, p['perm'])) for p in params] if rflip else None
    flat = orc.flatten(params)
    info, _, tables = plan_host(K, L, hidden, scale, shift, perms)
    desc, _keep = _lib.make_desc(K, L, hidden, scale, shift, _lib.PREC_BF16_TC, perms)
    n_g = ctypes.c_int64(0)
    _lib.call('cnf_tc_gather_len', ctypes.byref(desc), ctypes.byref(n_g))
    assert info.tc_bytes > 0 and n_g.value > 0
    g = np.empty(n_g.value, dtype=np.int32)
    _lib.call('cnf_plan_build_tc', ctypes.byref(desc), g.ctypes.data_as(ctypes.c_void_p))
    d0, d1 = K // 2, K - K // 2
    m = len(hidden)
    Hp = [_rup(h, 16) for h in hidden]
    K1, N2 = _rup(d1 + 1, 16), _rup(d0, 16)
    n_nets = int(scale) + int(shift)
    b1_el = 128 * K1
    # blocks of one phase: block b = [B1 if b == 0] | Bm_(b+1) | [B2 if b == m - 2]
    mid_off, off = [], 0
    for b in range(m - 1):
        if b == 0:
            off += b1_el
        mid_off.append(off)
        off += Hp[b] * Hp[b + 1]
    b2_off = off
    ph_el = off + Hp[-1] * N2
    n_bf16 = L * n_nets * ph_el
    bm_floats = L * n_nets * (m - 1) * 128
    assert g.size == n_bf16 + bm_floats + L * 2 * N2
    assert info.tc_bytes == 2 * n_bf16 + 4 * (g.size - n_bf16)
    used = g[g >= 0]
    assert used.size == np.unique(used).size and used.max() < flat.size
    img = np.where(g >= 0, flat[np.maximum(g, 0)], 0.0)

    def kmajor(block, n_rows, n_k, lbo_el, kstep_el):
        """element (n, k) of a K-major image: k-steps of 16, two k-halves lbo apart, 8-row groups of 64 elements"""
        n, k = np.meshgrid(np.arange(n_rows), np.arange(n_k), indexing='ij')
        return block[(k // 16) * kstep_el + ((k % 16) // 8) * lbo_el + (n // 8) * 64 + (n % 8) * 8 + (k % 8)]

    tab_cond = (L + 1) * K
    tab_trans = tab_cond + L * d1
    x, _ = orc.synth_logits(24, K, seed=2)
    a = x.astype(np.float64).copy()
    ld = np.zeros(x.shape[0])
    for l in range(L):
        cond = tables[tab_cond + l * d1: tab_cond + (l + 1) * d1]
        trans = tables[tab_trans + l * d0: tab_trans + (l + 1) * d0]
        a1 = np.zeros((x.shape[0], K1))
        a1[:, :d1] = a[:, cond]
        a1[:, d1] = 1.0
        outs = []
        for slot in range(n_nets):
            ph = img[(l * n_nets + slot) * ph_el:(l * n_nets + slot + 1) * ph_el]
            # A1 / B1: [k-block][row-block] core matrices, 128 rows: k-halves 1024 elements apart, k-step = 2 of them
            B1 = kmajor(ph[:b1_el], 128, K1, 1024, 2048)[:Hp[0]]
            h = np.maximum(a1 @ B1.T, 0)
            for b in range(m - 1):
                Bm = kmajor(ph[mid_off[b]:mid_off[b] + Hp[b] * Hp[b + 1]], Hp[b + 1], Hp[b], Hp[b + 1] * 8, Hp[b + 1] * 16)
                o0 = n_bf16 + ((l * n_nets + slot) * (m - 1) + b) * 128
                h = np.maximum(h @ Bm.T + img[o0:o0 + Hp[b + 1]], 0)
            B2 = kmajor(ph[b2_off:], N2, Hp[-1], N2 * 8, N2 * 16)
            b2 = img[n_bf16 + bm_floats + (l * 2 + slot) * N2: n_bf16 + bm_floats + (l * 2 + slot + 1) * N2]
            o = h @ B2.T + b2
            assert np.all(o[:, d0:] == 0)            # padded outputs are exact zeros (the kernel's EPI2 relies on it)
            outs.append(o[:, :d0])
        s = outs[0] if scale else np.zeros_like(outs[0])
        t = outs[-1] if shift else np.zeros_like(outs[0])
        a[:, trans] = a[:, trans] * np.exp(s) + t
        ld += s.sum(axis=1)
    z = a[:, tables[L * K:(L + 1) * K]]
    zs, ld0 = orc.flow_forward(params, x.astype(np.float64))
    assert rel_err(z, zs[-1]) < 1e-12
    assert np.max(np.abs(ld - ld0)) < 1e-10


def test_tcm_shapes_outside_coverage_have_no_tc_blob():
    import cnf_b200  # noqa: F401
    from cnf_b200 import _lib
    for K, hidden in ((10, [129, 64]), (10, [64, 200]), (80, [64, 64]), (10, [32, 32, 8]), (10, [5, 5]), (10, [64, 15])):
        desc, _keep = _lib.make_desc(K, 2, hidden, True, True, _lib.PREC_FP32)
        info = _lib.PlanInfo()
        _lib.call('cnf_plan_info_get', ctypes.byref(desc), ctypes.byref(info))
        assert info.tc_bytes == 0
