"""Guard-band canaries around every output buffer the C ABI writes (no memory-safety tool is
available on this pool, compute-sanitizer answers "closed").  Every output lives inside a larger allocation filled
with a sentinel; after the call the bands on both sides must be untouched and the payload fully written (no
sentinel left).  Ragged sizes exercise the tail paths of every kernel; calls go straight through ctypes."""
import ctypes

import numpy as np
import pytest

import flow_oracle as orc
from conftest import load_golden
from helpers import build_flow_from_golden

pytestmark = pytest.mark.gpu
G = 256            # guard elements on each side (keeps 16-byte alignment of the payload for 4- and 8-byte types)
SENT = -7.25e33    # sentinel (exactly representable in float32 and float64)


def guarded(n, dtype, dev, misalign=0):
    import torch
    buf = torch.full((n + 2 * G + 8,), SENT, dtype=dtype, device=dev)
    return buf, buf[G + misalign:G + misalign + n]


def check(buf, view, n, misalign=0, written=True):
    import torch
    lo, hi = buf[:G + misalign], buf[G + misalign + n:]
    assert bool((lo == SENT).all()) and bool((hi == SENT).all()), 'write outside the output buffer'
    if written and n:
        assert not bool((view == SENT).any()), 'output not fully written'
        assert bool(torch.isfinite(view).all())


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
@pytest.mark.parametrize('N', [1, 33, 1000, 40_001, 70_003, 262_147])
@pytest.mark.parametrize('misalign', [0, 1])
def test_flow_outputs_stay_inside_their_buffers(precision, N, misalign, cuda_device):
    import torch
    from cnf_b200 import _lib
    from cnf_b200._engine import _ptr, _stream
    g = load_golden('flow_c2_nvp_k10')
    flow = build_flow_from_golden(g, cuda_device, precision=precision)
    eng = flow.engine()
    eng.ensure(cuda_device)
    eng.pack(tc=True)
    K = 10
    x, y = orc.synth_logits(N, K, seed=N)
    xt, yt = torch.from_numpy(x).to(cuda_device), torch.from_numpy(y).to(cuda_device)
    desc = eng.desc_tc if precision == 'bf16' else eng.desc
    packed = eng.packed_tc if precision == 'bf16' else eng.packed
    st = _stream(cuda_device)
    for fn in ('cnf_flow_forward', 'cnf_flow_inverse'):
        zb, z = guarded(N * K, torch.float32, cuda_device, misalign)
        lb, ld = guarded(N, torch.float32, cuda_device, misalign)
        _lib.call(fn, ctypes.byref(desc), _ptr(packed), _ptr(eng.tables), _ptr(xt), _ptr(z), _ptr(ld), None,
                  ctypes.c_int64(N), st)
        torch.cuda.synchronize()
        check(zb, z, N * K, misalign)
        check(lb, ld, N, misalign)
    if precision == 'fp32':            # intermediate outputs zs [L, N, K]
        zb, z = guarded(N * K, torch.float32, cuda_device)
        lb, ld = guarded(N, torch.float32, cuda_device)
        ab, allz = guarded(eng.L * N * K, torch.float32, cuda_device)
        _lib.call('cnf_flow_forward', ctypes.byref(desc), _ptr(packed), _ptr(eng.tables), _ptr(xt), _ptr(z), _ptr(ld),
                  _ptr(allz), ctypes.c_int64(N), st)
        torch.cuda.synchronize()
        check(ab, allz, eng.L * N * K)
    # fused predict: z, log-det, float64 probabilities, statistics
    lp = torch.as_tensor(orc.log_priors(orc.onehot_encode(np.concatenate([y, np.arange(K)])))).to(cuda_device)
    edges = torch.tensor([i * (1. / 15) for i in range(16)], dtype=torch.float64, device=cuda_device)
    zb, z = guarded(N * K, torch.float32, cuda_device, misalign)
    lb, ld = guarded(N, torch.float32, cuda_device)
    pb, pr = guarded(N * K, torch.float64, cuda_device)
    sb, acc = guarded(48, torch.float64, cuda_device)
    acc.zero_()
    _lib.call('cnf_flow_predict', ctypes.byref(desc), _ptr(packed), _ptr(eng.tables), _ptr(xt), ctypes.c_int64(N),
              ctypes.c_int32(1), ctypes.c_int32(_lib.METRICS_CALIBRATED), _ptr(lp), _ptr(z), _ptr(ld), _ptr(pr),
              _ptr(yt), ctypes.c_int32(15), _ptr(edges), _ptr(acc), st)
    torch.cuda.synchronize()
    check(zb, z, N * K, misalign)
    check(lb, ld, N)
    check(pb, pr, N * K)
    check(sb, acc, 48)
    assert float(acc[47]) == N
    # training step: flat gradient and loss sums
    eng.flat_grad = None
    fb, fg = guarded(eng.n_flat, torch.float32, cuda_device)
    eng._want_partials()
    eng.flat_grad = fg
    ab, la = guarded(4, torch.float64, cuda_device)
    la.zero_()
    eng.nll_step(xt, yt, la, precision=precision)
    torch.cuda.synchronize()
    check(fb, fg, eng.n_flat)
    check(ab, la, 4)


@pytest.mark.parametrize('K,N', [(3, 1), (10, 4097), (100, 513), (7, 30_001)])
def test_streaming_layers_and_metrics_stay_inside_their_buffers(K, N, cuda_device):
    import torch
    from cnf_b200 import _lib
    from cnf_b200._engine import _ptr, _stream
    rs = np.random.RandomState(K + N)
    x = torch.from_numpy(rs.randn(N, K).astype(np.float32)).to(cuda_device)
    y = torch.from_numpy(rs.randint(0, K, N)).to(cuda_device)
    s = torch.from_numpy(0.1 * rs.randn(K).astype(np.float32)).to(cuda_device)
    t = torch.from_numpy(rs.randn(K).astype(np.float32)).to(cuda_device)
    one = torch.tensor([0.3], device=cuda_device)
    st = _stream(cuda_device)
    n, k = ctypes.c_int64(N), ctypes.c_int32(K)
    zb, z = guarded(N * K, torch.float32, cuda_device)
    _lib.call('cnf_affine_const', _ptr(x), _ptr(s), _ptr(t), _ptr(z), n, k, ctypes.c_int32(0), st)
    torch.cuda.synchronize()
    check(zb, z, N * K)
    gxb, gx = guarded(N * K, torch.float32, cuda_device)
    gsb, gs = guarded(K, torch.float32, cuda_device)
    gtb, gt = guarded(K, torch.float32, cuda_device)
    _lib.call('cnf_affine_const_backward', _ptr(x), _ptr(x), _ptr(s), _ptr(gx), _ptr(gs), _ptr(gt), n, k, st)
    torch.cuda.synchronize()
    check(gxb, gx, N * K); check(gsb, gs, K); check(gtb, gt, K)
    zb, z = guarded(N * K, torch.float32, cuda_device)
    lb, ld = guarded(N, torch.float32, cuda_device)
    _lib.call('cnf_planar_forward', _ptr(x), _ptr(s), _ptr(t), _ptr(one), _ptr(z), _ptr(ld), n, k, st)
    torch.cuda.synchronize()
    check(zb, z, N * K); check(lb, ld, N, written=False)
    gxb, gx = guarded(N * K, torch.float32, cuda_device)
    gwb, gw = guarded(K, torch.float32, cuda_device)
    gub, gu = guarded(K, torch.float32, cuda_device)
    gbb, gb = guarded(1, torch.float32, cuda_device)
    _lib.call('cnf_planar_backward', _ptr(x), _ptr(x), None, _ptr(s), _ptr(t), _ptr(one), _ptr(gx), _ptr(gw), _ptr(gu),
              _ptr(gb), n, k, st)
    torch.cuda.synchronize()
    check(gxb, gx, N * K); check(gwb, gw, K); check(gub, gu, K); check(gbb, gb, 1)
    zb, z = guarded(N * K, torch.float32, cuda_device)
    _lib.call('cnf_radial_forward', _ptr(x), _ptr(t), _ptr(one), _ptr(one), _ptr(z), n, k, st)
    torch.cuda.synchronize()
    check(zb, z, N * K)
    gxb, gx = guarded(N * K, torch.float32, cuda_device)
    gzb, gz0 = guarded(K, torch.float32, cuda_device)
    gab, ga = guarded(1, torch.float32, cuda_device)
    gbb, gb = guarded(1, torch.float32, cuda_device)
    _lib.call('cnf_radial_backward', _ptr(x), _ptr(x), _ptr(t), _ptr(one), _ptr(one), _ptr(gx), _ptr(gz0), _ptr(ga),
              _ptr(gb), n, k, st)
    torch.cuda.synchronize()
    check(gxb, gx, N * K); check(gzb, gz0, K); check(gab, ga, 1); check(gbb, gb, 1)
    # metrics: 3*bins+3 statistics; calibrated probabilities
    p = torch.softmax(x, dim=1).contiguous()
    for bins in (1, 15, 100):
        edges = torch.tensor([i * (1. / bins) for i in range(bins + 1)], dtype=torch.float64, device=cuda_device)
        ab, acc = guarded(3 * bins + 3, torch.float64, cuda_device)
        acc.zero_()
        _lib.call('cnf_metrics', _ptr(p), ctypes.c_int32(0), _ptr(y), n, k, ctypes.c_int32(bins), ctypes.c_int32(0), None,
                  _ptr(edges), _ptr(acc), st)
        torch.cuda.synchronize()
        check(ab, acc, 3 * bins + 3)
        assert float(acc[3 * bins + 2]) == N and float(acc[:bins].sum()) == N
    lp = torch.log(torch.full((K,), 1.0 / K, dtype=torch.float64, device=cuda_device))
    pb, pr = guarded(N * K, torch.float64, cuda_device)
    _lib.call('cnf_calibrated_probs', _ptr(x), n, k, _ptr(lp), _ptr(pr), st)
    torch.cuda.synchronize()
    check(pb, pr, N * K)


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_host_buffer_call_stays_inside_pinned_buffers(precision, cuda_device):
    import torch
    g = load_golden('flow_c2_nvp_k10')
    flow = build_flow_from_golden(g, cuda_device, precision=precision)
    eng = flow.engine()
    for N in (5, 70_001, 200_003):
        x, _ = orc.synth_logits(N, 10, seed=N)
        xh = torch.from_numpy(x).pin_memory()
        zbuf = torch.full((N * 10 + 2 * G,), SENT, dtype=torch.float32).pin_memory()
        lbuf = torch.full((N + 2 * G,), SENT, dtype=torch.float32).pin_memory()
        zh, lh = zbuf[G:G + N * 10].view(N, 10), lbuf[G:G + N]
        eng.apply_host(xh, zh, lh, precision=precision, device=cuda_device, chunk=65536)
        torch.cuda.synchronize()
        for b, n in ((zbuf, N * 10), (lbuf, N)):
            assert bool((b[:G] == SENT).all()) and bool((b[G + n:] == SENT).all())
            assert not bool((b[G:G + n] == SENT).any())
