"""CPU oracle for the coupling-flow calibration hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-numpy restatement of the reference algorithm.  It is the
checker for the CUDA path; nothing under ``calibration-normalizing-flows_b200/``
may import it.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it.

Pinning: the reference ships no tests or golden vectors, so
this oracle is pinned against outputs of the reference itself, run in the
authoring container by ``oracle/make_golden.py`` (imports ``/root/reference``)
and committed under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks
every function here against those fixtures.

Reference lines followed (paths relative to the reference root):
  * MLP                      flows/utils.py:6-31
  * coupling forward         flows/flows.py:101-112
  * coupling inverse         flows/flows.py:114-126
  * stack forward / inverse  flows/flows.py:17-37
  * calibrator loss head     calibrators.py:287-291 (eps=1e-7, gamma=1)
  * script loss head         run_experiment3D.py:107 (eps=0, gamma=det)
  * Calibrator.predict tail  calibrators.py:40-44, 350-353
  * log priors               calibrators.py:31-35
  * ECE / NLL / accuracy     utils/metrics.py:35-73, 6-15, 76-80
  * one-hot                  utils/ops.py:42-51
  * Adam / SGD               torch.optim semantics used at calibrators.py:259,
                             run_experiment3D.py:54-61 (third-party arithmetic)

Parameter container ("params"): list over coupling layers of dicts
  {'s': [(W, b), ...] | None, 't': [(W, b), ...] | None, 'perm': int array | None}
with W in torch layout [out, in].
"""
import numpy as np


# --------------------------------------------------------------------------- #
# parameters
# --------------------------------------------------------------------------- #
def init_params(K, L, hidden, scale=True, shift=True, rng=None, wscale=0.001,
                random_flip=False, dtype=np.float32):
    """nn.Linear-style U(-1/sqrt(in), 1/sqrt(in)) init times ``wscale``
    (flows/utils.py:17-22; flows/flows.py:76-79).  Same distribution as the
    reference, not the same random stream."""
    rng = rng or np.random.default_rng(0)
    units = [K] + list(hidden) + [K]

    def net():
        out = []
        for i, o in zip(units[:-1], units[1:]):
            bound = 1.0 / np.sqrt(i)
            W = rng.uniform(-bound, bound, size=(o, i)) * wscale
            b = rng.uniform(-bound, bound, size=(o,)) * wscale
            out.append((W.astype(dtype), b.astype(dtype)))
        return out

    params = []
    for _ in range(L):
        params.append({'s': net() if scale else None,
                       't': net() if shift else None,
                       'perm': rng.permutation(K) if random_flip else None})
    return params


def cast_params(params, dtype):
    out = []
    for lay in params:
        new = {'perm': lay.get('perm')}
        for name in ('s', 't'):
            net = lay[name]
            new[name] = None if net is None else [(W.astype(dtype), b.astype(dtype)) for W, b in net]
        out.append(new)
    return out


def flatten(params):
    """Flat trainable vector in the framework's canonical order: per layer, the
    s-net then the t-net, each Linear as weight.ravel() then bias."""
    chunks = []
    for lay in params:
        for name in ('s', 't'):
            if lay[name] is not None:
                for W, b in lay[name]:
                    chunks.append(np.asarray(W).ravel())
                    chunks.append(np.asarray(b).ravel())
    return np.concatenate(chunks) if chunks else np.zeros(0)


def unflatten(flat, like):
    out, off = [], 0
    for lay in like:
        new = {'perm': lay.get('perm')}
        for name in ('s', 't'):
            if lay[name] is None:
                new[name] = None
                continue
            net = []
            for W, b in lay[name]:
                w = flat[off:off + W.size].reshape(W.shape); off += W.size
                bb = flat[off:off + b.size].reshape(b.shape); off += b.size
                net.append((w, bb))
            new[name] = net
        out.append(new)
    assert off == flat.size
    return out


# --------------------------------------------------------------------------- #
# forward / inverse
# --------------------------------------------------------------------------- #
def mask_of(K, dtype):
    m = np.zeros((1, K), dtype=dtype)          # flows/flows.py:81-82
    m[:, K // 2:] = 1
    return m


def mlp(net, x, keep=None):
    """flows/utils.py:26-31 with the default ReLU activation."""
    h = x
    for W, b in net[:-1]:
        a = h @ W.T + b
        h = np.maximum(a, 0)
        if keep is not None:
            keep.append(h)
    W, b = net[-1]
    return h @ W.T + b


def layer_forward(lay, x):
    K = x.shape[1]
    m = mask_of(K, x.dtype)
    xb = m * x
    b1 = 1 - m
    s = mlp(lay['s'], xb) if lay['s'] is not None else np.zeros_like(x)
    t = mlp(lay['t'], xb) if lay['t'] is not None else np.zeros_like(x)
    z = xb + b1 * (x * np.exp(s) + t)
    ld = np.sum(b1 * s, axis=1)
    if lay.get('perm') is not None:
        z = z[:, np.asarray(lay['perm'])]
    return z[:, ::-1].copy(), ld


def layer_inverse(lay, z):
    K = z.shape[1]
    z = z[:, ::-1]
    if lay.get('perm') is not None:
        perm = np.asarray(lay['perm'])
        rev = np.zeros(K, dtype=np.int64)
        rev[perm] = np.arange(K)
        z = z[:, rev]
    m = mask_of(K, z.dtype)
    xb = m * z
    b1 = 1 - m
    s = mlp(lay['s'], xb) if lay['s'] is not None else np.zeros_like(z)
    t = mlp(lay['t'], xb) if lay['t'] is not None else np.zeros_like(z)
    x = xb + b1 * (z - t) * np.exp(-s)
    ld = np.sum(b1 * (-s), axis=1)
    return x, ld


def flow_forward(params, x):
    zs, cum = [], np.zeros(x.shape[0], dtype=x.dtype)
    for lay in params:
        x, ld = layer_forward(lay, x)
        zs.append(x)
        cum = cum + ld
    return zs, cum


def flow_inverse(params, z):
    xs, cum = [], np.zeros(z.shape[0], dtype=z.dtype)
    for lay in params[::-1]:
        z, ld = layer_inverse(lay, z)
        xs.append(z)
        cum = cum + ld
    return xs, cum


# --------------------------------------------------------------------------- #
# loss heads and analytic gradients (checked against reference autograd by
# oracle/make_golden.py)
# --------------------------------------------------------------------------- #
def softmax(z):
    e = np.exp(z - z.max(axis=1, keepdims=True))
    return e / e.sum(axis=1, keepdims=True)


def nll_head(z, ld, y, eps=1e-7, gamma=1.0, n_total=None):
    """Returns (loss, ce_mean, logdet_mean, g_z, g_ld).
    eps>0: calibrators.py:288-291; eps==0: CrossEntropyLoss, run_experiment3D.py:107."""
    N = z.shape[0]
    n_total = n_total or N
    p = softmax(z)
    py = p[np.arange(N), y]
    if eps == 0:
        zc = z - z.max(axis=1, keepdims=True)
        ce = zc[np.arange(N), y] - np.log(np.exp(zc).sum(axis=1))
        coef = np.ones_like(py)
    else:
        ce = np.log(py + eps)
        coef = py / (py + eps)
    loss = -(ce.sum() + gamma * ld.sum()) / n_total
    onehot = np.zeros_like(p)
    onehot[np.arange(N), y] = 1
    gz = -(coef[:, None] * (onehot - p)) / n_total
    gld = np.full(N, -gamma / n_total, dtype=z.dtype)
    return loss, -ce.sum() / n_total, ld.sum() / n_total, gz, gld


def _mlp_backward(net, x_in, hs, g_out):
    """Gradient of mlp(); returns (g_x_in, [(gW, gb), ...])."""
    grads = [None] * len(net)
    g = g_out
    for j in range(len(net) - 1, -1, -1):
        W, _ = net[j]
        inp = x_in if j == 0 else hs[j - 1]
        grads[j] = (g.T @ inp, g.sum(axis=0))
        g = g @ W
        if j > 0:
            g = g * (hs[j - 1] > 0)
    return g, grads


def flow_backward(params, x, g_z, g_ld):
    """Backprop of (z_L, sum ld) wrt parameters and x.  Returns (grads, g_x) where
    grads has the structure of params."""
    K = x.shape[1]
    m = mask_of(K, x.dtype)
    b1 = 1 - m
    # forward, keeping what autograd would keep
    tape, v = [], x
    for lay in params:
        xb = m * v
        hs_s, hs_t = [], []
        s = mlp(lay['s'], xb, hs_s) if lay['s'] is not None else np.zeros_like(v)
        t = mlp(lay['t'], xb, hs_t) if lay['t'] is not None else np.zeros_like(v)
        es = np.exp(s)
        yv = xb + b1 * (v * es + t)
        tape.append((v, xb, hs_s, hs_t, es))
        if lay.get('perm') is not None:
            yv = yv[:, np.asarray(lay['perm'])]
        v = yv[:, ::-1]
    grads = [None] * len(params)
    g = g_z
    for li in range(len(params) - 1, -1, -1):
        lay = params[li]
        v, xb, hs_s, hs_t, es = tape[li]
        g = g[:, ::-1]
        if lay.get('perm') is not None:
            perm = np.asarray(lay['perm'])
            gy = np.zeros_like(g)
            gy[:, perm] = g
            g = gy
        g_t = b1 * g
        g_s = b1 * (g * v * es + g_ld[:, None])
        gx = m * g + b1 * g * es
        lg = {'perm': lay.get('perm'), 's': None, 't': None}
        if lay['s'] is not None:
            gin, lg['s'] = _mlp_backward(lay['s'], xb, hs_s, g_s)
            gx = gx + m * gin
        if lay['t'] is not None:
            gin, lg['t'] = _mlp_backward(lay['t'], xb, hs_t, g_t)
            gx = gx + m * gin
        grads[li] = lg
        g = gx
    return grads, g


def train_step_grads(params, x, y, eps=1e-7, gamma=1.0, n_total=None):
    zs, ld = flow_forward(params, x)
    loss, ce, ldm, gz, gld = nll_head(zs[-1], ld, y, eps, gamma, n_total)
    grads, gx = flow_backward(params, x, gz, gld)
    return loss, ce, ldm, grads, gx


# --------------------------------------------------------------------------- #
# optimisers (torch.optim semantics)
# --------------------------------------------------------------------------- #
def adam_step(p, g, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, wd=0.0):
    """One torch.optim.Adam step (t is 1-based).  Returns new (p, m, v)."""
    if wd != 0:
        g = g + wd * p
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1 = 1 - b1 ** t
    bc2 = 1 - b2 ** t
    denom = np.sqrt(v) / np.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


def sgd_step(p, g, lr, wd=0.0):
    if wd != 0:
        g = g + wd * p
    return p - lr * g


# --------------------------------------------------------------------------- #
# calibrator pieces and metrics
# --------------------------------------------------------------------------- #
def onehot_encode(target):
    target = np.array(target).astype(np.int32)       # utils/ops.py:42-51
    n, k = len(target), np.max(target) + 1
    oh = np.zeros((n, k))
    oh[np.arange(n), target] = 1.
    return oh.astype(np.int32)


def log_priors(target_onehot):
    pri = np.sum(target_onehot, axis=0)               # calibrators.py:31-35
    pri = pri / np.sum(pri)
    return np.log(pri)


def center(logits):
    return logits - np.mean(logits, axis=1, keepdims=True)   # calibrators.py:17, 42


def calibrated_probs(z_last, logpri):
    """calibrators.py:350-353 then :44, in float64 on the given z."""
    z = np.asarray(z_last)
    e = np.exp(z - z.max(axis=1, keepdims=True))
    probs = e / e.sum(axis=1, keepdims=True)
    u = np.log(probs + 1e-7) - logpri
    e = np.exp(u - u.max(axis=1, keepdims=True))
    return e / e.sum(axis=1, keepdims=True)


def neg_log_likelihood(probs, target):
    if target.shape != probs.shape:                   # utils/metrics.py:6-15
        target = onehot_encode(target)
    return np.mean(-np.sum(target * np.log(probs + 1e-7), axis=1))


def accuracy(probs, target):
    if target.shape != probs.shape:                   # utils/metrics.py:76-80
        target = onehot_encode(target)
    return np.mean(np.argmax(probs, axis=1) == np.argmax(target, axis=1))


def ece_bins(probs, target, bins=15):
    """Per-bin (count, sum_conf, sum_acc) with the reference's right-closed bins
    (utils/metrics.py:56-70); edges are i*width in Python floats, compared in the
    array's dtype (NumPy weak-scalar promotion)."""
    preds = np.argmax(probs, axis=1)
    if target.shape == probs.shape:
        target = np.argmax(target, axis=1)
    conf = probs[np.arange(probs.shape[0]), preds]
    accs = (preds == target).astype(np.float64)
    width = 1. / bins
    cnt = np.zeros(bins, dtype=np.int64)
    sconf = np.zeros(bins)
    sacc = np.zeros(bins)
    for i in range(bins):
        low, high = i * width, (i + 1) * width
        idx = np.where((low < conf) & (conf <= high))
        cnt[i] = idx[0].size
        sconf[i] = conf[idx].astype(np.float64).sum()
        sacc[i] = accs[idx].sum()
    return cnt, sconf, sacc


def expected_calibration_error(probs, target, bins=15):
    cnt, sconf, sacc = ece_bins(probs, target, bins)   # utils/metrics.py:35-73
    e = 0.0
    for i in range(bins):
        if cnt[i] > 0:
            e += abs(sacc[i] / cnt[i] - sconf[i] / cnt[i]) * cnt[i]
    return e / probs.shape[0]


# --------------------------------------------------------------------------- #
# PlanarLayer / RadialLayer (flows/flows.py:129-193), forward direction
# --------------------------------------------------------------------------- #
def planar_forward(w, u, b, x):
    """flows/flows.py:148-164.  Returns (z, log_det, u_hat)."""
    w, u, x = np.asarray(w), np.asarray(u), np.asarray(x)
    b = np.asarray(b).reshape(-1)[0]
    wtu = w @ u
    m = -1 + np.log1p(np.exp(wtu))
    u_hat = u + (m - wtu) * w / np.linalg.norm(w)
    h = np.tanh(x @ w + b)
    z = x + h[:, None] * u_hat[None, :]
    hp = 1 - h ** 2
    log_det = np.log(np.abs(1 + hp * (w @ u_hat)))
    return z, log_det, u_hat


def planar_backward(w, u_hat, b, x, g_z, g_ld):
    """Gradients w.r.t. (x, w, u_hat, b) treating u_hat as an input (the caller chains u_hat -> (w, u))."""
    w, u_hat, x, g_z, g_ld = (np.asarray(v, dtype=np.float64) for v in (w, u_hat, x, g_z, g_ld))
    b = float(np.asarray(b).reshape(-1)[0])
    h = np.tanh(x @ w + b)
    hp = 1 - h ** 2
    wu = w @ u_hat
    D = 1 + hp * wu
    g_a = (g_z @ u_hat) * hp - 2 * h * hp * (g_ld * wu / D)
    c = (g_ld * hp / D).sum()
    return g_z + g_a[:, None] * w[None, :], x.T @ g_a + c * u_hat, g_z.T @ h + c * w, g_a.sum()


def radial_forward(z0, a, b, x):
    """flows/flows.py:180-193.  Returns (z, log_det) with the reference's constant log_det = log(1.0)."""
    z0, x = np.asarray(z0), np.asarray(x)
    a = np.asarray(a).reshape(-1)[0]
    b = np.asarray(b).reshape(-1)[0]
    b_hat = -a + np.log1p(np.exp(b))
    d = x - z0
    h = 1.0 / (a + np.linalg.norm(d, axis=1, keepdims=True))
    return x + b_hat * h * d, np.log(1.0)


# --------------------------------------------------------------------------- #
# index-map restatement (SURVEY.md Appendix A "folding flips/perms"); used to
# check the C++ planner in csrc/cnf_plan.cpp
# --------------------------------------------------------------------------- #
def pi_maps(K, L, perms=None):
    """pi[l][j] = physical slot of logical dim j at the input of layer l."""
    pi = [np.arange(K)]
    for l in range(L):
        perm = np.arange(K) if perms is None or perms[l] is None else np.asarray(perms[l])
        prev = pi[-1]
        pi.append(np.array([prev[perm[K - 1 - j]] for j in range(K)]))
    return pi


# --------------------------------------------------------------------------- #
# synthetic data shared by oracle and kernels
# --------------------------------------------------------------------------- #
def synth_logits(N, K, seed, dtype=np.float32):
    rng = np.random.default_rng(seed)
    y = rng.integers(0, K, size=N)
    hit = rng.random(N) < 0.8
    x = 1.5 * rng.standard_normal((N, K))
    x[np.arange(N), y] += 3.0 * hit
    x = x - x.mean(axis=1, keepdims=True)
    return x.astype(dtype), y.astype(np.int64)
