"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

TEST INFRASTRUCTURE ONLY.  Usage (authoring container, needs /root/reference):

    python oracle/make_golden.py

The reference has no golden vectors of its own, so these
fixtures are the parity pin: outputs of the reference's own ``Flow`` /
``NvpCouplingLayer`` / ``TorchFlowCalibrator`` / ``utils.metrics`` on seeded
inputs and seeded weights.  The GPU box has no /root/reference; tests there read
only the committed .npz files.

Shims applied in memory (never written to disk):
  * utils/metrics.py:54 ``np.equal(..., dtype=np.float32)`` fails on NumPy>=2
     -> ``np.equal(...).astype(np.float32)``.
  * ``flows.nice_torch.NiceFlow`` / ``flows.realNVP_torch.RealNvpFlow`` do not
    exist in the tree -> adapter built from the in-tree
    ``Flow`` + ``NvpCouplingLayer`` with the contract calibrators.py:251,287 expects.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get('CNF_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '..', 'tests', 'golden')
sys.path.insert(0, REF)
sys.path.insert(0, HERE)

from flows.flows import Flow, NvpCouplingLayer  # noqa: E402  (reference)
import calibrators as ref_cal                    # noqa: E402  (reference)
import flow_oracle as orc                        # noqa: E402


def load_ref_metrics():
    src = open(os.path.join(REF, 'utils', 'metrics.py')).read()
    src = src.replace('np.equal(preds, target, dtype=np.float32)',
                      'np.equal(preds, target).astype(np.float32)')
    src = src.replace('from .ops import onehot_encode', 'from utils.ops import onehot_encode')
    mod = types.ModuleType('ref_metrics')
    exec(compile(src, 'utils/metrics.py', 'exec'), mod.__dict__)
    return mod


def build_flow(K, L, hidden, scale, shift, random_flip, seed, wmult):
    torch.manual_seed(seed)
    np.random.seed(seed)
    flow = Flow([NvpCouplingLayer(K, hidden_size=list(hidden), scale=scale, shift=shift,
                                  random_flip=random_flip) for _ in range(L)])
    with torch.no_grad():
        for p in flow.parameters():
            if p.requires_grad:
                p.mul_(wmult)
    return flow


def flat_of(flow, grad=False):
    """Framework canonical flat order: per layer s-net then t-net, weight then bias."""
    chunks = []
    for lay in flow.layers:
        for net in (lay.s, lay.t):
            if isinstance(net, torch.nn.Module):
                for lin in net.layers:
                    for p in (lin.weight, lin.bias):
                        t = p.grad if grad else p
                        chunks.append(t.detach().reshape(-1).numpy().copy())
    return np.concatenate(chunks)


def perms_of(flow, K):
    out = []
    for lay in flow.layers:
        if lay.random_flip:
            out.append(lay.perm.detach().numpy().reshape(-1).astype(np.int32))
        else:
            out.append(np.arange(K, dtype=np.int32))
    return np.stack(out)


FLOW_CASES = [
    # name, K, L, hidden, scale, shift, random_flip, wmult, N
    ('c1_nice_k3', 3, 4, [32], False, True, False, 300, 64),
    ('nvp_k3_h33', 3, 5, [3, 3], True, True, False, 300, 64),
    ('nvp_k5_oddL', 5, 3, [5, 5], True, True, False, 300, 64),
    ('c2_nvp_k10', 10, 6, [128], True, True, False, 300, 48),
    ('c2_nvp_k10_init', 10, 6, [128], True, True, False, 1, 32),
    ('nvp_k10_nohidden', 10, 2, [], True, True, False, 300, 64),
    ('nvp_k7_randflip', 7, 5, [16], True, True, True, 300, 64),
    ('nvp_k40_wide', 40, 2, [24], True, True, False, 200, 32),
    ('scaleonly_k4_h888', 4, 3, [8, 8, 8], True, False, False, 300, 64),
    ('nvp_k2', 2, 4, [7], True, True, False, 300, 33),
]


def gen_flow_cases():
    for i, (name, K, L, hidden, scale, shift, rf, wmult, N) in enumerate(FLOW_CASES):
        flow = build_flow(K, L, hidden, scale, shift, rf, seed=100 + i, wmult=wmult)
        x, y = orc.synth_logits(N, K, seed=1234 + i)
        xt = torch.as_tensor(x)
        with torch.no_grad():
            zs, ld = flow(xt)
            xs, ldi = flow.backward(zs[-1])
        out = dict(K=K, L=L, hidden=np.array(hidden, dtype=np.int32), scale=int(scale), shift=int(shift),
                   random_flip=int(rf), perms=perms_of(flow, K), flat=flat_of(flow), x=x, y=y,
                   zs=np.stack([z.numpy() for z in zs]), logdet=ld.numpy(),
                   xs=np.stack([v.numpy() for v in xs]), logdet_inv=ldi.numpy())

        # float64 run of the same reference modules (tolerance floor / bf16 comparisons)
        flow64 = build_flow(K, L, hidden, scale, shift, rf, seed=100 + i, wmult=wmult).double()
        with torch.no_grad():
            zs64, ld64 = flow64(torch.as_tensor(x).double())
        out['z64'] = zs64[-1].numpy()
        out['logdet64'] = ld64.numpy()

        # one training step through reference autograd, both loss heads
        yt = torch.as_tensor(y)
        for tag, eps, gamma in (('cal', 1e-7, 1.0), ('script', 0.0, 1.0), ('script_nodet', 0.0, 0.0)):
            flow.zero_grad()
            xg = torch.as_tensor(x).clone().requires_grad_(True)
            zs, ld = flow(xg)
            if eps > 0:   # calibrators.py:287-291
                probs = torch.softmax(zs[-1], dim=1)
                ce = torch.log(probs.gather(1, yt.view(-1, 1)) + 1e-7)
                loss = -torch.mean(ce.squeeze() + ld)
            else:         # run_experiment3D.py:107
                loss = torch.nn.CrossEntropyLoss()(zs[-1], yt) - gamma * torch.mean(ld)
            loss.backward()
            out['loss_' + tag] = float(loss)
            g = []
            for lay in flow.layers:
                for net in (lay.s, lay.t):
                    if isinstance(net, torch.nn.Module):
                        for lin in net.layers:
                            for p in (lin.weight, lin.bias):
                                g.append((torch.zeros_like(p) if p.grad is None else p.grad)
                                         .reshape(-1).numpy().copy())
            out['grad_' + tag] = np.concatenate(g)
            out['gx_' + tag] = xg.grad.numpy().copy()

        # three Adam steps, calibrator head, torch.optim.Adam defaults (calibrators.py:259)
        flow = build_flow(K, L, hidden, scale, shift, rf, seed=100 + i, wmult=wmult)
        opt = torch.optim.Adam([p for p in flow.parameters() if p.requires_grad])
        losses = []
        for _ in range(3):
            zs, ld = flow(xt)
            probs = torch.softmax(zs[-1], dim=1)
            ce = torch.log(probs.gather(1, yt.view(-1, 1)) + 1e-7)
            loss = -torch.mean(ce.squeeze() + ld)
            flow.zero_grad()
            loss.backward()
            opt.step()
            losses.append(float(loss))
        out['adam3_losses'] = np.array(losses)
        out['adam3_flat'] = flat_of(flow)

        # two SGD steps with weight decay, script head (run_experiment3D.py:59-61)
        flow = build_flow(K, L, hidden, scale, shift, rf, seed=100 + i, wmult=wmult)
        opt = torch.optim.SGD([p for p in flow.parameters() if p.requires_grad], lr=1e-2, weight_decay=1e-2)
        for _ in range(2):
            zs, ld = flow(xt)
            loss = torch.nn.CrossEntropyLoss()(zs[-1], yt) - torch.mean(ld)
            opt.zero_grad()
            loss.backward()
            opt.step()
        out['sgd2_flat'] = flat_of(flow)

        np.savez_compressed(os.path.join(OUT, 'flow_%s.npz' % name), **out)
        print('flow case', name, 'loss', out['loss_cal'])


class RefAdapter(torch.nn.Module):
    """The (pred, log_det) contract calibrators.py:251,287,304,339 expects."""

    def __init__(self, dim, layers=4, hidden_size=None, scale=True, seed=0, **ignored):
        super().__init__()
        hidden_size = [dim] if hidden_size is None else list(hidden_size)
        torch.manual_seed(seed)
        self.flow = Flow([NvpCouplingLayer(dim, hidden_size=hidden_size, scale=scale)
                          for _ in range(layers)])
        self.layers = self.flow.layers

    def forward(self, x):
        zs, ld = self.flow(x)
        return zs[-1], ld


def gen_calibrator_case():
    for name, K, layers, hidden, scale, N, epochs in (
            ('cal_nice_k3', 3, 4, [32], False, 400, 6),
            ('cal_nvp_k10', 10, 3, [16], True, 300, 5)):
        x, y = orc.synth_logits(N, K, seed=77)
        x = x + 0.37          # un-centred on purpose: Calibrator.__init__ must centre
        xt, _ = orc.synth_logits(200, K, seed=78)
        init = RefAdapter(K, layers=layers, hidden_size=hidden, scale=scale, seed=5)
        flat0 = flat_of(init.flow)
        torch.manual_seed(9)
        cal = ref_cal.TorchFlowCalibrator(RefAdapter, x, y, layers=layers, hidden_size=hidden,
                                          scale=scale, seed=5, epochs=epochs, dev=torch.device('cpu'))
        hist = {k: np.array([float(v) for v in cal.history[k]]) for k in ('loss', 'ce', 'log_det')}
        pred = cal.predict(xt)
        pred_logits = cal.predict_logits(orc.center(xt))
        np.savez_compressed(os.path.join(OUT, 'calibrator_%s.npz' % name), K=K, layers=layers,
                            hidden=np.array(hidden, dtype=np.int32), scale=int(scale), epochs=epochs,
                            x=x, y=y, x_test=xt, flat0=flat0, flat_end=flat_of(cal.flow.flow),
                            log_priors=cal.log_priors, pred=pred, pred_logits=pred_logits,
                            **{'hist_' + k: v for k, v in hist.items()})
        print('calibrator case', name, hist['loss'])


def gen_metrics_cases():
    met = load_ref_metrics()
    rng = np.random.default_rng(3)
    cases = {}
    for name, N, K, dtype in (('f64_k10', 5000, 10, np.float64), ('f32_k10', 5000, 10, np.float32),
                              ('f64_k3', 777, 3, np.float64), ('f32_k100', 512, 100, np.float32)):
        x, y = orc.synth_logits(N, K, seed=500 + K)
        p = orc.softmax(x.astype(np.float64) * 1.7).astype(dtype)
        cases[name] = (p, y)
    # confidences sitting exactly on bin edges and at 1.0 (right-closed bins)
    edges = np.array([i * (1. / 15) for i in range(1, 16)])
    p = np.zeros((15, 2))
    p[:, 0] = edges
    p[:, 1] = 1 - edges
    y = (rng.random(15) < 0.5).astype(np.int64)
    cases['edges_f64'] = (p, y)
    cases['edges_f32'] = (p.astype(np.float32), y)
    out = {}
    for name, (p, y) in cases.items():
        out[name + '_probs'] = p
        out[name + '_y'] = y
        out[name + '_ece15'] = met.expected_calibration_error(p, y, bins=15)
        out[name + '_ece10'] = met.expected_calibration_error(p, y, bins=10)
        yy = y.copy()
        if yy.max() < p.shape[1] - 1:      # onehot_encode sizes by max label (utils/ops.py:46)
            yy_oh = np.zeros(p.shape, dtype=np.int32)
            yy_oh[np.arange(len(y)), y] = 1
        else:
            yy_oh = yy
        out[name + '_nll'] = met.neg_log_likelihood(p, yy_oh)
        out[name + '_acc'] = met.accuracy(p, yy_oh)
    np.savez_compressed(os.path.join(OUT, 'metrics.npz'), **out)
    print('metrics cases', {k: float(v) for k, v in out.items() if k.endswith('ece15')})


def gen_affine_case():
    """AffineConstantLayer / TempScaler (flows/flows.py:40-65, flows/utils.py:34-48) inside a Flow with
    coupling layers on both sides, as notebooks/train-flows-det.ipynb cell 11 stacks them."""
    from flows.flows import AffineConstantLayer
    from flows.utils import TempScaler
    torch.manual_seed(31)
    K, N = 6, 50
    aff = AffineConstantLayer(K)
    with torch.no_grad():
        aff.s.copy_(0.3 * torch.randn(1, K))
        aff.t.copy_(torch.randn(1, K))
    c0 = NvpCouplingLayer(K, hidden_size=[8])
    c1 = NvpCouplingLayer(K, hidden_size=[8])
    with torch.no_grad():
        for lay in (c0, c1):
            for p in lay.parameters():
                if p.requires_grad:
                    p.mul_(300.0)
    flow = Flow([c0, aff, c1])
    x, y = orc.synth_logits(N, K, seed=9)
    xt = torch.as_tensor(x).requires_grad_(True)
    zs, ld = flow(xt)
    loss = torch.nn.CrossEntropyLoss()(zs[-1], torch.as_tensor(y)) - torch.mean(ld)
    loss.backward()
    with torch.no_grad():
        xs, ldi = flow.backward(zs[-1])
    ts = TempScaler()
    with torch.no_grad():
        ts.T.fill_(-1.7)
    zt = ts(torch.as_tensor(x).requires_grad_(True))
    lt = (zt ** 2).sum()
    lt.backward()
    sd = {k: v.detach().numpy() for k, v in flow.state_dict().items()}
    np.savez_compressed(os.path.join(OUT, 'affine.npz'), K=K, x=x, y=y, z=zs[-1].detach().numpy(), z_mid=zs[1].detach().numpy(),
                        logdet=ld.detach().numpy(), loss=float(loss.detach()), gx=xt.grad.numpy(),
                        g_s=aff.s.grad.numpy(), g_t=aff.t.grad.numpy(), g_w=c0.s.layers[0].weight.grad.numpy(),
                        x_rec=xs[-1].numpy(), logdet_inv=ldi.numpy(), temp_z=zt.detach().numpy(),
                        temp_gT=ts.T.grad.numpy(), **{'sd_' + k: v for k, v in sd.items()})
    print('affine case loss', float(loss.detach()))


def gen_planar_radial_case():
    """PlanarLayer / RadialLayer (flows/flows.py:129-193): forward outputs and autograd gradients of a
    seeded linear functional of (z, log_det), and both inside a Flow with a coupling layer."""
    from flows.flows import PlanarLayer, RadialLayer
    out = {}
    for K in (3, 10, 40):
        torch.manual_seed(100 + K)
        N = 77
        x, _ = orc.synth_logits(N, K, seed=20 + K)
        cz = torch.randn(N, K)
        cl = torch.randn(N)
        pl = PlanarLayer(K)
        xt = torch.as_tensor(x).requires_grad_(True)
        z, ld = pl(xt)
        ((z * cz).sum() + (ld * cl).sum()).backward()
        tag = 'planar_k%d_' % K
        out.update({tag + 'x': x, tag + 'cz': cz.numpy(), tag + 'cl': cl.numpy(), tag + 'w': pl.w.detach().numpy(),
                    tag + 'u': pl.u.detach().numpy(), tag + 'b': pl.b.detach().numpy(), tag + 'z': z.detach().numpy(),
                    tag + 'ld': ld.detach().numpy(), tag + 'gx': xt.grad.numpy(), tag + 'gw': pl.w.grad.numpy(),
                    tag + 'gu': pl.u.grad.numpy(), tag + 'gb': pl.b.grad.numpy()})
        rl = RadialLayer(K)
        xt = torch.as_tensor(x).requires_grad_(True)
        z, ld = rl(xt)
        (z * cz).sum().backward()
        tag = 'radial_k%d_' % K
        out.update({tag + 'x': x, tag + 'cz': cz.numpy(), tag + 'z0': rl.z0.detach().numpy(),
                    tag + 'a': rl.a.detach().numpy(), tag + 'b': rl.b.detach().numpy(), tag + 'z': z.detach().numpy(),
                    tag + 'ld': ld.detach().numpy(), tag + 'gx': xt.grad.numpy(), tag + 'gz0': rl.z0.grad.numpy(),
                    tag + 'ga': rl.a.grad.numpy(), tag + 'gb': rl.b.grad.numpy()})
    # both inside a Flow, as the notebooks stack arbitrary layers
    torch.manual_seed(7)
    K, N = 6, 33
    pl, rl, cp = PlanarLayer(K), RadialLayer(K), NvpCouplingLayer(K, hidden_size=[8])
    with torch.no_grad():
        for p in cp.parameters():
            if p.requires_grad:
                p.mul_(300.0)
    flow = Flow([pl, cp, rl])
    x, _ = orc.synth_logits(N, K, seed=3)
    with torch.no_grad():
        zs, ld = flow(torch.as_tensor(x))
    sd = {k: v.detach().numpy() for k, v in flow.state_dict().items()}
    out.update({'flow_x': x, 'flow_z': zs[-1].numpy(), 'flow_ld': ld.numpy(), 'flow_K': K})
    out.update({'flow_sd_' + k: v for k, v in sd.items()})
    np.savez_compressed(os.path.join(OUT, 'planar_radial.npz'), **out)
    print('planar/radial cases written')


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)          # bit-stable reductions
    cases = {'flow': gen_flow_cases, 'calibrator': gen_calibrator_case, 'metrics': gen_metrics_cases,
             'affine': gen_affine_case, 'planar_radial': gen_planar_radial_case}
    for name in (sys.argv[1:] or list(cases)):
        cases[name]()
