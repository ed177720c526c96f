"""Drop-in for the flow calibrator of the reference's ``calibrators.py``: the ``Calibrator``
base class (:13-44) and ``TorchFlowCalibrator`` (:239-353), same constructor signature,
methods and attributes (numpy in, numpy out).

What runs where:
  * ``fit``: samples stay resident in HBM; every optimisation step is the fused
    forward + NLL head + backward kernel, a gradient reduce, (with torch.distributed
    initialised: one NCCL all-reduce of the flat gradient and of the four loss sums) and the
    fused Adam kernel.  No DataLoader, no autograd graph, no host synchronisation inside the
    epoch loop.  The loss is the reference's ``-mean(log(softmax(z)[y] + 1e-7) + log_det)``
    (calibrators.py:288-291), Adam with torch defaults (:259).
  * ``predict``: centre -> H2D -> fused flow forward -> fused ``softmax(log(softmax(z)+1e-7)
    - log_priors)`` kernel in float64 -> D2H (calibrators.py:40-44, 330-353).
Reference quirks kept: the per-epoch history takes the statistics of the LAST evaluation batch
times its length over N (``_loss =`` overwrites, calibrators.py:309-317) -- exact only for the
default full batch; history entries are 0-d torch tensors; the class handed in as ``Flow`` is
called as ``Flow(n_classes, **kwargs)`` with the calibrator's own kwargs included (:251).
"""
import ctypes

import numpy as np
import torch
from scipy.special import softmax

from . import _lib
from ._engine import _ptr, _stream
from .utils.ops import onehot_encode


class Calibrator:
    """Base class: centres the logits, one-hot encodes the targets, estimates log priors."""

    def __init__(self, logits, target):
        logits = np.asarray(logits)
        self.logits = logits - logits.mean(axis=1, keepdims=True)
        target = np.asarray(target)
        self.target = target if target.shape == logits.shape else onehot_encode(target)
        self.n_classes = self.target.shape[1]
        self.log_priors = self._get_log_priors(self.target)

    def __call__(self, logits):
        return self.predict(logits)

    def _get_log_priors(self, target):
        counts = np.sum(target, axis=0)
        return np.log(counts / np.sum(counts))

    def predict_post(self, logits):
        raise NotImplementedError

    def predict(self, logits):
        logits = np.asarray(logits)
        probs = self.predict_post(logits - logits.mean(axis=1, keepdims=True))
        return softmax(np.log(probs + 1e-7) - self.log_priors, axis=1)


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def shard_bounds(n, rank, world):
    """Contiguous row block [lo, hi) of rank `rank` when n samples are split over `world` ranks."""
    return (n * rank) // world, (n * (rank + 1)) // world


class FusedNLLTrainer:
    """The calibrator's optimisation loop on device buffers.  Data-parallel when
    torch.distributed is initialised: every rank holds a contiguous shard of the samples, the
    flat gradient (and the loss sums) are all-reduced once per step, and every rank applies the
    same optimiser update, so parameters stay bit-identical across ranks without a broadcast."""

    def __init__(self, engine, x, y, n_total=None, eps=1e-7, gamma=1.0, lr=1e-3, betas=(0.9, 0.999),
                 adam_eps=1e-8, weight_decay=0.0, optim='adam', precision='fp32'):
        self.engine, self.x, self.y = engine, x, y
        self.precision = precision
        self.n_local = x.shape[0]
        self.dist = _dist()
        if n_total is None:
            n_total = self.n_local
            if self.dist is not None:
                t = torch.tensor([self.n_local], dtype=torch.int64, device=x.device)
                self.dist.all_reduce(t)
                n_total = int(t.item())
        self.n_total = n_total
        self.eps, self.gamma = eps, gamma
        self.lr, self.betas, self.adam_eps, self.wd, self.optim = lr, betas, adam_eps, weight_decay, optim
        engine.ensure(x.device)
        engine.pack(tc=(precision == 'bf16'), fp32=(precision != 'bf16'))
        self.loss_acc = torch.zeros(4, dtype=torch.float64, device=x.device)

    def step(self, xb=None, yb=None, n_batch_total=None):
        """One optimiser step on (xb, yb) (default: all local samples).  Returns nothing;
        the summed loss statistics of the batch are left in self.loss_acc (device)."""
        e = self.engine
        xb = self.x if xb is None else xb
        yb = self.y if yb is None else yb
        n_tot = self.n_total if n_batch_total is None else n_batch_total
        self.loss_acc.zero_()
        e.nll_step(xb, yb, self.loss_acc, self.eps, self.gamma, n_tot, with_grad=True, precision=self.precision)
        if self.dist is not None:
            self.dist.all_reduce(e.flat_grad)
            self.dist.all_reduce(self.loss_acc)
        if self.optim == 'adam':
            e.adam(self.lr, self.betas, self.adam_eps, self.wd)
        else:
            e.sgd(self.lr, self.wd)
        e.pack(tc=(self.precision == 'bf16'), fp32=(self.precision != 'bf16'))

    # ------------------------------------------------------------------ CUDA-graph epoch
    def _epoch_body(self):
        e = self.engine
        self.loss_acc.zero_()
        e.nll_step(self.x, self.y, self.loss_acc, self.eps, self.gamma, self.n_total, with_grad=True,
                   precision=self.precision)
        if self.optim == 'adam':
            e.adam_dev(self.lr, self.betas, self.adam_eps, self.wd)
        else:
            e.sgd(self.lr, self.wd)
        e.pack(tc=(self.precision == 'bf16'), fp32=(self.precision != 'bf16'))
        self.eval_acc.zero_()
        e.nll_step(self.x, self.y, self.eval_acc, self.eps, self.gamma, self.n_total, with_grad=False,
                   precision=self.precision)

    def epoch_graph(self):
        """One full-batch epoch of ``fit`` -- optimiser step, then the evaluation pass into ``self.eval_acc``
        -- as ONE CUDA-graph launch.  The first call runs the body eagerly (allocating every buffer) and
        captures it; later calls replay (~10 launches per epoch become one).  Single-process only; with torch.distributed the eager ``step`` / ``evaluate`` pair is used."""
        if self.dist is not None:
            raise RuntimeError('epoch_graph is single-process; use step() / evaluate() under torch.distributed')
        if getattr(self, '_graph', None) is None:
            self.eval_acc = torch.zeros(4, dtype=torch.float64, device=self.x.device)
            self._epoch_body()                                   # epoch 0: eager, for real
            torch.cuda.current_stream(self.x.device).synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._epoch_body()                               # recorded, not executed
            self._graph = g
            return
        self._graph.replay()

    def _step_body(self):
        e = self.engine
        self.loss_acc.zero_()
        e.nll_step(self.x, self.y, self.loss_acc, self.eps, self.gamma, self.n_total, with_grad=True,
                   precision=self.precision)
        if self.optim == 'adam':
            e.adam_dev(self.lr, self.betas, self.adam_eps, self.wd)
        else:
            e.sgd(self.lr, self.wd)
        e.pack(tc=(self.precision == 'bf16'), fp32=(self.precision != 'bf16'))

    def step_graph(self):
        """``step()`` on all local samples as ONE CUDA-graph launch (no evaluation pass: ``fit`` takes an epoch's
        evaluation from the next step's forward, see there).  First call: eager + capture; later calls replay.
        Single-process only."""
        if self.dist is not None:
            raise RuntimeError('step_graph is single-process; use step() under torch.distributed')
        if getattr(self, '_sgraph', None) is None:
            self._step_body()                                    # step 0: eager, for real
            torch.cuda.current_stream(self.x.device).synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step_body()                                # recorded, not executed
            self._sgraph = g
            return
        self._sgraph.replay()

    def evaluate(self, xb=None, yb=None, out=None):
        """Loss statistics (sum(ce+gamma*ld), sum ce, sum ld, #non-finite) of a batch, summed
        over ranks, as a float64 device tensor."""
        e = self.engine
        xb = self.x if xb is None else xb
        yb = self.y if yb is None else yb
        acc = torch.zeros(4, dtype=torch.float64, device=xb.device) if out is None else out.zero_()
        e.nll_step(xb, yb, acc, self.eps, self.gamma, self.n_total, with_grad=False, precision=self.precision)
        if self.dist is not None:
            self.dist.all_reduce(acc)
        return acc


class TorchFlowCalibrator(Calibrator):

    def __init__(self, Flow, logits, target, **kwargs):
        super().__init__(logits, target)
        self.target = np.argmax(self.target, axis=1)
        self.logits = torch.as_tensor(self.logits, dtype=torch.float)
        self.target = torch.as_tensor(self.target, dtype=torch.long)

        self.flow = Flow(self.n_classes, **kwargs)

        self.dev = kwargs.get('dev', torch.device('cuda') if torch.cuda.is_available() else torch.device('cpu'))
        self.dev = torch.device(self.dev)
        if self.dev.type != 'cuda':
            raise RuntimeError('cnf_b200.TorchFlowCalibrator needs a CUDA device (dev=%s); the flow kernels '
                               'have no CPU fallback' % self.dev)
        if self.dev.index is None:
            self.dev = torch.device('cuda', torch.cuda.current_device())

        # extension: precision='bf16' trains on the tcgen05 kernels (stated bf16 tolerance); default fp32
        self.precision = kwargs.get('precision', 'fp32')
        # extension: cuda_graph=True replays each full-batch optimiser step as one captured CUDA graph.  Off by
        # default: the steady state is only ~5-10 % faster than the eager launches (K=10, N=5,000, bf16: 93 vs
        # 104 us per epoch including the capture; K=3, N=10,000: 59 vs 61 us), and the first capture in a
        # process costs seconds (profiles/microbench/c1_fit_speed.py, c2_fit_speed.py).
        self.cuda_graph = kwargs.get('cuda_graph', False)
        self.CE = torch.nn.CrossEntropyLoss()
        self.optimizer = torch.optim.Adam(self.flow.parameters())
        self.history = self.fit(self.logits, self.target,
                                epochs=kwargs.get('epochs', 1000),
                                batch_size=kwargs.get('batch_size', np.asarray(logits).shape[0]))

    # ------------------------------------------------------------------ training
    def _engine(self):
        eng = self.flow.engine() if hasattr(self.flow, 'engine') else None
        if eng is None:
            raise NotImplementedError(
                'cnf_b200.TorchFlowCalibrator trains flows that are one homogeneous stack of coupling layers '
                '(NiceFlow / RealNvpFlow / CouplingStack / Flow([NvpCouplingLayer...])); got %r' % type(self.flow))
        return eng

    def fit(self, logits, target, epochs, batch_size):
        dist = _dist()
        n_all = logits.shape[0]
        if dist is not None:
            # contiguous row blocks per rank; every rank was given the full arrays
            rank, world = dist.get_rank(), dist.get_world_size()
            lo, hi = shard_bounds(n_all, rank, world)
            logits, target = logits[lo:hi], target[lo:hi]
        x = logits.to(self.dev).contiguous()
        y = target.to(self.dev).contiguous()
        self.flow.to(self.dev)
        eng = self._engine()
        group = self.optimizer.param_groups[0]
        trainer = FusedNLLTrainer(eng, x, y, n_total=n_all, eps=1e-7, gamma=1.0, lr=group['lr'],
                                  betas=group['betas'], adam_eps=group['eps'], weight_decay=group['weight_decay'],
                                  precision=getattr(self, 'precision', 'fp32'))
        self.trainer = trainer
        n_local = x.shape[0]
        world = dist.get_world_size() if dist is not None else 1
        local_bs = max(1, min(n_local, int(batch_size) // world if dist is not None else int(batch_size)))
        full_batch = local_bs >= n_local
        hist = torch.zeros((max(epochs, 0), 4), dtype=torch.float64, device=self.dev)
        gen = torch.Generator(device=self.dev)
        gen.manual_seed(int(torch.initial_seed()) & 0x7fffffff)
        use_graph = full_batch and dist is None and bool(getattr(self, 'cuda_graph', False))
        for epoch in range(epochs):
            if use_graph:
                # the same scheme as the eager branch below, with the step replayed as one graph launch
                trainer.step_graph()
                if epoch > 0:
                    hist[epoch - 1].copy_(trainer.loss_acc)
                if epoch == epochs - 1:
                    trainer.evaluate(out=hist[epoch])
            elif full_batch:
                # The reference evaluates the whole set after every update (calibrators.py:297-317).  With the
                # full batch and no shuffling effect that number IS the loss the next step's forward computes
                # on the same weights and samples, so it is taken from there; only the last epoch needs its
                # own evaluation pass.  Same kernels, same arithmetic: the history is unchanged.
                trainer.step()
                if epoch > 0:
                    hist[epoch - 1].copy_(trainer.loss_acc)
                if epoch == epochs - 1:
                    trainer.evaluate(out=hist[epoch])
            else:
                perm = torch.randperm(n_local, device=self.dev, generator=gen)
                for s in range(0, n_local, local_bs):
                    idx = perm[s:s + local_bs]
                    nb = idx.numel() * world
                    trainer.step(x.index_select(0, idx), y.index_select(0, idx), n_batch_total=nb)
                # reference quirk: only the last evaluation batch survives (calibrators.py:309-317)
                perm = torch.randperm(n_local, device=self.dev, generator=gen)
                last = perm[(n_local - 1) // local_bs * local_bs:]
                trainer.evaluate(x.index_select(0, last), y.index_select(0, last), out=hist[epoch])
        hist = hist / float(n_all)
        history = {
            'loss': [v for v in (-hist[:, 0]).to(torch.float32).unbind(0)],
            'ce': [v for v in (-hist[:, 1]).to(torch.float32).unbind(0)],
            'log_det': [v for v in hist[:, 2].to(torch.float32).unbind(0)],
        }
        self.nonfinite = hist[:, 3] * float(n_all)
        self.flow.cpu()
        torch.cuda.empty_cache()
        return history

    # ------------------------------------------------------------------ inference
    def _forward_device(self, logits):
        x = torch.as_tensor(np.ascontiguousarray(logits), dtype=torch.float)
        self.flow.to(self.dev)
        x = x.to(self.dev)
        with torch.no_grad():
            preds, _ = self.flow(x)
        return preds

    def predict_logits(self, logits):
        preds = self._forward_device(logits)
        out = preds.cpu().detach().numpy()
        self.flow.cpu()
        torch.cuda.empty_cache()
        return out

    def predict_post(self, logits):
        return softmax(self.predict_logits(logits), axis=1)

    def predict(self, logits):
        """Same result as the base-class ``predict`` (calibrators.py:40-44) with the whole tail
        fused on the device; returns float64 probabilities."""
        logits = np.asarray(logits)
        z = self._forward_device(logits - logits.mean(axis=1, keepdims=True)).contiguous()
        n, k = z.shape
        lp = torch.as_tensor(np.asarray(self.log_priors, dtype=np.float64)).to(self.dev)
        out = torch.empty((n, k), dtype=torch.float64, device=self.dev)
        _lib.call('cnf_calibrated_probs', _ptr(z), ctypes.c_int64(n), ctypes.c_int32(k), _ptr(lp), _ptr(out),
                  _stream(self.dev))
        res = out.cpu().numpy()
        self.flow.cpu()
        return res
