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
from ._engine import _ptr, _stream, on_device, check_logits, check_labels
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


def plan_fit(n_all, world, batch_size):
    """Rank-invariant schedule of one ``fit`` epoch over `n_all` samples sharded over `world` ranks:
    (full_batch, local_bs, steps_per_epoch, batch_totals).  Depends on (n_all, world, batch_size) only --
    never on a rank's own shard size -- so every rank issues the same sequence of collectives even when
    n_all % world != 0 (shards then differ by one row; the shorter ones contribute an empty slice to the
    last step).  batch_totals[s] = true number of samples all ranks together feed into step s (the
    gradient of step s is the mean over exactly those)."""
    world = max(1, int(world))
    shards = [shard_bounds(n_all, r, world) for r in range(world)]
    max_shard = max(hi - lo for lo, hi in shards)
    bs = max(1, int(batch_size))
    full_batch = bs >= n_all
    local_bs = max(1, bs // world)
    steps = 1 if full_batch else max(1, -(-max_shard // local_bs))
    if full_batch:
        totals = [n_all]
    else:
        totals = [sum(min(max((hi - lo) - s * local_bs, 0), local_bs) for lo, hi in shards) for s in range(steps)]
    return full_batch, local_bs, steps, totals


class FusedNLLTrainer:
    """The calibrator's optimisation loop on device buffers.  Data-parallel when
    torch.distributed is initialised: every rank holds a contiguous shard of the samples; per step the
    flat gradient and the four loss sums travel in ONE all-reduce (a float64 buffer [n_flat + 4]: the
    loss sums keep their precision and the gradient sum is rounded once), and every rank applies the
    same optimiser update.  The initial parameters, optimiser state and random_flip permutations are
    taken from rank 0 at construction (each rank builds its flow from its own RNG state), so
    parameters are bit-identical across ranks from the first step on."""

    def __init__(self, engine, x, y, n_total=None, eps=1e-7, gamma=1.0, lr=1e-3, betas=(0.9, 0.999),
                 adam_eps=1e-8, weight_decay=0.0, optim='adam', precision='fp32'):
        if x.is_cuda:                # (host-logic tests inject a CPU engine; the product engine is CUDA-only)
            x = check_logits(x, engine.K, cast=False)
            y = check_labels(y, x.shape[0], x.device)
        elif x.dim() != 2 or x.shape[1] != engine.K or y.shape[0] != x.shape[0]:
            raise ValueError('cnf_b200: x must be [N, %d] and y [N], got %s / %s' % (engine.K, tuple(x.shape), tuple(y.shape)))
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
        if self.dist is not None:
            self._sync_from_rank0()
        engine.pack(tc=(precision == 'bf16'), fp32=(precision != 'bf16'))
        self.loss_acc = torch.zeros(4, dtype=torch.float64, device=x.device)
        self._comm = None

    def _sync_from_rank0(self):
        """Every rank must start from the same model: compare the index tables (random_flip permutations are
        structure, they cannot be patched in) and broadcast the parameters and Adam state from rank 0."""
        e, d = self.engine, self.dist
        tab = e.tables.to(torch.int64).clone()
        ref = tab.clone()
        d.broadcast(ref, src=0)
        same = torch.tensor([1 if (ref.numel() == tab.numel() and bool((ref == tab).all())) else 0],
                            dtype=torch.int64, device=tab.device)
        d.all_reduce(same, op=d.ReduceOp.MIN)
        if int(same.item()) != 1:
            raise RuntimeError('cnf_b200: the flow\'s random_flip permutations differ between ranks; seed numpy '
                               'identically on every rank before building the flow (np.random.seed)')
        d.broadcast(e.flat, src=0)
        has_state = torch.tensor([0 if e.adam_m is None else 1], dtype=torch.int64, device=e.flat.device)
        d.broadcast(has_state, src=0)
        if int(has_state.item()):
            if e.adam_m is None:
                e.adam_m, e.adam_v = torch.zeros_like(e.flat), torch.zeros_like(e.flat)
            t = torch.tensor([e.adam_t], dtype=torch.int64, device=e.flat.device)
            d.broadcast(e.adam_m, src=0)
            d.broadcast(e.adam_v, src=0)
            d.broadcast(t, src=0)
            e.adam_t = int(t.item())

    def _all_reduce_step(self, loss_acc):
        """Gradient + loss sums in one collective."""
        e = self.engine
        n = e.n_flat
        if self._comm is None or self._comm.numel() != n + 4 or self._comm.device != e.flat_grad.device:
            self._comm = torch.empty(n + 4, dtype=torch.float64, device=e.flat_grad.device)
        c = self._comm
        c[:n].copy_(e.flat_grad)
        c[n:].copy_(loss_acc)
        self.dist.all_reduce(c)
        e.flat_grad.copy_(c[:n])
        loss_acc.copy_(c[n:])

    def _optim_and_pack(self, dev_counter=False):
        e = self.engine
        if self.optim == 'adam':
            (e.adam_dev if dev_counter else e.adam)(self.lr, self.betas, self.adam_eps, self.wd)
        else:
            e.sgd(self.lr, self.wd)
        e.pack(tc=(self.precision == 'bf16'), fp32=(self.precision != 'bf16'))

    def step(self, xb=None, yb=None, n_batch_total=None, acc=None):
        """One optimiser step on (xb, yb) (default: all local samples; an empty batch is a valid
        contribution of zero under torch.distributed).  Returns nothing; the summed loss statistics of the
        batch are left in self.loss_acc (device) -- or ADDED to `acc` (float64 [4], device, zeroed by the caller)
        when one is given: ``fit`` passes the history row the sums belong to, which saves a clear and a copy
        launch per epoch."""
        e = self.engine
        xb = self.x if xb is None else xb
        yb = self.y if yb is None else yb
        n_tot = self.n_total if n_batch_total is None else n_batch_total
        nvtx = torch.cuda.nvtx
        loss_acc = self.loss_acc if acc is None else acc
        if acc is None:
            loss_acc.zero_()
        # single process, Adam without weight decay (the reference's defaults, calibrators.py:259): the partial
        # reduction, the update and the repack are ONE launch behind the training kernel
        fused_tail = (self.dist is None and self.optim == 'adam' and self.wd == 0.0 and xb.shape[0] > 0
                      and getattr(e, 'adam_step_dev', None) is None and hasattr(e, 'reduce_adam_pack'))
        if fused_tail:
            fused_tail = (e.tc_tail_maps() is not None) if self.precision == 'bf16' else getattr(e, 'gather_one_to_one', False)
        nvtx.range_push('cnf.fwd_bwd')
        if fused_tail:
            e.nll_step(xb, yb, loss_acc, self.eps, self.gamma, n_tot, with_grad=True, precision=self.precision,
                       reduce=False)
            nvtx.range_pop()
            nvtx.range_push('cnf.optim')
            if self.precision == 'bf16':
                e.reduce_adam_pack_tc(self.lr, self.betas, self.adam_eps)
            else:
                e.reduce_adam_pack(self.lr, self.betas, self.adam_eps)
            nvtx.range_pop()
            return
        e.nll_step(xb, yb, loss_acc, self.eps, self.gamma, n_tot, with_grad=True, precision=self.precision)
        nvtx.range_pop()
        if self.dist is not None:
            nvtx.range_push('cnf.all_reduce')
            self._all_reduce_step(loss_acc)
            nvtx.range_pop()
        nvtx.range_push('cnf.optim')
        self._optim_and_pack()
        nvtx.range_pop()

    # ------------------------------------------------------------------ CUDA-graph epoch
    def _epoch_body(self):
        e = self.engine
        self.loss_acc.zero_()
        e.nll_step(self.x, self.y, self.loss_acc, self.eps, self.gamma, self.n_total, with_grad=True,
                   precision=self.precision)
        if self.dist is not None:
            self._all_reduce_step(self.loss_acc)
        self._optim_and_pack(dev_counter=True)
        self.eval_acc.zero_()
        e.nll_step(self.x, self.y, self.eval_acc, self.eps, self.gamma, self.n_total, with_grad=False,
                   precision=self.precision)
        if self.dist is not None:
            self.dist.all_reduce(self.eval_acc)

    def epoch_graph(self):
        """One full-batch epoch of ``fit`` -- optimiser step, then the evaluation pass into ``self.eval_acc``
        -- as ONE CUDA-graph launch.  The first call runs the body eagerly (allocating every buffer) and
        captures it; later calls replay (~10 launches per epoch become one).  Single-process only: capturing the
        NCCL all-reduce inside the graph hung on this stack (torch 2.11 / NCCL 2.28.9, 2 x B200; NOTES.md), so
        under torch.distributed the eager ``step`` / ``evaluate`` pair is used."""
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
        self.engine.version += 1         # the replayed graph updated the weights

    def _step_body(self):
        e = self.engine
        self.loss_acc.zero_()
        e.nll_step(self.x, self.y, self.loss_acc, self.eps, self.gamma, self.n_total, with_grad=True,
                   precision=self.precision)
        if self.dist is not None:
            self._all_reduce_step(self.loss_acc)
        self._optim_and_pack(dev_counter=True)

    def step_graph(self):
        """``step()`` on all local samples as ONE CUDA-graph launch (no evaluation pass: ``fit`` takes an epoch's
        evaluation from the next step's forward, see there).  First call: eager + capture; later calls replay.
        Single-process only (see epoch_graph)."""
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
        self.engine.version += 1         # the replayed graph updated the weights

    def fit_loop(self, epochs, batch_size, perm_fn, cuda_graph=False):
        """The epoch loop of ``TorchFlowCalibrator.fit`` (calibrators.py:283-317) on the resident local
        samples.  Returns the per-epoch loss sums [epochs, 4] (float64, summed over ranks; divide by n_total).
        `perm_fn(n_local)` yields the shuffled order of the local samples for one pass."""
        x, y = self.x, self.y
        n_local, n_all = self.n_local, self.n_total
        world = self.dist.get_world_size() if self.dist is not None else 1
        # Every decision that shapes the sequence of collectives comes from (n_all, world, batch_size) only
        # (plan_fit), never from the rank-local shard size.
        full_batch, local_bs, steps_per_epoch, batch_totals = plan_fit(n_all, world, batch_size)
        use_graph = full_batch and cuda_graph and self.dist is None
        e = self.engine
        # single process, full batch, fp32, Adam without weight decay (the reference's defaults, calibrators.py:259,
        # 263-265): the whole epoch loop is enqueued by one library call -- the same two launches per epoch as step(),
        # without the interpreter between them (K = 3, N = 1,500 notebook fits: 39 -> ~25 us per epoch)
        if (full_batch and not use_graph and self.dist is None and epochs > 0 and n_local > 0 and self.precision != 'bf16'
                and self.optim == 'adam' and self.wd == 0.0 and type(self).step is FusedNLLTrainer.step
                and getattr(e, 'adam_step_dev', None) is None and getattr(e, 'gather_one_to_one', False)
                and hasattr(e, 'fit_full_batch')):
            return e.fit_full_batch(x, y, epochs, self.eps, self.gamma, self.n_total, self.lr, self.betas, self.adam_eps)
        hist = torch.zeros((max(epochs, 0), 4), dtype=torch.float64, device=x.device)
        for epoch in range(epochs):
            if full_batch:
                # The reference evaluates the whole set after every update (calibrators.py:297-317).  With the
                # full batch and no shuffling effect that number IS the loss the next step's forward computes
                # on the same weights and samples, so it is taken from there; only the last epoch needs its
                # own evaluation pass.  Same kernels, same arithmetic: the history is unchanged.
                if use_graph:
                    self.step_graph()       # the same step replayed as one captured graph launch
                    if epoch > 0:
                        hist[epoch - 1].copy_(self.loss_acc)
                else:                       # the step's forward adds its loss sums straight into the history row
                    self.step(acc=hist[epoch - 1] if epoch > 0 else None)
                if epoch == epochs - 1:
                    self.evaluate(out=hist[epoch])
            else:
                perm = perm_fn(n_local)
                for s in range(steps_per_epoch):
                    idx = perm[s * local_bs:(s + 1) * local_bs]        # may be empty on the shorter shards
                    self.step(x.index_select(0, idx), y.index_select(0, idx), n_batch_total=batch_totals[s])
                # reference quirk: only the last evaluation batch survives (calibrators.py:309-317)
                perm = perm_fn(n_local)
                last = perm[(steps_per_epoch - 1) * local_bs:steps_per_epoch * local_bs]
                self.evaluate(x.index_select(0, last), y.index_select(0, last), out=hist[epoch])
        return hist

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


class HostStreamNLLTrainer(FusedNLLTrainer):
    """The same optimisation loop for samples that live in (pinned) HOST memory -- the on-disk logit sets of
    utils/data.py (reference utils/data.py:170-210) loaded by ``cnf_b200.utils.data`` -- instead of HBM.

    * ``resident=False`` (data larger than the HBM budget): every full-batch step streams the samples through two
      device staging buffers of ``chunk_rows`` rows; the H2D copy of chunk c+1 (copy stream) overlaps the fused
      forward+backward kernel of chunk c, per-chunk gradients are summed on the device (``inv_n_total`` makes the
      sum the full-batch gradient), then one all-reduce / optimiser update as usual.
    * ``resident=True`` (default when the shard fits): the FIRST step streams the chunks straight into their final
      place in HBM and computes on each chunk as it lands, so the load is hidden behind the first epoch; later steps
      run on the resident copy in one launch, exactly as ``FusedNLLTrainer``.
    Full-batch only (the reference's default ``batch_size = N``, calibrators.py:263-265)."""

    def __init__(self, engine, x_host, y_host, device, n_total=None, chunk_rows=1 << 22, resident=None, **kw):
        device = torch.device(device)
        if device.index is None:
            device = torch.device('cuda', torch.cuda.current_device())
        x_host = torch.as_tensor(x_host)
        y_host = torch.as_tensor(y_host)
        if x_host.is_cuda or y_host.is_cuda:
            raise ValueError('cnf_b200: HostStreamNLLTrainer expects host tensors')
        x_host = check_logits(x_host, engine.K, 'x_host', cast=True)
        if y_host.dim() != 1 or y_host.shape[0] != x_host.shape[0] or y_host.dtype.is_floating_point:
            raise ValueError('cnf_b200: y_host must be an integer tensor of shape [%d]' % x_host.shape[0])
        y_host = y_host.to(torch.int64).contiguous()
        if not x_host.is_pinned():
            x_host = x_host.pin_memory()
        if not y_host.is_pinned():
            y_host = y_host.pin_memory()
        self.x_host, self.y_host = x_host, y_host
        n, K = x_host.shape
        self.chunk_rows = max(1024, int(chunk_rows))
        if resident is None:
            free, _ = torch.cuda.mem_get_info(device)
            resident = n * (4 * K + 8) < 0.6 * free
        self.resident = bool(resident)
        self._loaded = False
        rows = n if self.resident else min(n, 2 * self.chunk_rows)
        xbuf = torch.empty((rows, K), dtype=torch.float32, device=device)
        ybuf = torch.empty(rows, dtype=torch.int64, device=device)
        if n_total is None:          # the size of the whole set: the sum of the ranks' host shards
            n_total = n
            d = _dist()
            if d is not None:
                t = torch.tensor([n], dtype=torch.int64, device=device)
                d.all_reduce(t)
                n_total = int(t.item())
        super().__init__(engine, xbuf, ybuf, n_total=n_total, **kw)
        self.n_local = n
        self._copy_stream = torch.cuda.Stream(device=device)
        self._copied = [torch.cuda.Event(), torch.cuda.Event()]
        self._freed = [torch.cuda.Event(), torch.cuda.Event()]
        self._grad_acc = None

    def _chunks(self):
        n = self.n_local
        return [(lo, min(n, lo + self.chunk_rows)) for lo in range(0, n, self.chunk_rows)] or [(0, 0)]

    def _streamed_pass(self, with_grad, acc):
        """One pass over the host samples: H2D of the next chunk overlaps the kernel of the current one."""
        e = self.engine
        dev = self.x.device
        cur = torch.cuda.current_stream(dev)
        chunks = self._chunks()
        if with_grad:
            if self._grad_acc is None:
                self._grad_acc = torch.zeros(e.n_flat, dtype=torch.float32, device=dev)
            self._grad_acc.zero_()

        def place(c):      # device rows that receive chunk c
            lo, hi = chunks[c]
            if self.resident:
                return self.x[lo:hi], self.y[lo:hi]
            base = (c % 2) * self.chunk_rows
            return self.x[base:base + (hi - lo)], self.y[base:base + (hi - lo)]

        def issue_copy(c):
            lo, hi = chunks[c]
            xd, yd = place(c)
            with torch.cuda.stream(self._copy_stream):
                if not self.resident and c >= 2:
                    self._copy_stream.wait_event(self._freed[c % 2])     # the kernel that last read this staging half
                xd.copy_(self.x_host[lo:hi], non_blocking=True)
                yd.copy_(self.y_host[lo:hi], non_blocking=True)
                self._copied[c % 2].record(self._copy_stream)

        self._copy_stream.wait_stream(cur)
        issue_copy(0)
        for c in range(len(chunks)):
            cur.wait_event(self._copied[c % 2])
            if c + 1 < len(chunks):
                issue_copy(c + 1)
            xd, yd = place(c)
            e.nll_step(xd, yd, acc, self.eps, self.gamma, self.n_total, with_grad=with_grad, precision=self.precision)
            if with_grad:
                self._grad_acc.add_(e.flat_grad)
            self._freed[c % 2].record(cur)
        if with_grad:
            e.flat_grad.copy_(self._grad_acc)
        if self.resident:
            self._loaded = True

    def step(self, xb=None, yb=None, n_batch_total=None, acc=None):
        if xb is not None or yb is not None:
            raise NotImplementedError('HostStreamNLLTrainer is full-batch (mini-batches need the samples resident)')
        if self.resident and self._loaded:
            return super().step(acc=acc)
        nvtx = torch.cuda.nvtx
        loss_acc = self.loss_acc if acc is None else acc
        if acc is None:
            loss_acc.zero_()
        nvtx.range_push('cnf.fwd_bwd(streamed)')
        self._streamed_pass(True, loss_acc)
        nvtx.range_pop()
        if self.dist is not None:
            self._all_reduce_step(loss_acc)
        self._optim_and_pack()

    def step_graph(self):
        raise RuntimeError('CUDA-graph steps need resident samples; use FusedNLLTrainer')

    epoch_graph = step_graph

    def evaluate(self, xb=None, yb=None, out=None):
        if xb is not None or yb is not None:
            raise NotImplementedError('HostStreamNLLTrainer evaluates the whole local shard')
        if self.resident and self._loaded:
            return super().evaluate(out=out)
        acc = torch.zeros(4, dtype=torch.float64, device=self.x.device) if out is None else out.zero_()
        self._streamed_pass(False, acc)
        if self.dist is not None:
            self.dist.all_reduce(acc)
        return acc

    def fit_loop(self, epochs, batch_size, perm_fn, cuda_graph=False):
        if int(batch_size) < self.n_total:
            raise NotImplementedError('HostStreamNLLTrainer trains full-batch (batch_size >= N)')
        return super().fit_loop(epochs, batch_size, perm_fn, cuda_graph=False)


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
        self.host_stream = kwargs.get('host_stream', 'auto')
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
        self.flow.to(self.dev)
        eng = self._engine()
        group = self.optimizer.param_groups[0]
        opt = dict(n_total=n_all, eps=1e-7, gamma=1.0, lr=group['lr'], betas=group['betas'], adam_eps=group['eps'],
                   weight_decay=group['weight_decay'], precision=getattr(self, 'precision', 'fp32'))
        # extension: host_stream=True keeps the samples in pinned host memory and streams them through the GPU
        # (first epoch overlapped with the load when they fit HBM, every epoch when they do not); 'auto' switches to
        # it when the local shard would not fit comfortably.  Full-batch only.
        hs = getattr(self, 'host_stream', 'auto')
        n_loc = logits.shape[0]
        if hs == 'auto':
            free, _ = torch.cuda.mem_get_info(self.dev)
            hs = int(batch_size) >= n_all and n_loc * (4 * self.n_classes + 8) > 0.5 * free
        if hs:
            trainer = HostStreamNLLTrainer(eng, logits, target, self.dev, **opt)
        else:
            x = logits.to(self.dev).contiguous()
            y = target.to(self.dev).contiguous()
            trainer = FusedNLLTrainer(eng, x, y, **opt)
        self.trainer = trainer
        gen = torch.Generator(device=self.dev)
        gen.manual_seed(int(torch.initial_seed()) & 0x7fffffff)
        gen_holder = [gen]
        hist = trainer.fit_loop(epochs, batch_size, lambda n: self._epoch_permutation(n, gen_holder[0]),
                                cuda_graph=bool(getattr(self, 'cuda_graph', False)))
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

    def _epoch_permutation(self, n_local, gen):
        """Shuffled order of the local samples for one pass (the reference's DataLoader(shuffle=True),
        calibrators.py:274).  A hook: tests pin it to compare the mini-batch branch with the oracle."""
        return torch.randperm(n_local, device=self.dev, generator=gen)

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

    def _fused(self, logits, target=None, bins=15, want_probs=True):
        """The fused pass (cnf_flow_predict): raw logits -> row-mean centring (calibrators.py:42) -> flow ->
        ``softmax(log(softmax(z)+1e-7) - log_priors)`` (:44) [-> ECE / NLL / accuracy statistics] in ONE
        kernel launch.  Returns the engine's result dict, or None when the flow / shape is outside the fused
        kernels (float64 input, K > 128, heterogeneous flows): the caller then composes the separate kernels."""
        eng = self.flow.engine() if hasattr(self.flow, 'engine') else None
        logits = np.asarray(logits) if not isinstance(logits, torch.Tensor) else logits
        if eng is None or eng.K > 128 or logits.dtype not in (np.float32, torch.float32):
            return None          # numpy centres float64 logits in float64: keep that arithmetic on the host
        if getattr(self, 'precision', 'fp32') != 'bf16' and logits.shape[0] <= 32768 and not getattr(self, 'force_fused', False):
            # fp32 at calibration-set sizes: the 32-sample-tile forward kernel plus the tail kernel is quicker than the
            # one-thread-per-sample fused kernel (0.08 vs ~0.2 ms at N = 10,000); the fused pass pays from ~10^5 rows
            return None
        x = torch.as_tensor(np.ascontiguousarray(logits) if not isinstance(logits, torch.Tensor) else logits)
        self.flow.to(self.dev)
        x = x.to(self.dev)
        y = None
        if target is not None:
            t = np.asarray(target) if not isinstance(target, torch.Tensor) else target
            if t.ndim == 2:
                t = t.argmax(1)
            y = torch.as_tensor(t).to(torch.int64).to(self.dev)
        try:
            return eng.predict(x, center=True, log_priors=self.log_priors, y=y, bins=bins, want_probs=want_probs,
                               precision=getattr(self, 'precision', 'fp32'))
        except NotImplementedError:
            return None

    def predict(self, logits):
        """Same result as the base-class ``predict`` (calibrators.py:40-44): float64 calibrated probabilities.
        Centring, flow and the predict tail run as one kernel launch (see ``_fused``)."""
        res = self._fused(logits)
        if res is not None:
            out = res['probs'].cpu().numpy()
            self.flow.cpu()
            return out
        logits = np.asarray(logits)
        z = self._forward_device(logits - logits.mean(axis=1, keepdims=True)).contiguous()
        n, k = z.shape
        lp = torch.as_tensor(np.asarray(self.log_priors, dtype=np.float64)).to(self.dev)
        out = torch.empty((n, k), dtype=torch.float64, device=self.dev)
        with on_device(self.dev):
            _lib.call('cnf_calibrated_probs', _ptr(z), ctypes.c_int64(n), ctypes.c_int32(k), _ptr(lp), _ptr(out),
                      _stream(self.dev))
        res = out.cpu().numpy()
        self.flow.cpu()
        return res

    def evaluate(self, logits, target, bins=15, reduce=True):
        """ECE / NLL / accuracy of the calibrated probabilities ``self.predict(logits)`` against `target`
        (what the notebooks compute with utils/metrics.py:35-73, 6-15, 76-80 on the calibrator's output) in
        one fused pass: the probabilities are never materialised.  Under torch.distributed every rank passes
        its own shard of (logits, target) and, with `reduce`, the 3*bins+3 sufficient statistics are
        all-reduced before the metrics are formed, so every rank returns the metrics of the whole set."""
        from .utils import metrics as M
        res = self._fused(logits, target=target, bins=bins, want_probs=False)
        if res is not None:
            stats = res['stats']
        else:
            logits = np.asarray(logits)
            z = self._forward_device(logits - logits.mean(axis=1, keepdims=True)).contiguous()
            stats = M.statistics(z, target, bins=bins, mode=_lib.METRICS_CALIBRATED, log_priors=self.log_priors,
                                 device=self.dev)
        if reduce:
            stats = M.reduce_statistics(stats)
        self.flow.cpu()
        return M.metrics_from_statistics(stats, bins)
