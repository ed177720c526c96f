#!/usr/bin/env python
"""Benchmark of the coupling-flow hot path (see BASELINE.json, DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Own arm.  A step = one fused forward + log-det pass of the RealNVP calibrator of BASELINE config C2
(K=10 classes, 6 affine couplings, hidden 128) over 10^8 synthetic logits per GPU -- the size
BASELINE.json's north_star quotes the headline on (configs[1] is the same shape at N=10^6; it is timed
too, as `legs.c2_1m`).  Samples shard over GPUs with no collective (weak scaling).
  value  samples/s with the logits resident in HBM, bf16 tcgen05 kernel (stated tolerance 1e-2); the fp32
         kernel (the API default, 1e-5 of the reference) is timed on the same workload: `summary.fwd_fp32`.
  e2e    the same step through the public host-buffer API (`transform_host` -> cnf_flow_apply_host):
         pinned host logits -> H2D -> kernel -> D2H of z and log-det, all inside the timed region.
Other legs (all in `legs`, headline numbers repeated in `summary`, the LAST key of the line):
C3 training step (bf16 and fp32), C5 as one job (inverse sampling + fused flow->ECE/NLL/accuracy pass +
all-reduce of the 48 statistics), the HBM-bound metrics / affine kernels, C4 (K=100, hidden 512), C1
calibrator fit, a calibration-set-sized epoch, and at N>1 a data-parallel consistency check.

Reference arm (--impl reference): the UNMODIFIED reference (`Flow.forward`, flows/flows.py:17-25) from
oracle/_ref/reference.zip (packed there by __graft_entry__.build(); the GPU box has no /root/reference) on
all host threads, rank 0 only, on a bounded sample of the same workload.  It imports nothing of the
product.  Falls back to the torch-CPU port (oracle/ref_port_torch.py, kind "port") when the copy is absent.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'calibrated samples/sec (flow fwd+logdet)'
UNIT = 'samples/s'
K, L, HIDDEN = 10, 6, [128]
N_HEAD = 100_000_000            # north-star size: 10^8 K=10 logits on one GPU (4 GB in, 4.4 GB out)
N_C2 = 1_000_000                # BASELINE configs[1]
N_ROT = 8                       # rotating input batches of the 10^6 leg: 8 x 84 MB in+out > 126 MB L2
N_E2E_CHUNK = 10_000_000        # pinned host buffers of the end-to-end leg (one step = N_HEAD / this many calls)
BYTES_PER_SAMPLE = 8 * K + 4    # read x, write z, write log-det
FLOPS_PER_SAMPLE = L * 2 * 2 * (5 * 128 + 128 * 5)   # minimal, mask-exploiting: 30 720
WORKLOAD = ('C2 shape at the north-star size: RealNVP K=10 L=6 hidden=[128], fused fwd+logdet, '
            'N=100,000,000 synthetic logits per GPU per step')


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return p['hbm_gbs'], p['bf16_tflops'], p.get('bf16_tflops_sustained'), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 1590.0, None, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML, every 5 ms; falls back
    to polling nvidia-smi when pynvml is unavailable)."""

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical GPUs; honour CUDA_VISIBLE_DEVICES when it is a plain list
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = index
            if vis:
                ids = [v for v in vis.split(',') if v.strip() != '']
                if index < len(ids) and ids[index].strip().isdigit():
                    phys = int(ids[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _run(self):
        n = self.nvml
        while not self.stop.is_set():
            try:
                if n is not None:
                    mhz = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                    r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(
                        n, 'nvmlDeviceGetCurrentClocksEventReasons') else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    self.rows.append((mhz, self.max_mhz, r))
                    self.stop.wait(0.005)
                else:
                    out = subprocess.run(['nvidia-smi', '-i', str(self.index),
                                          '--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active',
                                          '--format=csv,noheader,nounits'], capture_output=True, text=True,
                                         timeout=5).stdout.strip().split(',')
                    self.rows.append((float(out[0]), float(out[1]), int(out[2].strip(), 16)))
                    self.stop.wait(0.05)
            except Exception:
                self.stop.wait(0.05)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        # NVML clocks-event-reason bits
        bits = {'sw_power_cap': 0x4, 'hw_slowdown': 0x8, 'sw_thermal_slowdown': 0x20,
                'hw_thermal_slowdown': 0x40, 'hw_power_brake_slowdown': 0x80}
        sm = sorted(r[0] for r in self.rows)
        mx = max((r[1] for r in self.rows), default=None)
        reasons = sorted({name for r in self.rows for name, b in bits.items() if r[2] & b})
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': reasons,
                'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
# synthetic data: over-confident 80%-accurate classifier logits, row-centred
# ------------------------------------------------------------------------------------------------
def synth(n, seed, device=None, k=K):
    import torch
    g = torch.Generator(device='cpu').manual_seed(seed)
    y = torch.randint(0, k, (n,), generator=g)
    x = 1.5 * torch.randn((n, k), generator=g)
    hit = (torch.rand(n, generator=g) < 0.8).float()
    x[torch.arange(n), y] += 3.0 * hit
    x -= x.mean(dim=1, keepdim=True)
    if device is not None:
        return x.to(device), y.to(device)
    return x, y


def synth_dev(n, seed, dev, k=K, block=12_500_000):
    """The same distribution generated on the device in blocks of 12.5 M rows with per-block seeds, so that
    block b of a sharded run equals block b of the single-GPU run (config C5 splits 10^8 rows 8 ways)."""
    import torch
    x = torch.empty((n, k), dtype=torch.float32, device=dev)
    y = torch.empty(n, dtype=torch.int64, device=dev)
    for b, lo in enumerate(range(0, n, block)):
        hi = min(n, lo + block)
        synth_block(x[lo:hi], y[lo:hi], seed + b)
    return x, y


def synth_block(xb, yb, seed):
    import torch
    m, k = xb.shape
    g = torch.Generator(device=xb.device).manual_seed(seed)
    yb.copy_(torch.randint(0, k, (m,), generator=g, device=xb.device))
    torch.randn((m, k), generator=g, device=xb.device, out=xb)
    xb.mul_(1.5)
    hit = (torch.rand(m, generator=g, device=xb.device) < 0.8).float() * 3.0
    xb.scatter_add_(1, yb.view(-1, 1), hit.view(-1, 1))
    xb.sub_(xb.mean(dim=1, keepdim=True))


def init_flat(seed=1, wmult=300.0, k=K, layers=L, hidden=HIDDEN):
    """Reference init (nn.Linear default x 0.001, flows/flows.py:76-79; same RNG consumption as the reference's
    constructors) x wmult ('trained-like' scale for throughput runs), as the canonical flat vector.  Plain
    torch: shared by both arms, imports nothing of the product."""
    import torch
    torch.manual_seed(seed)
    chunks = []
    widths = [k] + list(hidden) + [k]
    for _ in range(layers):
        for _net in range(2):
            for fan_in, fan_out in zip(widths, widths[1:]):
                lin = torch.nn.Linear(fan_in, fan_out)
                chunks += [lin.weight.detach().reshape(-1) * 0.001 * wmult, lin.bias.detach().reshape(-1) * 0.001 * wmult]
    return torch.cat(chunks)


def make_model(seed=1, wmult=300.0, k=K, layers=L, hidden=HIDDEN):
    import torch
    import cnf_b200
    model = cnf_b200.RealNvpFlow(k, layers=layers, hidden_size=hidden)
    flat = init_flat(seed, wmult, k, layers, hidden)
    off = 0
    with torch.no_grad():
        for lay in model.layers:
            for p in lay.canonical_parameters():
                p.copy_(flat[off:off + p.numel()].view(p.shape))
                off += p.numel()
    assert off == flat.numel()
    return model


# ------------------------------------------------------------------------------------------------
# the reference on the host cores
# ------------------------------------------------------------------------------------------------
class HostReference:
    """The reference's CPU implementation of the path: the unmodified modules when oracle/_ref/reference.zip (or
    /root/reference) exists -> kind 'reference'; else the torch-CPU port -> kind 'port'."""

    def __init__(self):
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import ref_loader
        self.ref = ref_loader.load()
        self.kind = 'reference' if self.ref is not None else 'port'
        if self.ref is None:
            import ref_port_torch
            self.port = ref_port_torch

    def flow(self, flat, k=K, layers=L, hidden=HIDDEN, scale=True):
        """A callable bundle (forward, inverse, module-or-None) with the weights of `flat`."""
        import torch
        if self.ref is not None:
            f = self.ref.Flow([self.ref.NvpCouplingLayer(k, hidden_size=list(hidden), scale=scale) for _ in range(layers)])
            off = 0
            with torch.no_grad():
                for lay in f.layers:
                    for net in (lay.s, lay.t):
                        if isinstance(net, torch.nn.Module):
                            for lin in net.layers:
                                for p in (lin.weight, lin.bias):
                                    p.copy_(flat[off:off + p.numel()].view(p.shape))
                                    off += p.numel()
            assert off == flat.numel()
            return f
        return self.port.unflatten(flat, k, layers, hidden, scale=scale)

    def forward(self, f, x):
        import torch
        with torch.no_grad():
            if self.ref is not None:
                zs, ld = f(x)              # Flow.forward, flows/flows.py:17-25
                return zs[-1], ld
            zs, ld = self.port.forward(f, x)
            return zs[-1], ld

    def inverse(self, f, z):
        import torch
        with torch.no_grad():
            if self.ref is not None:
                xs, ld = f.backward(z)     # Flow.backward, flows/flows.py:27-37
                return xs[-1], ld
            return self.port.inverse(f, z)

    def train_stepper(self, flat, k=K, layers=L, hidden=HIDDEN, scale=True):
        """step(x, y) = one optimiser step of calibrators.py:287-295 (loss, zero_grad, backward, Adam.step)."""
        import torch
        if self.ref is None:
            st = self.port.TrainState(flat, k, layers, hidden, scale=scale, shift=True)
            return st.step
        f = self.flow(flat, k, layers, hidden, scale)
        opt = torch.optim.Adam(f.parameters())
        softmx = torch.nn.Softmax(dim=1)

        def step(xb, yb):
            zs, log_det = f(xb)
            probs = softmx(zs[-1])
            ce = torch.log(probs.gather(1, yb.view(-1, 1)) + 1e-7)
            loss = -torch.mean(ce.squeeze() + log_det)
            f.zero_grad()
            loss.backward()
            opt.step()
            return float(loss.detach())
        return step

    def metrics(self, probs, y):
        """ECE(15) + NLL + accuracy: utils/metrics.py:35-73, 6-15, 76-80 (NumPy-2 shim, see ref_loader)."""
        import numpy as np
        if self.ref is not None:
            m = self.ref.metrics
            oh = np.eye(probs.shape[1], dtype=np.int32)[y]
            return m.expected_calibration_error(probs, y, bins=15), m.neg_log_likelihood(probs, oh), m.accuracy(probs, oh)
        import flow_oracle as orc
        return orc.expected_calibration_error(probs, y, 15), orc.neg_log_likelihood(probs, y), orc.accuracy(probs, y)

    def what(self):
        return ('unmodified reference modules (%s)' % os.path.relpath(self.ref.root, ROOT) if self.ref is not None
                else 'torch CPU ops restating flows/flows.py:101-126 (oracle/ref_port_torch.py)')


def median_time(fn, reps, warm=1):
    ts = []
    for i in range(warm + reps):
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if i >= warm:
            ts.append(dt)
    ts.sort()
    return ts[len(ts) // 2], ts


def run_reference(args):
    import torch
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    host = HostReference()
    flat = init_flat()
    f = host.flow(flat)
    # size the per-step sample so that the whole run stays within ~2 minutes
    probe, _ = synth(100_000, 7)
    t_probe, _ = median_time(lambda: host.forward(f, probe), reps=1, warm=1)
    budget = 120.0 / max(1, args.steps + args.warmup)
    n = int(min(N_C2, max(50_000, 100_000 * budget / max(t_probe, 1e-6))))
    x, _ = synth(n, 11)
    _, times = median_time(lambda: host.forward(f, x), reps=args.steps, warm=args.warmup)
    total = sum(times)
    value = n * len(times) / total
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'sample_per_step': n, 'device': 'cpu',
                   'weights': 'reference init x300 (trained-like), seed 1'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': host.kind,
                         'sample': '%d samples per step, %d steps: Flow.forward under no_grad, %s'
                                   % (n, len(times), host.what())},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


def bind_to_gpu_cpus(local):
    """Pin this rank to the CPUs NVML reports as local to its GPU, so the pinned host buffers of the
    end-to-end leg are first-touched on the GPU's own NUMA node.  (On this pool's single-NUMA-node VMs it
    changes nothing: 1.20 vs 1.21 G samples/s end to end on 2 GPUs.)  Returns the number of CPUs, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import cnf_b200  # noqa: F401  (fails loudly if the CUDA library is missing)
    from cnf_b200 import _lib
    from cnf_b200._engine import _ptr, _stream
    from cnf_b200.utils import metrics as M

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product path has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa = bind_to_gpu_cpus(local) if world > 1 and not os.environ.get('CNF_NO_AFFINITY') else None
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    hbm_peak, tf_peak, tf_sustained, peak_src = peaks()
    G = 1e9

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warm=3, sampler=None):
        """ms per step of fn(i): `warm` untimed calls, then exactly `steps` calls between CUDA events on the
        launching stream, a barrier + synchronize on both sides, max over ranks."""
        for i in range(max(3, warm)):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler is not None:
            sampler.__enter__()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        if sampler is not None:
            sampler.__exit__()
        return max_over_ranks(e0.elapsed_time(e1)) / steps

    def wall_timed(fn, steps, warm=3):
        for i in range(max(3, warm)):
            fn(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            fn(i)
        barrier()
        return max_over_ranks(time.perf_counter() - t0) * 1e3 / steps

    model = make_model().to(dev)
    eng = model.engine()
    eng.ensure(dev)
    has_tc = eng.tc_bytes > 0
    eng.pack(tc=has_tc)
    precisions = (['bf16'] if has_tc else []) + ['fp32']
    if args.precision != 'auto':
        precisions = [args.precision]
    head_prec = precisions[0]
    n_head = args.head_n or N_HEAD
    few = max(3, min(args.steps, 5))             # step count of the slower side legs
    legs = {}

    def fwd_call(prec, x, z, ld, n, inverse=False):
        desc = eng.desc_tc if prec == 'bf16' else eng.desc
        packed = eng.packed_tc if prec == 'bf16' else eng.packed
        _lib.call('cnf_flow_inverse' if inverse else 'cnf_flow_forward', ctypes.byref(desc), _ptr(packed),
                  _ptr(eng.tables), _ptr(x), _ptr(z), _ptr(ld), None, ctypes.c_int64(n), _stream(dev))

    # ---- headline: 10^8 resident logits per step, one launch per step -----------------------------
    xh_dev, yh_dev = synth_dev(n_head, 1000 * (rank + 1), dev)
    zh_dev = torch.empty_like(xh_dev)
    ldh_dev = torch.empty(n_head, dtype=torch.float32, device=dev)
    clocks = ClockSampler(local)
    fwd_ms = {}
    for prec in precisions:
        steps = args.steps if prec == head_prec else few
        fwd_ms[prec] = timed(lambda i, p=prec: fwd_call(p, xh_dev, zh_dev, ldh_dev, n_head), steps, args.warmup,
                             sampler=clocks if prec == head_prec else None)
    ms = fwd_ms[head_prec]
    value = world * n_head / (ms * 1e-3)
    launches = args.steps

    # ---- configs[1] as written: N = 10^6 per launch, rotating inputs (same kernels) ------------------
    xs = [xh_dev[i * N_C2:(i + 1) * N_C2] for i in range(N_ROT)] if n_head >= N_ROT * N_C2 else [synth(N_C2, 100 + i, dev)[0] for i in range(N_ROT)]
    z1 = torch.empty((N_C2, K), dtype=torch.float32, device=dev)
    l1 = torch.empty(N_C2, dtype=torch.float32, device=dev)
    c2 = {}
    for prec in precisions:
        t = timed(lambda i, p=prec: fwd_call(p, xs[i % N_ROT], z1, l1, N_C2), max(args.steps, 20), args.warmup)
        c2[prec] = {'value': world * N_C2 / (t * 1e-3), 'ms_per_step': t}
    legs['c2_1m'] = {'what': 'BASELINE configs[1]: the same pass at N=1,000,000 per GPU per launch, 8 rotating inputs',
                     'unit': UNIT, **{p: c2[p] for p in c2}}

    # ---- end to end through the public host-buffer API ---------------------------------------------
    n_chunk = min(N_E2E_CHUNK, n_head)
    calls_per_step = max(1, n_head // n_chunk)
    xh = torch.empty((n_chunk, K), dtype=torch.float32, pin_memory=True)
    xh.copy_(xh_dev[:n_chunk])
    zh = torch.empty((n_chunk, K), dtype=torch.float32, pin_memory=True)
    lh = torch.empty(n_chunk, dtype=torch.float32, pin_memory=True)
    e2e_by = {}
    for prec in precisions:
        model.flow.precision = prec

        def e2e_step(i):
            for _ in range(calls_per_step):
                model.transform_host(xh, zh, lh, device=dev)      # synchronises: the result is on the host
        steps = args.steps if prec == head_prec else few
        t = wall_timed(e2e_step, steps, 3)
        e2e_by[prec] = world * calls_per_step * n_chunk / (t * 1e-3)
    model.flow.precision = 'fp32'
    # platform ceiling of this leg: pinned cudaMemcpyAsync of the same buffers, both directions at once, on every rank
    # at the same time (what the box's host-memory / PCIe path carries, whatever kernel sits in between)
    d_in = torch.empty_like(xh, device=dev)
    d_out = torch.empty((n_chunk, K + 1), dtype=torch.float32, device=dev)
    h_out = torch.empty((n_chunk, K + 1), dtype=torch.float32, pin_memory=True)
    cs1, cs2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def copy_round(i):
        with torch.cuda.stream(cs1):
            d_in.copy_(xh, non_blocking=True)
        with torch.cuda.stream(cs2):
            h_out.copy_(d_out, non_blocking=True)
        cs1.synchronize()
        cs2.synchronize()
    t_copy = wall_timed(copy_round, 5, 3)
    ceiling = world * n_chunk / (t_copy * 1e-3)
    del d_in, d_out, h_out
    e2e = {'value': e2e_by[head_prec], 'unit': UNIT,
           'h2d_bytes_per_step': calls_per_step * xh.numel() * 4,
           'd2h_bytes_per_step': calls_per_step * (zh.numel() * 4 + lh.numel() * 4),
           'api': 'RealNvpFlow.transform_host -> cnf_flow_apply_host: pinned host logits in, host z + log-det out; '
                  'one step = %d calls of %d samples through the same pinned buffers' % (calls_per_step, n_chunk),
           'by_precision': e2e_by,
           'platform_ceiling': {'value': ceiling, 'unit': UNIT, 'frac': e2e_by[head_prec] / ceiling,
                                'what': 'pinned cudaMemcpyAsync H2D + D2H of the same 84 B/sample at the same time on all '
                                        '%d rank(s): the host-memory / PCIe path of the box' % world}}
    del xh, zh, lh

    # ---- config C5 as ONE job: inverse sampling + fused flow -> ECE/NLL/accuracy + all-reduce ---------
    # 10^8 samples in total at every N (strong scaling): rank r owns blocks r, r+world, ... of 12.5 M rows.
    c5 = None
    if not args.no_extra:
        n5 = (args.c5_n or 100_000_000) // world
        if n5 > n_head:
            n5 = n_head
        # logits of this rank's blocks (block seeds are global, so the union over ranks is the same set at every N)
        blocks = list(range(rank, 8, world)) if (n5 * world == 100_000_000 and world in (1, 2, 4, 8)) else None
        x5, y5 = synth_dev(n5, 5000 + rank, dev)
        if blocks is not None:
            for j, b in enumerate(blocks):
                synth_block(x5[j * 12_500_000:(j + 1) * 12_500_000], y5[j * 12_500_000:(j + 1) * 12_500_000], 5000 + b)
        lp = torch.log(torch.bincount(y5, minlength=K).double() / n5)
        if world > 1:
            cnt = torch.bincount(y5, minlength=K).double()
            dist.all_reduce(cnt)
            lp = torch.log(cnt / cnt.sum())
        lp_np = lp.cpu().numpy()
        zbase = zh_dev[:n5]                    # "sampling": base-space points to invert (here the forward images)
        fwd_call(head_prec, x5, zbase, ldh_dev[:n5], n5)
        xrec = torch.empty_like(zbase)
        ldi = torch.empty(n5, dtype=torch.float32, device=dev)
        stats_box = [None]

        def c5_job(i, prec=head_prec):
            fwd_call(prec, zbase, xrec, ldi, n5, inverse=True)                       # inverse-pass sampling
            res = eng.predict(x5, center=False, log_priors=lp_np, y=y5, bins=15, precision=prec, repack=False)
            stats_box[0] = M.reduce_statistics(res['stats'])                        # 48 doubles over NCCL
        c5 = {}
        for prec in precisions:
            t = timed(lambda i, p=prec: c5_job(i, p), few, 3)
            st = stats_box[0]
            mt = M.metrics_from_statistics(st, 15)
            c5[prec] = {'value': world * n5 / (t * 1e-3), 'ms_per_step': t, 'ece': mt['ece'], 'nll': mt['nll'],
                        'accuracy': mt['accuracy'], 'n': mt['n']}
        # the fused statistics pass alone: reads 4K+8 B/sample, writes nothing
        fused = {}
        for prec in precisions:
            t = timed(lambda i, p=prec: eng.predict(x5, center=False, log_priors=lp_np, y=y5, bins=15, precision=p,
                                                    repack=False), few, 3)
            fused[prec] = {'value': world * n5 / (t * 1e-3), 'ms_per_step': t,
                           'hbm_gbs': (4 * K + 8) * n5 / (t * 1e-3) / 1e9}
        rt = float((xrec - x5).abs().max())
        legs['c5_job'] = {'what': 'C5 as one job: Flow.backward (inverse + log-det) on this rank\'s share of 10^8 base points, '
                                  'then ONE fused pass flow -> softmax(log(softmax(z)+1e-7)-log_priors) -> ECE(15)/NLL/accuracy '
                                  '(4K+8 B/sample read, nothing written), then all-reduce of the 48 statistics; value = '
                                  'samples/s of the whole job', 'unit': UNIT, 'samples_total': n5 * world,
                          'scaling': 'strong (10^8 samples over all GPUs)', 'roundtrip_max_abs': rt, **c5}
        legs['c5_fused_stats'] = {'what': 'the fused flow -> calibrated ECE/NLL/accuracy pass alone (cnf_flow_predict, statistics only)',
                                  'unit': UNIT, 'algorithmic_bytes_per_sample': 4 * K + 8, **fused,
                                  'note': 'bound by the flow arithmetic, not by HBM: 48 B/sample at this rate is a few % of the HBM roof'}
        del xrec, ldi

    # ---- training step, config C3 shape ---------------------------------------------------------------
    train = {}
    if not args.no_train:
        n_loc = args.train_n or (64 * (1 << 20)) // world
        xt, yt = xh_dev[:n_loc], yh_dev[:n_loc]
        if n_loc > n_head:
            xt, yt = synth_dev(n_loc, 7000 + rank, dev)
        for prec in precisions:
            if prec == 'bf16' and eng.tc_train is None:
                continue
            tmodel = make_model(seed=2, wmult=1.0).to(dev)          # reference init for training
            tr = cnf_b200.FusedNLLTrainer(tmodel.engine(), xt, yt, n_total=n_loc * world, precision=prec)
            t = timed(lambda i: tr.step(), few, 3)
            train[prec] = {'value': world * n_loc / (t * 1e-3), 'unit': UNIT, 'ms_per_step': t,
                           'samples_per_gpu': n_loc, 'samples_total': n_loc * world, 'steps': few,
                           'scaling': 'strong (C3: 64 Mi samples over all GPUs)' if not args.train_n else 'weak',
                           'loss': -float(tr.loss_acc[0]) / (n_loc * world),
                           'what': ('tcgen05 forward (+tape) and backward kernels per 2^20-sample chunk' if prec == 'bf16'
                                    else 'fused NLL fwd+bwd kernel') + ' + grad reduce' +
                                   (' + ONE NCCL all-reduce (grad + loss sums)' if world > 1 else '') + ' + Adam + repack per step'}
            if prec == 'bf16':
                train[prec]['tensor_tflops_minimal'] = train[prec]['value'] * 4 * FLOPS_PER_SAMPLE / 1e12
            else:
                # forward + recompute + dgrad + wgrad = 4 x the forward's minimal flops, against the fp32 FMA peak
                tfl = train[prec]['value'] / world * 4 * FLOPS_PER_SAMPLE / 1e12
                train[prec]['roofline'] = {'bound': 'fp32 FMA issue', 'achieved': tfl, 'peak': 148 * 128 * 2 * 1.965e9 / 1e12,
                                           'unit': 'TFLOP/s', 'frac': tfl / (148 * 128 * 2 * 1.965e9 / 1e12)}
                train[prec]['what'] = train[prec]['what'].replace('fused NLL fwd+bwd kernel', 'register-resident fused NLL fwd+bwd kernel (train_reg10_kernel)')
            del tr, tmodel
        legs['train_step'] = train
        # the same step with the samples in pinned HOST memory, streamed through the GPU every step (SURVEY.md 8f rank 4:
        # sets larger than HBM): H2D of chunk c+1 overlaps the kernel of chunk c
        if not args.no_extra:
            n_st = min(n_loc, 16 * (1 << 20))
            xs_h = torch.empty((n_st, K), dtype=torch.float32, pin_memory=True)
            ys_h = torch.empty(n_st, dtype=torch.int64, pin_memory=True)
            xs_h.copy_(xt[:n_st])
            ys_h.copy_(yt[:n_st])
            streamed = {}
            for prec in precisions:
                if prec == 'bf16' and eng.tc_train is None:
                    continue
                tmodel = make_model(seed=2, wmult=1.0).to(dev)
                trs = cnf_b200.HostStreamNLLTrainer(tmodel.engine(), xs_h, ys_h, dev, n_total=n_st * world, resident=False,
                                                    chunk_rows=1 << 21, precision=prec)
                t = wall_timed(lambda i: (trs.step(), torch.cuda.current_stream(dev).synchronize()), few, 3)
                streamed[prec] = {'value': world * n_st / (t * 1e-3), 'ms_per_step': t}
                del trs, tmodel
            legs['train_step_host_streamed'] = {
                'what': 'the same Adam step with the %d samples per GPU resident in pinned host memory and streamed over PCIe '
                        'every step (HostStreamNLLTrainer, 2^21-row chunks double-buffered; 48 B/sample H2D)' % n_st,
                'unit': UNIT, 'h2d_bytes_per_step': n_st * (4 * K + 8), **streamed}
            del xs_h, ys_h

    # ---- data-parallel consistency (N > 1): identical parameters on every rank, same loss as one rank -----
    dp_check = None
    if world > 1 and not args.no_train:
        n_dp = 1 << 20
        xg, yg = synth_dev(n_dp, 4242, dev)                       # the same global set on every rank
        lo, hi = cnf_b200.calibrators.shard_bounds(n_dp, rank, world)
        torch.manual_seed(100 + rank)                             # ranks start from DIFFERENT weights on purpose
        m_dp = cnf_b200.RealNvpFlow(K, layers=L, hidden_size=HIDDEN).to(dev)
        tr = cnf_b200.FusedNLLTrainer(m_dp.engine(), xg[lo:hi].contiguous(), yg[lo:hi].contiguous(), n_total=n_dp)
        start = m_dp.engine().flat.clone()                        # rank 0's weights after the broadcast
        for _ in range(4):
            tr.step()
        flat = m_dp.engine().flat
        sig = torch.stack([flat.view(torch.int32).to(torch.int64).sum(), (flat.view(torch.int32).to(torch.int64) *
                           torch.arange(1, flat.numel() + 1, device=dev)).sum()])
        sigs = [torch.zeros_like(sig) for _ in range(world)]
        dist.all_gather(sigs, sig)
        identical = all(bool((s == sigs[0]).all()) for s in sigs)
        loss_dp = -float(tr.loss_acc[0]) / n_dp
        # the same four steps on ONE rank over the whole set, from the same start
        m_1 = cnf_b200.RealNvpFlow(K, layers=L, hidden_size=HIDDEN).to(dev)
        e1 = m_1.engine()
        e1.ensure(dev)
        e1.flat.copy_(start)
        saved = cnf_b200.calibrators._dist
        cnf_b200.calibrators._dist = lambda: None
        try:
            tr1 = cnf_b200.FusedNLLTrainer(e1, xg, yg, n_total=n_dp)
            for _ in range(4):
                tr1.step()
        finally:
            cnf_b200.calibrators._dist = saved
        loss_1 = -float(tr1.loss_acc[0]) / n_dp
        dflat = float((e1.flat - flat).abs().max())
        dp_check = {'ok': bool(identical and abs(loss_dp - loss_1) <= 1e-6 * max(1.0, abs(loss_1))),
                    'params_bit_identical_across_ranks': identical, 'loss_dp': loss_dp, 'loss_single_rank': loss_1,
                    'max_abs_param_diff_vs_single_rank': dflat,
                    'what': '4 fp32 Adam steps on 2^20 samples sharded over the ranks (ranks built from different '
                            'seeds, start state broadcast from rank 0) vs the same steps on one rank'}
        del xg, yg, tr, tr1

    # ---- HBM-bound kernels of the path: metrics and the constant affine layer ----------------------------
    if not args.no_extra:
        n_m = 12_500_000
        pm = torch.softmax(xh_dev[:n_m], dim=1).contiguous()
        ym = yh_dev[:n_m]
        edges = torch.from_numpy(M.bin_edges(15)).to(dev)
        macc = torch.zeros(48, dtype=torch.float64, device=dev)

        def mstep(i):
            _lib.call('cnf_metrics', _ptr(pm), ctypes.c_int32(0), _ptr(ym), ctypes.c_int64(n_m), ctypes.c_int32(K),
                      ctypes.c_int32(15), ctypes.c_int32(0), None, _ptr(edges), _ptr(macc), _stream(dev))
        met_ms = timed(mstep, 20, 3)
        n_a = 10_000_000
        xa = xh_dev[:n_a]
        za = zh_dev[:n_a]
        sa = torch.full((K,), 0.1, device=dev)
        ta = torch.full((K,), -0.2, device=dev)

        def astep(i):
            _lib.call('cnf_affine_const', _ptr(xa), _ptr(sa), _ptr(ta), _ptr(za), ctypes.c_int64(n_a), ctypes.c_int32(K),
                      ctypes.c_int32(0), _stream(dev))
        aff_ms = timed(astep, 20, 3)
        ab = 8 * K * n_a / (aff_ms * 1e-3) / 1e9
        mb = (4 * K + 8) * n_m / (met_ms * 1e-3) / 1e9
        legs['affine_const'] = {'value': world * n_a / (aff_ms * 1e-3), 'unit': UNIT, 'ms_per_step': aff_ms, 'samples_per_gpu': n_a,
                                'what': 'AffineConstantLayer.forward z = x*exp(s)+t (flows/flows.py:53-58), fp32',
                                'roofline': {'bound': 'hbm', 'achieved': ab, 'peak': hbm_peak, 'unit': 'GB/s',
                                             'frac': ab / hbm_peak, 'note': 'algorithmic bytes 8K = 80 B/sample'}}
        legs['metrics'] = {'value': world * n_m / (met_ms * 1e-3), 'unit': UNIT, 'ms_per_step': met_ms, 'samples_per_gpu': n_m,
                           'what': 'ECE(15 bins)+NLL+accuracy in one pass over fp32 probabilities',
                           'roofline': {'bound': 'hbm', 'achieved': mb, 'peak': hbm_peak, 'unit': 'GB/s',
                                        'frac': mb / hbm_peak, 'note': 'algorithmic bytes 4K+8 = 48 B/sample'}}
        del pm

    # the 8.8 GB of headline buffers are no longer needed
    inv_keep = None
    if not args.no_extra:
        t = timed(lambda i: fwd_call(head_prec, zh_dev, xh_dev, ldh_dev, min(n_head, 12_500_000), inverse=True), few, 3)
        inv_keep = {'value': world * min(n_head, 12_500_000) / (t * 1e-3), 'unit': UNIT, 'ms_per_step': t, 'dtype': head_prec,
                    'what': 'Flow.backward (inverse + log-det) on 12,500,000 samples per GPU (config C5 share at 8 GPUs)'}
        legs['inverse'] = inv_keep
    flat_c2 = eng.flat.detach().cpu().clone()
    del xh_dev, zh_dev, ldh_dev, yh_dev, xs, z1, l1
    torch.cuda.empty_cache()

    # ---- config C4: K=100, 8 couplings, hidden 512 --------------------------------------------------------
    if not args.no_extra:
        c4 = make_model(seed=4, wmult=60.0, k=100, layers=8, hidden=[512]).to(dev)
        e4 = c4.engine()
        e4.ensure(dev)
        e4.pack(tc=True)
        n4 = args.c4_n or 10_000_000
        g4 = torch.Generator(device=dev).manual_seed(40 + rank)
        x4 = torch.empty((n4, 100), dtype=torch.float32, device=dev)
        for lo in range(0, n4, 1_000_000):
            blk = x4[lo:lo + 1_000_000]
            torch.randn(blk.shape, generator=g4, device=dev, out=blk)
            blk.mul_(1.5)
            blk.sub_(blk.mean(dim=1, keepdim=True))
        z4 = torch.empty_like(x4)
        l4 = torch.empty(n4, dtype=torch.float32, device=dev)

        def c4step(i):
            _lib.call('cnf_flow_forward', ctypes.byref(e4.desc_tc), _ptr(e4.packed_tc), _ptr(e4.tables), _ptr(x4),
                      _ptr(z4), _ptr(l4), None, ctypes.c_int64(n4), _stream(dev))
        c4_ms = timed(c4step, few, 3)
        c4_flops = 8 * 2 * 2 * (50 * 512 + 512 * 50)
        c4_tf = c4_flops * n4 / (c4_ms * 1e-3) / 1e12
        legs['c4_forward'] = {'value': world * n4 / (c4_ms * 1e-3), 'unit': UNIT, 'ms_per_step': c4_ms, 'dtype': 'bf16',
                              'what': 'C4: RealNVP K=100 L=8 hidden=[512] fwd+logdet, %d samples per GPU, '
                                      'streamed-weight tcgen05 kernel' % n4,
                              'roofline': {'bound': 'tensor', 'achieved': c4_tf, 'peak': tf_peak, 'unit': 'TFLOP/s',
                                           'frac': c4_tf / tf_peak, 'note': 'minimal 1,638,400 flop/sample'}}
        # the same shape through the fp32 kernels (the 1e-5 parity path and the training path of wide shapes)
        n4f = 1_000_000
        c4.flow.precision = 'fp32'
        with torch.no_grad():
            c4f_ms = timed(lambda i: c4(x4[:n4f]), 3, 3)
        n4t = 100_000
        y4 = torch.randint(0, 100, (n4t,), generator=g4, device=dev)
        tr4 = cnf_b200.FusedNLLTrainer(e4, x4[:n4t].contiguous(), y4, n_total=n4t * world)
        c4t_ms = timed(lambda i: tr4.step(), 3, 3)
        legs['c4_fp32'] = {'forward': {'value': world * n4f / (c4f_ms * 1e-3), 'unit': UNIT, 'ms_per_step': c4f_ms},
                           'train_step': {'value': world * n4t / (c4t_ms * 1e-3), 'unit': UNIT, 'ms_per_step': c4t_ms},
                           'dtype': 'f32',
                           'what': 'C4 shape on the lean fp32 kernels: forward on 1,000,000 samples, NLL + Adam step on 100,000'}
        del x4, z4, l4, tr4, c4, e4
        torch.cuda.empty_cache()

    # ---- conditioners with two wide hidden layers (flows/utils.py:6-31): the H x H middle Linear on tcgen05 -----------
    if not args.no_extra:
        n6 = 4_000_000
        m6 = make_model(seed=7, wmult=30.0, hidden=[128, 128]).to(dev)
        e6 = m6.engine()
        e6.ensure(dev)
        e6.pack(tc=True)
        e6.pack()
        x6, _ = synth_dev(n6, 6000 + rank, dev)
        z6 = torch.empty_like(x6)
        l6 = torch.empty(n6, dtype=torch.float32, device=dev)

        def d6step(i):
            _lib.call('cnf_flow_forward', ctypes.byref(e6.desc_tc), _ptr(e6.packed_tc), _ptr(e6.tables), _ptr(x6),
                      _ptr(z6), _ptr(l6), None, ctypes.c_int64(n6), _stream(dev))
        d6_ms = timed(d6step, few, 3)
        n6f = 400_000
        with torch.no_grad():
            d6f_ms = timed(lambda i: e6.apply(x6[:n6f], repack=False), 3, 3)
        f6 = 6 * 2 * 2 * (5 * 128 + 128 * 128 + 128 * 5)          # minimal flops per sample
        d6_tf = f6 * n6 / (d6_ms * 1e-3) / 1e12
        legs['two_hidden_layers'] = {
            'what': 'K=10, 6 couplings, hidden_size=[128, 128], fwd+logdet: bf16 tcgen05 kernel (cnf_flow_tcm.cu) on %d '
                    'samples per GPU; the fp32 kernel (flow_apply_deep_kernel) on %d' % (n6, n6f),
            'value': world * n6 / (d6_ms * 1e-3), 'unit': UNIT, 'ms_per_step': d6_ms, 'dtype': 'bf16',
            'fp32': {'value': world * n6f / (d6f_ms * 1e-3), 'unit': UNIT, 'ms_per_step': d6f_ms},
            'roofline': {'bound': 'tensor', 'achieved': d6_tf, 'peak': tf_peak, 'unit': 'TFLOP/s', 'frac': d6_tf / tf_peak,
                         'note': 'minimal %d flop/sample' % f6}}
        del x6, z6, l6, m6, e6
        torch.cuda.empty_cache()

    # ---- the reference's DEFAULT conditioner (NvpCouplingLayer(dim, hidden_size=[5, 5]), flows/flows.py:69) ----------
    if not args.no_extra:
        n5 = 10_000_000
        m5 = make_model(seed=5, wmult=300.0, hidden=[5, 5]).to(dev)
        e5 = m5.engine()
        e5.ensure(dev)
        e5.pack()
        x5, y5 = synth_dev(n5, 5000 + rank, dev)
        z5 = torch.empty_like(x5)
        l5 = torch.empty(n5, dtype=torch.float32, device=dev)

        def d5step(i):
            _lib.call('cnf_flow_forward', ctypes.byref(e5.desc), _ptr(e5.packed), _ptr(e5.tables), _ptr(x5), _ptr(z5),
                      _ptr(l5), None, ctypes.c_int64(n5), _stream(dev))
        d5_ms = timed(d5step, few, 3)
        n5t = min(n5, 4_000_000)
        m5t = make_model(seed=6, wmult=1.0, hidden=[5, 5]).to(dev)
        tr5 = cnf_b200.FusedNLLTrainer(m5t.engine(), x5[:n5t], y5[:n5t], n_total=n5t * world)
        d5t_ms = timed(lambda i: tr5.step(), few, 3)
        f5 = 6 * 2 * 2 * (5 * 5 + 5 * 5 + 5 * 5)                  # minimal flops per sample
        legs['default_conditioner'] = {
            'what': 'K=10, 6 couplings, hidden_size=[5, 5] (the reference default), fp32 register-resident kernels: '
                    'forward + log-det on %d samples per GPU, NLL + Adam step on %d' % (n5, n5t),
            'unit': UNIT, 'dtype': 'f32',
            'forward': {'value': world * n5 / (d5_ms * 1e-3), 'ms_per_step': d5_ms,
                        'hbm_gbs': 84 * n5 / (d5_ms * 1e-3) / 1e9, 'frac_of_hbm_peak': 84 * n5 / (d5_ms * 1e-3) / 1e9 / hbm_peak,
                        'fma_tflops_minimal': f5 * n5 / (d5_ms * 1e-3) / 1e12},
            'train_step': {'value': world * n5t / (d5t_ms * 1e-3), 'ms_per_step': d5t_ms}}
        del x5, y5, z5, l5, tr5, m5, m5t, e5
        torch.cuda.empty_cache()

    # ---- calibrator API at calibration-set sizes (rank 0, N=1 only) -----------------------------------------
    if not args.no_extra and rank == 0 and world == 1:
        rs = np.random.RandomState(3)
        y1 = rs.randint(0, 3, size=10_000)
        x1 = (1.5 * rs.randn(10_000, 3)).astype(np.float32)
        x1[np.arange(10_000), y1] += 3.0 * (rs.rand(10_000) < 0.8)
        t1 = np.eye(3, dtype=np.float32)[y1]

        def c1_run():
            cal = cnf_b200.TorchFlowCalibrator(cnf_b200.NiceFlow, x1, t1, layers=4, hidden_size=[32], epochs=50, dev=dev)
            return cal.predict(x1)
        c1_run()
        torch.cuda.synchronize()
        c1_times = []
        for _ in range(5):
            t0 = time.perf_counter()
            p1 = c1_run()
            torch.cuda.synchronize()
            c1_times.append(time.perf_counter() - t0)
        c1_s = sorted(c1_times)[2]
        legs['c1_calibrator'] = {'value': 10_000 * 50 / c1_s, 'unit': UNIT, 'wall_s': c1_s, 'dtype': 'f32', 'repeats': 5,
                                 'what': 'C1: TorchFlowCalibrator(NiceFlow, K=3, N=10,000, 4 couplings, hidden 32): fit 50 '
                                         'full-batch epochs + fused predict, host numpy in/out; median wall time of 5 runs',
                                 'finite': bool(np.isfinite(p1).all())}
        # the fits whose wall-clock prints are the reference's only performance evidence
        rsn = np.random.RandomState(1)
        yn = rsn.randint(0, 3, size=1500)
        xn = (1.5 * rsn.randn(1500, 3)).astype(np.float32)
        xn[np.arange(1500), yn] += 3.0 * (rsn.rand(1500) < 0.8)
        tn = np.eye(3, dtype=np.float32)[yn]
        nb = {}
        # (RealNVP: this objective, -mean(log p_y + log_det), is unbounded below on separable synthetic logits -- the
        #  unmodified reference reaches NaN after 489 / 744 epochs on such data, and so does this implementation -- so
        #  its fit is timed over the first 300 epochs and compared per epoch)
        for name, fac, pub, n_ep in (('RealNvpFlow', cnf_b200.RealNvpFlow, 188.335, 300), ('NiceFlow', cnf_b200.NiceFlow, 160.681, 5000)):
            def nb_fit(n_epochs):
                best_t, pred = None, None
                for _rep in range(2):
                    torch.manual_seed(1)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    caln = cnf_b200.TorchFlowCalibrator(fac, xn, tn, layers=5, hidden_size=[3, 3], epochs=n_epochs, dev=dev)
                    pred = caln.predict(xn)
                    torch.cuda.synchronize()
                    dtn = time.perf_counter() - t0
                    best_t = dtn if best_t is None else min(best_t, dtn)
                return best_t, pred
            best, pn = nb_fit(n_ep)
            short, _ = nb_fit(n_ep // 3)            # the same call with a third of the epochs: the difference is epochs only
            marginal = (best - short) / (n_ep - n_ep // 3)
            nb[name] = {'epochs': n_ep, 'wall_s': best, 'us_per_epoch': best / n_ep * 1e6,
                        'us_per_additional_epoch': marginal * 1e6,
                        'published_reference_us_per_epoch': pub / 5000 * 1e6,
                        'published_over_ours_per_epoch': (pub / 5000) / (best / n_ep), 'finite': bool(np.isfinite(pn).all())}
        legs['notebook_fit'] = {
            'what': 'TorchFlowCalibrator(<flow>, N=1500, K=3, layers=5, hidden_size=[3,3], epochs=...) + predict, host numpy '
                    'in/out, best of 2: the fits of notebooks/simulated-predictions-flows.ipynb:214 / :224, whose printed '
                    'wall times for 5000 epochs (:165, :166: 188.3 s / 160.7 s; author\'s CPU, their data) are the '
                    'reference\'s only published timings; synthetic 3-class logits here',
            'unit': 's', **nb}
        xe, ye = synth(5000, 77, dev)
        ep = {}
        for prec in precisions:
            me = make_model(seed=2, wmult=1.0).to(dev)
            if prec == 'bf16' and me.engine().tc_train is None:
                continue
            tre = cnf_b200.FusedNLLTrainer(me.engine(), xe, ye, precision=prec)

            def epoch(i):
                tre.step()
                tre.evaluate()
            ep[prec] = wall_timed(epoch, 200, 5) * 1e3
        legs['calibration_set_epoch'] = {
            'value': 5000 / (ep['fp32'] * 1e-6), 'unit': UNIT, 'us_per_epoch': ep, 'dtype': 'f32',
            'what': 'C2 shape, N=5,000: full-batch Adam step + evaluation pass per epoch, wall time over 200 epochs; '
                    'value = samples per second on the fp32 (1e-5 parity) kernels'}

    # ---- rooflines of the dominant kernel (measured above with CUDA events on the launching stream) ----------
    kname = 'flow_tc_kernel' if head_prec == 'bf16' else 'flow_apply_kernel'
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):       # dram__bytes_read.sum + dram__bytes_write.sum per launch, from ncu --set full
        traffic = json.load(open(tpath)).get(kname + '_1e8' if n_head == N_HEAD else kname)
    ach_tf = FLOPS_PER_SAMPLE * n_head / (ms * 1e-3) / 1e12
    ach_gb = BYTES_PER_SAMPLE * n_head / (ms * 1e-3) / 1e9
    if head_prec == 'bf16':
        roof = {'bound': 'tensor', 'achieved': ach_tf, 'peak': tf_peak, 'unit': 'TFLOP/s', 'frac': ach_tf / tf_peak,
                'traffic': traffic, 'kernel': kname, 'peak_source': peak_src + ', bf16 burst',
                'frac_of_sustained_peak': (ach_tf / tf_sustained) if tf_sustained else None,
                'hbm_note': {'achieved_gbs': ach_gb, 'frac_of_hbm_peak': ach_gb / hbm_peak},
                'note': 'minimal (mask-exploiting) 30,720 flop/sample x %d samples per launch; the tensor pipe binds '
                        '(SURVEY.md 8d / F13), algorithmic HBM bytes 84 B/sample' % n_head}
    else:
        fma_peak = 148 * 128 * 2 * 1.965e9 / 1e12
        roof = {'bound': 'hbm', 'achieved': ach_gb, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach_gb / hbm_peak,
                'traffic': traffic, 'kernel': kname, 'peak_source': peak_src,
                'fma_note': {'achieved_tflops': ach_tf, 'fp32_fma_peak_tflops': fma_peak, 'frac': ach_tf / fma_peak},
                'note': 'fp32 path is FMA-issue bound; HBM figure from 84 B/sample'}
    if 'fp32' in fwd_ms:
        tf32 = FLOPS_PER_SAMPLE * n_head / (fwd_ms['fp32'] * 1e-3) / 1e12
        legs['fwd_fp32'] = {'value': world * n_head / (fwd_ms['fp32'] * 1e-3), 'unit': UNIT, 'ms_per_step': fwd_ms['fp32'],
                            'what': 'the headline workload on the fp32 CUDA-core kernel (API default, 1e-5 of the reference)',
                            'roofline': {'bound': 'fp32 FMA issue', 'achieved': tf32, 'peak': 148 * 128 * 2 * 1.965e9 / 1e12,
                                         'unit': 'TFLOP/s', 'frac': tf32 / (148 * 128 * 2 * 1.965e9 / 1e12)}}

    # ---- the reference on this box's host cores (rank 0, N=1) -------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        torch.set_num_threads(os.cpu_count() or 1)
        host = HostReference()
        cores = torch.get_num_threads()
        f = host.flow(flat_c2)
        xc, yc = synth(N_C2, 11)
        med, _ = median_time(lambda: host.forward(f, xc), reps=3, warm=1)
        cpu = {'value': N_C2 / med, 'unit': UNIT, 'cores': cores, 'kind': host.kind,
               'sample': '1,000,000 samples x 3 passes (median): Flow.forward under no_grad, ' + host.what()}
        if train:
            n_t = 200_000
            xt_c, yt_c = synth(n_t, 13)
            step = host.train_stepper(flat_c2 / 300.0)
            step(xt_c, yt_c)
            t0 = time.perf_counter()
            for _ in range(2):
                step(xt_c, yt_c)
            dt = (time.perf_counter() - t0) / 2
            legs['train_step']['cpu_baseline'] = {
                'value': n_t / dt, 'unit': UNIT, 'cores': cores, 'kind': host.kind,
                'sample': '200,000 samples x 2 full-batch steps of calibrators.py:287-295 (loss, backward, Adam.step), ' + host.what()}
        if not args.no_extra:
            zc, _ = host.forward(f, xc)
            med_i, _ = median_time(lambda: host.inverse(f, zc), reps=1, warm=1)
            legs['inverse']['cpu_baseline'] = {'value': N_C2 / med_i, 'unit': UNIT, 'cores': cores, 'kind': host.kind,
                                               'sample': '1,000,000 samples, one pass of Flow.backward, ' + host.what()}
            # C5 job on the host: inverse + forward + predict tail + metrics on 1,000,000 samples
            from scipy.special import softmax as sp_softmax
            lp_c = np.log(np.bincount(yc.numpy(), minlength=K) / float(N_C2))

            def c5_host():
                host.inverse(f, zc)
                zz, _ = host.forward(f, xc)
                pr = sp_softmax(np.log(sp_softmax(zz.numpy(), axis=1) + 1e-7) - lp_c, axis=1)     # calibrators.py:44, 352
                return host.metrics(pr, yc.numpy())
            t0 = time.perf_counter()
            mt_c = c5_host()
            dt5 = time.perf_counter() - t0
            legs['c5_job']['cpu_baseline'] = {'value': N_C2 / dt5, 'unit': UNIT, 'cores': cores, 'kind': host.kind,
                                              'ece': float(mt_c[0]), 'sample': '1,000,000 samples, one run of the same job, ' + host.what()}
            n_mc = 2_000_000
            xm, ym_c = synth(n_mc, 17)
            pm_c = torch.softmax(xm, dim=1).numpy()
            t0 = time.perf_counter()
            host.metrics(pm_c, ym_c.numpy())
            dt = time.perf_counter() - t0
            legs['metrics']['cpu_baseline'] = {'value': n_mc / dt, 'unit': UNIT, 'cores': 1, 'kind': host.kind,
                                               'sample': '2,000,000 samples: utils/metrics.py:35-73, 6-15, 76-80 (numpy), ' + host.what()}
            if 'c1_calibrator' in legs:
                st1 = host.train_stepper(_nice_flat(), 3, 4, [32], scale=False)
                g1 = torch.Generator().manual_seed(5)
                xs1 = 1.5 * torch.randn(10_000, 3, generator=g1)
                ys1 = torch.randint(0, 3, (10_000,), generator=g1)
                st1(xs1, ys1)
                t0 = time.perf_counter()
                for _ in range(50):
                    st1(xs1, ys1)
                dt1 = time.perf_counter() - t0
                legs['c1_calibrator']['cpu_baseline'] = {
                    'value': 10_000 * 50 / dt1, 'unit': UNIT, 'cores': cores, 'kind': host.kind,
                    'sample': '50 full-batch steps on 10,000 samples WITHOUT the reference\'s DataLoader (calibrators.py:'
                              '268-283 collates per sample, which dominates the reference at this size, SURVEY.md 6), ' + host.what()}
            if 'calibration_set_epoch' in legs:
                ge = torch.Generator().manual_seed(7)
                xse = 1.5 * torch.randn(5000, 10, generator=ge)
                yse = torch.randint(0, 10, (5000,), generator=ge)
                ste = host.train_stepper(flat_c2 / 300.0)
                ste(xse, yse)
                t0 = time.perf_counter()
                for _ in range(20):
                    ste(xse, yse)
                dte = (time.perf_counter() - t0) / 20
                legs['calibration_set_epoch']['cpu_baseline'] = {
                    'value': 5000 / dte, 'unit': UNIT, 'cores': cores, 'kind': host.kind, 'us_per_epoch': dte * 1e6,
                    'sample': '20 full-batch steps on 5,000 samples (no evaluation pass), ' + host.what()}

    if rank == 0:
        def g(v):
            return None if v is None else round(v / G, 4)
        tr_ = legs.get('train_step', {})
        summary = {
            'unit': 'G samples/s', 'n_gpus': world,
            'fwd_bf16': g(world * n_head / (fwd_ms['bf16'] * 1e-3)) if 'bf16' in fwd_ms else None,
            'fwd_fp32': g(world * n_head / (fwd_ms['fp32'] * 1e-3)) if 'fp32' in fwd_ms else None,
            'e2e_bf16': g(e2e_by.get('bf16')), 'e2e_fp32': g(e2e_by.get('fp32')), 'e2e_frac_of_pcie_ceiling': round(e2e_by[head_prec] / ceiling, 3),
            'train_bf16': g(tr_.get('bf16', {}).get('value')), 'train_fp32': g(tr_.get('fp32', {}).get('value')),
            'train_streamed_bf16': g(legs.get('train_step_host_streamed', {}).get('bf16', {}).get('value')),
            'c5_job': g(legs.get('c5_job', {}).get(head_prec, {}).get('value')),
            'c5_ece': legs.get('c5_job', {}).get(head_prec, {}).get('ece'),
            'c4_fwd_bf16': g(legs.get('c4_forward', {}).get('value')),
            'fwd_bf16_hidden_128_128': g(legs.get('two_hidden_layers', {}).get('value')),
            'fwd_fp32_hidden_5_5': g(legs.get('default_conditioner', {}).get('forward', {}).get('value')),
            'train_fp32_hidden_5_5': g(legs.get('default_conditioner', {}).get('train_step', {}).get('value')),
            'tensor_frac': round(ach_tf / tf_peak, 4) if head_prec == 'bf16' else None,
            'dp_check_ok': None if dp_check is None else dp_check['ok'],
        }
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'bf16' if head_prec == 'bf16' else 'f32', 'data': 'synthetic',
                'config': {'workload': WORKLOAD if n_head == N_HEAD else WORKLOAD.replace('100,000,000', str(n_head)),
                           'l2': 'inputs larger than L2: %.1f GB of logits per launch' % (n_head * K * 4 / 1e9),
                           'precision_path': ('bf16 tcgen05 kernel (stated tolerance 1e-2 on z / log-det / probabilities, '
                                              'measured 4e-4 / 2e-3 / 7e-4); the fp32 kernel (API default, 1e-5) on the '
                                              'same workload is summary.fwd_fp32 / e2e_fp32') if head_prec == 'bf16'
                           else 'fp32 CUDA-core kernel (1e-5 of the reference)',
                           'weights': 'reference init x300 (trained-like), seed 1',
                           'cpu_affinity': ('%d GPU-local CPUs per rank' % numa) if numa else 'default'},
                'clocks': clocks.summary(), 'e2e': e2e, 'gpu_launches': launches, 'roofline': roof,
                'cpu_baseline': cpu, 'legs': legs, 'dp_check': dp_check, 'summary': summary}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def _nice_flat():
    """Reference init of the C1 NICE flow (K=3, 4 additive couplings, hidden 32): t-nets only."""
    import torch
    torch.manual_seed(5)
    chunks = []
    for _ in range(4):
        for fan_in, fan_out in ((3, 32), (32, 3)):
            lin = torch.nn.Linear(fan_in, fan_out)
            chunks += [lin.weight.detach().reshape(-1) * 0.001, lin.bias.detach().reshape(-1) * 0.001]
    return torch.cat(chunks)


_JSON_OUT = None


def emit(line):
    """The one JSON line goes to the process's original stdout; everything else that libraries
    write to fd 1 (NCCL prints its version banner there) is redirected to stderr in main()."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def main():
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default='auto', choices=['auto', 'fp32', 'bf16'])
    ap.add_argument('--head-n', type=int, default=0, help='samples per GPU per headline step (default 10^8)')
    ap.add_argument('--train-n', type=int, default=0, help='training samples per GPU (default: 64 Mi / n_gpus, config C3)')
    ap.add_argument('--c5-n', type=int, default=0, help='total samples of the C5 job (default 10^8)')
    ap.add_argument('--c4-n', type=int, default=0, help='samples per GPU of the C4 forward leg (default 10^7)')
    ap.add_argument('--no-train', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-extra', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == 'reference':
        return run_reference(args)
    if args.gpus > 1 and 'WORLD_SIZE' not in os.environ:
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(args.gpus),
               '--master-addr', '127.0.0.1', '--master-port', '29511', os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd, stdout=_JSON_OUT.fileno()))
    run_ours(args)


if __name__ == '__main__':
    main()
