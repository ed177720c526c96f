#!/usr/bin/env python
"""Benchmark of the coupling-flow hot path (see BASELINE.json, DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Own arm.  Workload at every N: BASELINE config C2 -- RealNVP affine flow, K=10 classes, 6 coupling
layers, hidden 128, fused forward + log-det over a batch of 1,000,000 synthetic logits per GPU per
step (weak scaling: samples shard with no collective).  `value` = samples/s with inputs resident in
HBM; `e2e` = the same pass through the public host-buffer API (pinned host logits -> H2D -> kernel ->
D2H of z and log-det inside the timed region).  Extra keys: `roofline` (dominant kernel, CUDA
events), `cpu_baseline` (the reference's arithmetic on the box's host cores), `train_step`
(config C3 shape: fused NLL forward+backward+Adam, NCCL all-reduce of the flat gradient at N>1).

Reference arm (--impl reference): the reference's own CPU implementation of the same pass
(oracle/ref_port_torch.py: the reference's torch op sequence; /root/reference does not exist on
the GPU box) on all host threads, rank 0 only.
"""
import argparse
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
N_STEP = 1_000_000
N_ROT = 8                       # rotating input batches: 8 x 84 MB in+out > 126 MB L2
BYTES_PER_SAMPLE = 8 * K + 4    # read x, write z, write log-det
FLOPS_PER_SAMPLE = L * 2 * 2 * (5 * 128 + 128 * 5)   # minimal, mask-exploiting
WORKLOAD = 'C2: RealNVP K=10 L=6 hidden=[128], fused fwd+logdet, N=1,000,000 synthetic logits per GPU per step'


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return p['hbm_gbs'], p['bf16_tflops'], 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 1590.0, 'fallback (B200_PROFILING.md)'


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


def make_weights(seed=1):
    """Reference init (nn.Linear default x 0.001, flows/flows.py:76-79) x 300 = 'trained-like' scale."""
    import torch
    import cnf_b200
    torch.manual_seed(seed)
    model = cnf_b200.RealNvpFlow(K, layers=L, hidden_size=HIDDEN)
    with torch.no_grad():
        for p in model.parameters():
            if p.requires_grad:
                p.mul_(300.0)
    return model


def synth(n, seed, device=None):
    """Over-confident 80%-accurate synthetic classifier logits, row-centred."""
    import torch
    g = torch.Generator(device='cpu').manual_seed(seed)
    y = torch.randint(0, K, (n,), generator=g)
    x = 1.5 * torch.randn((n, K), generator=g)
    hit = (torch.rand(n, generator=g) < 0.8).float()
    x[torch.arange(n), y] += 3.0 * hit
    x -= x.mean(dim=1, keepdim=True)
    if device is not None:
        return x.to(device), y.to(device)
    return x, y


def cpu_reference_pass(flat, x, reps, warm=1):
    """Seconds per forward+log-det pass of the reference's torch-CPU arithmetic on x."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_port_torch as rp
    layers = rp.unflatten(flat, K, L, HIDDEN)
    times = []
    with torch.no_grad():
        for i in range(warm + reps):
            t0 = time.perf_counter()
            zs, ld = rp.forward(layers, x)
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
    times.sort()
    return times[len(times) // 2], times


def run_reference(args):
    import torch
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    model = make_weights()
    flat = torch.cat([p.detach().reshape(-1) for lay in model.layers for p in lay.canonical_parameters()])
    # size the per-step sample so the whole run stays within ~2 minutes
    probe, _ = synth(100_000, 7)
    t_probe, _ = cpu_reference_pass(flat, probe, reps=1, warm=1)
    budget = 120.0 / max(1, args.steps + args.warmup)
    n = int(min(N_STEP, max(50_000, 100_000 * budget / max(t_probe, 1e-6))))
    x, _ = synth(n, 11)
    med, times = cpu_reference_pass(flat, x, reps=args.steps, warm=args.warmup)
    total = sum(times)
    value = n * len(times) / total
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'sample_per_step': n, 'device': 'cpu'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
                         'sample': '%d samples per step, %d steps, torch CPU ops restating flows/flows.py:101-112 '
                                   '(oracle/ref_port_torch.py)' % (n, len(times))},
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
    import torch
    import torch.distributed as dist
    import cnf_b200  # noqa: F401  (fails loudly if the CUDA library is missing)

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

    model = make_weights().to(dev)
    eng = model.engine()
    precision = args.precision
    if precision == 'auto':
        precision = 'bf16' if eng.tc_bytes > 0 else 'fp32'
    eng.ensure(dev)
    eng.pack(tc=(precision == 'bf16'))

    # ---- kernel-resident measurement: inputs already in HBM ---------------------------------
    xs = [synth(N_STEP, 100 + rank * N_ROT + i, dev)[0] for i in range(N_ROT)]
    import ctypes
    from cnf_b200 import _lib
    from cnf_b200._engine import _ptr, _stream
    z = torch.empty_like(xs[0])
    ld = torch.empty(N_STEP, dtype=torch.float32, device=dev)
    desc = eng.desc_tc if precision == 'bf16' else eng.desc
    packed = eng.packed_tc if precision == 'bf16' else eng.packed

    def step(i):
        _lib.call('cnf_flow_forward', ctypes.byref(desc), _ptr(packed), _ptr(eng.tables), _ptr(xs[i % N_ROT]),
                  _ptr(z), _ptr(ld), None, ctypes.c_int64(N_STEP), _stream(dev))

    for i in range(args.warmup):
        step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    value = world * N_STEP * args.steps / (ms * 1e-3)
    kern_ms = e0.elapsed_time(e1) / args.steps        # one kernel per step on this stream
    launches = args.steps

    # ---- end to end through the public host-buffer API ---------------------------------------
    xh = torch.empty((N_STEP, K), dtype=torch.float32, pin_memory=True)
    xh.copy_(synth(N_STEP, 999 + rank)[0])
    zh = torch.empty((N_STEP, K), dtype=torch.float32, pin_memory=True)
    lh = torch.empty(N_STEP, dtype=torch.float32, pin_memory=True)
    model.flow.precision = precision
    for _ in range(max(3, args.warmup)):
        model.transform_host(xh, zh, lh, device=dev)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        model.transform_host(xh, zh, lh, device=dev)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = {'value': world * N_STEP * args.steps / e2e_s, 'unit': UNIT,
           'h2d_bytes_per_step': xh.numel() * 4, 'd2h_bytes_per_step': zh.numel() * 4 + lh.numel() * 4,
           'api': 'RealNvpFlow.transform_host (pinned host logits in, host z + log-det out)'}

    # ---- training step, config C3 shape: tensor-core path (headline) and fp32 path -------------
    train = train_fp32 = None
    if not args.no_train:
        n_loc = args.train_n or (64 * (1 << 20)) // world
        xt, yt = synth(n_loc, 5000 + rank, dev)

        def time_training(prec):
            tmodel = make_weights(seed=2)
            with torch.no_grad():
                for p in tmodel.parameters():
                    if p.requires_grad:
                        p.mul_(1.0 / 300.0)          # reference init for training
            tmodel.to(dev)
            tr = cnf_b200.FusedNLLTrainer(tmodel.engine(), xt, yt, n_total=n_loc * world, precision=prec)
            for _ in range(3):
                tr.step()
            barrier()
            t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0e.record()
            tsteps = max(3, min(args.steps, 5))
            for _ in range(tsteps):
                tr.step()
            t1e.record()
            barrier()
            tms = max_over_ranks(t0e.elapsed_time(t1e))
            kern = ('tcgen05 forward (+tape) and backward kernels per 2^20-sample chunk' if prec == 'bf16'
                    else 'fused NLL fwd+bwd kernel')
            return {'value': world * n_loc * tsteps / (tms * 1e-3), 'unit': UNIT, 'ms_per_step': tms / tsteps,
                    'samples_per_gpu': n_loc, 'samples_total': n_loc * world,
                    'scaling': 'strong (C3: 64 Mi samples over all GPUs)' if not args.train_n else 'weak',
                    'steps': tsteps, 'dtype': 'bf16' if prec == 'bf16' else 'f32',
                    'what': kern + ' + grad reduce' + (' + NCCL all-reduce' if world > 1 else '') +
                            ' + Adam + repack per step', 'loss': -float(tr.loss_acc[0]) / (n_loc * world)}

        if precision == 'bf16':
            train = time_training('bf16')
            train['tensor_tflops_minimal'] = train['value'] * 4 * FLOPS_PER_SAMPLE / 1e12   # fwd + recompute + dgrad + wgrad
        train_fp32 = time_training('fp32')
        if train is None:
            train = train_fp32
        del xt, yt

    # ---- config C5 pieces: inverse pass and the metrics kernel (HBM-bound) -------------------
    extra = None
    if not args.no_extra:
        from cnf_b200.utils import metrics as M
        for i in range(3):
            _lib.call('cnf_flow_inverse', ctypes.byref(desc), _ptr(packed), _ptr(eng.tables), _ptr(xs[i % N_ROT]),
                      _ptr(z), _ptr(ld), None, ctypes.c_int64(N_STEP), _stream(dev))
        barrier()
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        i0.record()
        for i in range(20):
            _lib.call('cnf_flow_inverse', ctypes.byref(desc), _ptr(packed), _ptr(eng.tables), _ptr(xs[i % N_ROT]),
                      _ptr(z), _ptr(ld), None, ctypes.c_int64(N_STEP), _stream(dev))
        i1.record()
        barrier()
        inv_ms = max_over_ranks(i0.elapsed_time(i1)) / 20
        n_m = 12_500_000                      # 10^8 samples over 8 GPUs (config C5)
        pm = torch.softmax(synth(n_m, 31 + rank, dev)[0], dim=1).contiguous()
        ym = synth(n_m, 31 + rank)[1].to(dev)
        edges = torch.from_numpy(M.bin_edges(15)).to(dev)
        macc = torch.zeros(48, dtype=torch.float64, device=dev)

        def mstep():
            _lib.call('cnf_metrics', _ptr(pm), ctypes.c_int32(0), _ptr(ym), ctypes.c_int64(n_m), ctypes.c_int32(K),
                      ctypes.c_int32(15), ctypes.c_int32(0), None, _ptr(edges), _ptr(macc), _stream(dev))
        for _ in range(3):
            mstep()
        barrier()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record()
        for _ in range(20):
            mstep()
        m1.record()
        barrier()
        met_ms = max_over_ranks(m0.elapsed_time(m1)) / 20
        # AffineConstantLayer / TempScaler streaming kernel: pure 8K B/sample
        n_a = 10_000_000
        xa = synth(n_a, 41 + rank, dev)[0]
        za = torch.empty_like(xa)
        sa = torch.full((K,), 0.1, device=dev)
        ta = torch.full((K,), -0.2, device=dev)

        def astep():
            _lib.call('cnf_affine_const', _ptr(xa), _ptr(sa), _ptr(ta), _ptr(za), ctypes.c_int64(n_a), ctypes.c_int32(K),
                      ctypes.c_int32(0), _stream(dev))
        for _ in range(3):
            astep()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(20):
            astep()
        a1.record()
        barrier()
        aff_ms = max_over_ranks(a0.elapsed_time(a1)) / 20
        del xa, za
        hbm_peak = peaks()[0]
        ab = 8 * K * n_a / (aff_ms * 1e-3) / 1e9
        mb = (4 * K + 8) * n_m / (met_ms * 1e-3) / 1e9
        extra = {'affine_const': {'value': world * n_a / (aff_ms * 1e-3), 'unit': UNIT, 'ms_per_step': aff_ms,
                                  'samples_per_gpu': n_a,
                                  'what': 'AffineConstantLayer.forward z = x*exp(s)+t (flows/flows.py:53-58), fp32',
                                  'roofline': {'bound': 'hbm', 'achieved': ab, 'peak': hbm_peak, 'unit': 'GB/s',
                                               'frac': ab / hbm_peak, 'note': 'algorithmic bytes 8K = 80 B/sample'}},
                 'inverse': {'value': world * N_STEP / (inv_ms * 1e-3), 'unit': UNIT, 'ms_per_step': inv_ms,
                             'what': 'Flow.backward (inverse + log-det) on 1,000,000 samples per GPU, same path as value'},
                 'metrics': {'value': world * n_m / (met_ms * 1e-3), 'unit': UNIT, 'ms_per_step': met_ms,
                             'samples_per_gpu': n_m, 'what': 'ECE(15 bins)+NLL+accuracy in one pass over fp32 probabilities',
                             'roofline': {'bound': 'hbm', 'achieved': mb, 'peak': hbm_peak, 'unit': 'GB/s',
                                          'frac': mb / hbm_peak, 'note': 'algorithmic bytes 4K+8 = 48 B/sample'}}}
        del pm, ym
        # config C4: K=100, 8 couplings, hidden 512 on the streamed-weight tensor-core kernel
        torch.manual_seed(4)
        c4 = cnf_b200.RealNvpFlow(100, layers=8, hidden_size=[512])
        with torch.no_grad():
            for prm in c4.parameters():
                if prm.requires_grad:
                    prm.mul_(60.0)
        c4.to(dev)
        e4 = c4.engine()
        e4.ensure(dev)
        e4.pack(tc=True)
        n4 = 1_000_000
        g4 = torch.Generator(device=dev).manual_seed(40 + rank)
        x4 = 1.5 * torch.randn((n4, 100), generator=g4, device=dev)
        x4 -= x4.mean(dim=1, keepdim=True)
        z4 = torch.empty_like(x4)
        l4 = torch.empty(n4, dtype=torch.float32, device=dev)

        def c4step():
            _lib.call('cnf_flow_forward', ctypes.byref(e4.desc_tc), _ptr(e4.packed_tc), _ptr(e4.tables), _ptr(x4),
                      _ptr(z4), _ptr(l4), None, ctypes.c_int64(n4), _stream(dev))
        for _ in range(2):
            c4step()
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(5):
            c4step()
        c1.record()
        barrier()
        c4_ms = max_over_ranks(c0.elapsed_time(c1)) / 5
        c4_flops = 8 * 2 * 2 * (50 * 512 + 512 * 50)
        extra['c4_forward'] = {'value': world * n4 / (c4_ms * 1e-3), 'unit': UNIT, 'ms_per_step': c4_ms, 'dtype': 'bf16',
                               'what': 'C4: RealNVP K=100 L=8 hidden=[512] fwd+logdet, 1,000,000 samples per GPU, '
                                       'streamed-weight tcgen05 kernel',
                               'tensor_tflops_minimal': c4_flops * n4 / (c4_ms * 1e-3) / 1e12,
                               'frac_of_bf16_peak': c4_flops * n4 / (c4_ms * 1e-3) / 1e12 / peaks()[1]}
        # the same shape through the fp32 kernels (the 1e-5 parity path and the only training path for wide
        # shapes): forward on 200,000 samples, full NLL + Adam training step on 100,000
        n4f = 200_000
        c4.flow.precision = 'fp32'
        with torch.no_grad():
            c4(x4[:n4f])
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(3):
                c4(x4[:n4f])
            f1.record()
        barrier()
        c4f_ms = max_over_ranks(f0.elapsed_time(f1)) / 3
        n4t = 100_000
        y4 = torch.randint(0, 100, (n4t,), generator=g4, device=dev)
        tr4 = cnf_b200.FusedNLLTrainer(e4, x4[:n4t].contiguous(), y4, n_total=n4t * world)
        tr4.step()
        t40, t41 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t40.record()
        for _ in range(3):
            tr4.step()
        t41.record()
        barrier()
        c4t_ms = max_over_ranks(t40.elapsed_time(t41)) / 3
        extra['c4_fp32'] = {'forward': {'value': world * n4f / (c4f_ms * 1e-3), 'unit': UNIT, 'ms_per_step': c4f_ms},
                            'train_step': {'value': world * n4t / (c4t_ms * 1e-3), 'unit': UNIT, 'ms_per_step': c4t_ms},
                            'dtype': 'f32',
                            'what': 'C4 shape on the lean fp32 kernels (16-unit hidden chunks, chunk weights staged in '
                                    'shared memory, no tape): forward on 200,000 samples, NLL + Adam step on 100,000'}
        del x4, z4, l4, tr4
        # config C1 through the calibrator API (rank 0, N=1 only): NICE, K=3, N=10,000, 4 additive couplings,
        # hidden 32, fit (50 full-batch epochs) + predict, host numpy in and out (calibrators.py:241-353)
        if rank == 0 and world == 1:
            import numpy as np
            rs = np.random.RandomState(3)
            y1 = rs.randint(0, 3, size=10_000)
            x1 = (1.5 * rs.randn(10_000, 3)).astype(np.float32)
            x1[np.arange(10_000), y1] += 3.0 * (rs.rand(10_000) < 0.8)
            t1 = np.eye(3, dtype=np.float32)[y1]

            def c1_run():
                cal = cnf_b200.TorchFlowCalibrator(cnf_b200.NiceFlow, x1, t1, layers=4, hidden_size=[32], epochs=50,
                                                   dev=dev)
                return cal.predict(x1)
            c1_run()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            p1 = c1_run()
            torch.cuda.synchronize()
            c1_s = time.perf_counter() - t0
            extra['c1_calibrator'] = {'value': 10_000 * 50 / c1_s, 'unit': UNIT, 'wall_s': c1_s, 'dtype': 'f32',
                                      'what': 'C1: TorchFlowCalibrator(NiceFlow, K=3, N=10,000, 4 couplings, hidden 32): '
                                              'fit 50 full-batch epochs + predict, host numpy in/out; value = '
                                              'training samples per second of wall time',
                                      'finite': bool(np.isfinite(p1).all())}

        # C2 shape at a real calibration-set size: one full-batch epoch (optimiser step + evaluation pass) at
        # N = 5,000, the regime the reference's calibrators run in (calibrators.py:267-328)
        if rank == 0 and world == 1:
            xe, ye = synth(5000, 77, dev)
            ep = {}
            for prec in (('fp32', 'bf16') if precision == 'bf16' else ('fp32',)):
                me = make_weights(seed=2).to(dev)
                tre = cnf_b200.FusedNLLTrainer(me.engine(), xe, ye, precision=prec)
                for _ in range(5):
                    tre.step(); tre.evaluate()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(200):
                    tre.step(); tre.evaluate()
                torch.cuda.synchronize()
                ep[prec] = (time.perf_counter() - t0) / 200
            extra['calibration_set_epoch'] = {
                'value': 5000 / ep['fp32'], 'unit': UNIT, 'us_per_epoch': {k: v * 1e6 for k, v in ep.items()},
                'dtype': 'f32', 'what': 'C2 shape, N=5,000: full-batch Adam step + evaluation pass per epoch, wall time '
                                        'over 200 epochs; value = samples per second on the fp32 (1e-5 parity) kernels'}

    hbm, tf, src = peaks()
    ach = BYTES_PER_SAMPLE * N_STEP / (kern_ms * 1e-3) / 1e9
    kname = 'flow_tc_kernel' if precision == 'bf16' else 'flow_apply_kernel'
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):       # dram__bytes_read.sum + dram__bytes_write.sum per launch, from ncu --set full
        traffic = json.load(open(tpath)).get(kname)
    roof = {'bound': 'hbm', 'achieved': ach, 'peak': hbm, 'unit': 'GB/s', 'frac': ach / hbm, 'traffic': traffic,
            'kernel': kname, 'peak_source': src,
            'note': 'algorithmic bytes 84 B/sample x 1e6 samples per launch'}
    ach_tf = FLOPS_PER_SAMPLE * N_STEP / (kern_ms * 1e-3) / 1e12
    roof_tensor = {'bound': 'tensor', 'achieved': ach_tf, 'peak': tf, 'unit': 'TFLOP/s', 'frac': ach_tf / tf,
                   'note': 'minimal (mask-exploiting) 30720 flop/sample; peak = measured bf16 burst'}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        torch.set_num_threads(os.cpu_count() or 1)
        flat = eng.flat.detach().cpu()
        xc, _ = synth(N_STEP, 11)
        med, times = cpu_reference_pass(flat, xc, reps=3, warm=1)
        cpu = {'value': N_STEP / med, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
               'sample': '1,000,000 samples x 3 passes (median), torch CPU ops restating flows/flows.py:101-112'}
        # the other legs of SURVEY.md 8(d) on the same host cores, bounded samples
        import ref_port_torch as rp
        import flow_oracle as orc
        cores = torch.get_num_threads()
        if train is not None:
            n_t = 200_000
            xt_c, yt_c = synth(n_t, 13)
            st = rp.TrainState(flat / 300.0, K, L, HIDDEN)
            st.step(xt_c, yt_c)
            t0 = time.perf_counter()
            for _ in range(2):
                st.step(xt_c, yt_c)
            dt = (time.perf_counter() - t0) / 2
            cb = {'value': n_t / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                  'sample': '200,000 samples x 2 full-batch steps: torch autograd + torch.optim.Adam on the '
                            'reference op sequence (calibrators.py:287-295)'}
            train['cpu_baseline'] = cb
            if train_fp32 is not None:
                train_fp32['cpu_baseline'] = cb
        if extra is not None:
            layers = rp.unflatten(flat, K, L, HIDDEN)
            with torch.no_grad():
                rp.inverse(layers, xc[:200_000])
                t0 = time.perf_counter()
                rp.inverse(layers, xc)
                dt = time.perf_counter() - t0
            extra['inverse']['cpu_baseline'] = {'value': N_STEP / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                                'sample': '1,000,000 samples, one pass, flows/flows.py:114-126 in torch CPU ops'}
            if 'c1_calibrator' in extra:
                g1 = torch.Generator().manual_seed(5)
                xs1 = 1.5 * torch.randn(10_000, 3, generator=g1)
                ys1 = torch.randint(0, 3, (10_000,), generator=g1)
                n1 = 4 * (32 * 3 + 32 + 3 * 32 + 3)
                st1 = rp.TrainState(0.001 * torch.randn(n1, generator=g1), 3, 4, [32], scale=False, shift=True)
                st1.step(xs1, ys1)
                t0 = time.perf_counter()
                for _ in range(50):
                    st1.step(xs1, ys1)
                dt1 = time.perf_counter() - t0
                extra['c1_calibrator']['cpu_baseline'] = {
                    'value': 10_000 * 50 / dt1, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                    'sample': '50 full-batch steps on 10,000 samples: torch autograd + Adam on the reference op '
                              'sequence WITHOUT its DataLoader (calibrators.py:268-283 collates per sample, which '
                              'dominates the reference at this size, SURVEY.md 6)'}
            if 'calibration_set_epoch' in extra:
                ge = torch.Generator().manual_seed(7)
                xse = 1.5 * torch.randn(5000, 10, generator=ge)
                yse = torch.randint(0, 10, (5000,), generator=ge)
                npe = 6 * 2 * (128 * 10 + 128 + 10 * 128 + 10)
                ste = rp.TrainState(0.001 * torch.randn(npe, generator=ge), 10, 6, [128])
                ste.step(xse, yse)
                t0 = time.perf_counter()
                for _ in range(20):
                    ste.step(xse, yse)
                dte = (time.perf_counter() - t0) / 20
                extra['calibration_set_epoch']['cpu_baseline'] = {
                    'value': 5000 / dte, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'us_per_epoch': dte * 1e6,
                    'sample': '20 full-batch steps on 5,000 samples (no evaluation pass): torch autograd + Adam on the '
                              'reference op sequence'}
            if 'c4_fp32' in extra:
                g4c = torch.Generator().manual_seed(6)
                n4c = 5_000
                x4c = 1.5 * torch.randn(n4c, 100, generator=g4c)
                y4c = torch.randint(0, 100, (n4c,), generator=g4c)
                n4p = 8 * 2 * (512 * 100 + 512 + 100 * 512 + 100)
                st4 = rp.TrainState(0.001 * torch.randn(n4p, generator=g4c), 100, 8, [512])
                st4.step(x4c, y4c)
                t0 = time.perf_counter()
                for _ in range(2):
                    st4.step(x4c, y4c)
                dt4 = (time.perf_counter() - t0) / 2
                extra['c4_fp32']['train_step']['cpu_baseline'] = {
                    'value': n4c / dt4, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                    'sample': '5,000 samples x 2 full-batch steps, torch autograd + Adam on the reference op sequence'}
            n_mc = 2_000_000
            xm, ym_c = synth(n_mc, 17)
            pm_c = torch.softmax(xm, dim=1).numpy()
            ym_n = ym_c.numpy()
            t0 = time.perf_counter()
            orc.expected_calibration_error(pm_c, ym_n, 15)
            orc.neg_log_likelihood(pm_c, ym_n)
            orc.accuracy(pm_c, ym_n)
            dt = time.perf_counter() - t0
            extra['metrics']['cpu_baseline'] = {'value': n_mc / dt, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                                                'sample': '2,000,000 samples: numpy restatement of utils/metrics.py:35-73, 6-15, 76-80 '
                                                          '(15 masked passes + one-hot NLL + argmax accuracy)'}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'bf16' if precision == 'bf16' else 'f32', 'data': 'synthetic',
                'config': {'workload': WORKLOAD, 'l2': 'inputs larger than L2: %d rotating batches' % N_ROT,
                           'precision_path': precision, 'weights': 'reference init x300 (trained-like), seed 1',
                           'cpu_affinity': ('%d GPU-local CPUs per rank' % numa) if numa else 'default'},
                'clocks': clocks.summary(), 'e2e': e2e, 'gpu_launches': launches, 'roofline': roof,
                'roofline_tensor': roof_tensor, 'cpu_baseline': cpu, 'train_step': train, 'train_step_fp32': train_fp32, 'c5': extra}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default='auto', choices=['auto', 'fp32', 'bf16'])
    ap.add_argument('--train-n', type=int, default=0, help='training samples per GPU (default: 64 Mi / n_gpus, config C3)')
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
