"""Host-side engine: owns the device buffers of one fused coupling stack and drives
libcnf_b200 through the C ABI.  torch is used for device memory, streams and autograd
plumbing only; every arithmetic step is a kernel from the library.

Buffers (all on the flow's CUDA device):
  flat        float32 [n_flat]     canonical parameter vector; the nn.Parameters of the stack
                                   are views into it, so load_state_dict / .to() / optimisers
                                   keep working (reference state_dict keys are unchanged)
  gather      int32   [n_packed]   packed <- flat index map (host-planned, csrc/cnf_plan.cpp)
  tables      int32   [n_tables]   physical-slot index tables (flips / permutations folded)
  packed      float32 [n_packed]   kernel-side weights, refreshed by cnf_pack_weights
  packed_tc   bf16 blob            tensor-core kernel weights (only when supported)
  partials    float32 [rows, n_packed]  per-CTA gradient partials (allocated on first use)
"""
import ctypes

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, what='input'):
    if not t.is_cuda:
        raise RuntimeError('cnf_b200: %s must be a CUDA tensor (got %s). The flow kernels are '
                           'sm_100a CUDA only; there is no CPU fallback.' % (what, t.device))


def on_device(device):
    """Context manager: make `device` the current CUDA device for the ctypes calls inside.  libcnf
    launches on the caller's stream, but cudaFuncSetAttribute / cudaGetDevice and the launch itself act
    on the *current* device, so every call into the library runs under this guard (a flow may live on
    cuda:1 while the process's current device is cuda:0, as the reference's ``dev=`` argument allows)."""
    return torch.cuda.device(device)


def _guarded(fn):
    """Run an engine method with the engine's device current (no-op when it already is)."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        dev = self.device
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(self, *args, **kwargs)
        with torch.cuda.device(dev):
            return fn(self, *args, **kwargs)
    return wrapper


def check_logits(x, K, what='x', cast=True):
    """[N, K] float32 contiguous view of `x` for the C ABI (which reads N*K floats blindly).  Wrong rank or
    column count raises ValueError like the reference's matmul would; other float dtypes are cast when
    `cast`, else rejected."""
    if x.dim() != 2 or x.shape[1] != K:
        raise ValueError('cnf_b200: %s must have shape [N, %d], got %s' % (what, K, tuple(x.shape)))
    if x.dtype != torch.float32:
        if not cast or not x.dtype.is_floating_point:
            raise ValueError('cnf_b200: %s must be float32, got %s' % (what, x.dtype))
        x = x.to(torch.float32)
    return x.contiguous()


def check_labels(y, N, device, what='y'):
    """int64 [N] contiguous labels on `device` (the kernels read 8 bytes per sample)."""
    if y.dim() != 1 or y.shape[0] != N:
        raise ValueError('cnf_b200: %s must have shape [%d], got %s' % (what, N, tuple(y.shape)))
    if y.dtype != torch.int64:
        if y.dtype.is_floating_point or y.dtype == torch.bool:
            raise ValueError('cnf_b200: %s must be an integer tensor, got %s' % (what, y.dtype))
        y = y.to(torch.int64)
    if y.device != device:
        raise ValueError('cnf_b200: %s lives on %s, the samples on %s' % (what, y.device, device))
    return y.contiguous()


class StackEngine:
    """One homogeneous stack of coupling layers (same K, hidden sizes, scale/shift flags)."""

    def __init__(self, K, hidden, scale, shift, perms, param_lists):
        """param_lists: per layer, the list of nn.Parameters in canonical order
        (s-net then t-net; per Linear weight then bias)."""
        self.K, self.hidden, self.scale, self.shift = int(K), [int(h) for h in hidden], bool(scale), bool(shift)
        self.L = len(param_lists)
        self.perms = perms
        self.params = [p for lay in param_lists for p in lay]
        self.desc, self._keep = _lib.make_desc(self.K, self.L, self.hidden, scale, shift, _lib.PREC_FP32, perms)
        self.desc_tc, self._keep_tc = _lib.make_desc(self.K, self.L, self.hidden, scale, shift, _lib.PREC_BF16_TC, perms)
        info = _lib.PlanInfo()
        _lib.call('cnf_plan_info_get', ctypes.byref(self.desc), ctypes.byref(info))
        self.info = info
        self.n_flat, self.n_packed = int(info.n_flat), int(info.n_packed)
        self.n_rows = int(info.n_grad_rows)
        self.tc_bytes = int(info.tc_bytes)
        assert self.n_flat == sum(p.numel() for p in self.params), 'parameter list does not match the plan'
        gather = np.empty(self.n_packed, dtype=np.int32)
        tables = np.empty(int(info.n_tables), dtype=np.int32)
        _lib.call('cnf_plan_build', ctypes.byref(self.desc), gather.ctypes.data_as(ctypes.c_void_p),
                  tables.ctypes.data_as(ctypes.c_void_p))
        self._gather_host, self._tables_host = gather, tables
        self._gather_tc_host = None
        if self.tc_bytes > 0:
            n_g = ctypes.c_int64(0)
            _lib.call('cnf_tc_gather_len', ctypes.byref(self.desc_tc), ctypes.byref(n_g))
            g = np.empty(int(n_g.value), dtype=np.int32)
            _lib.call('cnf_plan_build_tc', ctypes.byref(self.desc_tc), g.ctypes.data_as(ctypes.c_void_p))
            self._gather_tc_host = g
        # tensor-core training (cnf_flow_tcb.cu): partial-row length, row count, workspace per sample
        self.tc_train = None
        if self.tc_bytes > 0:
            ng, rows, wsb = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
            _lib.call('cnf_tc_train_info', ctypes.byref(self.desc_tc), ctypes.byref(ng), ctypes.byref(rows),
                      ctypes.byref(wsb))
            if ng.value > 0:
                g = np.empty(int(ng.value), dtype=np.int32)
                _lib.call('cnf_plan_build_tcgrad', ctypes.byref(self.desc_tc), g.ctypes.data_as(ctypes.c_void_p))
                self.tc_train = (int(ng.value), int(rows.value), int(wsb.value), g)
        self.device = None
        self.flat = None
        self.partials = None
        self.flat_grad = None
        self.adam_m = self.adam_v = None
        self.adam_t = 0
        self.version = 0        # bumped whenever the library changes the weights (torch's own counter does not see that)

    def weights_version(self):
        """Identity of the current weights: changes on every update through this engine or through torch."""
        if self.flat is None:
            return None
        # the nn.Parameters are views into `flat` with version counters of their own: torch optimisers and in-place
        # edits bump those, the fused optimiser kernels bump self.version
        return (self.version, self.flat.data_ptr(), sum(p._version for p in self.params))

    # ------------------------------------------------------------------ buffers
    def _bind(self, device):
        """(Re)create device buffers on `device` and alias the parameters into `flat`."""
        with torch.no_grad():
            vals = [p.detach().to(device=device, dtype=torch.float32).reshape(-1) for p in self.params]
            self.flat = torch.cat(vals) if vals else torch.zeros(0, device=device)
            off = 0
            for p in self.params:
                n = p.numel()
                p.data = self.flat[off:off + n].view(p.shape)
                off += n
        self.device = device
        self.version += 1
        self.gather = torch.from_numpy(self._gather_host).to(device)
        live = self._gather_host[self._gather_host >= 0]
        # the one-launch optimiser tail (reduce_adam_pack) needs every live flat entry at exactly one packed position
        self.gather_one_to_one = bool(live.size == np.unique(live).size)
        self.tables = torch.from_numpy(self._tables_host).to(device)
        self.packed = torch.empty(self.n_packed, dtype=torch.float32, device=device)
        self.packed_tc = None
        if self._gather_tc_host is not None:
            self.gather_tc = torch.from_numpy(self._gather_tc_host).to(device)
            self.packed_tc = torch.empty(self.tc_bytes, dtype=torch.uint8, device=device)
        self.partials = None
        self.flat_grad = None
        self.partials_tc = self.gather_tcgrad = self._train_ws = None
        # optimiser state follows the parameters (the reference keeps it in self.optimizer)
        if self.adam_m is not None:
            self.adam_m = self.adam_m.to(device)
            self.adam_v = self.adam_v.to(device)

    def ensure(self, device):
        """Make sure the parameters still alias `flat` on `device` (module.to(), a replaced
        nn.Parameter or load_state_dict(assign=True) break the aliasing)."""
        device = torch.device(device)
        if device.type != 'cuda':
            raise RuntimeError('cnf_b200: the flow must live on a CUDA device (got %s); no CPU fallback' % device)
        ok = self.flat is not None and self.device == device
        if ok:
            base, off = self.flat.data_ptr(), 0
            for p in self.params:
                if p.data_ptr() != base + 4 * off or p.device != device:
                    ok = False
                    break
                off += p.numel()
        if not ok:
            self._bind(device)

    def pack(self, tc=False, fp32=True):
        """Refresh the kernel-side weight copies from `flat`: the fp32 blob (fp32=True) and/or the bf16
        tensor-core blob (tc=True)."""
        with on_device(self.device):
            st = _stream(self.device)
            if fp32:
                _lib.call('cnf_pack_weights', ctypes.byref(self.desc), _ptr(self.flat), _ptr(self.gather), _ptr(self.packed), st)
            if tc and self.packed_tc is not None:
                _lib.call('cnf_pack_weights_tc', ctypes.byref(self.desc_tc), _ptr(self.flat), _ptr(self.gather_tc),
                          _ptr(self.packed_tc), st)

    def _want_partials(self):
        if self.partials is None:
            self.partials = torch.empty(self.n_rows * self.n_packed, dtype=torch.float32, device=self.device)
            self.flat_grad = torch.zeros(self.n_flat, dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------ kernels
    def apply(self, x, inverse=False, want_all=False, precision='fp32', repack=True):
        """Returns (out [N,K], logdet [N], all_or_None [L,N,K])."""
        require_cuda(x)
        x = check_logits(x.detach(), self.K)
        self.ensure(x.device)
        N = x.shape[0]
        use_tc = precision == 'bf16'
        if use_tc and self.packed_tc is None:
            raise NotImplementedError('cnf_b200: the bf16 tensor-core path does not cover this flow shape '
                                      '(K=%d, hidden=%s); use precision="fp32"' % (self.K, self.hidden))
        if repack:
            self.pack(tc=use_tc)
        out = torch.empty_like(x)
        ld = torch.empty(N, dtype=torch.float32, device=x.device)
        allz = torch.empty((self.L, N, self.K), dtype=torch.float32, device=x.device) if want_all else None
        fn = 'cnf_flow_inverse' if inverse else 'cnf_flow_forward'
        desc = self.desc_tc if use_tc else self.desc
        packed = self.packed_tc if use_tc else self.packed
        with on_device(x.device):
            _lib.call(fn, ctypes.byref(desc), _ptr(packed), _ptr(self.tables), _ptr(x), _ptr(out), _ptr(ld),
                      _ptr(allz), ctypes.c_int64(N), _stream(x.device))
        return out, ld, allz

    def predict(self, x, center=False, log_priors=None, y=None, bins=15, want_z=False, want_probs=False,
                precision='fp32', repack=True):
        """One fused launch (cnf_flow_predict): [row-mean centring] -> flow forward -> calibrated probabilities
        ``softmax(log(softmax(z)+1e-7) - log_priors)`` (or the statistics of plain ``softmax(z)`` when
        `log_priors` is None) and / or the ECE-NLL-accuracy statistics against labels `y`, without z making a
        round trip through HBM.  Returns a dict with the requested entries: 'z', 'logdet' (want_z), 'probs'
        float64 (want_probs, needs log_priors), 'stats' float64 [3*bins+3] (when y is given).  Shapes outside
        the fused kernels raise NotImplementedError (callers then compose apply() + the metrics kernel)."""
        require_cuda(x)
        x = check_logits(x.detach(), self.K)
        dev = x.device
        self.ensure(dev)
        N = x.shape[0]
        use_tc = precision == 'bf16'
        if use_tc and self.packed_tc is None:
            raise NotImplementedError('cnf_b200: the bf16 tensor-core path does not cover this flow shape')
        if want_probs and log_priors is None:
            raise ValueError('cnf_b200: calibrated probabilities need log_priors')
        if repack:
            self.pack(tc=use_tc)
        mode = _lib.METRICS_CALIBRATED if log_priors is not None else _lib.METRICS_LOGITS
        lp = None
        if log_priors is not None:
            lp = torch.as_tensor(np.asarray(log_priors, dtype=np.float64)).to(dev).contiguous()
            if lp.numel() != self.K:
                raise ValueError('cnf_b200: log_priors must have %d entries' % self.K)
        z = ld = probs = acc = edges = None
        if want_z:
            z = torch.empty_like(x)
            ld = torch.empty(N, dtype=torch.float32, device=dev)
        if want_probs:
            probs = torch.empty((N, self.K), dtype=torch.float64, device=dev)
        if y is not None:
            y = check_labels(y, N, dev)
            acc = torch.zeros(3 * bins + 3, dtype=torch.float64, device=dev)
            cache = self.__dict__.setdefault('_edges', {})
            edges = cache.get((bins, dev))
            if edges is None:        # bin edges i * (1 / bins) as the reference forms them (utils/metrics.py:57, 61)
                edges = cache[(bins, dev)] = torch.tensor([i * (1. / bins) for i in range(bins + 1)], dtype=torch.float64, device=dev)
        desc = self.desc_tc if use_tc else self.desc
        packed = self.packed_tc if use_tc else self.packed
        with on_device(dev):
            try:
                _lib.call('cnf_flow_predict', ctypes.byref(desc), _ptr(packed), _ptr(self.tables), _ptr(x),
                          ctypes.c_int64(N), ctypes.c_int32(1 if center else 0), ctypes.c_int32(mode), _ptr(lp),
                          _ptr(z), _ptr(ld), _ptr(probs), _ptr(y), ctypes.c_int32(bins), _ptr(edges), _ptr(acc),
                          _stream(dev))
            except _lib.CnfError as e:
                if e.code in (-3, -4):       # CNF_E_SMEM / CNF_E_UNSUPPORTED: shape outside the fused kernels
                    raise NotImplementedError(str(e))
                raise
        out = {}
        if want_z:
            out['z'], out['logdet'] = z, ld
        if want_probs:
            out['probs'] = probs
        if acc is not None:
            out['stats'] = acc
        return out

    def apply_host(self, x_host, out_host=None, ld_host=None, inverse=False, precision='fp32',
                   chunk=None, slots=4, repack=True, device=None):
        """Host buffers in, host buffers out: the samples stream through the GPU in chunks on
        `slots` CUDA streams so that the H2D copy, the flow kernel and the D2H copy of
        neighbouring chunks overlap.  x_host: CPU float32 [N,K] (pinned memory gives full PCIe
        speed).  Returns (out_host, ld_host) as CPU tensors (pinned when allocated here)."""
        device = torch.device(device) if device is not None else (self.device or torch.device('cuda', torch.cuda.current_device()))
        x_host = torch.as_tensor(x_host, dtype=torch.float32)
        if x_host.is_cuda:
            raise RuntimeError('apply_host expects host memory')
        x_host = check_logits(x_host, self.K, 'x_host')
        N, K = x_host.shape
        if chunk is None:
            # rows per H2D -> kernel -> D2H stage: 2^17 keeps a 10^6-row call pipelined over 4 streams, 2^19 is the
            # measured optimum once a call has >= 2^23 rows (1.00 vs 0.81 G samples/s at 10^7 rows on one B200)
            chunk = min(1 << 19, max(1 << 17, N // 16))
        if out_host is None:
            out_host = torch.empty((N, K), dtype=torch.float32, pin_memory=True)
        if ld_host is None:
            ld_host = torch.empty(N, dtype=torch.float32, pin_memory=True)
        for t, shape, what in ((out_host, (N, K), 'out_host'), (ld_host, (N,), 'ld_host')):
            if (not isinstance(t, torch.Tensor) or t.is_cuda or t.dtype != torch.float32 or tuple(t.shape) != shape
                    or not t.is_contiguous()):
                raise ValueError('cnf_b200: %s must be a contiguous float32 CPU tensor of shape %s' % (what, shape))
        self.ensure(device)
        use_tc = precision == 'bf16'
        if use_tc and self.packed_tc is None:
            raise NotImplementedError('cnf_b200: bf16 tensor-core path not available for this flow shape')
        if repack:
            self.pack(tc=use_tc)
        ws = getattr(self, '_host_ws', None)
        need = slots * chunk * (2 * K + 1) * 4
        if ws is None or ws.device != device or ws.numel() < need:
            ws = self._host_ws = torch.empty(need, dtype=torch.uint8, device=device)
        desc = self.desc_tc if use_tc else self.desc
        packed = self.packed_tc if use_tc else self.packed
        with on_device(device):
            _lib.call('cnf_flow_apply_host', ctypes.byref(desc), _ptr(packed), _ptr(self.tables), _ptr(x_host),
                      _ptr(out_host), _ptr(ld_host), ctypes.c_int64(N), ctypes.c_int32(1 if inverse else 0), _ptr(ws),
                      ctypes.c_int64(need), ctypes.c_int64(chunk), _stream(device))
        return out_host, ld_host

    def backward(self, x, g_z, g_ld, need_gx=True):
        """Generic autograd backward: returns (g_x or None, flat_grad)."""
        self.ensure(x.device)
        with on_device(x.device):
            return self._backward(x, g_z, g_ld, need_gx)

    def _backward(self, x, g_z, g_ld, need_gx):
        self._want_partials()
        x = check_logits(x, self.K, cast=False)
        N = x.shape[0]
        g_z = g_z.to(torch.float32).contiguous()
        g_ld = g_ld.to(torch.float32).contiguous()
        if g_z.shape != x.shape or g_ld.numel() != N:
            raise ValueError('cnf_b200: upstream gradients of shape %s / %s do not match the output [%d, %d] / [%d]'
                             % (tuple(g_z.shape), tuple(g_ld.shape), N, self.K, N))
        gx = torch.empty_like(x) if need_gx else None
        st = _stream(x.device)
        used = ctypes.c_int64(0)     # only the partial rows the launch wrote are cleared and reduced
        _lib.call('cnf_flow_backward_rows', ctypes.byref(self.desc), _ptr(self.packed), _ptr(self.tables), _ptr(x),
                  _ptr(g_z), _ptr(g_ld), _ptr(gx), _ptr(self.partials), ctypes.c_int64(N), ctypes.byref(used), st)
        flat_grad = torch.empty(self.n_flat, dtype=torch.float32, device=x.device)
        _lib.call('cnf_grad_reduce_rows', ctypes.byref(self.desc), _ptr(self.partials), used, _ptr(self.gather),
                  _ptr(flat_grad), st)
        return gx, flat_grad

    TRAIN_CHUNK = 1 << 20   # samples per forward/backward kernel pair on the tensor-core training path

    def _check_batch(self, x, y, loss_acc):
        """The training kernels read N*K floats and N int64 labels blindly."""
        if x.dim() != 2 or x.shape[1] != self.K or x.dtype != torch.float32 or not x.is_contiguous():
            raise ValueError('cnf_b200: x must be a contiguous float32 [N, %d] tensor, got %s %s'
                             % (self.K, x.dtype, tuple(x.shape)))
        if y.dim() != 1 or y.shape[0] != x.shape[0] or y.dtype != torch.int64 or not y.is_contiguous():
            raise ValueError('cnf_b200: y must be a contiguous int64 [%d] tensor, got %s %s'
                             % (x.shape[0], y.dtype, tuple(y.shape)))
        if not x.is_cuda or y.device != x.device or loss_acc.device != x.device:
            raise ValueError('cnf_b200: x, y and loss_acc must live on the same CUDA device')
        if loss_acc.dtype != torch.float64 or loss_acc.numel() < 4:
            raise ValueError('cnf_b200: loss_acc must be float64 [4]')

    def _nll_step_tc(self, x, y, loss_acc, eps, gamma, n_total, with_grad, reduce=True):
        if self.tc_train is None:
            raise NotImplementedError('cnf_b200: the bf16 tensor-core training path does not cover this flow shape '
                                      '(K=%d, hidden=%s, L=%d); use precision="fp32"' % (self.K, self.hidden, self.L))
        n_grad, rows, wsb, g_host = self.tc_train
        N = x.shape[0]
        st = _stream(x.device)
        if self.flat_grad is None:
            self.flat_grad = torch.zeros(self.n_flat, dtype=torch.float32, device=self.device)
        if with_grad and self.partials_tc is None:
            self.partials_tc = torch.empty(rows * n_grad, dtype=torch.float32, device=self.device)
            self.gather_tcgrad = torch.from_numpy(g_host).to(self.device)
        chunk = max(1024, min(self.TRAIN_CHUNK, (N + 1023) // 1024 * 1024))
        need = chunk * wsb
        if self._train_ws is None or self._train_ws.numel() < need:
            self._train_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        used = ctypes.c_int64(0)
        _lib.call('cnf_nll_train_step_tc', ctypes.byref(self.desc_tc), _ptr(self.packed_tc), _ptr(self.tables),
                  _ptr(x), _ptr(y), ctypes.c_int64(N), ctypes.c_float(eps), ctypes.c_float(gamma),
                  ctypes.c_float(1.0 / max(n_total, 1)), _ptr(self.partials_tc) if with_grad else None,
                  _ptr(loss_acc), _ptr(self._train_ws), ctypes.c_int64(need), ctypes.byref(used), st)
        self.rows_used_tc = int(used.value)
        if with_grad and reduce:
            _lib.call('cnf_grad_reduce_tc', ctypes.byref(self.desc_tc), _ptr(self.partials_tc), used,
                      _ptr(self.gather_tcgrad), _ptr(self.flat_grad), st)

    def tc_tail_maps(self):
        """scatter_tc (flat entry -> its position in the tensor-core blob's gather map) on the device, or None when
        the one-launch tail does not apply (no tensor-core training path, or a map that is not one-to-one)."""
        if getattr(self, '_scatter_tc_dev', None) is not None and self._scatter_tc_dev.device == self.flat.device:
            return self._scatter_tc_dev
        if self.tc_train is None or self._gather_tc_host is None:
            return None
        if getattr(self, '_scatter_tc_host', None) is None:
            gt, gg = np.asarray(self._gather_tc_host), np.asarray(self.tc_train[3])
            pos = np.nonzero(gt >= 0)[0]
            live_g = gg[gg >= 0]
            ok = (np.unique(gt[pos]).size == pos.size) and (np.unique(live_g).size == live_g.size)
            inv = np.full(self.n_flat, -1, dtype=np.int32)
            inv[gt[pos]] = pos.astype(np.int32)
            # every packed entry must receive its gradient through the tail (else its blob copy would go stale)
            ok = ok and bool(np.all(np.isin(gt[pos], live_g)))
            self._scatter_tc_host = inv if ok else False
        if self._scatter_tc_host is False:
            return None
        self._scatter_tc_dev = torch.from_numpy(self._scatter_tc_host).to(self.flat.device)
        return self._scatter_tc_dev

    @_guarded
    def reduce_adam_pack_tc(self, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        """One-launch tail of a single-process bf16 Adam step (cnf_reduce_adam_pack_tc) behind
        nll_step(precision='bf16', reduce=False)."""
        if self.adam_m is None:
            self.adam_m = torch.zeros_like(self.flat)
            self.adam_v = torch.zeros_like(self.flat)
            self.adam_t = 0
        self.adam_t += 1
        self.version += 1
        _lib.call('cnf_reduce_adam_pack_tc', ctypes.byref(self.desc_tc), _ptr(self.partials_tc),
                  ctypes.c_int64(self.rows_used_tc), _ptr(self.gather_tcgrad), _ptr(self.tc_tail_maps()),
                  _ptr(self.flat), _ptr(self.flat_grad), _ptr(self.adam_m), _ptr(self.adam_v), _ptr(self.packed_tc),
                  ctypes.c_int64(self.adam_t), ctypes.c_float(lr), ctypes.c_float(betas[0]), ctypes.c_float(betas[1]),
                  ctypes.c_float(eps), _stream(self.device))

    @_guarded
    def nll_step(self, x, y, loss_acc, eps=1e-7, gamma=1.0, n_total=None, with_grad=True, precision='fp32',
                 reduce=True):
        """One fused forward+loss(+backward) pass over the local samples.  Accumulates the
        loss sums into loss_acc (float64 [4], device) and, with_grad, leaves
        d(sum loss)/n_total in self.flat_grad.  precision='bf16' runs the tcgen05 forward and
        backward kernels (weights must have been packed with pack(tc=True))."""
        N = x.shape[0]
        n_total = N if n_total is None else n_total
        self._check_batch(x, y, loss_acc)
        if precision == 'bf16':
            return self._nll_step_tc(x, y, loss_acc, eps, gamma, n_total, with_grad, reduce)
        st = _stream(x.device)
        if with_grad:
            self._want_partials()
        used = ctypes.c_int64(0)     # only the partial rows the launch wrote are cleared and reduced
        _lib.call('cnf_nll_train_step_rows', ctypes.byref(self.desc), _ptr(self.packed), _ptr(self.tables), _ptr(x),
                  _ptr(y), ctypes.c_int64(N), ctypes.c_float(eps), ctypes.c_float(gamma),
                  ctypes.c_float(1.0 / max(n_total, 1)), _ptr(self.partials) if with_grad else None,
                  _ptr(loss_acc), ctypes.byref(used), st)
        self.rows_used = int(used.value)
        if with_grad and reduce:          # (reduce=False: the caller follows with reduce_adam_pack)
            _lib.call('cnf_grad_reduce_rows', ctypes.byref(self.desc), _ptr(self.partials), used, _ptr(self.gather),
                      _ptr(self.flat_grad), st)

    @_guarded
    def reduce_adam_pack(self, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        """The tail of a single-process fp32 Adam step in ONE launch (cnf_reduce_adam_pack_rows): the partial rows of
        the nll_step(reduce=False) before it are summed, Adam (no weight decay) updates the parameters and the packed
        fp32 blob is refreshed; bitwise what cnf_grad_reduce_rows + adam + pack produce."""
        if self.adam_m is None:
            self.adam_m = torch.zeros_like(self.flat)
            self.adam_v = torch.zeros_like(self.flat)
            self.adam_t = 0
        self.adam_t += 1
        self.version += 1
        _lib.call('cnf_reduce_adam_pack_rows', ctypes.byref(self.desc), _ptr(self.partials),
                  ctypes.c_int64(self.rows_used), _ptr(self.gather), _ptr(self.flat), _ptr(self.flat_grad),
                  _ptr(self.adam_m), _ptr(self.adam_v), _ptr(self.packed), ctypes.c_int64(self.adam_t),
                  ctypes.c_float(lr), ctypes.c_float(betas[0]), ctypes.c_float(betas[1]), ctypes.c_float(eps),
                  _stream(self.device))

    @_guarded
    def fit_full_batch(self, x, y, epochs, eps=1e-7, gamma=1.0, n_total=None, lr=1e-3, betas=(0.9, 0.999), adam_eps=1e-8):
        """`epochs` full-batch fp32 Adam steps on (x, y) enqueued by ONE library call (cnf_fit_full_batch: per epoch the
        fused NLL pass and the one-launch optimiser tail, no Python between the launches).  Returns the per-epoch loss
        sums [epochs, 4] (float64, device) as ``FusedNLLTrainer.fit_loop`` defines them.  Needs a one-to-one gather map
        (``gather_one_to_one``) and host-counted Adam steps; the parameters come out bitwise as from nll_step(reduce=False) +
        reduce_adam_pack repeated."""
        N = x.shape[0]
        n_total = N if n_total is None else n_total
        hist = torch.empty((epochs, 4), dtype=torch.float64, device=x.device)
        scratch = torch.empty(4, dtype=torch.float64, device=x.device)
        self._check_batch(x, y, scratch)
        if epochs <= 0:
            return hist
        if self.adam_m is None:
            self.adam_m = torch.zeros_like(self.flat)
            self.adam_v = torch.zeros_like(self.flat)
            self.adam_t = 0
        self._want_partials()
        _lib.call('cnf_fit_full_batch', ctypes.byref(self.desc), _ptr(self.packed), _ptr(self.tables), _ptr(x), _ptr(y),
                  ctypes.c_int64(N), ctypes.c_float(eps), ctypes.c_float(gamma), ctypes.c_float(1.0 / max(n_total, 1)),
                  _ptr(self.partials), _ptr(self.gather), _ptr(self.flat), _ptr(self.flat_grad), _ptr(self.adam_m),
                  _ptr(self.adam_v), ctypes.c_int64(self.adam_t), ctypes.c_float(lr), ctypes.c_float(betas[0]),
                  ctypes.c_float(betas[1]), ctypes.c_float(adam_eps), ctypes.c_int64(epochs), _ptr(hist), _ptr(scratch),
                  _stream(x.device))
        self.adam_t += epochs
        self.version += epochs
        return hist

    @_guarded
    def adam(self, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if getattr(self, 'adam_step_dev', None) is not None:      # steps were taken on the device-counted path
            return self.adam_dev(lr, betas, eps, weight_decay)
        if self.adam_m is None:
            self.adam_m = torch.zeros_like(self.flat)
            self.adam_v = torch.zeros_like(self.flat)
            self.adam_t = 0
        self.adam_t += 1
        self.version += 1
        _lib.call('cnf_adam_step', _ptr(self.flat), _ptr(self.flat_grad), _ptr(self.adam_m), _ptr(self.adam_v),
                  ctypes.c_int64(self.n_flat), ctypes.c_int64(self.adam_t), ctypes.c_float(lr),
                  ctypes.c_float(betas[0]), ctypes.c_float(betas[1]), ctypes.c_float(eps),
                  ctypes.c_float(weight_decay), _stream(self.device))

    @_guarded
    def adam_dev(self, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        """Adam with the step count kept on the device (cnf_adam_step_dev): no host state, so the call can
        sit inside a captured CUDA graph.  Continues from the host-counted steps taken so far."""
        if self.adam_m is None:
            self.adam_m = torch.zeros_like(self.flat)
            self.adam_v = torch.zeros_like(self.flat)
            self.adam_t = 0
        if getattr(self, 'adam_step_dev', None) is None or self.adam_step_dev.device != self.flat.device:
            self.adam_step_dev = torch.full((1,), int(self.adam_t), dtype=torch.int64, device=self.flat.device)
            self.adam_coef = torch.zeros(2, dtype=torch.float32, device=self.flat.device)
        self.version += 1
        _lib.call('cnf_adam_step_dev', _ptr(self.flat), _ptr(self.flat_grad), _ptr(self.adam_m), _ptr(self.adam_v),
                  ctypes.c_int64(self.n_flat), _ptr(self.adam_step_dev), _ptr(self.adam_coef), ctypes.c_float(lr),
                  ctypes.c_float(betas[0]), ctypes.c_float(betas[1]), ctypes.c_float(eps),
                  ctypes.c_float(weight_decay), _stream(self.device))

    @_guarded
    def sgd(self, lr, weight_decay=0.0):
        self.version += 1
        _lib.call('cnf_sgd_step', _ptr(self.flat), _ptr(self.flat_grad), ctypes.c_int64(self.n_flat),
                  ctypes.c_float(lr), ctypes.c_float(weight_decay), _stream(self.device))


class _StackFunction(torch.autograd.Function):
    """Autograd node for one fused stack: saves only x; the backward kernel recomputes."""

    @staticmethod
    def forward(ctx, engine, x, *params):
        out, ld, _ = engine.apply(x)
        ctx.engine = engine
        ctx.save_for_backward(x.detach().to(torch.float32).contiguous())
        ctx.x_needs_grad = x.requires_grad
        return out, ld

    @staticmethod
    def backward(ctx, g_z, g_ld):
        engine = ctx.engine
        (x,) = ctx.saved_tensors
        if g_z is None:
            g_z = torch.zeros_like(x)
        if g_ld is None:
            g_ld = torch.zeros(x.shape[0], dtype=torch.float32, device=x.device)
        engine.ensure(x.device)
        engine.pack()
        gx, flat_grad = engine.backward(x, g_z, g_ld, need_gx=ctx.x_needs_grad)
        grads, off = [], 0
        for p in engine.params:
            n = p.numel()
            grads.append(flat_grad[off:off + n].view(p.shape))
            off += n
        return (None, gx) + tuple(grads)


def stack_forward(engine, x):
    """Differentiable fused forward of one stack."""
    need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in engine.params))
    if need_grad:
        engine.ensure(x.device)
        return _StackFunction.apply(engine, x, *engine.params)
    out, ld, _ = engine.apply(x)
    return out, ld
