"""Drop-in for the reference's ``flows/flows.py``: ``Flow`` (:8-37) and ``NvpCouplingLayer``
(:68-126) with the same constructor arguments, attributes and ``state_dict`` keys.

The modules hold parameters; the arithmetic is one fused CUDA kernel per run of consecutive
coupling layers (``_engine.StackEngine`` -> libcnf_b200).  Quirks of the reference that are
preserved on purpose:
  * fixed half-split mask ``mask[:, dim//2:] = 1`` and a data flip after every layer, so for odd
    K the middle logit is never transformed and for odd L the class order comes out reversed
    (flows/flows.py:81-86, 112);
  * ``Flow.forward`` returns ``(zs, cum_log_det)`` with ``zs`` a list of the L intermediate
    outputs (flows/flows.py:17-25) -- here a lazy sequence: ``zs[-1]`` is the fused result,
    other entries are materialised by a second launch only when indexed;
  * ``log_det`` is squeezed, i.e. 0-d when N == 1 (flows/flows.py:109, 125);
  * ``Flow.backward`` raises ``ValueError('Flow inverse not tractable!')`` when a layer is not
    invertible (flows/flows.py:28-29).
Difference: ``random_flip=True`` also works on CUDA (the reference calls ``.numpy()`` on the
permutation and is CPU-only there).
"""
import numpy as np
import torch
from torch import nn

from .utils import MLP
import ctypes

from .. import _lib
from .._engine import StackEngine, stack_forward, require_cuda, on_device, check_logits, _ptr, _stream


def _zero_net(x):
    return x.new_zeros(x.size())


def _vec(p, K, what):
    """float32 contiguous [K] device vector for the C ABI; a length mismatch raises instead of reading out of bounds."""
    if p is None:
        return None
    v = p.detach().to(torch.float32).contiguous().view(-1)
    if v.numel() != K:
        raise ValueError('cnf_b200: %s has %d entries, the input has %d columns' % (what, v.numel(), K))
    return v


class _AffineConstFn(torch.autograd.Function):
    """z = x*exp(s)+t per dimension through cnf_affine_const; s/t are [K] tensors or None."""

    @staticmethod
    def forward(ctx, x, s, t):
        if x.dim() != 2:
            raise ValueError('cnf_b200: expected [N, K] input, got %s' % (tuple(x.shape),))
        x = x.to(torch.float32).contiguous()
        N, K = x.shape
        z = torch.empty_like(x)
        sv, tv = _vec(s, K, 's'), _vec(t, K, 't')
        with on_device(x.device):
            _lib.call('cnf_affine_const', _ptr(x), _ptr(sv), _ptr(tv), _ptr(z), ctypes.c_int64(N), ctypes.c_int32(K),
                      ctypes.c_int32(0), _stream(x.device))
        ctx.save_for_backward(x, sv)
        ctx.has = (s is not None, t is not None)
        return z

    @staticmethod
    def backward(ctx, g_z):
        x, sv = ctx.saved_tensors
        N, K = x.shape
        g_z = g_z.to(torch.float32).contiguous()
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gs = torch.empty(K, dtype=torch.float32, device=x.device) if ctx.has[0] else None
        gt = torch.empty(K, dtype=torch.float32, device=x.device) if ctx.has[1] else None
        with on_device(x.device):
            _lib.call('cnf_affine_const_backward', _ptr(x), _ptr(g_z), _ptr(sv), _ptr(gx), _ptr(gs), _ptr(gt),
                      ctypes.c_int64(N), ctypes.c_int32(K), _stream(x.device))
        return gx, gs, gt


def affine_const_apply(x, s=None, t=None, inverse=False):
    """Non-differentiable helper: forward or inverse of the per-dimension affine map."""
    require_cuda(x)
    if x.dim() != 2:
        raise ValueError('cnf_b200: expected [N, K] input, got %s' % (tuple(x.shape),))
    x = x.detach().to(torch.float32).contiguous()
    z = torch.empty_like(x)
    sv, tv = _vec(s, x.shape[1], 's'), _vec(t, x.shape[1], 't')
    with on_device(x.device):
        _lib.call('cnf_affine_const', _ptr(x), _ptr(sv), _ptr(tv), _ptr(z), ctypes.c_int64(x.shape[0]),
                  ctypes.c_int32(x.shape[1]), ctypes.c_int32(1 if inverse else 0), _stream(x.device))
    return z


class AffineConstantLayer(nn.Module):
    """Drop-in for reference ``flows/flows.py:40-65``: learned per-dimension scale and shift.
    As in the reference, ``log_det`` has shape [1] when ``scale`` (it is ``sum(s, dim=1)`` of a
    [1, dim] parameter) and [N] zeros otherwise."""

    def __init__(self, dim, scale=True, shift=True):
        super().__init__()
        self.s = nn.Parameter(torch.zeros(1, dim)) if scale else None
        self.t = nn.Parameter(torch.zeros(1, dim)) if shift else None
        self.invertible = True

    def forward(self, x):
        require_cuda(x)
        s = None if self.s is None else self.s.view(-1)
        t = None if self.t is None else self.t.view(-1)
        if torch.is_grad_enabled() and (x.requires_grad or s is not None or t is not None):
            z = _AffineConstFn.apply(x, s, t)
        else:
            z = affine_const_apply(x, s, t)
        log_det = torch.sum(self.s, dim=1) if self.s is not None else x.new_zeros(x.shape[0])
        return z, log_det

    def backward(self, z):
        require_cuda(z)
        x = affine_const_apply(z, None if self.s is None else self.s, None if self.t is None else self.t, inverse=True)
        log_det = torch.sum(-self.s, dim=1).detach() if self.s is not None else z.new_zeros(z.shape[0])
        return x, log_det


class _PlanarFn(torch.autograd.Function):
    """(x, w, u_hat, b) -> (z, log_det) through cnf_planar_forward / cnf_planar_backward."""

    @staticmethod
    def forward(ctx, x, w, u_hat, b):
        if x.dim() != 2:
            raise ValueError('cnf_b200: expected [N, K] input, got %s' % (tuple(x.shape),))
        x = x.to(torch.float32).contiguous()
        N, K = x.shape
        wv, uv, bv = _vec(w, K, 'w'), _vec(u_hat, K, 'u'), _vec(b, 1, 'b')
        z = torch.empty_like(x)
        ld = torch.empty(N, dtype=torch.float32, device=x.device)
        with on_device(x.device):
            _lib.call('cnf_planar_forward', _ptr(x), _ptr(wv), _ptr(uv), _ptr(bv), _ptr(z), _ptr(ld), ctypes.c_int64(N),
                      ctypes.c_int32(K), _stream(x.device))
        ctx.save_for_backward(x, wv, uv, bv)
        ctx.shapes = (w.shape, u_hat.shape, b.shape)
        return z, ld

    @staticmethod
    def backward(ctx, g_z, g_ld):
        x, wv, uv, bv = ctx.saved_tensors
        N, K = x.shape
        g_z = torch.zeros_like(x) if g_z is None else g_z.to(torch.float32).contiguous()
        g_ld = None if g_ld is None else g_ld.to(torch.float32).contiguous()
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gw = torch.empty(K, dtype=torch.float32, device=x.device)
        gu = torch.empty(K, dtype=torch.float32, device=x.device)
        gb = torch.empty(1, dtype=torch.float32, device=x.device)
        with on_device(x.device):
            _lib.call('cnf_planar_backward', _ptr(x), _ptr(g_z), _ptr(g_ld), _ptr(wv), _ptr(uv), _ptr(bv), _ptr(gx),
                      _ptr(gw), _ptr(gu), _ptr(gb), ctypes.c_int64(N), ctypes.c_int32(K), _stream(x.device))
        ws, us, bs = ctx.shapes
        return gx, gw.view(ws), gu.view(us), gb.view(bs)


class PlanarLayer(nn.Module):
    """Drop-in for reference ``flows/flows.py:129-164`` (forward direction only; ``invertible`` is
    False, so a ``Flow`` containing it raises on ``backward`` exactly as the reference does).  The
    K-vector re-parametrisation of ``u`` (:150-153) stays in torch ops; the per-sample work is one
    streaming kernel, differentiable through ``cnf_planar_backward``."""

    def __init__(self, dim=0, params=None):
        super().__init__()
        if params is not None:
            self.w = params['w'].squeeze()
            self.u = params['u'].squeeze()
            self.b = params['b'].squeeze()
        else:
            if dim < 1:
                raise ValueError('Either dim of params must be provided!')
            self.w = nn.Parameter(torch.rand(dim))
            self.u = nn.Parameter(torch.rand(dim))
            self.b = nn.Parameter(torch.rand(1))
        self.invertible = False

    def forward(self, x):
        require_cuda(x)
        wtu = torch.dot(self.w, self.u)
        m = -1 + torch.log1p(torch.exp(wtu))
        u_hat = self.u + (m - wtu) * self.w / torch.norm(self.w)
        z, log_det = _PlanarFn.apply(x, self.w, u_hat, self.b.reshape(1))
        return z, log_det.squeeze()


class _RadialFn(torch.autograd.Function):
    """(x, z0, a, b_hat) -> z through cnf_radial_forward / cnf_radial_backward."""

    @staticmethod
    def forward(ctx, x, z0, a, b_hat):
        if x.dim() != 2:
            raise ValueError('cnf_b200: expected [N, K] input, got %s' % (tuple(x.shape),))
        x = x.to(torch.float32).contiguous()
        N, K = x.shape
        zv, av, bv = _vec(z0, K, 'z0'), _vec(a, 1, 'a'), _vec(b_hat, 1, 'b')
        z = torch.empty_like(x)
        with on_device(x.device):
            _lib.call('cnf_radial_forward', _ptr(x), _ptr(zv), _ptr(av), _ptr(bv), _ptr(z), ctypes.c_int64(N),
                      ctypes.c_int32(K), _stream(x.device))
        ctx.save_for_backward(x, zv, av, bv)
        ctx.shapes = (z0.shape, a.shape, b_hat.shape)
        return z

    @staticmethod
    def backward(ctx, g_z):
        x, zv, av, bv = ctx.saved_tensors
        N, K = x.shape
        g_z = g_z.to(torch.float32).contiguous()
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gz0 = torch.empty(K, dtype=torch.float32, device=x.device)
        ga = torch.empty(1, dtype=torch.float32, device=x.device)
        gb = torch.empty(1, dtype=torch.float32, device=x.device)
        with on_device(x.device):
            _lib.call('cnf_radial_backward', _ptr(x), _ptr(g_z), _ptr(zv), _ptr(av), _ptr(bv), _ptr(gx), _ptr(gz0),
                      _ptr(ga), _ptr(gb), ctypes.c_int64(N), ctypes.c_int32(K), _stream(x.device))
        zs, as_, bs = ctx.shapes
        return gx, gz0.view(zs), ga.view(as_), gb.view(bs)


class RadialLayer(nn.Module):
    """Drop-in for reference ``flows/flows.py:167-193``.  As in the reference the returned log-det
    is the constant ``log(tensor(1.0))`` (a 0-d CPU tensor there; :188-190), not the true one."""

    def __init__(self, dim):
        super().__init__()
        self.z0 = nn.Parameter(torch.rand(dim))
        self.a = nn.Parameter(torch.rand(1))
        self.b = nn.Parameter(torch.rand(1))
        self.invertible = False

    def forward(self, x):
        require_cuda(x)
        b_hat = -self.a + torch.log1p(torch.exp(self.b))
        z = _RadialFn.apply(x, self.z0, self.a, b_hat)
        log_det = torch.log(torch.tensor(1.0))
        return z, log_det


class NvpCouplingLayer(nn.Module):
    def __init__(self, dim, hidden_size=[5, 5], scale=True, shift=True, random_flip=False):
        super().__init__()
        self.dim = int(dim)
        self.hidden_size = list(hidden_size)
        # Plain callables returning zeros stand in for an absent net, as in the reference.
        self.s = MLP(dim, self.hidden_size, wscale=0.001) if scale else _zero_net
        self.t = MLP(dim, self.hidden_size, wscale=0.001) if shift else _zero_net
        mask = torch.zeros(1, dim)
        mask[:, dim // 2:] = 1
        self.mask = nn.Parameter(mask, requires_grad=False)
        self.invertible = True
        self.random_flip = random_flip
        if random_flip:
            perm = np.random.permutation(dim)
            rev = np.empty(dim, dtype=np.int64)
            rev[perm] = np.arange(dim)
            self.perm = nn.Parameter(torch.as_tensor(perm, dtype=torch.long).view(1, -1), requires_grad=False)
            self.rev_perm = nn.Parameter(torch.as_tensor(rev, dtype=torch.long).view(1, -1), requires_grad=False)
        self._engine = None
        self._perm_cache = None

    def __getstate__(self):
        state = dict(self.__dict__)
        state['_engine'] = None          # device handles are rebuilt lazily, never pickled
        return state

    # -- structure helpers ---------------------------------------------------------------
    @property
    def has_scale(self):
        return isinstance(self.s, nn.Module)

    @property
    def has_shift(self):
        return isinstance(self.t, nn.Module)

    @property
    def coupling_func(self):
        """NICE-style accessor the reference's notebooks used on additive layers."""
        return self.t

    def shape_signature(self):
        return (self.dim, tuple(self.hidden_size), self.has_scale, self.has_shift)

    def signature(self):
        return self.shape_signature() + (tuple(self.perm_list()) if self.random_flip else None,)

    def canonical_parameters(self):
        out = []
        if self.has_scale:
            out += self.s.canonical_parameters()
        if self.has_shift:
            out += self.t.canonical_parameters()
        return out

    def perm_list(self):
        if not self.random_flip:
            return list(range(self.dim))
        if self._perm_cache is None:      # one host read; refreshed after load_state_dict
            self._perm_cache = [int(v) for v in self.perm.detach().view(-1).tolist()]
        return self._perm_cache

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._perm_cache = None
        self._engine = None

    # -- reference API -------------------------------------------------------------------
    def _own_engine(self):
        if self._engine is None:
            self._engine = build_engine([self])
        return self._engine

    def forward(self, x):
        z, ld = stack_forward(self._own_engine(), x)
        return z, ld.squeeze()

    def backward(self, z):
        x, ld, _ = self._own_engine().apply(z, inverse=True)
        return x, ld.squeeze()


def _check_conditioner(net):
    """The fused kernels evaluate Linear-ReLU-...-Linear conditioners (what the reference's coupling layer builds,
    flows/flows.py:76-79 with flows/utils.py:10).  Anything else must fail loudly, not silently run as ReLU."""
    import torch.nn.functional as F
    if not isinstance(net, nn.Module):
        return                                    # absent net (scale=False / shift=False)
    act = getattr(net, 'activation', None)
    if not (act is F.relu or act is torch.relu or isinstance(act, nn.ReLU)):
        raise NotImplementedError('cnf_b200: the fused coupling kernels evaluate ReLU conditioners only (the '
                                  'reference\'s NvpCouplingLayer never builds another one); got activation=%r' % (act,))
    if not hasattr(net, 'canonical_parameters') or not all(isinstance(l, nn.Linear) for l in net.layers):
        raise NotImplementedError('cnf_b200: a coupling layer\'s s / t must be cnf_b200.flows.utils.MLP instances '
                                  '(stacks of nn.Linear), got %s' % type(net).__name__)


def build_engine(layers):
    first = layers[0]
    for lay in layers:
        _check_conditioner(lay.s)
        _check_conditioner(lay.t)
    perms = [lay.perm_list() for lay in layers] if any(lay.random_flip for lay in layers) else None
    return StackEngine(first.dim, first.hidden_size, first.has_scale, first.has_shift, perms,
                       [lay.canonical_parameters() for lay in layers])


class _LazyOutputs:
    """List-like view of the per-layer outputs of a Flow pass.  ``[-1]`` is available at once;
    any other index triggers (once per fused segment) a launch that writes all intermediates.

    Differences from the reference's plain list (flows/flows.py:17-25), by design: the intermediates are
    recomputed from the segment's input when first indexed, by the fp32 kernel (also when ``zs[-1]`` came from
    the bf16 kernel), and they carry no autograd graph -- only ``zs[-1]`` is differentiable.  They must be read
    before the weights change: the pass remembers the version of every stack's parameter buffer and indexing an
    intermediate after an optimiser step / ``load_state_dict`` raises instead of silently returning the
    outputs of the NEW weights."""

    def __init__(self, flow, segments, seg_inputs, seg_outputs, inverse):
        self._flow, self._segments = flow, segments
        self._inputs, self._outputs = seg_inputs, seg_outputs
        self._inverse = inverse
        self._cache = {}
        self._n = sum(seg[2] - seg[1] for seg in segments)
        self._versions = [seg[3].weights_version() if seg[0] == 'stack' else None for seg in segments]

    def __len__(self):
        return self._n

    def __iter__(self):
        return (self[i] for i in range(self._n))

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self._n))]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError('list index out of range')
        pos = 0
        for si, (kind, a, b, eng) in enumerate(self._segments):
            n = b - a
            if i < pos + n:
                if i == pos + n - 1:
                    return self._outputs[si]
                if si not in self._cache:
                    v = self._versions[si]
                    if v is not None and eng.weights_version() != v:
                        raise RuntimeError('cnf_b200: the flow\'s weights changed since this pass; intermediate outputs '
                                           '(zs[i], i < L-1) are materialised lazily and must be read before the next '
                                           'optimiser step / load_state_dict (zs[-1] is always available)')
                    with torch.no_grad():
                        _, _, allz = eng.apply(self._inputs[si], inverse=self._inverse, want_all=True)
                    self._cache[si] = allz
                return self._cache[si][i - pos]
            pos += n
        raise IndexError(i)


class Flow(nn.Module):
    def __init__(self, layers, **kwargs):
        super().__init__()
        self.layers = nn.ModuleList(layers)
        self.invertible = all([layer.invertible for layer in self.layers])
        self.precision = kwargs.get('precision', 'fp32')
        self._plan = None

    def __getstate__(self):
        state = dict(self.__dict__)
        state['_plan'] = None
        return state

    # Consecutive coupling layers with the same shape fuse into one kernel launch; any other
    # module that follows the (z, log_det) layer protocol is called as is.
    def _segments(self):
        key = tuple((id(l), l.signature() if isinstance(l, NvpCouplingLayer) else None) for l in self.layers)
        if self._plan is not None and self._plan[0] == key:
            return self._plan[1]
        segs, i, layers = [], 0, list(self.layers)
        while i < len(layers):
            lay = layers[i]
            if isinstance(lay, NvpCouplingLayer):
                j = i + 1
                while (j < len(layers) and isinstance(layers[j], NvpCouplingLayer)
                       and layers[j].shape_signature() == lay.shape_signature()
                       and len({id(l) for l in layers[i:j + 1]}) == j + 1 - i):
                    j += 1
                segs.append(('stack', i, j, build_engine(layers[i:j])))
                i = j
            else:
                segs.append(('module', i, i + 1, lay))
                i += 1
        self._plan = (key, segs)
        return segs

    def engine(self):
        """The single fused engine when the whole flow is one homogeneous coupling stack."""
        segs = self._segments()
        if len(segs) == 1 and segs[0][0] == 'stack':
            return segs[0][3]
        return None

    def forward(self, x):
        require_cuda(x)
        segs = self._segments()
        cum, ins, outs = 0.0, [], []
        for kind, a, b, obj in segs:
            ins.append(x)
            if kind == 'stack':
                torch.cuda.nvtx.range_push('cnf.flow_forward')
                if self.precision == 'bf16' and not torch.is_grad_enabled():
                    x, ld, _ = obj.apply(x, precision='bf16')
                else:
                    x, ld = stack_forward(obj, x)
                torch.cuda.nvtx.range_pop()
                ld = ld.squeeze()
            else:
                x, ld = obj(x)
            outs.append(x)
            cum = cum + ld
        return _LazyOutputs(self, segs, ins, outs, False), cum

    def backward(self, z):
        if not self.invertible:
            raise ValueError('Flow inverse not tractable!')
        require_cuda(z)
        segs = self._segments()[::-1]
        cum, ins, outs = 0.0, [], []
        for kind, a, b, obj in segs:
            ins.append(z)
            if kind == 'stack':
                torch.cuda.nvtx.range_push('cnf.flow_inverse')
                z, ld, _ = obj.apply(z, inverse=True, precision=self.precision)
                torch.cuda.nvtx.range_pop()
                ld = ld.squeeze()
            else:
                z, ld = obj.backward(z)
            outs.append(z)
            cum = cum + ld
        return _LazyOutputs(self, segs, ins, outs, True), cum


class CouplingStack(nn.Module):
    """Factory-style flow with the contract ``TorchFlowCalibrator`` expects from the class it is
    given (calibrators.py:251, 287, 304, 339): ``Flow(dim, **kwargs)`` must swallow the
    calibrator's own kwargs (dev, epochs, batch_size), ``flow(x)`` returns ``(z, log_det)`` and
    ``.layers`` iterates the coupling layers.  Defaults follow the Keras predecessors
    (code-old/nice.py:104-113, code-old/realNVP.py:47-52): 4 layers, hidden ``[dim]``."""

    def __init__(self, dim, layers=4, hidden_size=None, scale=True, shift=True, random_flip=False,
                 precision='fp32', **ignored):
        super().__init__()
        hidden_size = [dim] if hidden_size is None else list(hidden_size)
        self.flow = Flow([NvpCouplingLayer(dim, hidden_size=hidden_size, scale=scale, shift=shift,
                                           random_flip=random_flip) for _ in range(layers)],
                         precision=precision)
        self.invertible = True

    @property
    def layers(self):
        return self.flow.layers

    def engine(self):
        return self.flow.engine()

    def forward(self, x):
        zs, ld = self.flow(x)
        return zs[-1], ld

    def transform_host(self, x_host, out_host=None, ld_host=None, device=None, **kw):
        """numpy / CPU tensor in, CPU tensors out: (z, log_det), streamed through the GPU in
        overlapping chunks (see StackEngine.apply_host).  Synchronises before returning."""
        eng = self.engine()
        if eng is None:
            raise NotImplementedError('transform_host needs a homogeneous coupling stack')
        dev = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        if dev.index is None:
            dev = torch.device('cuda', torch.cuda.current_device())
        if eng.device != dev or eng.params[0].device != dev:
            self.to(dev)                     # module traversal is ~0.1 ms: only when something moved
        z, ld = eng.apply_host(x_host, out_host, ld_host, precision=self.flow.precision, device=dev, **kw)
        torch.cuda.current_stream(dev).synchronize()
        return z, ld

    def backward(self, z):
        xs, ld = self.flow.backward(z)
        return xs[-1], ld
