"""Import the UNMODIFIED reference modules for timing / checking.  TEST / BASELINE INFRASTRUCTURE ONLY
(bench.py's reference arm and cpu_baseline legs, oracle/make_golden.py-style checks); never imported by
the product package.

Search order: oracle/_ref/reference.zip (packed by oracle/fetch_reference.py, travels to the GPU box; imported
straight from the archive), then /root/reference (authoring container).  Returns None when neither exists (callers fall back to the
port in oracle/ref_port_torch.py and say kind == "port").

Shims applied in memory only (the files stay byte-identical):
  * utils/metrics.py:54 ``np.equal(..., dtype=np.float32)`` fails on NumPy >= 2
    -> ``np.equal(...).astype(np.float32)``;
  * ``flows.nice_torch.NiceFlow`` / ``flows.realNVP_torch.RealNvpFlow`` never existed in the tree
     -> ``FlowAdapter`` below gives the (pred, log_det) contract calibrators.py:251, 287 expects.
"""
import importlib
import os
import sys
import types
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = [os.path.join(HERE, '_ref', 'reference.zip'), '/root/reference']


class Reference:
    def __init__(self, root):
        self.root = root
        for name in [m for m in sys.modules if m == 'flows' or m.startswith('flows.') or m == 'utils' or
                     m.startswith('utils.') or m == 'calibrators']:
            del sys.modules[name]
        sys.path.insert(0, root)
        try:
            self.flows = importlib.import_module('flows.flows')
            self.calibrators = importlib.import_module('calibrators')
        finally:
            sys.path.remove(root)
        self.Flow, self.NvpCouplingLayer = self.flows.Flow, self.flows.NvpCouplingLayer
        if zipfile.is_zipfile(root):
            src = zipfile.ZipFile(root).read('utils/metrics.py').decode()
        else:
            src = open(os.path.join(root, 'utils', 'metrics.py')).read()
        src = src.replace('np.equal(preds, target, dtype=np.float32)', 'np.equal(preds, target).astype(np.float32)')
        src = src.replace('from .ops import onehot_encode', 'from utils.ops import onehot_encode')
        self.metrics = types.ModuleType('ref_metrics')
        exec(compile(src, 'utils/metrics.py', 'exec'), self.metrics.__dict__)
        import torch
        Flow, Layer = self.Flow, self.NvpCouplingLayer

        class FlowAdapter(torch.nn.Module):
            def __init__(self, dim, layers=4, hidden_size=None, scale=True, **ignored):
                super().__init__()
                hidden_size = [dim] if hidden_size is None else list(hidden_size)
                self.flow = Flow([Layer(dim, hidden_size=hidden_size, scale=scale) for _ in range(layers)])
                self.layers = self.flow.layers

            def forward(self, x):
                zs, ld = self.flow(x)
                return zs[-1], ld
        self.FlowAdapter = FlowAdapter

    def build_flow(self, K, L, hidden, scale=True, shift=True, seed=1, wmult=1.0):
        """Reference init (flows/flows.py:76-79) under torch.manual_seed(seed), weights x wmult."""
        import torch
        torch.manual_seed(seed)
        flow = self.Flow([self.NvpCouplingLayer(K, hidden_size=list(hidden), scale=scale, shift=shift) for _ in range(L)])
        with torch.no_grad():
            for p in flow.parameters():
                if p.requires_grad:
                    p.mul_(wmult)
        return flow


def load():
    for root in CANDIDATES:
        if (os.path.isfile(root) and zipfile.is_zipfile(root)) or os.path.isfile(os.path.join(root, 'flows', 'flows.py')):
            return Reference(root)
    return None
