"""``RealNvpFlow`` factory -- counterpart of the deleted ``flows.realNVP_torch`` module
(notebooks/simulated-predictions-flows.ipynb:214)."""
from .flows import CouplingStack


class RealNvpFlow(CouplingStack):
    def __init__(self, dim, layers=4, hidden_size=None, **ignored):
        super().__init__(dim, layers=layers, hidden_size=hidden_size, scale=True, shift=True, **ignored)
