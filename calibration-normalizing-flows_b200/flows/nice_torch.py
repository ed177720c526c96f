"""``NiceFlow`` factory -- the class the reference's notebooks imported from the (deleted)
``flows.nice_torch`` module and handed to ``TorchFlowCalibrator`` (SURVEY.md F7;
notebooks/simulated-predictions-flows.ipynb:224).  NICE additive coupling in this code base
is ``NvpCouplingLayer(scale=False)`` (flows/flows.py:76-79)."""
from .flows import CouplingStack


class NiceFlow(CouplingStack):
    def __init__(self, dim, layers=4, hidden_size=None, **ignored):
        super().__init__(dim, layers=layers, hidden_size=hidden_size, scale=False, shift=True, **ignored)
