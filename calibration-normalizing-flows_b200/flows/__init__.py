"""Drop-in for the reference's ``flows`` package (flows/flows.py, flows/utils.py) plus the
factory adapters its notebooks imported (flows.nice_torch / flows.realNVP_torch)."""
from .flows import Flow, NvpCouplingLayer, AffineConstantLayer, PlanarLayer, RadialLayer
from .utils import MLP, TempScaler
from .nice_torch import NiceFlow
from .realNVP_torch import RealNvpFlow

__all__ = ['Flow', 'NvpCouplingLayer', 'AffineConstantLayer', 'PlanarLayer', 'RadialLayer', 'MLP', 'TempScaler', 'NiceFlow', 'RealNvpFlow']
