"""cnf_b200 -- B200-native (sm_100a) coupling-flow calibration hot path.

Drop-in for the reference's ``flows.flows`` / ``flows.utils`` / ``calibrators`` (flow calibrator)
/ ``utils.metrics`` on this path; the arithmetic lives in ``libcnf_b200.so`` (C ABI in
``include/cnf.h``).  Importing the package loads the library and fails loudly if it is missing.
"""
from . import _lib

_lib.load()

from .flows.flows import (Flow, NvpCouplingLayer, CouplingStack, AffineConstantLayer, PlanarLayer,  # noqa: E402
                          RadialLayer)
from .flows.utils import MLP, TempScaler  # noqa: E402
from .flows.nice_torch import NiceFlow  # noqa: E402
from .flows.realNVP_torch import RealNvpFlow  # noqa: E402
from .calibrators import Calibrator, TorchFlowCalibrator, FusedNLLTrainer, HostStreamNLLTrainer  # noqa: E402
from .utils.metrics import expected_calibration_error, neg_log_likelihood, accuracy  # noqa: E402
from .utils.ops import onehot_encode  # noqa: E402

__all__ = ['Flow', 'NvpCouplingLayer', 'CouplingStack', 'AffineConstantLayer', 'PlanarLayer', 'RadialLayer', 'TempScaler', 'MLP', 'NiceFlow', 'RealNvpFlow', 'Calibrator',
           'TorchFlowCalibrator', 'FusedNLLTrainer', 'HostStreamNLLTrainer', 'expected_calibration_error', 'neg_log_likelihood',
           'accuracy', 'onehot_encode']
