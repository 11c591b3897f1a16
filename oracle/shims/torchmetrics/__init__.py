"""Import shim (golden-vector generation only): routes ``torchmetrics`` to the oracle restatement."""
from oracle.tm_pearson import Metric, PearsonCorrCoef  # noqa: F401
from . import regression  # noqa: F401
