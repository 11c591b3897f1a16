from oracle.tm_pearson import PearsonCorrCoef  # noqa: F401
