"""Estimators on the hot path: same names and constructor arguments as ``cca_zoo.linear``."""
from ._rcca import CCA, PLS, rCCA
from ._mcca import MCCA
from ._gcca import GCCA
from ._partialcca import PartialCCA
from ._grcca import GRCCA

__all__ = ["CCA", "rCCA", "PLS", "MCCA", "GCCA", "PartialCCA", "GRCCA"]
