"""B200-native (sm_100a) implementation of ProxyTransformation's point-cloud preshaping hot path.

Public surface = the reference's: ``ProxyTransformationNormReverse`` registered in ``MODELS``
(embodiedscan/models/necks/preshape_norm_reverse_drop.py:280-281).
"""
from .registry import MODELS  # noqa: F401
from .necks import ProxyTransformationNormReverse  # noqa: F401

__all__ = ["MODELS", "ProxyTransformationNormReverse"]
