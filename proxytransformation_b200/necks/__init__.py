from .preshape import ProxyTransformationNormReverse

__all__ = ["ProxyTransformationNormReverse"]
