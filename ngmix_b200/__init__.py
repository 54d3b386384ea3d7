"""
ngmix_b200: B200-native batched Gaussian-mixture image fitting, a drop-in for
the fitting hot path of esheldon/ngmix.
"""
__version__ = "0.1.0"
