"""B200-native batched log-posterior + gradient for the NMGP stationary / separable /
nonseparable multivariate Gaussian-process models (drop-in for the reference's Utility/logpos.py path)."""
__version__ = "0.1.0"
