"""triumvirate_b200 -- B200-native three-point clustering estimators.

A from-scratch sm_100a implementation of the FFT-based bispectrum / 3PCF
estimator of MikeSWang/Triumvirate behind the reference's own API surface.
The compute path is hand-written CUDA + cuFFT in ``libtrvb.so``; there is no
CPU fallback.
"""
from . import core  # noqa: F401

__version__ = "0.1.0"
