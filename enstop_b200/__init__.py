"""enstop_b200 — B200-native pLSA EM engine behind the enstop ``PLSA`` / ``EnsembleTopics`` API.

Host code is numpy/scipy + ctypes; the EM hot path is ``libplsa_b200.so`` (hand-written
sm_100a CUDA, C ABI in ``include/plsa_b200.h``).  Importing the package does not need a GPU;
fitting does, and fails loudly without one (there is no CPU fallback).
"""
from ._lib import release_device_memory  # noqa: F401
from .plsa import PLSA, plsa_fit, plsa_init, plsa_refit  # noqa: F401

__all__ = ["PLSA", "EnsembleTopics", "plsa_fit", "plsa_init", "plsa_refit",
           "release_device_memory"]
__version__ = "0.1.0"


def __getattr__(name):
    if name == "EnsembleTopics":
        from .enstop_ import EnsembleTopics
        return EnsembleTopics
    raise AttributeError(name)
