"""vegasafterglow_b200 -- B200-native implementation of VegasAfterglow's model-evaluation path.

* ``Engine``            batched C-ABI front end (ctypes) -- flux_density_grid / flux_density / chi2
* ``VegasAfterglowC_b200`` pybind11 mirror of the reference's ``VegasAfterglowC`` surface
  (``Model``, ``TophatJet``, ``GaussianJet``, ``PowerLawJet``, ``ISM``, ``Wind``, ``Observer``,
  ``Radiation``) -- import it as ``from vegasafterglow_b200 import VegasAfterglowC_b200 as va``
* ``fitting.BatchedLikelihood`` drop-in for the emcee ``log_prob_batch`` of the reference
* ``parallel``          walker partition over the GPUs of a box

GPU only: there is no CPU fallback anywhere in this package.
"""
from . import abi, configs  # noqa: F401

__all__ = ["abi", "configs", "Engine"]


def __getattr__(name):
    if name == "Engine":
        from .engine import Engine

        return Engine
    raise AttributeError(name)
