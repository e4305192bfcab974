"""Batched likelihood: the drop-in for ``log_prob_batch`` of the reference's emcee driver
(VegasAfterglow/fitting/samplers.py:72-91) and for ``Fitter._evaluate`` (fitter.py:503-533).

The reference evaluates each walker on a host thread (``pool.map(eval_one, valid_samples)``);
here the whole ensemble goes to the GPU as ONE batch of ``vag_params`` records, and -- when
``torch.distributed`` is initialised -- the batch is partitioned by walker across the ranks and
only the float64[n_walkers] log-likelihood vector is gathered (vegasafterglow_b200/parallel.py).

Same data conventions as the reference (fitter.py:407-437): points sorted by time, weights
normalised to sum N, ln-flux chi-squared, non-finite chi2 -> logL = -inf.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np

from . import abi

# sampler-space parameter name -> (record field path, ) for the typed jet/medium variants
# (fitting/config.py:99-137 registries; Fitter._build_model fitter.py:455-495)
_FIELDS = {
    "theta_c": ("theta_c",), "E_iso": ("E_iso",), "Gamma0": ("Gamma0",), "k_e": ("k_e",), "k_g": ("k_g",),
    "duration": ("duration",), "tau": ("duration",), "theta_w": ("theta_w",), "E_iso_w": ("E_iso_w",),
    "Gamma0_w": ("Gamma0_w",), "sigma0": ("sigma0",), "n_ism": ("n_ism",), "A_star": ("A_star",), "n0": ("n0",),
    "theta_v": ("theta_obs",), "eps_e": ("fwd", "eps_e"), "eps_B": ("fwd", "eps_B"), "p": ("fwd", "p"),
    "xi_e": ("fwd", "xi_e"), "eps_e_r": ("rvs", "eps_e"), "eps_B_r": ("rvs", "eps_B"), "p_r": ("rvs", "p"),
    "xi_e_r": ("rvs", "xi_e"),
}


def consolidate_data(t, nu, flux, err, weights=None):
    """fitter.py:407-437: sort by time, normalise weights to sum N, ln-flux and relative error."""
    t, nu, flux, err = (np.asarray(a, dtype=np.float64).reshape(-1) for a in (t, nu, flux, err))
    w = np.ones_like(t) if weights is None else np.asarray(weights, dtype=np.float64).reshape(-1)
    if np.any(flux <= 0) or np.any(err <= 0):
        raise ValueError("flux and err must be positive")
    order = np.argsort(t, kind="stable")
    t, nu, flux, err, w = t[order], nu[order], flux[order], err[order], w[order]
    if w.sum() > 0:
        w = w * (len(w) / w.sum())
    return t, nu, np.log(flux), err / flux, w


class BatchedLikelihood:
    """``log_prob_batch(samples[n, ndim]) -> logp[n]`` on the GPU.

    names / log_scale describe the sampler-space columns (log10 parameters are exponentiated like
    ``_build_transformer``, fitting/utils.py:110-135); ``template`` is a 1-record ``vag_params``
    array carrying every fixed setting (jet/medium type, distance, redshift, switches).
    """

    def __init__(self, engine, template, names: Sequence[str], log_scale: Sequence[bool], t, nu, flux, err,
                 weights=None, lower=None, upper=None, log_prior: Optional[Callable] = None,
                 log_likelihood_fn: Callable = lambda chi2: -0.5 * chi2, distributed: bool = False):
        self.engine = engine
        self.template = np.ascontiguousarray(template, dtype=abi.PARAMS_DTYPE).reshape(-1)[:1].copy()
        self.names = list(names)
        self.log_scale = np.asarray(log_scale, dtype=bool)
        for n in self.names:
            if n not in _FIELDS:
                raise ValueError(f"unknown parameter '{n}' (supported: {sorted(_FIELDS)})")
        self.t, self.nu, self.lnF, self.sig, self.w = consolidate_data(t, nu, flux, err, weights)
        self.lower = None if lower is None else np.asarray(lower, dtype=np.float64)
        self.upper = None if upper is None else np.asarray(upper, dtype=np.float64)
        self.log_prior = log_prior
        self.log_likelihood_fn = log_likelihood_fn
        self.distributed = distributed

    def to_params(self, samples: np.ndarray) -> np.ndarray:
        samples = np.atleast_2d(np.asarray(samples, dtype=np.float64))
        P = np.repeat(self.template, samples.shape[0])
        for i, name in enumerate(self.names):
            col = 10.0 ** samples[:, i] if self.log_scale[i] else samples[:, i]
            path = _FIELDS[name]
            if len(path) == 1:
                P[path[0]] = col
            else:
                P[path[0]][path[1]] = col
        return P

    def chi2(self, samples: np.ndarray) -> np.ndarray:
        P = self.to_params(samples)
        if self.distributed:
            from . import parallel

            return parallel.partitioned_chi2(self.engine, P, self.t, self.nu, self.lnF, self.sig, self.w)
        return self.engine.chi2_series(P, self.t, self.nu, self.lnF, self.sig, self.w)

    def __call__(self, samples: np.ndarray) -> np.ndarray:
        samples = np.atleast_2d(np.asarray(samples, dtype=np.float64))
        n = samples.shape[0]
        in_bounds = np.ones(n, dtype=bool)
        if self.lower is not None:
            in_bounds &= np.all(samples >= self.lower, axis=1)
        if self.upper is not None:
            in_bounds &= np.all(samples <= self.upper, axis=1)
        logp = np.full(n, -np.inf)
        idx = np.nonzero(in_bounds)[0]
        if idx.size:
            try:
                chi2 = self.chi2(samples[idx])
            except ValueError:
                # a walker outside the model's validity range (the reference maps the exception
                # of that walker to -inf, samplers.py:63-70): fall back to per-walker validation
                chi2 = np.full(idx.size, np.inf)
                ok = np.array([self._valid(samples[i:i + 1]) for i in idx])
                if ok.any():
                    chi2[ok] = self.chi2(samples[idx[ok]])
            ll = np.where(np.isfinite(chi2), self.log_likelihood_fn(chi2), -np.inf)
            ll[~np.isfinite(ll)] = -np.inf
            if self.log_prior is not None:
                ll = ll + self.log_prior(samples[idx])
            logp[idx] = ll
        return logp

    def _valid(self, sample):
        from . import _lib

        p = self.to_params(sample)
        return _lib.load().vag_params_validate(p.ctypes.data) == abi.VAG_OK
