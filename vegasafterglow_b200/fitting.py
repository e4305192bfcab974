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


def band_obs(t, flux, err, nu_min, nu_max, num_points=5, weights=None):
    """One band-integrated data set in the form ``Engine.chi2`` takes (fitter.py BandObs + :525-531: ln flux,
    relative error, weights as given -- the reference normalises only the point-data weights)."""
    t, flux, err = (np.asarray(a, dtype=np.float64).reshape(-1) for a in (t, flux, err))
    w = np.ones_like(t) if weights is None else np.asarray(weights, dtype=np.float64).reshape(-1)
    if np.any(flux <= 0) or np.any(err <= 0):
        raise ValueError("the log-flux likelihood requires strictly positive fluxes and errors")
    order = np.argsort(t, kind="stable")
    return {"t": t[order], "lnF_obs": np.log(flux[order]), "sigma_ln": err[order] / flux[order], "w": w[order],
            "nu_min": float(nu_min), "nu_max": float(nu_max), "num_nu": int(num_points)}


class BatchedLikelihood:
    """``log_prob_batch(samples[n, ndim]) -> logp[n]`` on the GPU.

    names / log_scale describe the sampler-space columns (log10 parameters are exponentiated like
    ``_build_transformer``, fitting/utils.py:110-135); ``template`` is a 1-record ``vag_params``
    array carrying every fixed setting (jet/medium type, distance, redshift, switches).
    """

    def __init__(self, engine, template, names: Sequence[str], log_scale: Sequence[bool], t, nu, flux, err,
                 weights=None, lower=None, upper=None, log_prior: Optional[Callable] = None,
                 log_likelihood_fn: Callable = lambda chi2: -0.5 * chi2, distributed: bool = False, bands=()):
        self.engine = engine
        self.template = np.ascontiguousarray(template, dtype=abi.PARAMS_DTYPE).reshape(-1)[:1].copy()
        self.names = list(names)
        self.log_scale = np.asarray(log_scale, dtype=bool)
        for n in self.names:
            if n not in _FIELDS:
                raise ValueError(f"unknown parameter '{n}' (supported: {sorted(_FIELDS)})")
        if len(np.atleast_1d(t)):
            self.t, self.nu, self.lnF, self.sig, self.w = consolidate_data(t, nu, flux, err, weights)
        else:
            self.t = self.nu = self.lnF = self.sig = self.w = np.zeros(0)
        # band-integrated data sets (Fitter.add_flux -> BandObs; fitter.py:525-531 evaluates Model.flux on each)
        self.bands = [band_obs(**b) for b in bands]
        if not self.t.size and not self.bands:
            raise ValueError("no data")
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

    def _evaluate(self, P: np.ndarray) -> np.ndarray:
        if not self.bands:
            return self.engine.chi2_series(P, self.t, self.nu, self.lnF, self.sig, self.w)
        pts = (self.t, self.nu, self.lnF, self.sig, self.w) if self.t.size else None
        return self.engine.chi2(P, pts, self.bands)

    def chi2(self, samples: np.ndarray) -> np.ndarray:
        """chi2 of every sample; +inf for a sample the model constructor rejects (the reference maps that exception to
        logL = -inf, samplers.py:63-70).  The validity mask is computed from the parameters alone, identically on every
        rank, BEFORE the ensemble is split: a rejected walker can therefore never desynchronise the collective."""
        from .engine import Engine

        P = self.to_params(samples)
        ok = Engine.valid_mask(P)
        if self.distributed:
            from . import parallel

            tt = self.t if self.t.size else np.concatenate([b["t"] for b in self.bands])
            return parallel.partitioned_chi2(self.engine, P, tt, self.nu, self.lnF, self.sig, self.w, evaluate=self._evaluate,
                                             valid=ok)
        out = np.full(P.size, np.inf)
        if ok.any():
            out[ok] = self._evaluate(P[ok])
        return out

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
            chi2 = self.chi2(samples[idx])
            ll = np.where(np.isfinite(chi2), self.log_likelihood_fn(chi2), -np.inf)
            ll[~np.isfinite(ll)] = -np.inf
            if self.log_prior is not None:
                ll = ll + self.log_prior(samples[idx])
            logp[idx] = ll
        return logp

