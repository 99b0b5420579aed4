"""Optimality-criteria update with bisection on the *physical* volume.

Semantics of reference ``core/optimizers/oc.py``: sensitivity scale (max or
percentile, floored, EMA 0.6/0.4 after iteration 1, :187-196); per bisection
step the candidate rho_c = clip(rho * clip((-dC/(lam+eps))^eta, smin, smax),
rho -/+ move) (:50-65) is filtered and projected and the design-volume error
decides the bracket (:67-93); exits on |vol_err| < vol_tol, max_iter, or
|l2-l1| <= tolerance (:81-86); KKT residual on interior elements (:230-240).

Device version: one fused candidate kernel (K12), the filter's device apply,
the projection kernel and one single-kernel volume reduction (K13) per step;
only the scalar volume error returns to the host, which keeps the reference's
branching bit-for-bit.
"""
from __future__ import annotations

import os
from contextlib import contextmanager
from dataclasses import dataclass, field
from typing import Literal

import numpy as np
import torch

import sktopt
from sktopt._b200 import device as dev
from sktopt.core import projection
from sktopt.core.optimizers import common_density
from sktopt.tools.history import ArrayStats
from sktopt.tools.logconf import mylogger

logger = mylogger(__name__)


def rates_all_on_lower_clip(neg_dc_max: float, lmid: float, eps: float, eta: float,
                            scaling_rate_min: float) -> bool:
    """True when (-dC_e / (lmid + eps))^eta <= scaling_rate_min for EVERY element, i.e.
    the OC candidate does not depend on ``lmid`` any more (all scaling rates sit on
    their lower clip; elements with -dC_e < 0 give NaN for any lmid).  ``neg_dc_max`` =
    max_e(-dC_e).  A relative margin of 1e-9 on both sides keeps the answer False
    anywhere near the boundary, where ``pow`` rounding could differ by an ulp, so a
    True answer means the candidate is bit-identical to any other such midpoint's."""
    if not (eta > 0.0 and scaling_rate_min > 0.0 and neg_dc_max >= 0.0):
        return False
    bound = float(scaling_rate_min) ** (1.0 / float(eta)) * (lmid + eps)
    return bool(neg_dc_max * (1.0 + 1e-9) <= bound * (1.0 - 1e-9))


def bisection_with_physical_volume(
    dC, rho_e, rho_full, design_elements, filter_obj, rho_min, rho_max,
    move_limit, eta, eps, vol_frac, beta, beta_eta, scaling_rate,
    rho_design_eles, rho_clip_lower, rho_clip_upper, elements_volume,
    elements_volume_sum, scaling_rate_min, scaling_rate_max,
    rho_full_candidate, rho_filtered_candidate, rho_projected_candidate,
    max_iter: int = 100, tolerance: float = 1e-4, vol_tol: float = 1e-4,
    l1: float = 1e-7, l2: float = 1e+7,
):
    """Bisection on the Lagrange multiplier; all arrays are CUDA tensors,
    ``design_elements`` an int32 index tensor.  Returns (lmid, vol_error).

    ``rho_clip_lower`` / ``rho_clip_upper`` are kept in the signature for
    parity with the reference; the clip bounds are formed inside the kernel.
    """
    del rho_clip_lower, rho_clip_upper
    # non-design entries of the candidate never change during the bisection
    rho_full_candidate.copy_(rho_full)
    iter_num = 0
    lmid = 0.5 * (l1 + l2)
    vol_error = 0.0
    steps = 0
    # Two shortcuts that leave every evaluated quantity as it is:
    # * The bracket starts at [1e-7, 1e7]: for the first ~20 midpoints EVERY scaling
    #   rate (-dC/lmid)^eta sits on its lower clip, so the candidate -- and with it
    #   the filtered / projected field and the volume error -- is bit-identical to
    #   the one just evaluated.  max(-dC) tells that on the host (with a relative
    #   margin of 1e-9 on both sides of the comparison); such a step is counted and
    #   takes its bisection decision, but launches nothing.
    # * candidate = clip(rho_e c_e t) with t = (lmid + eps)^-eta is piecewise LINEAR
    #   in t and the filter is linear, so the filter's solution is too: filters that
    #   accept a ``hint`` start their solve from the secant through the last two
    #   solutions in t (exact while no element changes its clip status).  Their
    #   solves also stop at the filter's bisection tolerance (1e-9): the filtered
    #   candidate only decides the sign of a volume error against 1e-4 thresholds and
    #   never enters the next iteration (the accepted design is filtered again, at the
    #   filter's full tolerance, when the iteration starts).
    neg_dc_max = -dev.reduce_stats(dC)[0]
    hinted = bool(getattr(filter_obj, "accepts_hint", False))
    if hinted:
        filter_obj.reset_hint()          # the last bisection's secant is stale
    last_eval_saturated = None
    evaluated = 0
    while True:
        steps += 1
        saturated = rates_all_on_lower_clip(neg_dc_max, lmid, eps, float(eta),
                                            float(scaling_rate_min))
        if not (saturated and last_eval_saturated):
            dev.oc_candidate(dC, rho_e, lmid, eps, eta, move_limit, rho_min, rho_max,
                             scaling_rate_min, scaling_rate_max, design_elements,
                             scaling_rate, rho_design_eles, rho_full_candidate)
            if hinted:
                filter_obj.forward(rho_full_candidate, out=rho_filtered_candidate,
                                   hint=(lmid + eps) ** (-float(eta)),
                                   rtol=filter_obj.bisection_rtol)
            else:
                filter_obj.forward(rho_full_candidate, out=rho_filtered_candidate)
            projection.heaviside_projection_inplace(
                rho_filtered_candidate, beta=beta, eta=beta_eta,
                out=rho_projected_candidate)
            vol_error = dev.reduce_wsum(
                rho_projected_candidate, design_elements, elements_volume
            ) / elements_volume_sum - vol_frac
            last_eval_saturated = saturated
            evaluated += 1

        if abs(vol_error) < vol_tol:
            break
        if iter_num >= max_iter:
            break
        if abs(l2 - l1) <= tolerance:
            break
        if vol_error > 0:
            l1 = lmid
        else:
            l2 = lmid
        iter_num += 1
        lmid = 0.5 * (l1 + l2)
    # candidate evaluations of this call (the return value keeps the reference's
    # two-tuple); read by OC_Optimizer.rho_update for its bisection_steps log
    bisection_with_physical_volume.last_steps = steps
    bisection_with_physical_volume.last_evaluated = evaluated
    return lmid, vol_error


bisection_with_physical_volume.last_steps = 0
bisection_with_physical_volume.last_evaluated = 0


@dataclass
class OC_Config(common_density.DensityMethod_OC_Config):
    interpolation: Literal["SIMP"] = "SIMP"
    eta: sktopt.tools.SchedulerConfig = field(
        default_factory=lambda: sktopt.tools.SchedulerConfig.constant(target_value=0.5))
    scaling_rate_min: float = 0.7
    scaling_rate_max: float = 1.3
    sensitivity_scale_floor: float = 1e-12


class OC_Optimizer(common_density.DensityMethod):
    def __init__(self, cfg: OC_Config, tsk):
        assert cfg.lambda_lower < cfg.lambda_upper
        super().__init__(cfg, tsk)
        self.recorder = self.add_recorder(tsk)
        self.recorder.add("-dC", plot_type="min-max-mean-std", ylog=True)
        self.recorder.add("lmid", ylog=True)
        self.running_scale = 0
        self.bisection_steps = []

    def init_schedulers(self, export: bool = True):
        super().init_schedulers(False)
        if export:
            self.schedulers.export()

    @contextmanager
    def _scaled_sensitivity_mode(self):
        previous = os.environ.get("SCITOPT_SENSITIVITY_MODE")
        os.environ["SCITOPT_SENSITIVITY_MODE"] = "current"
        try:
            yield
        finally:
            if previous is None:
                os.environ.pop("SCITOPT_SENSITIVITY_MODE", None)
            else:
                os.environ["SCITOPT_SENSITIVITY_MODE"] = previous

    def optimize(self):
        with self._scaled_sensitivity_mode():
            super().optimize()

    def optimize_steps(self, num_steps: int):
        with self._scaled_sensitivity_mode():
            super().optimize_steps(num_steps)

    def rho_update(self, iter_num, rho_design_eles, rho_projected,
                   dC_drho_design_eles, u_dofs, strain_energy_mean,
                   scaling_rate, move_limit, eta, beta, rho_clip_lower,
                   rho_clip_upper, percentile, elements_volume_design,
                   elements_volume_design_sum, vol_frac):
        del rho_projected, u_dofs, strain_energy_mean
        cfg = self.cfg
        state = self._state
        if state is None:
            raise RuntimeError("Optimizer state is not initialized.")
        if self._rho_e_buffer is None:
            self._rho_e_buffer = torch.empty_like(rho_design_eles)
            self._dC_raw_buffer = torch.empty_like(dC_drho_design_eles)
        if not hasattr(self, "_rho_full_candidate"):
            self._rho_full_candidate = torch.empty_like(state.rho)
            self._rho_filtered_candidate = torch.empty_like(state.rho)
            self._rho_projected_candidate = torch.empty_like(state.rho)

        with self._timed_section("copy_buffers"):
            self._dC_raw_buffer.copy_(dC_drho_design_eles)
            self._rho_e_buffer.copy_(rho_design_eles)

        eps = 1e-12
        with self._timed_section("percentile_scale"):
            if isinstance(percentile, float):
                scale = dev.abs_percentile(dC_drho_design_eles, percentile)
            else:
                scale = dev.reduce_absmax(dC_drho_design_eles)
            scale = max(scale, cfg.sensitivity_scale_floor)
            self.running_scale = 0.6 * self.running_scale + \
                (1 - 0.6) * scale if iter_num > 1 else scale
            # dC /= running_scale
            dev.axpby(0.0, dC_drho_design_eles, 1.0 / self.running_scale,
                      dC_drho_design_eles)
            kkt_scale = self.running_scale

        with self._timed_section("bisection"):
            lmid, vol_error = bisection_with_physical_volume(
                dC_drho_design_eles, self._rho_e_buffer, state.rho,
                self._design_idx, self.filter, cfg.rho_min, cfg.rho_max,
                move_limit, eta, eps, vol_frac, beta, cfg.beta_eta,
                scaling_rate, rho_design_eles, rho_clip_lower, rho_clip_upper,
                elements_volume_design, elements_volume_design_sum,
                cfg.scaling_rate_min, cfg.scaling_rate_max,
                self._rho_full_candidate, self._rho_filtered_candidate,
                self._rho_projected_candidate,
                max_iter=1000, tolerance=1e-5,
                l1=cfg.lambda_lower, l2=cfg.lambda_upper,
            )
            self.bisection_steps.append(bisection_with_physical_volume.last_steps)

        with self._timed_section("kkt"):
            res, n_int = dev.kkt_residual(
                rho_design_eles, self._dC_raw_buffer, self._dV_drho_design,
                lmid * kkt_scale, cfg.rho_min + 1e-6, cfg.rho_max - 1e-6)
            self.kkt_residual = float(res) if n_int > 0 else 0.0

        self.recorder.feed_data("lmid", lmid)
        self.recorder.feed_data("vol_error", vol_error)
        self.recorder.feed_data(
            "-dC", ArrayStats(*dev.reduce_stats(dC_drho_design_eles)).negated())
        self.recorder.feed_data("kkt_residual", self.kkt_residual)
