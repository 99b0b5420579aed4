"""Log-space method-of-centres style update (LogMOC) on the device.

Semantics of reference ``core/optimizers/logmoc.py``: volume-constraint chain
dV = filter.gradient(dH * v/sum v) with the fall-back to ``filter.forward`` when
the Helmholtz adjoint's <= 0 clamp wipes it out (:128-153);
volume error from the current projected field (:160-163); dual variable by EMA
or augmented update (:185-200); dL = dC + coeff * dV (:204-205) with optional
filtering / centring (:206-215); percentile scale with EMA 0.2/0.8 (:217-225);
KKT residual on interior elements (:227-236); log-space clipped step (:36-66).

All arrays are CUDA tensors; the host sees only scalars.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Literal

import numpy as np
import torch

from sktopt._b200 import device as dev
from sktopt.core import projection
from sktopt.core.optimizers import common_density
from sktopt.tools.history import ArrayStats
from sktopt.tools.logconf import mylogger

logger = mylogger(__name__)


@dataclass
class LogMOC_Config(common_density.DensityMethod_OC_Config):
    interpolation: Literal["SIMP"] = "SIMP"
    mu_p: float = 5.0
    augmented_lagrangian_mu: float = 0.0
    lambda_v: float = 0.1
    lambda_decay: float = 0.90
    lambda_lower: float = -1e+7
    lambda_upper: float = 1e+7
    lagrangian_clip: float = 1.0
    lagrangian_percentile: float = 95.0
    lagrangian_scale_floor: float = 1e-8
    normalize_volume_chain: bool = False
    volume_chain_percentile: float = 95.0
    volume_chain_scale_floor: float = 1e-8
    volume_chain_gain: float = 1.0
    volume_chain_gain_under: float | None = None
    dual_update: Literal["ema", "augmented"] = "ema"
    filter_lagrangian: bool = False
    center_lagrangian: bool = False
    center_objective: bool = False


def lagrangian_log_update(rho, dL, scaling_rate, eta, move_limit,
                          rho_clip_lower, rho_clip_upper, rho_min, rho_max,
                          lagrangian_clip):
    """g = clip(dL, +/-clip); ln rho <- clip(ln rho - eta g, ln rho -/+
    ln(1 + move/rho)); rho = clip(exp(.), rho_min, rho_max) -- one kernel (K14);
    ``rho`` is updated in place, the other three arrays receive g and the
    log-space bounds like the reference's work buffers."""
    dev.logmoc_update(rho, dL, eta, move_limit, rho_min, rho_max,
                      lagrangian_clip, scaling_rate, rho_clip_lower,
                      rho_clip_upper)


class LogMOC_Optimizer(common_density.DensityMethod):
    def __init__(self, cfg: LogMOC_Config, tsk):
        super().__init__(cfg, tsk)
        rec = self.recorder = self.add_recorder(tsk)
        rec.add("dL", plot_type="min-max-mean-std", ylog=False)
        rec.add("-dC", plot_type="min-max-mean-std", ylog=True)
        rec.add("lambda_v", ylog=False)
        rec.add("constraint_coeff", ylog=False)
        rec.add("dV_chain", plot_type="min-max-mean-std", ylog=True)
        rec.add("volume_chain_scale", ylog=True)
        self.lambda_v = cfg.lambda_v
        self._dL_buffer = None
        self._dV_chain_design = None
        self._dV_unit_full = None
        self._dV_filtered_full = None
        self._dL_full = None

    def _weighted_mean(self, x, w, wsum):
        return dev.reduce_wsum(x, None, w) / wsum

    # -- volume chain behind the state solve ------------------------------------
    # dV = filter.gradient(dH * v / sum v), with the reference's fall-back to
    # filter.forward when the Helmholtz adjoint clamps it to zero (logmoc.py:117-122),
    # depends only on the filtered density, not on the displacements: on one GPU the
    # two filter solves (1.5 ms at C2) run on a side stream, driven by their own host
    # thread (the PCG polls its convergence flag), while the main thread runs the
    # elasticity solve.  Same kernels, same inputs: results are bit-identical.
    # SKTOPT_B200_PREFETCH_DV=0 keeps everything on one stream.
    def _ensure_volume_buffers(self, state):
        if self._dV_unit_full is None:
            self._dV_unit_full = torch.zeros_like(state.rho)
            dev.scatter(self._dV_drho_design, self._design_idx, self._dV_unit_full)
            self._dV_filtered_full = torch.zeros_like(state.rho)
            self._dL_full = torch.zeros_like(state.rho)

    def _prefetch_start(self, state, beta):
        import os
        import threading
        from sktopt._b200 import dist as bdist
        from sktopt.filters import HelmholtzFilterNodal
        self._pf = None
        if (os.environ.get("SKTOPT_B200_PREFETCH_DV", "1") == "0"
                or bdist.default_comm() is not None
                or not isinstance(self.filter, HelmholtzFilterNodal)):
            return
        self._ensure_volume_buffers(state)
        if getattr(self, "_pf_stream", None) is None:
            self._pf_stream = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        start = torch.cuda.Event()
        start.record(main)
        pf = dict(error=None, grad=None, fwd=None, done=torch.cuda.Event())

        def work():
            try:
                with torch.cuda.stream(self._pf_stream):
                    self._pf_stream.wait_event(start)
                    projection.heaviside_projection_derivative_inplace(
                        state.rho_filtered, beta=beta, eta=self.cfg.beta_eta,
                        out=self._dV_filtered_full)
                    dev.hadamard(1.0, self._dV_filtered_full, self._dV_unit_full,
                                 self._dV_filtered_full)
                    pf["grad"] = self.filter.gradient(self._dV_filtered_full)
                    pf["fwd"] = self.filter.forward(self._dV_filtered_full)   # speculative
                    pf["done"].record(self._pf_stream)
            except Exception as e:                    # re-raised at the join
                pf["error"] = e

        pf["thread"] = threading.Thread(target=work)
        pf["thread"].start()
        self._pf = pf

    def _prefetch_join(self):
        pf = getattr(self, "_pf", None)
        if pf is None:
            return
        pf["thread"].join()
        if pf["error"] is not None:
            self._pf = None
            raise pf["error"]
        torch.cuda.current_stream().wait_event(pf["done"])

    def rho_update(self, iter_num, rho_design_eles, rho_projected,
                   dC_drho_design_eles, u_dofs, strain_energy_mean,
                   scaling_rate, move_limit, eta, beta, rho_clip_lower,
                   rho_clip_upper, percentile, elements_volume_design,
                   elements_volume_design_sum, vol_frac):
        del u_dofs, strain_energy_mean
        cfg = self.cfg
        state = self._state
        if state is None:
            raise RuntimeError("Optimizer state is not initialized.")
        design = self._design_idx
        if self._dC_raw_buffer is None:
            self._dC_raw_buffer = torch.empty_like(dC_drho_design_eles)
            self._dL_buffer = torch.empty_like(dC_drho_design_eles)
            self._dV_chain_design = torch.empty_like(dC_drho_design_eles)
        # v_e / sum v on the design elements, zero elsewhere
        self._ensure_volume_buffers(state)

        self._dC_raw_buffer.copy_(dC_drho_design_eles)
        if cfg.center_objective:
            mean = self._weighted_mean(self._dC_raw_buffer, elements_volume_design,
                                       elements_volume_design_sum)
            dev.affine(1.0, self._dC_raw_buffer, 0.0, None, -mean, self._dC_raw_buffer)

        # volume chain: dH/drho~ * v/sum(v), pushed back through the filter
        pf, self._pf = getattr(self, "_pf", None), None
        if pf is not None:
            # both filter solves already ran behind the state solve (_prefetch_start)
            dV_backprop = pf["grad"]
            if dev.reduce_absmax(dV_backprop) <= 1e-8 and \
                    dev.reduce_stats(self._dV_filtered_full)[2] > 0.0:
                dV_backprop = pf["fwd"]
        else:
            projection.heaviside_projection_derivative_inplace(
                state.rho_filtered, beta=beta, eta=cfg.beta_eta,
                out=self._dV_filtered_full)
            dev.hadamard(1.0, self._dV_filtered_full, self._dV_unit_full,
                         self._dV_filtered_full)
            dV_backprop = self.filter.gradient(self._dV_filtered_full)
            if dev.reduce_absmax(dV_backprop) <= 1e-8 and \
                    dev.reduce_stats(self._dV_filtered_full)[2] > 0.0:
                # Helmholtz adjoint clamps positives to zero: fall back to forward
                dV_backprop = self.filter.forward(self._dV_filtered_full)
        dev.gather(dV_backprop, design, out=self._dV_chain_design)

        volume = dev.reduce_wsum(rho_projected, design, elements_volume_design) \
            / elements_volume_design_sum
        vol_error = volume - vol_frac

        volume_chain_scale = 1.0
        if cfg.normalize_volume_chain:
            obj_scale = max(dev.abs_percentile(self._dC_raw_buffer, cfg.lagrangian_percentile),
                            cfg.lagrangian_scale_floor)
            vol_scale = max(dev.abs_percentile(self._dV_chain_design, cfg.volume_chain_percentile),
                            cfg.volume_chain_scale_floor)
            gain = cfg.volume_chain_gain
            if cfg.volume_chain_gain_under is not None and vol_error < 0.0:
                gain = cfg.volume_chain_gain_under
            volume_chain_scale = gain * (obj_scale / vol_scale)
            dev.axpby(0.0, self._dV_chain_design, volume_chain_scale,
                      self._dV_chain_design)

        penalty = cfg.mu_p * vol_error
        if cfg.dual_update == "augmented":
            self.lambda_v = self.lambda_v + penalty if iter_num > 1 else penalty
        else:
            self.lambda_v = (cfg.lambda_decay * self.lambda_v
                             + (1.0 - cfg.lambda_decay) * penalty) if iter_num > 1 else penalty
        self.lambda_v = float(np.clip(self.lambda_v, cfg.lambda_lower, cfg.lambda_upper))
        constraint_coeff = self.lambda_v + cfg.augmented_lagrangian_mu * vol_error

        # dL = dC + coeff * dV
        dev.affine(1.0, self._dC_raw_buffer, constraint_coeff, self._dV_chain_design,
                   0.0, self._dL_buffer)
        if cfg.filter_lagrangian:
            dev.fill(self._dL_full, 0.0)
            dev.scatter(self._dL_buffer, design, self._dL_full)
            filtered = self.filter.forward(self._dL_full)
            dev.gather(filtered, design, out=self._dL_buffer)
        if cfg.center_lagrangian:
            mean = self._weighted_mean(self._dL_buffer, elements_volume_design,
                                       elements_volume_design_sum)
            dev.affine(1.0, self._dL_buffer, 0.0, None, -mean, self._dL_buffer)

        scale_percentile = percentile if isinstance(percentile, float) \
            else cfg.lagrangian_percentile
        scale = max(dev.abs_percentile(self._dL_buffer, scale_percentile),
                    cfg.lagrangian_scale_floor)
        self.running_scale = 0.2 * self.running_scale + \
            (1.0 - 0.2) * scale if iter_num > 1 else scale
        dev.axpby(0.0, self._dL_buffer, 1.0 / self.running_scale, self._dL_buffer)

        res, n_int = dev.kkt_residual(rho_design_eles, self._dL_buffer, None, 0.0,
                                      cfg.rho_min + 1e-6, cfg.rho_max - 1e-6)
        self.kkt_residual = float(res) if n_int > 0 else 0.0

        rec = self.recorder
        rec.feed_data("lambda_v", self.lambda_v)
        rec.feed_data("constraint_coeff", constraint_coeff)
        rec.feed_data("vol_error", vol_error)
        rec.feed_data("-dC", ArrayStats(*dev.reduce_stats(self._dC_raw_buffer)).negated())
        # |dV_chain|: the chain is single-signed in practice; statistics of |x|
        mn, mean, mx, sd = dev.reduce_stats(self._dV_chain_design)
        if mn >= 0.0:
            rec.feed_data("dV_chain", ArrayStats(mn, mean, mx, sd))
        elif mx <= 0.0:
            rec.feed_data("dV_chain", ArrayStats(-mx, -mean, -mn, sd))
        else:
            rec.feed_data("dV_chain", dev.absval(self._dV_chain_design))
        rec.feed_data("volume_chain_scale", volume_chain_scale)
        rec.feed_data("dL", self._dL_buffer)
        rec.feed_data("kkt_residual", self.kkt_residual)

        lagrangian_log_update(
            rho_design_eles, self._dL_buffer, scaling_rate, eta, move_limit,
            rho_clip_lower, rho_clip_upper, cfg.rho_min, cfg.rho_max,
            cfg.lagrangian_clip)
