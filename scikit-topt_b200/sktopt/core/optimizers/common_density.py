"""Density-method driver with a device-resident iteration loop.

Same configuration dataclasses, state layout, schedules and per-iteration
control flow as reference ``core/optimizers/common_density.py`` (config
:62-261, :348-389; ``DensityState`` :32-59; loop ``_optimize_impl`` :966-1243):

    filter -> project -> assemble K(rho) -> solve -> element energy ->
    dC/drho (chain through projection) -> filter adjoint -> rho_update

but every array of ``DensityState`` is a CUDA fp64 tensor that never leaves the
GPU between iterations; only scalars (compliance, volume error, recorder
statistics) and the checkpoint written on export ticks cross to the host.
"""
from __future__ import annotations

import inspect
import json
import os
import shutil
from abc import ABC, abstractmethod
from dataclasses import asdict, dataclass, field
from typing import Literal

import numpy as np
import torch

import sktopt
from sktopt import fea, filters, tools
from sktopt._b200 import device as dev
from sktopt._b200 import dist as bdist
from sktopt.core import derivatives, misc, projection, visualization
from sktopt.fea import composer
from sktopt.fea._petsc_compat import (
    PETScOptions, normalize_petsc_options, petsc_options_for_solver,
)
from sktopt.fea.solver_elastic import LinearSolverConfig, normalize_linear_solver_config
from sktopt.tools.history import ArrayStats
from sktopt.tools.logconf import mylogger

logger = mylogger(__name__)


@dataclass
class DensityState:
    """Work arrays of the loop (CUDA tensors) and the latest scalars."""
    rho: torch.Tensor
    rho_prev: torch.Tensor
    rho_filtered: torch.Tensor
    rho_projected: torch.Tensor
    dH_drho: torch.Tensor
    grad_filtered: torch.Tensor
    dC_drho_projected: torch.Tensor
    energy_mean: torch.Tensor
    dC_drho_full: torch.Tensor
    dC_drho_design_eles: torch.Tensor
    scaling_rate: torch.Tensor
    rho_design_eles: torch.Tensor
    rho_clip_lower: torch.Tensor
    rho_clip_upper: torch.Tensor
    u_dofs: torch.Tensor
    filter_radius: float
    elements_volume_design: torch.Tensor
    elements_volume_design_sum: float
    iter_begin: int
    iter_end: int
    last_iter: int | None = None
    compliance: float | None = None
    u_max: float | np.ndarray | None = None
    rho_change_max: float | None = None
    kkt_residual: float | None = None
    vol_error: float | None = None


@dataclass
class DensityMethodConfig():
    """Numerical settings shared by the OC / MOC style optimisers (same fields
    and defaults as the reference, ``common_density.py:208-261``)."""

    dst_path: str = "./result/pytests"
    interpolation: Literal["SIMP", "RAMP"] = "SIMP"
    record_times: int = 20
    max_iters: int = 200
    beta_eta: float = 0.50
    eta: float = 0.6
    p: tools.SchedulerConfig = field(
        default_factory=lambda: tools.SchedulerConfig.step(
            init_value=1.0, target_value=3.0, num_steps=3))
    vol_frac: tools.SchedulerConfig = field(
        default_factory=lambda: tools.SchedulerConfig.constant(target_value=0.8))
    beta: tools.SchedulerConfig = field(
        default_factory=lambda: tools.SchedulerConfig.step_accelerating(
            init_value=1.0, target_value=2.0, num_steps=3, curvature=2.0))
    neumann_scale: tools.SchedulerConfig = field(
        default_factory=lambda: tools.SchedulerConfig.constant_one(name="neumann_scale"))
    filter_type: Literal["spacial", "helmholtz"] = "helmholtz"
    filter_radius: tools.SchedulerConfig = field(
        default_factory=lambda: tools.SchedulerConfig.constant(target_value=0.01))
    E_min_coeff: float = 1e-3
    rho_min: float = 1e-2
    rho_max: float = 1.0
    restart: bool = False
    restart_from: int = -1
    export_img: bool = False
    export_img_opaque: bool = False
    design_dirichlet: bool = False
    sensitivity_filter: bool = False
    solver_option: Literal["spsolve", "cg_pyamg", "petsc", "petsc_spdirect"] = "spsolve"
    petsc_options: PETScOptions = field(default_factory=PETScOptions)
    scaling: bool = False
    check_convergence: bool = False
    tol_rho_change: float = 2e-1
    tol_kkt_residual: float = 5e-3

    @classmethod
    def from_defaults(cls, **args) -> 'DensityMethodConfig':
        known = inspect.signature(cls).parameters.keys()
        return cls(**{k: v for k, v in args.items() if k in known})

    def __post_init__(self):
        if self.solver_option in ("petsc", "petsc_spdirect"):
            self.petsc_options = petsc_options_for_solver(
                self.solver_option, self.petsc_options)
        else:
            self.petsc_options = normalize_petsc_options(self.petsc_options)
        self.solver_config = normalize_linear_solver_config(
            self.solver_option, petsc_options=self.petsc_options)

    @classmethod
    def import_from(cls, path: str) -> 'DensityMethodConfig':
        with open(f"{path}/cfg.json", "r") as f:
            data = json.load(f)
        data.pop("record_timing", None)
        data.pop("solver_config", None)
        for k, v in list(data.items()):
            if isinstance(v, dict) and "scheduler_type" in v:
                data[k] = tools.SchedulerConfig(**v)
            elif k == "petsc_options" and isinstance(v, dict):
                data[k] = PETScOptions(**v)
        return cls(**data)

    def export(self, path: str):
        with open(f"{path}/cfg.json", "w") as f:
            json.dump(asdict(self), f, indent=2)

    def vtu_path(self, iter_num: int):
        return f"{self.dst_path}/mesh_rho/info_mesh-{iter_num:08d}.vtu"

    def image_path(self, iter_num: int, prefix: str):
        if self.export_img:
            return f"{self.dst_path}/mesh_rho/info_{prefix}-{iter_num:08d}.jpg"
        return None


@dataclass
class DensityMethod_OC_Config(DensityMethodConfig):
    """Adds the OC-family knobs (``common_density.py:380-389``)."""

    lambda_lower: float = 1e-7
    lambda_upper: float = 1e+7
    percentile: tools.SchedulerConfig = field(
        default_factory=lambda: tools.SchedulerConfig.none())
    move_limit: tools.SchedulerConfig = field(
        default_factory=lambda: tools.SchedulerConfig.sawtooth_decay(
            "move_limit", 0.3, 0.1, 3))


def interpolation_funcs(cfg: DensityMethodConfig):
    if cfg.interpolation == "SIMP":
        return [composer.simp_interpolation, derivatives.dC_drho_simp]
    if cfg.interpolation == "RAMP":
        return [composer.ramp_interpolation, derivatives.dC_drho_ramp]
    raise ValueError("Interpolation method must be SIMP or RAMP.")


class DensityMethodBase(ABC):
    @abstractmethod
    def add_recorder(self):
        pass

    @abstractmethod
    def scale(self):
        pass

    @abstractmethod
    def unscale(self):
        pass

    @abstractmethod
    def load_parameters(self):
        pass

    @abstractmethod
    def init_schedulers(self, export: bool = True):
        pass

    @abstractmethod
    def parameterize(self):
        pass

    @abstractmethod
    def initialize_density(self):
        pass

    @abstractmethod
    def initialize_params(self):
        pass

    @abstractmethod
    def optimize(self):
        pass

    @abstractmethod
    def rho_update(self, iter_num, rho_design_eles, rho_projected,
                   dC_drho_design_eles, u_dofs, energy_mean, scaling_rate,
                   move_limit, eta, beta, rho_clip_lower, rho_clip_upper,
                   percentile, elements_volume_design,
                   elements_volume_design_sum, vol_frac):
        pass


def _idx(a):
    return dev.to_dev(np.asarray(a, dtype=np.int64), dev.I32)


def _zeros(n):
    return torch.zeros(n, dtype=dev.F64, device="cuda")


class DensityMethod(DensityMethodBase):
    """Backbone of the sensitivity-based density optimisers; subclasses provide
    ``rho_update`` (OC bisection, log-space MOC)."""

    def __init__(self, cfg: DensityMethodConfig, tsk):
        dev.require_cuda()
        self.cfg = cfg
        self.tsk = tsk
        self.timer = tools.SectionTimer(hierarchical=True)
        if cfg.scaling is True:
            self.scale()
        # one process per GPU: every rank holds the same state, rank 0 alone owns
        # the run directory (set-up, cfg.json, checkpoints, histories)
        self._io = bdist.is_io_rank()
        if self._io:
            os.makedirs(cfg.dst_path, exist_ok=True)
            cfg.export(cfg.dst_path)
            if cfg.restart is not True:
                shutil.rmtree(f"{cfg.dst_path}/mesh_rho", ignore_errors=True)
                os.makedirs(f"{cfg.dst_path}/mesh_rho", exist_ok=True)
                os.makedirs(f"{cfg.dst_path}/data", exist_ok=True)
        bdist.io_barrier()
        if cfg.design_dirichlet is False:
            tsk.exlude_dirichlet_from_design()
        if cfg.restart is True:
            self.load_parameters()

        interp = interpolation_funcs(cfg)[0]
        if isinstance(tsk, sktopt.mesh.LinearElasticity):
            self.fem = fea.FEM_SimpLinearElasticity(
                tsk, cfg.E_min_coeff, density_interpolation=interp,
                solver_config=cfg.solver_config)
        elif isinstance(tsk, sktopt.mesh.LinearHeatConduction):
            self.fem = fea.FEM_SimpLinearHeatConduction(
                tsk, cfg.E_min_coeff, density_interpolation=interp,
                solver_config=cfg.solver_config)
        else:
            raise NotImplementedError("")
        self.schedulers = tools.Schedulers(cfg.dst_path)
        self._schedulers_initialized = False

        self._rho_e_buffer = None
        self._dC_raw_buffer = None
        vol_design = tsk.elements_volume[tsk.design_elements]
        self._dV_drho_design = dev.to_dev(vol_design / np.sum(vol_design))
        self.kkt_residual = None
        self._state: DensityState | None = None
        self._iter_next: int | None = None
        self._iter_end: int | None = None
        self._completed = False
        # export ticks (checkpoint npz, recorder print) can be switched off by a
        # driver that times the loop itself (metric excludes them)
        self.export_enabled = True
        # device index sets (fixed once the design set is final)
        self._design_idx = _idx(tsk.design_elements)
        pin = tsk.neumann_elements if cfg.design_dirichlet else tsk.dirichlet_neumann_elements
        self._pin_idx = _idx(pin) if pin is not None and len(pin) else None
        self._pin_ones = (torch.ones(self._pin_idx.numel(), dtype=dev.F64, device="cuda")
                          if self._pin_idx is not None else None)

    # ------------------------------------------------------------ recorder
    def add_recorder(self, tsk) -> tools.HistoryCollection:
        rec = tools.HistoryCollection(self.cfg.dst_path)
        rec.add("rho_projected", plot_type="min-max-mean-std")
        rec.add("energy", plot_type="min-max-mean-std")
        rec.add("vol_error")
        if tsk.n_tasks > 1:
            rec.add("u_max", plot_type="min-max-mean-std")
        else:
            rec.add("u_max")
        rec.add(self._objective_history_name(tsk), ylog=self._objective_history_ylog(tsk))
        rec.add("scaling_rate", plot_type="min-max-mean-std")
        rec.add("neumann_scale")
        rec.add("rho_change_max")
        rec.add("kkt_residual")
        return rec

    def _objective_history_name(self, tsk) -> str:
        if isinstance(tsk, sktopt.mesh.LinearHeatConduction):
            return str(tsk.objective)
        return "compliance"

    def _objective_history_ylog(self, tsk) -> bool:
        if isinstance(tsk, sktopt.mesh.LinearHeatConduction):
            return tsk.objective == "compliance"
        return True

    def params_latest(self):
        return self.recorder.as_object_latest()

    # ------------------------------------------------------------- scaling
    def scale(self):
        self.L_scale = np.max(np.ptp(self.tsk.mesh.p, axis=1))
        self.F_scale = 10 ** 5
        self.tsk.scale(1.0 / self.L_scale, 1.0 / self.F_scale)

    def unscale(self):
        self.tsk.scale(self.L_scale, self.F_scale)

    # ---------------------------------------------------------- schedulers
    def init_schedulers(self, export: bool = True):
        cfg = self.cfg
        for sc, name in ((cfg.p, "p"), (cfg.vol_frac, "vol_frac"),
                         (cfg.move_limit, "move_limit"), (cfg.beta, "beta"),
                         (cfg.percentile, "percentile"),
                         (cfg.filter_radius, "filter_radius")):
            self.schedulers.add_object_from_config(sc, name)
        allowed = {"ConstantOne", "StepToOne", "StepAcceleratingToOne",
                   "StepDeceleratingToOne"}
        if cfg.neumann_scale.scheduler_type not in allowed:
            raise ValueError(f"neumann_scale must use one of {allowed}")
        self.schedulers.add_object_from_config(cfg.neumann_scale, "neumann_scale")
        if isinstance(cfg.eta, tools.SchedulerConfig):
            self.schedulers.add_object_from_config(cfg.eta, "eta")
        else:
            self.schedulers.add("eta", cfg.eta, cfg.eta, -1, cfg.max_iters)
        self.schedulers.set_iters_max(cfg.max_iters)
        if export and getattr(self, "_io", True):
            self.schedulers.export()
        self._schedulers_initialized = True

    def parameterize(self):
        kinds = {"spacial": filters.SpacialFilter,
                 "helmholtz": filters.HelmholtzFilterNodal}
        if self.cfg.filter_type not in kinds:
            raise ValueError("should be spacial or helmholtz")
        self.filter = kinds[self.cfg.filter_type].from_defaults(
            self.tsk.mesh, self.tsk.elements_volume,
            self.cfg.filter_radius.init_value, design_mask=self.tsk.design_mask)

    def load_parameters(self):
        pass

    # --------------------------------------------------------------- state
    def initialize_density(self):
        """Host initial density (``common_density.py:711-745``)."""
        tsk, cfg = self.tsk, self.cfg
        val_init = cfg.vol_frac.init_value \
            if cfg.vol_frac.init_value is not None else cfg.vol_frac.target_value
        rho = np.zeros_like(tsk.all_elements, dtype=np.float64)
        iter_begin = 1
        iter_end = cfg.max_iters + 1
        if cfg.restart is True:
            if cfg.restart_from > 0:
                path = f"{cfg.dst_path}/data/{cfg.restart_from:06d}-rho.npz"
                iter_begin = cfg.restart_from + 1
            else:
                it, path = misc.find_latest_iter_file(f"{cfg.dst_path}/data")
                iter_begin = it + 1
            self.recorder.import_histories()
            with np.load(path) as data:
                rho[tsk.design_elements] = data["rho_design_elements"]
        else:
            rho += val_init
            np.clip(rho, cfg.rho_min, cfg.rho_max, out=rho)
        if cfg.design_dirichlet is True:
            rho[tsk.neumann_elements] = 1.0
        else:
            rho[tsk.dirichlet_neumann_elements] = 1.0
        rho[tsk.fixed_elements] = 1.0
        return rho, iter_begin, iter_end

    def initialize_params(self):
        """Allocates the device work arrays (``common_density.py:747-843``)."""
        tsk, cfg = self.tsk, self.cfg
        rho_h, iter_begin, iter_end = self.initialize_density()
        ne, nd = rho_h.size, int(self._design_idx.numel())
        rho = dev.to_dev(rho_h)
        u_dofs = torch.zeros((tsk.n_tasks, tsk.basis.N), dtype=dev.F64, device="cuda").t()
        filter_radius = cfg.filter_radius.init_value \
            if isinstance(cfg.filter_radius.num_steps, (int, float)) \
            else cfg.filter_radius.target_value
        return (iter_begin, iter_end, rho, _zeros(ne), _zeros(ne), _zeros(ne),
                _zeros(ne), _zeros(ne), _zeros(ne), _zeros(ne), _zeros(ne),
                _zeros(nd), _zeros(nd), _zeros(nd), _zeros(nd), _zeros(nd),
                u_dofs, filter_radius)

    def _ensure_state_initialized(self):
        if self._completed:
            logger.info("Optimization already completed; skipping.")
            return False
        if not self._schedulers_initialized:
            self.init_schedulers()
        if self._state is None:
            (iter_begin, iter_end, rho, rho_prev, rho_filtered, rho_projected,
             dH_drho, grad_filtered, dC_drho_projected, energy_mean,
             dC_drho_full, dC_drho_design_eles, scaling_rate, rho_design_eles,
             rho_clip_lower, rho_clip_upper, u_dofs, filter_radius
             ) = self.initialize_params()
            vol_design_h = self.tsk.elements_volume[self.tsk.design_elements]
            self.filter.update_radius(filter_radius)
            self._state = DensityState(
                rho=rho, rho_prev=rho_prev, rho_filtered=rho_filtered,
                rho_projected=rho_projected, dH_drho=dH_drho,
                grad_filtered=grad_filtered,
                dC_drho_projected=dC_drho_projected, energy_mean=energy_mean,
                dC_drho_full=dC_drho_full,
                dC_drho_design_eles=dC_drho_design_eles,
                scaling_rate=scaling_rate, rho_design_eles=rho_design_eles,
                rho_clip_lower=rho_clip_lower, rho_clip_upper=rho_clip_upper,
                u_dofs=u_dofs, filter_radius=filter_radius,
                elements_volume_design=dev.to_dev(vol_design_h),
                elements_volume_design_sum=float(np.sum(vol_design_h)),
                iter_begin=iter_begin, iter_end=iter_end)
            self._iter_next = iter_begin
            self._iter_end = iter_end
        return True

    def _timed_section(self, name: str):
        return self.timer.section(name)

    def _report_timing(self):
        self.timer.report(logger_instance=logger)

    # --------------------------------------------------------------- public
    def optimize(self):
        """Run until ``cfg.max_iters``."""
        self._optimize_impl()

    def optimize_steps(self, num_steps: int):
        """Run only ``num_steps`` further iterations."""
        if num_steps <= 0:
            logger.info("optimize_steps called with non-positive num_steps; skipping.")
            return
        self._optimize_impl(num_steps)

    def _finalize(self):
        if self._completed or self._state is None:
            return
        if self.cfg.scaling is True:
            self.unscale()
        if self._io:
            self.recorder.export_histories(fname="histories.npz")
        self._completed = True

    def _export_iteration(self, iter_num, state, energy_mean):
        cfg = self.cfg
        self.recorder.print()
        self.recorder.export_progress()
        # info_mesh-XXXXXXXX.vtu with the cell fields the reference writes
        # (common_density.py:1199-1204); SKTOPT_EXPORT_VTU=0 skips it
        if os.environ.get("SKTOPT_EXPORT_VTU", "1") != "0":
            visualization.export_mesh_with_info(
                self.tsk.mesh, cell_data_names=["rho_projected", "energy"],
                cell_data_values=[state.rho_projected.cpu().numpy(),
                                  energy_mean.cpu().numpy()],
                filepath=cfg.vtu_path(iter_num))
        rho_design = dev.gather(state.rho, self._design_idx).cpu().numpy()
        np.savez_compressed(
            f"{cfg.dst_path}/data/{str(iter_num).zfill(6)}-rho.npz",
            rho_design_elements=rho_design)

    # ----------------------------------------------------------------- loop
    def _optimize_impl(self, max_steps: int | None = None):
        tsk, cfg = self.tsk, self.cfg
        if not getattr(self, "_condition_exported", False):
            # the reference rewrites this file on every call; once is enough
            if self._io:
                tsk.export_analysis_condition_on_mesh(cfg.dst_path)
            self._condition_exported = True
        if not self._ensure_state_initialized():
            return
        _, dC_drho_func = interpolation_funcs(cfg)
        ramp = cfg.interpolation == "RAMP"
        st = self._state
        design = self._design_idx
        n_tasks = tsk.n_tasks
        c_max = tsk.material_coef
        c_min = tsk.material_coef * cfg.E_min_coeff

        iter_start = self._iter_next if self._iter_next is not None else st.iter_begin
        if iter_start >= st.iter_end:
            self._finalize()
            return
        iter_limit = st.iter_end if max_steps is None \
            else min(iter_start + max_steps, st.iter_end)

        conv_rho = conv_kkt = converged = False
        iter_num = None
        for iter_num in range(iter_start, iter_limit):
            (neumann_scale, p, vol_frac, beta, move_limit, eta, percentile,
             filter_radius) = self.schedulers.values_as_list(
                iter_num,
                ['neumann_scale', 'p', 'vol_frac', 'beta', 'move_limit',
                 'eta', 'percentile', 'filter_radius'],
                export_log=True, precision=6)
            st.last_iter = iter_num
            if filter_radius != self.filter.radius:
                self.filter.update_radius(filter_radius)

            with self._timed_section("filter_and_project"):
                st.rho_prev.copy_(st.rho)
                self.filter.forward(st.rho, out=st.rho_filtered)
                projection.heaviside_projection_inplace(
                    st.rho_filtered, beta=beta, eta=cfg.beta_eta, out=st.rho_projected)

            dev.fill(st.dC_drho_full, 0.0)
            u_max = []
            # optimisers whose update needs filter solves that depend only on the
            # filtered field (LogMOC's volume chain) may start them now, on a side
            # stream, behind the state solve (joined before the filter is used again)
            prefetch = getattr(self, "_prefetch_start", None)
            if prefetch is not None:
                prefetch(st, beta)
            with self._timed_section("objective_and_energy"):
                with self._timed_section("objective"):
                    compliance_avg = self.fem.objectives_multi_load(
                        st.rho_projected, p, st.u_dofs, timer=self.timer,
                        force_scale=neumann_scale).mean()
                if prefetch is not None:
                    self._prefetch_join()
                with self._timed_section("energy"):
                    energy = self.fem.energy_multi_load(st.rho_projected, p, st.u_dofs)
                    if n_tasks == 1:
                        st.energy_mean.copy_(energy[:, 0])
                    else:
                        dev.fill(st.energy_mean, 0.0)
                        for i in range(n_tasks):
                            dev.axpby(1.0 / n_tasks, energy[:, i].contiguous(), 1.0,
                                      st.energy_mean)
                st.compliance = float(compliance_avg)

            with self._timed_section("sensitivity"):
                custom = None
                if hasattr(self.fem, "compliance_sensitivity_multi_load"):
                    custom = self.fem.compliance_sensitivity_multi_load(
                        st.rho_projected, p, st.u_dofs)
                projection.heaviside_projection_derivative_inplace(
                    st.rho_filtered, beta=beta, eta=cfg.beta_eta, out=st.dH_drho)
                for load in range(n_tasks):
                    with self._timed_section("task_loop"):
                        u_max.append(dev.reduce_absmax(st.u_dofs[:, load]))
                        if custom is not None:
                            st.dC_drho_projected.copy_(custom[:, load])
                            dev.hadamard(1.0, st.dC_drho_projected, st.dH_drho,
                                         st.grad_filtered)
                        else:
                            # K8: dC/drho_hat and the projection chain rule fused
                            dev.dc_drho(st.rho_projected, energy[:, load].contiguous(),
                                        c_max, c_min, p, ramp=ramp, dH=st.dH_drho,
                                        out=st.grad_filtered)
                        back = self.filter.gradient(st.grad_filtered)
                        dev.axpby(1.0 / n_tasks, back, 1.0, st.dC_drho_full)

            if cfg.sensitivity_filter:
                with self._timed_section("sensitivity_filter"):
                    st.dC_drho_full.copy_(self.filter.forward(st.dC_drho_full))

            dev.gather(st.dC_drho_full, design, out=st.dC_drho_design_eles)
            dev.gather(st.rho, design, out=st.rho_design_eles)
            with self._timed_section("rho_update"):
                self.rho_update(
                    iter_num, st.rho_design_eles, st.rho_projected,
                    st.dC_drho_design_eles, st.u_dofs, st.energy_mean,
                    st.scaling_rate, move_limit, eta, beta, st.rho_clip_lower,
                    st.rho_clip_upper, percentile, st.elements_volume_design,
                    st.elements_volume_design_sum, vol_frac)
            dev.scatter(st.rho_design_eles, design, st.rho)
            if self._pin_idx is not None:
                dev.scatter(self._pin_ones, self._pin_idx, st.rho)

            with self._timed_section("record_metrics"):
                rho_change_max = dev.reduce_maxdiff(st.rho, st.rho_prev, design)
                st.rho_change_max = rho_change_max
                rec = self.recorder
                rec.feed_data("rho_change_max", rho_change_max)
                rec.feed_data("rho_projected",
                              ArrayStats(*dev.reduce_stats(st.rho_projected, design)))
                rec.feed_data("energy", st.energy_mean)
                rec.feed_data(self._objective_history_name(tsk), compliance_avg)
                rec.feed_data("scaling_rate", st.scaling_rate)
                u_max = u_max[0] if len(u_max) == 1 else np.array(u_max)
                rec.feed_data("u_max", u_max)
                rec.feed_data("neumann_scale", neumann_scale)
                st.u_max = u_max
                if cfg.check_convergence:
                    conv_rho = rho_change_max < cfg.tol_rho_change
                    kkt = self.kkt_residual
                    if kkt is None:
                        raise ValueError("kkt_residual is not computed in rho_update")
                    st.kkt_residual = kkt
                    conv_kkt = abs(kkt) < cfg.tol_kkt_residual if np.isfinite(kkt) else True
                    st.vol_error = rec.latest("vol_error")

            export_now = (
                iter_num % (cfg.max_iters // cfg.record_times) == 0
                or iter_num == 1
                or (conv_rho and conv_kkt)
                or iter_num == iter_limit - 1
            )
            if export_now:
                if self.export_enabled and self._io:
                    with self._timed_section("export_iteration"):
                        self._export_iteration(iter_num, st, st.energy_mean)
                if conv_rho and conv_kkt:
                    converged = True
                    break

        self._iter_next = (iter_num + 1) if iter_num is not None else iter_start
        if converged or self._iter_next >= st.iter_end:
            self._finalize()
        self._report_timing()

    def rho_update(self, iter_num, rho_design_eles, rho_projected,
                   dC_drho_design_eles, u_dofs, energy_mean, scaling_rate,
                   move_limit, eta, beta, rho_clip_lower, rho_clip_upper,
                   percentile, elements_volume_design,
                   elements_volume_design_sum, vol_frac):
        raise NotImplementedError("")
