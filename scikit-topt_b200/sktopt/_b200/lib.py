"""ctypes binding of ``libsktopt_b200.so`` (C ABI in ``include/sktopt_b200.h``).

There is no CPU fallback: if the shared library is missing the import of any
GPU entry point raises, and every call checks the status code and raises
``RuntimeError`` with the library's message.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SKTOPT_B200_LIB: another build of the same library (kernel A/B experiments)
LIB_PATH = os.environ.get("SKTOPT_B200_LIB") or os.path.join(_HERE, "libsktopt_b200.so")

_lib = None

c_i32p = C.c_void_p
c_f64p = C.c_void_p
c_u8p = C.c_void_p
c_stream = C.c_void_p
i64 = C.c_int64
f64 = C.c_double
i32 = C.c_int

# name -> argtypes (restype is int unless listed in _RESTYPE)
SIGNATURES = {
    "sktb_last_error": [],
    "sktb_version": [],
    "sktb_mesh_create": [C.POINTER(C.c_void_p), i32, i64, i64, C.c_void_p, C.c_void_p, i32],
    "sktb_mesh_destroy": [C.c_void_p],
    "sktb_mesh_node_nnz": [C.c_void_p],
    "sktb_mesh_node_graph_h": [C.c_void_p, C.c_void_p, C.c_void_p],
    "sktb_mesh_dof_pattern": [C.c_void_p, i32, c_i32p, c_i32p, c_stream],
    "sktb_unit_ke": [C.c_void_p, i32, f64, i32, C.c_void_p, C.c_void_p, i64, C.c_void_p, c_f64p, c_stream],
    "sktb_assemble": [C.c_void_p, i32, c_f64p, c_i32p, c_f64p, c_u8p, c_f64p, c_stream],
    "sktb_csr_enforce": [i64, c_i32p, c_i32p, c_f64p, c_u8p, c_stream],
    "sktb_csr_inv_diag": [i64, c_i32p, c_i32p, c_f64p, c_f64p, c_stream],
    "sktb_spmv": [i64, i32, c_i32p, c_i32p, c_f64p, c_f64p, c_f64p, c_stream],
    "sktb_pcg_create": [C.POINTER(C.c_void_p), i64, i32],
    "sktb_pcg_destroy": [C.c_void_p],
    "sktb_spmv_bsr3": [i64, c_i32p, c_i32p, c_f64p, c_f64p, c_f64p, c_stream],
    "sktb_bsr3_inv_diag": [i64, i64, c_i32p, c_i32p, c_f64p, c_f64p, c_stream],
    "sktb_spmv_bsr3_tma": [i64, i64, i32, c_i32p, c_i32p, c_f64p, c_f64p, c_f64p, c_stream],
    "sktb_pcg_solve_bsr3": [C.c_void_p, c_i32p, c_i32p, i64, i32, c_f64p, c_f64p, c_f64p, c_f64p, i32, f64, i32, i32, C.c_void_p, C.c_void_p, c_stream],
    "sktb_mg_create": [C.POINTER(C.c_void_p), i32, i32],
    "sktb_mg_destroy": [C.c_void_p],
    "sktb_mg_set_params": [C.c_void_p, f64, i32],
    "sktb_mg_set_level": [C.c_void_p, i32, i64, i64, i32, c_i32p, c_i32p, c_f64p, c_f64p, c_u8p],
    "sktb_mg_set_level_omega": [C.c_void_p, i32, f64],
    "sktb_pcg_lambda_max_bsr3": [C.c_void_p, c_i32p, c_i32p, i64, i32, c_f64p, c_f64p, i32, C.c_void_p, c_stream],
    "sktb_mg_set_level0_range": [C.c_void_p, i64, i64],
    "sktb_mg_set_level_slab": [C.c_void_p, i32, i64, i64, i64, i32, i32],
    "sktb_mg_level_apply": [C.c_void_p, i32, C.c_void_p, c_f64p, c_f64p, c_stream],
    "sktb_elem_restrict_range": [i64, i64, i64, i64, c_i32p, c_u8p, c_f64p, c_f64p, c_f64p, c_i32p, c_f64p, c_f64p, c_stream],
    "sktb_galerkin_bsr3_lattice": [C.c_void_p, C.c_void_p, c_i32p, c_i32p, c_f64p, c_f64p, c_i32p, c_f64p, c_i32p, c_i32p, c_f64p, c_u8p, c_i32p, c_i32p, c_u8p, c_f64p, c_stream],
    "sktb_elem_combine_range": [i64, i64, i64, c_i32p, c_u8p, c_f64p, c_i32p, c_f64p, c_f64p, c_stream],
    "sktb_pcg_set_slab_halo": [C.c_void_p, i64, i32, i32],
    "sktb_pcg_set_first_batch": [C.c_void_p, i32],
    "sktb_element_stress": [C.c_void_p, i32, c_i32p, c_f64p, c_f64p, f64, c_f64p, c_f64p, c_stream],
    "sktb_dgemm_batched": [i32, i32, i32, c_f64p, i32, i64, c_f64p, i32, i64, c_f64p, i32, i64, i32, c_f64p, i64, c_stream],
    "sktb_smg_create": [C.POINTER(C.c_void_p), i32, C.c_void_p, i32],
    "sktb_smg_destroy": [C.c_void_p],
    "sktb_smg_set_mask": [C.c_void_p, i32, c_u8p],
    "sktb_smg_set_level_sweeps": [C.c_void_p, i32, i32],
    "sktb_smg_set_transfer": [C.c_void_p, i32, c_i32p, c_i32p, c_f64p, c_f64p, c_i32p, c_f64p],
    "sktb_smg_setup_csr": [C.c_void_p, c_i32p, c_i32p, c_f64p, c_stream],
    "sktb_smg_vcycle": [C.c_void_p, c_f64p, c_f64p, c_stream],
    "sktb_smg_apply": [C.c_void_p, i32, c_f64p, c_f64p, c_stream],
    "sktb_smg_level_values": [C.c_void_p, i32, c_f64p, c_stream],
    "sktb_pcg_solve_smg": [C.c_void_p, C.c_void_p, c_f64p, c_f64p, i32, f64, i32, i32, C.c_void_p, C.c_void_p, c_stream],
    "sktb_mg_set_transfer": [C.c_void_p, i32, C.c_void_p, C.c_void_p, c_i32p, c_i32p, c_f64p, c_f64p, c_i32p, c_f64p],
    "sktb_mg_vcycle": [C.c_void_p, c_f64p, c_f64p, c_stream],
    "sktb_gridop_create": [C.POINTER(C.c_void_p), i32, C.c_void_p, C.c_void_p, i32],
    "sktb_gridop_tile_shape": [C.c_void_p, C.c_void_p],
    "sktb_gridop_destroy": [C.c_void_p],
    "sktb_gridop_set_fields": [C.c_void_p, c_f64p, c_u8p],
    "sktb_gridop_apply": [C.c_void_p, i64, i64, c_f64p, c_f64p, c_stream],
    "sktb_gridop_inv_diag": [C.c_void_p, i64, i64, c_f64p, c_stream],
    "sktb_pcg_solve_grid": [C.c_void_p, C.c_void_p, C.c_void_p, c_f64p, c_f64p, c_f64p, i32, f64, i32, i32, C.c_void_p, C.c_void_p, c_stream],
    "sktb_pcg_lambda_max_grid": [C.c_void_p, C.c_void_p, c_f64p, i32, C.c_void_p, c_stream],
    "sktb_mg_set_precision": [C.c_void_p, i32],
    "sktb_mg_set_fused_tail": [C.c_void_p, i32],
    "sktb_mg_set_level_sweeps": [C.c_void_p, i32, i32],
    "sktb_mg_set_level_cheby": [C.c_void_p, i32, i32, C.c_void_p, C.c_void_p],
    "sktb_mg_factor_coarsest": [C.c_void_p, c_stream],
    "sktb_mg_share_coarsest": [C.c_void_p, C.c_void_p],
    "sktb_mg_set_level_vals32": [C.c_void_p, i32, C.c_void_p],
    "sktb_f64_to_f32": [i64, c_f64p, C.c_void_p, c_stream],
    "sktb_spmv_bsr3_tma_f32": [i64, i64, i32, c_i32p, c_i32p, C.c_void_p, c_f64p, c_f64p, c_stream],
    "sktb_mg_set_level0_grid": [C.c_void_p, C.c_void_p, i64, c_f64p, c_u8p],
    "sktb_elem_combine": [i64, c_i32p, c_u8p, c_f64p, c_i32p, c_f64p, c_f64p, c_stream],
    "sktb_elem_restrict": [i64, c_i32p, c_u8p, c_f64p, c_f64p, c_f64p, c_i32p, c_f64p, c_f64p, c_stream],
    "sktb_pcg_solve_bsr3_mg": [C.c_void_p, C.c_void_p, c_i32p, c_i32p, i64, i32, c_f64p, c_f64p, c_f64p, c_f64p, i32, f64, i32, i32, C.c_void_p, C.c_void_p, c_stream],
    "sktb_mesh_dof_pattern_rows": [C.c_void_p, i32, i64, i64, c_i32p, c_i32p, c_stream],
    "sktb_assemble_rows": [C.c_void_p, i32, i64, i64, c_f64p, c_i32p, c_f64p, c_u8p, c_f64p, c_stream],
    "sktb_csr_inv_diag_rows": [i64, i64, c_i32p, c_i32p, c_f64p, c_f64p, c_stream],
    "sktb_comm_unique_id": [C.c_void_p],
    "sktb_comm_create": [C.POINTER(C.c_void_p), C.c_void_p, i32, i32, i32],
    "sktb_comm_destroy": [C.c_void_p],
    "sktb_comm_rank": [C.c_void_p],
    "sktb_comm_world": [C.c_void_p],
    "sktb_comm_allreduce_sum": [C.c_void_p, c_f64p, c_f64p, i64, c_stream],
    "sktb_comm_arena_create": [C.c_void_p, i64, C.c_void_p],
    "sktb_comm_arena_open": [C.c_void_p, C.c_void_p],
    "sktb_comm_arena_status": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "sktb_comm_allgatherv": [C.c_void_p, c_f64p, C.c_void_p, C.c_void_p, c_stream],
    "sktb_pcg_create_dist": [C.POINTER(C.c_void_p), C.c_void_p, i64, i64, i64, i32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, i32],
    "sktb_pcg_set_profile": [C.c_void_p, i32],
    "sktb_pcg_get_profile": [C.c_void_p, C.c_void_p, C.c_void_p],
    "sktb_launch_count": [],
    "sktb_pcg_solve": [C.c_void_p, i32, c_i32p, c_i32p, c_f64p, c_f64p, c_f64p, c_f64p, i32, f64, i32, i32, C.c_void_p, C.c_void_p, c_stream],
    "sktb_interpolate_modulus": [i64, c_f64p, f64, f64, f64, i32, c_f64p, c_stream],
    "sktb_element_energy": [C.c_void_p, i32, c_f64p, c_i32p, c_f64p, c_f64p, c_f64p, c_stream],
    "sktb_element_energy_hex_uniform": [C.c_void_p, C.c_void_p, c_f64p, c_f64p, c_f64p, c_stream],
    "sktb_element_bilinear": [C.c_void_p, i32, c_f64p, c_i32p, c_f64p, c_f64p, c_f64p, f64, c_f64p, c_stream],
    "sktb_heat_exchange_local": [C.c_void_p, i32, c_i32p, c_f64p, c_f64p, c_f64p, c_f64p, c_f64p, f64, f64, f64, f64, c_f64p, c_f64p, c_f64p, c_stream],
    "sktb_dc_drho": [i64, c_f64p, c_f64p, f64, f64, f64, i32, c_f64p, c_f64p, c_stream],
    "sktb_heaviside": [i64, c_f64p, f64, f64, c_f64p, c_f64p, c_stream],
    "sktb_e2n": [C.c_void_p, c_f64p, c_f64p, c_u8p, f64, c_f64p, c_f64p, c_stream],
    "sktb_e2n_wsum": [C.c_void_p, c_f64p, c_f64p, c_stream],
    "sktb_n2e_mean": [C.c_void_p, c_f64p, i32, c_f64p, c_stream],
    "sktb_geom_tables": [C.c_void_p, i32, C.c_void_p, C.c_void_p, i64, C.c_void_p, c_f64p, c_f64p, c_f64p, c_stream],
    "sktb_unit_qp_mass": [C.c_void_p, i32, i64, c_f64p, c_f64p, c_f64p, c_stream],
    "sktb_robin_virtual_scale": [C.c_void_p, i32, c_i32p, c_f64p, c_f64p, c_f64p, c_f64p, f64, f64, f64, c_f64p, c_stream],
    "sktb_assemble_terms": [C.c_void_p, i32, c_f64p, c_i32p, c_f64p, c_f64p, c_stream],
    "sktb_robin_explicit_local": [C.c_void_p, i32, c_i32p, c_f64p, c_f64p, c_f64p, c_f64p, c_f64p, f64, f64, f64, f64, c_f64p, c_stream],
    "sktb_local_to_nodes": [C.c_void_p, c_f64p, c_f64p, c_f64p, c_stream],
    "sktb_oc_candidate": [i64, c_f64p, c_f64p, f64, f64, f64, f64, f64, f64, f64, f64, c_i32p, c_f64p, c_f64p, c_f64p, c_stream],
    "sktb_logmoc_update": [i64, c_f64p, c_f64p, f64, f64, f64, f64, f64, c_f64p, c_f64p, c_f64p, c_stream],
    "sktb_reduce_wsum_h": [i64, c_f64p, c_i32p, c_f64p, C.c_void_p, c_stream],
    "sktb_reduce_stats_h": [i64, c_f64p, c_i32p, C.c_void_p, c_stream],
    "sktb_reduce_absmax_h": [i64, c_f64p, C.c_void_p, c_stream],
    "sktb_dot_h": [i64, c_f64p, c_f64p, C.c_void_p, c_stream],
    "sktb_abs_percentile_h": [i64, c_f64p, f64, C.c_void_p, C.c_void_p, c_stream],
    "sktb_gather": [i64, c_f64p, c_i32p, c_f64p, c_stream],
    "sktb_scatter": [i64, c_f64p, c_i32p, c_f64p, c_stream],
    "sktb_axpby": [i64, f64, c_f64p, f64, c_f64p, c_stream],
    "sktb_affine": [i64, f64, c_f64p, f64, c_f64p, f64, c_f64p, c_stream],
    "sktb_hadamard": [i64, f64, c_f64p, c_f64p, c_f64p, c_stream],
    "sktb_fill_abs": [i64, c_f64p, f64, c_f64p, c_stream],
    "sktb_host_hex_volumes": [i64, i64, C.c_void_p, C.c_void_p, C.c_void_p],
    "sktb_host_lattice_facets": [i64, i64, C.c_void_p, C.c_void_p, i64, i64, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "sktb_kkt_residual_h": [i64, c_f64p, c_f64p, c_f64p, f64, f64, f64, C.c_void_p, c_stream],
    "sktb_reduce_maxdiff_h": [i64, c_f64p, c_f64p, c_i32p, C.c_void_p, c_stream],
    "sktb_enforce_rhs": [i64, c_f64p, c_f64p, c_u8p, c_f64p, c_f64p, c_stream],
    "sktb_flush_l2": [C.c_void_p, i64, c_stream],
    "sktb_fp64_probe": [i32, c_f64p, C.c_void_p, c_stream],
    "sktb_fp64_probe_const": [i32, c_f64p, C.c_void_p, c_stream],
}
_RESTYPE = {
    "sktb_last_error": C.c_char_p,
    "sktb_mesh_destroy": None,
    "sktb_pcg_destroy": None,
    "sktb_comm_destroy": None,
    "sktb_mg_destroy": None,
    "sktb_smg_destroy": None,
    "sktb_mesh_node_nnz": C.c_int64,
    "sktb_launch_count": C.c_int64,
}


def load():
    """Load the shared library (once) and declare every signature."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `make -C scikit-topt_b200/csrc` "
            "(or __graft_entry__.build()); there is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPE.get(name, C.c_int)
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().sktb_last_error()
        raise RuntimeError(
            f"sktopt_b200: {msg.decode() if msg else 'unknown error'} (status {rc})"
        )
