"""ctypes binding of libnosh_b200.so (the C ABI of include/nosh_b200.h).

There is no fallback: if the CUDA library is missing the import of this module raises,
and nosh_ctx_create fails (NOSH_ECUDA) on a machine without a GPU.
"""
import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libnosh_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "nosh_b200.h")

NOSH_OK, NOSH_EINVAL, NOSH_ECUDA, NOSH_ESTATE, NOSH_EMESH, NOSH_EKEY, NOSH_ECOMM, NOSH_EUNSUPPORTED = range(8)
NO_TRANS, TRANS, CONJ_TRANS = 0, 1, 2
LAYOUT_CSR, LAYOUT_SELL32 = 0, 1
MAT_KEO, MAT_DKEO = 0, 1
OP_JACOBIAN, OP_KEO, OP_KEOREG = 0, 1, 2
PREC_NONE, PREC_KEOREG_AMG = 0, 1
SOLVER_MINRES, SOLVER_CG, SOLVER_GMRES = 0, 1, 2
AMG_REUSE_NONE, AMG_REUSE_FULL = 0, 1
ARC_SCALING, ARC_HIT_BOUND = 1, 2
FVM_VERTEX_NONE, FVM_VERTEX_EXP, FVM_VERTEX_EXP_LINEARIZED = 0, 1, 2
FVM_DIRICHLET_NONE, FVM_DIRICHLET_IDENTITY, FVM_DIRICHLET_ZERO, FVM_DIRICHLET_VALUE = 0, 1, 2, 3
AMG_MAX_LEVELS = 16


class MeshInfo(C.Structure):
    _fields_ = [("dim", C.c_int32), ("n_global", C.c_int64), ("owned_begin", C.c_int64),
                ("n_owned", C.c_int64), ("n_ghost", C.c_int64), ("n_cells", C.c_int64),
                ("n_edges", C.c_int64), ("n_blocks", C.c_int64), ("n_stored", C.c_int64)]


class KrylovResult(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("converged", C.c_int32), ("relres", C.c_double),
                ("breakdown", C.c_int32), ("reserved", C.c_int32)]


class ContinuationStep(C.Structure):
    _fields_ = [("step", C.c_int32), ("converged", C.c_int32), ("newton_steps", C.c_int32),
                ("linear_iterations", C.c_int32), ("predictor_linear_iterations", C.c_int32),
                ("reserved", C.c_int32), ("param", C.c_double), ("gibbs_energy", C.c_double),
                ("norm", C.c_double), ("fnorm", C.c_double)]


class AmgInfo(C.Structure):
    _fields_ = [("levels", C.c_int32), ("degree", C.c_int32), ("coarse_degree", C.c_int32), ("reserved", C.c_int32),
                ("nodes", C.c_int64 * AMG_MAX_LEVELS),
                ("blocks", C.c_int64 * AMG_MAX_LEVELS), ("p_blocks", C.c_int64 * AMG_MAX_LEVELS),
                ("lambda_max", C.c_double * AMG_MAX_LEVELS), ("setup_seconds", C.c_double)]


class ArclengthOptions(C.Structure):
    _fields_ = [("initial_step_size", C.c_double), ("min_step_size", C.c_double), ("max_step_size", C.c_double),
                ("aggressiveness", C.c_double), ("max_steps", C.c_int32), ("nl_maxit", C.c_int32),
                ("nl_tol", C.c_double), ("lin_tol", C.c_double), ("lin_maxit", C.c_int32),
                ("flags", C.c_int32), ("min_value", C.c_double), ("max_value", C.c_double),
                ("goal_contribution", C.c_double), ("max_contribution", C.c_double), ("min_scale", C.c_double),
                ("initial_scale", C.c_double)]


class ArclengthStep(C.Structure):
    _fields_ = [("step", C.c_int32), ("converged", C.c_int32), ("newton_steps", C.c_int32),
                ("linear_iterations", C.c_int32), ("predictor_linear_iterations", C.c_int32),
                ("reserved", C.c_int32), ("param", C.c_double), ("gibbs_energy", C.c_double),
                ("norm", C.c_double), ("fnorm", C.c_double), ("step_size", C.c_double),
                ("dparam_ds", C.c_double), ("scale", C.c_double)]


class NewtonResult(C.Structure):
    _fields_ = [("steps", C.c_int32), ("converged", C.c_int32),
                ("total_linear_iterations", C.c_int32), ("linear_solve_status", C.c_int32),
                ("fnorm", C.c_double)]


# nosh_allgather_fn: int (*)(void *user, const void *send, void *recv, int64_t bytes_per_rank)
ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64)


# nosh_step_observer_fn: int (*)(void *user, int step, double param, double gibbs, double norm, const double *psi, int64 n)
STEP_OBSERVER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double),
                               C.c_int64)


def build(force=False):
    """Compile the library in-tree with nvcc for sm_100a (nosh_b200/csrc/Makefile)."""
    args = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    if force:
        args.append("-B")
    subprocess.check_call(args, stdout=subprocess.DEVNULL)
    return SO_PATH


def declared_symbols():
    """Every NOSH_API function name declared in include/nosh_b200.h."""
    txt = open(HEADER_PATH).read()
    return sorted(set(re.findall(r"NOSH_API\s+[\w\s\*]+?\b(nosh_\w+)\s*\(", txt)))


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            "nosh_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a). There is no CPU fallback." % SO_PATH)
    L = C.CDLL(SO_PATH)
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
    cpp = C.POINTER(C.c_char_p)
    sig = {
        "nosh_version": (C.c_char_p, []),
        "nosh_ctx_create": (C.c_int, [C.c_int, vp, C.POINTER(vp)]),
        "nosh_ctx_destroy": (None, [vp]),
        "nosh_last_error": (C.c_char_p, [vp]),
        "nosh_ctx_set_layout": (C.c_int, [vp, C.c_int]),
        "nosh_ctx_set_group_vertices": (C.c_int, [vp, i64]),
        "nosh_ctx_synchronize": (C.c_int, [vp]),
        "nosh_prefetch": (C.c_int, [vp, vp]),
        "nosh_ctx_set_async_output": (C.c_int, [vp, C.c_int]),
        "nosh_comm_unique_id": (C.c_int, [vp]),
        "nosh_ctx_comm_init": (C.c_int, [vp, vp, C.c_int, C.c_int]),
        "nosh_ctx_comm_init_host": (C.c_int, [vp, C.c_int, C.c_int, ALLGATHER_FN, vp]),
        "nosh_ctx_get_stat": (C.c_int, [vp, C.c_char_p, C.POINTER(dbl)]),
        "nosh_ctx_list_stats": (C.c_int, [vp, vp, i64]),
        "nosh_ctx_set_step_observer": (C.c_int, [vp, STEP_OBSERVER_FN, vp]),
        "nosh_partition_range": (C.c_int, [i64, C.c_int, C.c_int, i64, C.POINTER(i64), C.POINTER(i64),
                                           C.POINTER(i64)]),
        "nosh_mesh_set": (C.c_int, [vp, C.c_int, i64, vp, i64, vp]),
        "nosh_mesh_set_local": (C.c_int, [vp, C.c_int, i64, i64, vp, vp, i64, vp]),
        "nosh_mesh_tetgrid": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, vp, dbl, C.c_uint64]),
        "nosh_mesh_info": (C.c_int, [vp, C.POINTER(MeshInfo)]),
        "nosh_mesh_local_gids": (C.c_int, [vp, vp]),
        "nosh_mesh_get_coords": (C.c_int, [vp, vp]),
        "nosh_mesh_get_cells": (C.c_int, [vp, vp]),
        "nosh_mesh_get_edges": (C.c_int, [vp, vp, vp, vp]),
        "nosh_mesh_get_control_volumes": (C.c_int, [vp, vp]),
        "nosh_set_thickness": (C.c_int, [vp, vp, dbl]),
        "nosh_set_potential_constant": (C.c_int, [vp, dbl, C.c_char_p]),
        "nosh_set_potential_values": (C.c_int, [vp, vp]),
        "nosh_set_mvp_explicit": (C.c_int, [vp, vp]),
        "nosh_set_mvp_explicit_curl": (C.c_int, [vp, vp]),
        "nosh_set_mvp_constcurl": (C.c_int, [vp, vp, vp]),
        "nosh_get_alpha_cache": (C.c_int, [vp, vp]),
        "nosh_get_edge_projection": (C.c_int, [vp, C.c_int, cpp, vp, C.c_char_p, vp, vp]),
        "nosh_keo_fill": (C.c_int, [vp, C.c_int, cpp, vp]),
        "nosh_dkeo_fill": (C.c_int, [vp, C.c_int, cpp, vp, C.c_char_p]),
        "nosh_matrix_apply": (C.c_int, [vp, C.c_int, vp, i64, vp, i64, C.c_int, C.c_int, dbl, dbl]),
        "nosh_get_block_csr": (C.c_int, [vp, C.c_int, vp, vp, vp]),
        "nosh_jac_rebuild": (C.c_int, [vp, C.c_int, cpp, vp, vp]),
        "nosh_jac_apply": (C.c_int, [vp, vp, i64, vp, i64, C.c_int, C.c_int, dbl, dbl]),
        "nosh_jac_get_diags": (C.c_int, [vp, vp, vp]),
        "nosh_compute_f": (C.c_int, [vp, C.c_int, cpp, vp, vp, vp]),
        "nosh_compute_dfdp": (C.c_int, [vp, C.c_int, cpp, vp, C.c_char_p, vp, vp]),
        "nosh_keoreg_rebuild": (C.c_int, [vp, C.c_int, cpp, vp, vp]),
        "nosh_keoreg_matrix_apply": (C.c_int, [vp, vp, i64, vp, i64, C.c_int]),
        "nosh_keoreg_get_diags": (C.c_int, [vp, vp, vp]),
        "nosh_keoreg_apply": (C.c_int, [vp, vp, i64, vp, i64, C.c_int, C.c_int, dbl, dbl]),
        "nosh_amg_set_options": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
        "nosh_amg_setup": (C.c_int, [vp]),
        "nosh_amg_info": (C.c_int, [vp, C.POINTER(AmgInfo)]),
        "nosh_amg_get_aggregates": (C.c_int, [vp, C.c_int, vp]),
        "nosh_amg_get_matrix": (C.c_int, [vp, C.c_int, vp, vp, vp]),
        "nosh_amg_get_prolongator": (C.c_int, [vp, C.c_int, vp, vp, vp]),
        "nosh_dot": (C.c_int, [vp, vp, vp, C.POINTER(dbl)]),
        "nosh_norm2": (C.c_int, [vp, vp, C.POINTER(dbl)]),
        "nosh_minres": (C.c_int, [vp, C.c_int, vp, vp, dbl, C.c_int, C.POINTER(KrylovResult), vp]),
        "nosh_cg": (C.c_int, [vp, C.c_int, vp, vp, dbl, C.c_int, C.POINTER(KrylovResult), vp]),
        "nosh_minres_prec": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, dbl, C.c_int, C.POINTER(KrylovResult), vp]),
        "nosh_cg_prec": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, dbl, C.c_int, C.POINTER(KrylovResult), vp]),
        "nosh_gmres": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, dbl, C.c_int, C.c_int, C.POINTER(KrylovResult), vp]),
        "nosh_ctx_set_linear_solver": (C.c_int, [vp, C.c_int, C.c_int]),
        "nosh_ctx_set_preconditioner": (C.c_int, [vp, C.c_int]),
        "nosh_newton": (C.c_int, [vp, C.c_int, cpp, vp, vp, dbl, C.c_int, dbl, C.c_int,
                                  C.POINTER(NewtonResult), vp, vp]),
        "nosh_inner_product": (C.c_int, [vp, vp, vp, C.POINTER(dbl)]),
        "nosh_gibbs_energy": (C.c_int, [vp, vp, C.POINTER(dbl)]),
        "nosh_continuation": (C.c_int, [vp, C.c_int, cpp, vp, C.c_char_p, dbl, C.c_int, vp, dbl, C.c_int, dbl,
                                        C.c_int, vp]),
        "nosh_continuation_arclength": (C.c_int, [vp, C.c_int, cpp, vp, C.c_char_p, C.POINTER(ArclengthOptions), vp,
                                                  vp, C.POINTER(C.c_int32)]),
        "nosh_meshfile_last_error": (C.c_char_p, []),
        "nosh_meshfile_read": (C.c_int, [C.c_char_p, C.POINTER(vp)]),
        "nosh_meshfile_free": (None, [vp]),
        "nosh_meshfile_info": (C.c_int, [vp, C.POINTER(i32), C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]),
        "nosh_meshfile_get": (C.c_int, [vp, vp, vp]),
        "nosh_meshfile_field_name": (C.c_int, [vp, i32, C.POINTER(C.c_char_p), C.POINTER(i32)]),
        "nosh_meshfile_get_field": (C.c_int, [vp, C.c_char_p, C.POINTER(i32), vp]),
        "nosh_meshfile_write": (C.c_int, [C.c_char_p, i32, i64, vp, i64, vp, i32, cpp, vp, vp, i32]),
        "nosh_morton_order": (C.c_int, [i64, vp, vp]),
        "nosh_mesh_boundary_vertices": (C.c_int, [vp, vp]),
        "nosh_fvm_matrix_fill": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "nosh_fvm_matrix_apply": (C.c_int, [vp, vp, vp]),
        "nosh_fvm_get_csr": (C.c_int, [vp, vp, vp, vp]),
        "nosh_fvm_operator_apply": (C.c_int, [vp, C.c_int, C.c_int, dbl, vp, vp, C.c_int, vp, vp, vp]),
        "nosh_fvm_cg": (C.c_int, [vp, vp, vp, dbl, C.c_int, C.POINTER(KrylovResult)]),
        "nosh_scratch_vector": (C.c_int, [vp, C.c_int, C.POINTER(vp)]),
        "nosh_ctx_set_tuning": (C.c_int, [vp, C.c_char_p, C.c_int]),
        "nosh_launch_count": (i64, [vp]),
        "nosh_timer_start": (C.c_int, [vp]),
        "nosh_timer_stop": (C.c_int, [vp, C.POINTER(C.c_float)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    L._nosh_signatures = sig
    _lib = L
    return L
