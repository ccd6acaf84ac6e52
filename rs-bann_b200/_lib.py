"""ctypes binding of libbann_b200.so (the C ABI declared in include/bann.h).

There is no CPU fallback: if the shared library is missing the import fails loudly, and every
compute entry point needs a CUDA device."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BANN_LIB_PATH") or os.path.join(_HERE, "libbann_b200.so")   # override: kernel-variant experiments only

MAX_LAYERS = 8
STD_NORMAL, RIDGE_BASE, RIDGE_ARD, LASSO_BASE, LASSO_ARD = range(5)
TANH, RELU, LEAKY_RELU, SILU, IDENTITY = range(5)
STEP_UNIFORM, STEP_RANDOM, STEP_STD_SCALED, STEP_IZMAILOV = range(4)
HMC_REJECTED_EARLY, HMC_REJECTED, HMC_ACCEPTED = range(3)
COMM_HANDLE_BYTES = 128

MODEL_NAMES = {"std_normal": STD_NORMAL, "ridge_base": RIDGE_BASE, "ridge_ard": RIDGE_ARD,
               "lasso_base": LASSO_BASE, "lasso_ard": LASSO_ARD}
ACT_NAMES = {"tanh": TANH, "relu": RELU, "leaky_relu": LEAKY_RELU, "silu": SILU, "identity": IDENTITY}
STEP_NAMES = {"uniform": STEP_UNIFORM, "random": STEP_RANDOM, "std_scaled": STEP_STD_SCALED,
              "izmailov": STEP_IZMAILOV}


class BranchLayout(C.Structure):
    _fields_ = [("num_layers", C.c_uint32), ("widths", C.c_uint32 * MAX_LAYERS)]


class McmcCfg(C.Structure):
    _fields_ = [("hmc_step_size_factor", C.c_float), ("hmc_max_hamiltonian_error", C.c_float),
                ("hmc_integration_length", C.c_uint32), ("hmc_step_size_mode", C.c_int32),
                ("fixed_param_precisions", C.c_int32), ("joint_hmc", C.c_int32), ("gradient_descent", C.c_int32),
                ("gradient_descent_joint", C.c_int32), ("num_grad", C.c_int32), ("num_grad_traj", C.c_int32)]


class RngInject(C.Structure):
    _fields_ = [("momenta", C.POINTER(C.c_float)), ("accept_uniform", C.POINTER(C.c_float)),
                ("step_uniforms", C.POINTER(C.c_float)), ("std_gammas", C.POINTER(C.c_float)),
                ("num_std_gammas", C.c_uint32)]


class HmcResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("log_density", C.c_float), ("neg_h_init", C.c_float),
                ("neg_h_final", C.c_float), ("steps_done", C.c_uint32), ("u_turn_step", C.c_int32)]


class Trajectory(C.Structure):
    _fields_ = [("params", C.POINTER(C.c_float)), ("ldg", C.POINTER(C.c_float)),
                ("hamiltonian", C.POINTER(C.c_float))]


class TrajectoryJoint(C.Structure):
    _fields_ = [("params", C.POINTER(C.c_float)), ("precisions", C.POINTER(C.c_float)), ("ldg", C.POINTER(C.c_float)),
                ("hamiltonian", C.POINTER(C.c_float)), ("num_ldg", C.POINTER(C.c_float))]


class SweepStats(C.Structure):
    _fields_ = [("num_samples", C.c_uint64), ("num_accepted", C.c_uint64), ("num_early_rejected", C.c_uint64),
                ("mse_train", C.c_float), ("lpd", C.c_float), ("output_bias", C.c_float),
                ("error_precision", C.c_float), ("output_layer_precision", C.c_float)]


_fp = C.POINTER(C.c_float)
_vp = C.c_void_p
_u64 = C.c_uint64

# every symbol include/bann.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "bann_last_error": (C.c_char_p, []),
    "bann_cuda_available": (C.c_int, []),
    "bann_ctx_create": (C.c_int, [C.c_int, _vp, C.c_int, C.c_int, C.POINTER(_vp)]),
    "bann_ctx_destroy": (None, [_vp]),
    "bann_ctx_sync": (C.c_int, [_vp]),
    "bann_ctx_comm_handle": (C.c_int, [_vp, _vp]),
    "bann_ctx_comm_connect": (C.c_int, [_vp, _vp]),
    "bann_ctx_comm_connected": (C.c_int, [_vp]),
    "bann_genotypes_create": (C.c_int, [_vp, _vp, _u64, _u64, _u64, _fp, _fp, _u64, C.POINTER(_u64),
                                         C.POINTER(_u64), C.POINTER(_vp)]),
    "bann_genotypes_random": (C.c_int, [_vp, _u64, _u64, _u64, _u64, _u64, C.c_float, C.c_float, _u64, C.POINTER(_u64),
                                         C.POINTER(_u64), C.POINTER(_vp)]),
    "bann_genotypes_destroy": (None, [_vp]),
    "bann_genotypes_col_stats": (C.c_int, [_vp, _fp, _fp]),
    "bann_genotypes_col_counts": (C.c_int, [_vp, C.POINTER(_u64)]),
    "bann_genotypes_set_col_stats": (C.c_int, [_vp, _fp, _fp]),
    "bann_genotypes_decode_branch": (C.c_int, [_vp, _u64, C.c_int, _fp]),
    "bann_genotypes_decode_branch_tc": (C.c_int, [_vp, _u64, C.c_int, _fp]),
    "bann_genotypes_has_tc_store": (C.c_int, [_vp]),
    "bann_genotypes_has_byte_store": (C.c_int, [_vp]),
    "bann_genotypes_release_byte_store": (C.c_int, [_vp]),
    "bann_net_create": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.POINTER(BranchLayout), _fp, C.POINTER(_vp)]),
    "bann_net_destroy": (None, [_vp]),
    "bann_net_branch_sizes": (C.c_int, [_vp, _u64, C.POINTER(_u64), C.POINTER(_u64)]),
    "bann_net_set_branch": (C.c_int, [_vp, _u64, _fp, _fp]),
    "bann_net_get_branch": (C.c_int, [_vp, _u64, _fp, _fp]),
    "bann_net_set_all_params": (C.c_int, [_vp, _fp, _fp]),
    "bann_net_get_all_params": (C.c_int, [_vp, _fp, _fp]),
    "bann_net_set_globals": (C.c_int, [_vp, _fp]),
    "bann_net_get_globals": (C.c_int, [_vp, _fp]),
    "bann_net_set_targets": (C.c_int, [_vp, _fp]),
    "bann_net_get_residual": (C.c_int, [_vp, _fp]),
    "bann_net_set_residual": (C.c_int, [_vp, _fp]),
    "bann_net_init_residual": (C.c_int, [_vp]),
    "bann_branch_fwd_bwd": (C.c_int, [_vp, _u64, _fp, _fp, _fp, _fp, _fp]),
    "bann_branch_log_density": (C.c_int, [_vp, _u64, C.c_float, _fp]),
    "bann_branch_step_sizes": (C.c_int, [_vp, _u64, C.POINTER(McmcCfg), _fp, _fp]),
    "bann_hmc_step": (C.c_int, [_vp, _u64, _fp, C.POINTER(McmcCfg), C.POINTER(RngInject), C.POINTER(HmcResult),
                                 C.POINTER(Trajectory), _fp]),
    "bann_branch_joint": (C.c_int, [_vp, _u64, _fp, _fp, _fp, _fp, _fp]),
    "bann_hmc_step_joint": (C.c_int, [_vp, _u64, _fp, C.POINTER(McmcCfg), C.POINTER(RngInject), C.POINTER(HmcResult),
                                       C.POINTER(TrajectoryJoint), _fp]),
    "bann_gradient_descent": (C.c_int, [_vp, _u64, _fp, C.POINTER(McmcCfg), C.POINTER(HmcResult), _fp,
                                         C.POINTER(C.c_uint32), _fp]),
    "bann_gradient_descent_joint": (C.c_int, [_vp, _u64, _fp, C.POINTER(McmcCfg), C.POINTER(HmcResult), _fp]),
    "bann_gibbs_branch": (C.c_int, [_vp, _u64, C.POINTER(McmcCfg), C.POINTER(RngInject)]),
    "bann_visit_branch": (C.c_int, [_vp, _u64, C.POINTER(McmcCfg), C.POINTER(RngInject), C.POINTER(HmcResult)]),
    "bann_visit_branch_traj": (C.c_int, [_vp, _u64, C.POINTER(McmcCfg), _u64, C.POINTER(HmcResult), C.POINTER(TrajectoryJoint)]),
    "bann_sweep": (C.c_int, [_vp, C.POINTER(McmcCfg), C.POINTER(_u64), _u64, C.c_uint32, _u64,
                              C.POINTER(SweepStats)]),
    "bann_visit_group": (C.c_int, [_vp, C.POINTER(_u64), _u64, C.POINTER(McmcCfg), C.POINTER(RngInject), _u64,
                                    C.POINTER(HmcResult)]),
    "bann_predict": (C.c_int, [_vp, _vp, _fp]),
    "bann_branch_activations": (C.c_int, [_vp, _u64, _vp, _fp]),
    "bann_branch_effect_sizes": (C.c_int, [_vp, _u64, _vp, _fp, _fp]),
    "bann_net_stats": (C.c_int, [_vp, C.POINTER(SweepStats)]),
    "bann_net_gradient": (C.c_int, [_vp, _fp, _fp, _fp, _fp]),
    "bann_branch_numerical_ldg": (C.c_int, [_vp, _u64, _fp, _fp]),
    "bann_net_gradient_slice": (C.c_int, [_vp, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64)]),
    "bann_net_comm_handle": (C.c_int, [_vp, _vp]),
    "bann_net_comm_connect": (C.c_int, [_vp, _vp]),
    "bann_net_comm_connected": (C.c_int, [_vp]),
    "bann_grouped_allreduce": (C.c_int, [_vp]),
    "bann_net_select_hmc_path": (C.c_int, [_vp, C.c_int]),
    "bann_net_persistent_launches": (C.c_uint64, [_vp]),
    "bann_net_last_k1_kernel": (C.c_char_p, [_vp]),
    "bann_pinned_alloc": (C.c_int, [_u64, C.POINTER(_vp)]),
    "bann_pinned_free": (None, [_vp]),
    "bann_net_gradient_begin": (C.c_int, [_vp, _fp, _fp]),
    "bann_net_gradient_end": (C.c_int, [_vp, _fp, _fp]),
    "bann_grouped_begin": (C.c_int, [_vp, C.POINTER(McmcCfg), _u64, C.c_int]),
    "bann_grouped_leapfrog": (C.c_int, [_vp, C.POINTER(McmcCfg), C.c_uint32, C.c_int]),
    "bann_grouped_phase_a": (C.c_int, [_vp]),
    "bann_grouped_phase_b": (C.c_int, [_vp, C.POINTER(McmcCfg), C.c_int, C.c_int]),
    "bann_grouped_finish": (C.c_int, [_vp, _u64, C.POINTER(_u64), C.POINTER(_u64)]),
    "bann_grouped_state": (C.c_int, [_vp, _fp, _fp, C.POINTER(C.c_int32)]),
    "bann_allreduce_buffer": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_u64)]),
    "bann_net_force_generic": (C.c_int, [_vp, C.c_int]),
    "bann_net_lpd_terms": (C.c_int, [_vp, _fp, _fp, _fp]),
    "bann_net_select_k1": (C.c_int, [_vp, C.c_int]),
    "bann_net_select_k1_tc_variant": (C.c_int, [_vp, C.c_int]),
    "bann_launch_count": (_u64, [C.c_int]),
    "bann_net_algorithmic_bytes": (C.c_int, [_vp, C.POINTER(_u64)]),
}


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C rs-bann_b200/csrc`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here == header / library mismatch
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load()


class BannError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise BannError(lib.bann_last_error().decode("utf-8", "replace"))
