"""ctypes binding of librlc.so (include/rlc.h).  There is no CPU fallback: if the library is
missing or a call fails, this module raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RLC_LIB_PATH") or os.path.join(_HERE, "librlc.so")   # env override: A/B builds

TIN, TMID = 0, 1
LAYOUT_OUT_IN, LAYOUT_IN_OUT = 0, 1
ACT_SHARED, ACT_PER_STATE = 0, 1
PREC_FP32, PREC_FP16, PREC_BF16, PREC_AUTO, PREC_FP16X3, PREC_FP16C8 = 0, 1, 2, 3, 4, 5
ADAM_TORCH, ADAM_TF = 0, 1

PREC_BY_NAME = {"fp32": PREC_FP32, "fp16": PREC_FP16, "bf16": PREC_BF16, "auto": PREC_AUTO, "fp16x3": PREC_FP16X3,
                "fp16c8": PREC_FP16C8}


class RlcCritic(C.Structure):
    _fields_ = [("topology", C.c_int32), ("S", C.c_int32), ("A", C.c_int32), ("H1", C.c_int32),
                ("H2", C.c_int32), ("theta", C.c_void_p), ("smin", C.c_void_p),
                ("smax", C.c_void_p)]


class RlcMlp(C.Structure):
    _fields_ = [("inp", C.c_int32), ("H1", C.c_int32), ("H2", C.c_int32), ("O", C.c_int32),
                ("theta", C.c_void_p)]


ENV_PENDULUM, ENV_BIMODAL1D = 0, 1


class RlcEnv(C.Structure):
    _fields_ = [("kind", C.c_int32), ("S", C.c_int32), ("A", C.c_int32), ("episode_limit", C.c_int32),
                ("p", C.c_double * 8)]


class RlcError(RuntimeError):
    """Non-zero status from the C-ABI (map C status -> RuntimeError)."""


SB_MAX_NETS, SB_MAX_B = 8, 64
SB_ROLE_DOUT, SB_ROLE_V, SB_ROLE_Q, SB_ROLE_PI = 0, 1, 2, 3


class RlcSbNet(C.Structure):
    """rlc_sb_net (include/rlc.h): one forward pass of the small-minibatch fast path."""
    _fields_ = [("theta", C.c_void_p), ("inp", C.c_int), ("H1", C.c_int), ("H2", C.c_int), ("O", C.c_int),
                ("x0", C.c_void_p), ("n0", C.c_int), ("x1", C.c_void_p), ("n1", C.c_int),
                ("rows", C.c_int), ("x0_div", C.c_int), ("x1_mod", C.c_int),
                ("h1", C.c_void_p), ("h2", C.c_void_p), ("out", C.c_void_p), ("w3_snapshot", C.c_void_p),
                ("adam_state", C.c_void_p), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float),
                ("adam_variant", C.c_int), ("policy", C.c_int), ("eps", C.c_void_p), ("action_scale", C.c_float),
                ("log_std_min", C.c_float), ("log_std_max", C.c_float), ("action", C.c_void_p), ("logp", C.c_void_p),
                ("mean", C.c_void_p), ("mu_raw", C.c_void_p), ("log_std", C.c_void_p), ("z", C.c_void_p)]


class RlcSbTrain(C.Structure):
    """rlc_sb_train (include/rlc.h): one backward pass + optimiser step of the small-minibatch fast path."""
    _fields_ = [("theta", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("adam_state", C.c_void_p),
                ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("target", C.c_void_p),
                ("tau", C.c_float), ("inp", C.c_int), ("H1", C.c_int), ("H2", C.c_int), ("O", C.c_int),
                ("x0", C.c_void_p), ("n0", C.c_int), ("x1", C.c_void_p), ("n1", C.c_int),
                ("h1", C.c_void_p), ("h2", C.c_void_p), ("out", C.c_void_p), ("w3_snapshot", C.c_void_p),
                ("role", C.c_int), ("dout", C.c_void_p), ("r", C.c_void_p), ("gamma", C.c_void_p),
                ("v_next", C.c_void_p), ("logp", C.c_void_p), ("q_new", C.c_void_p), ("dmean", C.c_void_p),
                ("dlog_std", C.c_void_p), ("loss_b", C.c_void_p), ("log_std_min", C.c_float),
                ("log_std_max", C.c_float), ("entropy_scale", C.c_float), ("sac", C.c_int), ("loss_out", C.c_void_p)]


_p, _i, _f, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64
_cr = C.POINTER(RlcCritic)
_ml = C.POINTER(RlcMlp)
_en = C.POINTER(RlcEnv)

# name -> (restype, argtypes); must list every symbol include/rlc.h declares
SIGNATURES = {
    "rlc_version": (_i, []),
    "rlc_status_string": (C.c_char_p, [_i]),
    "rlc_last_cuda_error": (C.c_char_p, []),
    "rlc_create": (_i, [C.POINTER(_p), _i]),
    "rlc_destroy": (_i, [_p]),
    "rlc_launch_count": (_i64, [_p]),
    "rlc_theta_numel": (_i64, [_i, _i, _i, _i, _i]),
    "rlc_theta_offsets": (_i, [_i, _i, _i, _i, _i, C.POINTER(_i64)]),
    "rlc_pack_theta": (_i, [_i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "rlc_unpack_theta": (_i, [_i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "rlc_invalidate_pack": (_i, [_p, _p]),
    "rlc_critic_eval": (_i, [_p, _cr, _p, _i, _p, _i, _i, _i, _p, _p]),
    "rlc_critic_eval_reduce_policy": (_i, [_p, _cr, _p, _i, _p, _i, _p, _f, _p, _p, _p, _f, _i, _i, _i, _i, _i, _p, _p, _p, _p,
                                           _p]),
    "rlc_tmid_eval_grad": (_i, [_p, _cr, _p, _i, _p, _i, _i, _p, _p, _p]),
    "rlc_umma_last_error": (_i, [_p, _p]),
    "rlc_umma_mode": (_i, [_cr, _i]),
    "rlc_umma_mode_prec": (_i, [_cr, _i, _i]),
    "rlc_reduce_topk": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _i, _i, _p, _p]),
    "rlc_reduce_stats": (_i, [_p, _p, _i, _i, _p, _p, _p, _p]),
    "rlc_reduce_lse": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "rlc_reduce_fkl": (_i, [_p, _p, _p, _p, _i, _i, _f, _i, _p, _p, _p, _p]),
    "rlc_reduce_rkl": (_i, [_p, _p, _p, _p, _p, _i, _i, _f, _i, _i, _p, _p, _p]),
    "rlc_reduce_fkl_policy": (_i, [_p, _p, _p, _p, _i, _f, _p, _p, _i, _i, _f, _i, _p, _p, _p, _p, _p]),
    "rlc_reduce_rkl_policy": (_i, [_p, _p, _p, _p, _p, _i, _f, _p, _p, _i, _i, _f, _i, _i, _p, _p, _p, _p, _p]),
    "rlc_cem": (_i, [_p, _cr, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "rlc_gmm_refit": (_i, [_p, _p, _i, _i, _i, _i, _p, _f, _i, _p, _p, _p, _p, _p]),
    "rlc_critic_grad_action": (_i, [_p, _cr, _p, _p, _i, _p, _p, _p]),
    "rlc_critic_grads": (_i, [_p, _cr, _p, _p, _p, _i, _i, _p, _p, _p, _p]),
    "rlc_adam_step": (_i, [_p, _p, _p, _p, _p, _i64, _i, _f, _f, _f, _f, _i, _p, _f, _p]),
    "rlc_adam_step_dev": (_i, [_p, _p, _p, _p, _p, _i64, _p, _f, _f, _f, _f, _i, _p, _f, _p]),
    "rlc_soft_update": (_i, [_p, _p, _p, _i64, _f, _p]),
    "rlc_mlp_numel": (_i64, [_i, _i, _i, _i]),
    "rlc_mlp_offsets": (_i, [_i, _i, _i, _i, C.POINTER(_i64)]),
    "rlc_mlp_act_numel": (_i64, [_i, _i, _i]),
    "rlc_mlp_forward": (_i, [_p, _ml, _p, _i, _p, _p, _p]),
    "rlc_mlp_grads": (_i, [_p, _ml, _p, _p, _p, _i, _p, _p, _p]),
    "rlc_rows_gemm": (_i, [_p, _i, _i, _i, _i, _i, _p, _i, _p, _i, _p, _i, _p, _p, _i, _i, _f, _i, _i, _p]),
    "rlc_rows_gemm_force": (_i, [_i]),
    "rlc_tmid_tc_force": (_i, [_i]),
    "rlc_policy_evaluate": (_i, [_p, _p, _p, _i, _i, _f, _f, _f, _p, _p, _p, _p, _p, _p, _p]),
    "rlc_kl_targets": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _f, _i, _p, _p, _p, _p]),
    "rlc_policy_head_grad": (_i, [_p, _p, _i, _i, _f, _f, _i, _p, _p, _p, _p, _p, _p, _f, _i, _p, _p, _p]),
    "rlc_mixture_sample": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _i, _p, _p, _p, _p]),
    "rlc_ae_expert_step": (_i, [_p, _cr, _p, _i, _i, _i, _p, _p, _p, _i, _i, _p, _p, _p, _p, _i, _p, _p, _p, _p,
                                _p, _p, _p]),
    "rlc_mixture_nll": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "rlc_svgd_action_grads": (_i, [_p, _cr, _p, _i, _p, _i, _p, _i, _f, _f, _p, _p, _p, _p, _p, _p]),
    "rlc_replay_gather": (_i, [_p, _p, _p, _p, _p, _p, _i64, _i, _i, _p, _i, _p, _p, _p, _p, _p, _p]),
    "rlc_replay_rec_stride": (_i, [_i, _i]),
    "rlc_replay_gather_rec": (_i, [_p, _p, _i64, _i, _i, _i, _p, _i, _p, _p, _p, _p, _p, _p]),
    "rlc_replay_scatter_rec": (_i, [_p, _p, _i64, _i, _i, _i, _p, _i, _p, _p, _p, _p, _p, _p]),
    "rlc_replay_sample": (_i, [_p, _i64, _i, C.c_uint64, C.c_uint64, _i64, _i64, _p, _p, _p]),
    "rlc_replay_scatter": (_i, [_p, _p, _p, _p, _p, _p, _i64, _i, _i, _p, _i, _p, _p, _p, _p, _p, _p]),
    "rlc_env_reset": (_i, [_p, _en, _i, _p, _i64, _p, _p, _p, _p, _p, _p, _p]),
    "rlc_env_step_eval": (_i, [_p, _en, _i, _p, _p, _p, _p, _p, _p, _p]),
    "rlc_eval_store": (_i, [_p, _i, _p, _p, _p, _i64, _p, _p, _p]),
    "rlc_env_step_train": (_i, [_p, _en, _p, _p, _p, _p, _p, _p, _i64, _p, _p, _p, _p, _p, _i64, _f, _i64, _p, _p,
                                _i64, _p]),
    "rlc_loop_step": (_i, [_p, _en, _p, _p, _p, _p, _p, _p, _i64, _p, _p, _p, _p, _p, _i64, _f, _i64, _p, _p, _i, _p, _p, _p,
                           _p, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "rlc_sb_forward": (_i, [_p, C.POINTER(RlcSbNet), _i, _i, _p]),
    "rlc_sb_update": (_i, [_p, C.POINTER(RlcSbTrain), _i, _i, _i, _p]),
    "rlc_loop_stage": (_i, [_p, _p, _i, _i, _i64, _p, _p, _p, _i64, _p, _p, _p, _p]),
}

_lib = None


def load():
    """dlopen librlc.so (built in-tree by ``__graft_entry__.build()`` / ``make -C csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RlcError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                       f"g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header/library drift
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int):
    if status != 0:
        lib = load()
        msg = lib.rlc_status_string(status).decode()
        if status == -4:
            msg += ": " + lib.rlc_last_cuda_error().decode()
        raise RlcError(f"librlc status {status}: {msg}")
