"""ctypes binding of ``librnvp_b200.so`` (include/rnvp.h).

There is deliberately no fallback: if the CUDA library is missing the import of
this module raises, and every compute entry point of the package fails with it.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RNVP_B200_LIB", os.path.join(_HERE, "librnvp_b200.so"))

MATH_FP32 = 0
MATH_TF32 = 1
MATH_TF32X3 = 2          # 3xTF32 split operands on the tensor cores: the fp32-accurate tier (include/rnvp.h)
MATH_BY_NAME = {"fp32": MATH_FP32, "tf32": MATH_TF32, "tf32x3": MATH_TF32X3}


class RnvpError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("channels", C.c_int32), ("image_size", C.c_int32), ("base_dim", C.c_int32),
                ("res_blocks", C.c_int32), ("num_scales", C.c_int32),
                ("prior_loc", C.c_float), ("prior_scale", C.c_float)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with dl-normalizing-flows_b200/csrc/build.sh "
            "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, f32, u64, sz = C.c_void_p, C.c_int, C.c_float, C.c_uint64, C.c_size_t
    sigs = {
        "rnvp_last_error": (C.c_char_p, []),
        "rnvp_version": (C.c_char_p, []),
        "rnvp_device_ok": (i32, []),
        "rnvp_launch_count": (C.c_ulonglong, []),
        "rnvp_prof_enable": (i32, [i32]),
        "rnvp_prof_collect": (i32, [C.POINTER(C.c_double), i32]),
        "rnvp_plan_create": (i32, [C.POINTER(Config), C.POINTER(vp)]),
        "rnvp_plan_create_single": (i32, [i32, i32, i32, i32, i32, i32, C.POINTER(vp)]),
        "rnvp_plan_destroy": (i32, [vp]),
        "rnvp_plan_num_couplings": (i32, [vp]),
        "rnvp_plan_slots_per_coupling": (i32, [vp]),
        "rnvp_plan_slot_name": (C.c_char_p, [vp, i32]),
        "rnvp_plan_coupling_info": (i32, [vp, i32, C.c_char_p, i32] + [C.POINTER(i32)] * 5),
        "rnvp_plan_bind": (i32, [vp, C.POINTER(vp), C.POINTER(vp), vp]),
        "rnvp_plan_workspace_bytes": (sz, [vp, i32, i32]),
        "rnvp_plan_set_math": (i32, [vp, i32]),
        "rnvp_plan_forward_generation": (C.c_ulonglong, [vp]),
        "rnvp_flow_forward": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, vp, sz, vp]),
        "rnvp_flow_backward": (i32, [vp, vp, vp, vp, i32, vp, sz, vp]),
        "rnvp_flow_inverse": (i32, [vp, vp, vp, i32, i32, vp, sz, vp]),
        "rnvp_coupling_forward": (i32, [vp, i32, vp, vp, vp, i32, i32, vp, sz, vp]),
        "rnvp_coupling_inverse": (i32, [vp, i32, vp, vp, vp, i32, i32, vp, sz, vp]),
        "rnvp_coupling_backward": (i32, [vp, i32, vp, vp, vp, i32, vp, sz, vp]),
        "rnvp_logit_forward": (i32, [vp, vp, vp, vp, i32, i32, f32, u64, u64, vp]),
        "rnvp_logit_forward_u8": (i32, [vp, vp, vp, vp, i32, i32, f32, u64, u64, vp]),
        "rnvp_logit_inverse": (i32, [vp, vp, sz, f32, vp]),
        "rnvp_squeeze": (i32, [vp, vp, i32, i32, i32, i32, vp]),
        "rnvp_undo_squeeze": (i32, [vp, vp, i32, i32, i32, i32, vp]),
        "rnvp_factor_out": (i32, [vp, vp, vp, i32, i32, i32, i32, vp]),
        "rnvp_restore": (i32, [vp, vp, vp, i32, i32, i32, i32, vp]),
        "rnvp_weightnorm_forward": (i32, [vp, vp, vp, vp, i32, i32, i32, vp]),
        "rnvp_weightnorm_backward": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, vp]),
        "rnvp_conv_forward": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
        "rnvp_conv_wgrad": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
        "rnvp_conv_forward_bn": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, C.c_double,
                                       vp, vp, vp, vp, vp, i32, vp]),
        "rnvp_conv_wgrad_bn": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, i32, vp]),
        "rnvp_bn_relu_forward": (i32, [vp, vp, i32, i32, i32, vp, C.c_double, vp, vp, vp, vp, vp, i32, i32, vp]),
        "rnvp_conv_dgrad_bn": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]),
        "rnvp_bn_backward_apply": (i32, [vp, vp, vp, vp, i32, i32, i32, vp, vp, C.c_double, vp, vp, vp, i32, i32, vp]),
        "rnvp_dp_unique_id": (i32, [vp]),
        "rnvp_dp_init": (i32, [vp, vp, i32, i32]),
        "rnvp_dp_set_grad_layout": (i32, [vp, vp, C.POINTER(C.c_int64), C.c_int64]),
        "rnvp_dp_xchg_alloc": (i32, [vp, i32, vp]),
        "rnvp_dp_xchg_open": (i32, [vp, vp]),
        "rnvp_dp_xchg_errors": (i32, [vp]),
        "rnvp_dp_allreduce_stats": (i32, [vp, vp, sz, vp]),
        "rnvp_dp_finalize": (i32, [vp]),
        "rnvp_dp_allreduce": (i32, [vp, vp, sz, vp]),
        "rnvp_adam_create": (i32, [C.POINTER(vp), C.POINTER(C.c_int64), C.POINTER(C.c_int64), i32, C.POINTER(vp)]),
        "rnvp_adam_destroy": (i32, [vp]),
        "rnvp_adam_step": (i32, [vp, vp, vp, vp, C.c_int64] + [C.c_double] * 5 + [C.c_int64, i32, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype, fn.argtypes = res, args
    return lib, sorted(sigs)


lib, EXPORTS = _load()


def check(status: int) -> None:
    if status != 0:
        msg = lib.rnvp_last_error()
        raise RnvpError(f"rnvp status {status}: {msg.decode() if msg else '?'}")


def ptr(t):
    """Device (or host) pointer of a tensor, or NULL for None."""
    return None if t is None else C.c_void_p(t.data_ptr())
