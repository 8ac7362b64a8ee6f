"""ctypes binding of libachelous_b200.so (include/achelous_b200.h).

This is the stub a maintainer of the reference would add to call the sm_100a kernels
(INTEGRATION.md).  There is NO fallback: if the library is missing or a call fails this
module raises - the product path never silently computes on the CPU or through ATen.
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libachelous_b200.so")

c_float_p = C.c_void_p  # device pointers travel as integers
LL = C.c_longlong
I = C.c_int
F = C.c_float
VP = C.c_void_p

ACT_NONE, ACT_RELU, ACT_SILU, ACT_GELU, ACT_SIGMOID = 0, 1, 2, 3, 4


class AchPwConv(C.Structure):
    _fields_ = [("x0", VP), ("x1", VP), ("wt", VP), ("scale", VP), ("bias", VP), ("pbias", VP), ("res", VP),
                ("gamma", VP), ("out", VP),
                ("x0_bs", LL), ("x1_bs", LL), ("wt_bs", LL), ("res_bs", LL), ("out_bs", LL),
                ("c0", I), ("c1", I), ("ldw", I), ("B", I), ("O", I), ("P", I),
                ("ln", I), ("ln_eps", F), ("act", I), ("reduce_max", I)]


class AchMlp(C.Structure):
    _fields_ = [("x", VP), ("res", VP), ("b1", VP), ("b2", VP), ("gamma", VP), ("out", VP),
                ("x_bs", LL), ("res_bs", LL), ("out_bs", LL), ("B", I), ("C", I), ("P", I), ("ln_eps", F)]


class AchDwConv(C.Structure):
    _fields_ = [("x", VP), ("xadd", VP), ("w", VP), ("scale", VP), ("bias", VP), ("post", VP), ("out", VP),
                ("x_bs", LL), ("xadd_bs", LL), ("out_bs", LL),
                ("B", I), ("C", I), ("H", I), ("W", I), ("Ho", I), ("Wo", I), ("k", I), ("stride", I), ("act", I)]


class AchConvDense(C.Structure):
    _fields_ = [("x", VP), ("w", VP), ("scale", VP), ("bias", VP), ("ln_w", VP), ("ln_b", VP), ("out", VP),
                ("x_bs", LL), ("out_bs", LL),
                ("B", I), ("Cin", I), ("H", I), ("W", I), ("O", I), ("ldo", I), ("Ho", I), ("Wo", I), ("k", I),
                ("stride", I), ("pad", I), ("act", I), ("ln_out", I), ("ln_eps", F)]


class AchConv3x3Tc(C.Structure):
    _fields_ = [("x", VP), ("scale", VP), ("bias", VP), ("res", VP), ("out", VP), ("x_bs", LL), ("res_bs", LL), ("out_bs", LL),
                ("B", I), ("Cin", I), ("H", I), ("W", I), ("O", I), ("act", I)]


class AchRcDeform(C.Structure):
    _fields_ = [("x", VP), ("pooled", VP), ("w_om", VP), ("b_om", VP), ("w_reg", VP), ("w1", VP), ("scale", VP),
                ("bias", VP), ("out", VP), ("x_bs", LL), ("pooled_bs", LL), ("out_bs", LL),
                ("B", I), ("C", I), ("H", I), ("W", I), ("pooled_cl", I)]


class AchUpGhost(C.Structure):
    _fields_ = [("v", VP), ("b1", VP), ("w2", VP), ("s2", VP), ("b2", VP), ("out", VP), ("v_bs", LL), ("out_bs", LL),
                ("B", I), ("Ci", I), ("Cn", I), ("h", I), ("w", I)]


class AchUpGhostHead(C.Structure):
    _fields_ = [("v", VP), ("out", VP), ("b1", VP), ("w2", VP), ("s2", VP), ("b2", VP), ("w3", VP), ("b3", VP), ("w4", VP),
                ("s4", VP), ("b4", VP), ("v_bs", LL), ("out_bs", LL), ("B", I), ("C", I), ("init", I), ("K", I), ("h", I), ("w", I)]


class AchUpGhostPw2(C.Structure):
    _fields_ = [("v", VP), ("out", VP), ("b1", VP), ("w2", VP), ("s2", VP), ("b2", VP), ("w1t", VP), ("c1", VP), ("w2t", VP),
                ("v_bs", LL), ("out_bs", LL), ("B", I), ("Ci", I), ("C1", I), ("N2", I), ("h", I), ("w", I)]


_SIGNATURES = {
    "ach_version": ([], I),
    "ach_pw_conv": ([C.POINTER(AchPwConv), VP], I),
    "ach_pack_pw_tc_elems": ([I, I], LL),
    "ach_pack_pw_tc": ([VP, I, I, I, VP, VP, VP], I),
    "ach_pw_conv_tc": ([C.POINTER(AchPwConv), VP, VP, VP, VP], I),
    "ach_pack_pw_tc_nt_elems": ([I, I, I], LL),
    "ach_pack_pw_tc_nt": ([VP, I, I, I, I, VP, VP, VP], I),
    "ach_mlp_tc_supported": ([I], I),
    "ach_mlp_tc": ([C.POINTER(AchMlp), VP, VP, VP, VP, VP, VP], I),
    "ach_dw_conv": ([C.POINTER(AchDwConv), VP], I),
    "ach_conv_dense": ([C.POINTER(AchConvDense), VP], I),
    "ach_layernorm_cf": ([VP, LL, VP, VP, VP, LL, I, I, I, F, VP], I),
    "ach_ln_s2d": ([VP, LL, VP, VP, VP, LL, I, I, I, I, F, VP], I),
    "ach_upsample2x": ([VP, LL, VP, LL, I, I, I, I, VP], I),
    "ach_spp_maxpool": ([VP, LL, VP, VP, VP, LL, I, I, I, I, VP], I),
    "ach_shuffle_attention": ([VP, LL, VP, LL, VP, VP, VP, VP, VP, VP, I, I, I, I, F, VP], I),
    "ach_plane_mean": ([VP, LL, VP, LL, VP, I, I, I, VP], I),
    "ach_eca_fuse": ([VP, LL, VP, LL, VP, VP, I, VP, VP, VP, LL, I, I, I, VP], I),
    "ach_avgpool3": ([VP, LL, VP, LL, I, I, I, I, VP], I),
    "ach_avgpool3_cl": ([VP, LL, VP, LL, I, I, I, I, VP], I),
    "ach_rc_deform": ([C.POINTER(AchRcDeform), VP], I),
    "ach_rc_deform_tc_supported": ([I], I),
    "ach_rc_deform_tc": ([C.POINTER(AchRcDeform), VP, VP, VP, VP, VP], I),
    "ach_xca_fold": ([VP, LL, VP, VP, I, VP, LL, I, I, I, I, VP], I),
    "ach_mvit_attention": ([VP, LL, VP, LL, I, I, I, I, I, VP], I),
    "ach_mvit_attention_tc": ([VP, LL, VP, LL, I, I, I, I, I, VP], I),
    "ach_fc": ([VP, LL, VP, VP, VP, VP, LL, I, I, I, I, VP], I),
    "ach_logsoftmax_t": ([VP, LL, VP, LL, I, I, I, VP], I),
    "ach_copy_add": ([VP, LL, VP, VP, LL, I, I, I, VP], I),
    "ach_add": ([VP, LL, VP, LL, VP, LL, I, I, I, VP], I),
    "ach_fill": ([VP, LL, F, VP], I),
    "ach_up_ghost": ([C.POINTER(AchUpGhost), VP], I),
    "ach_up_ghost_pw2_supported": ([I, I, I], I),
    "ach_up_ghost_pw2": ([C.POINTER(AchUpGhostPw2), VP], I),
    "ach_up_ghost_pw2_tc_supported": ([I, I, I], I),
    "ach_up_ghost_pw2_tc": ([C.POINTER(AchUpGhostPw2), VP, VP, VP, VP, VP, VP], I),
    "ach_up_ghost_head_supported": ([I, I, I], I),
    "ach_up_ghost_head": ([C.POINTER(AchUpGhostHead), VP], I),
    "ach_up_ghost_head_argmax": ([C.POINTER(AchUpGhostHead), VP, LL, C.c_uint, VP], I),
    "ach_pn2_fps": ([VP, LL, I, I, I, VP, VP, LL, VP], I),
    "ach_pn2_group": ([VP, LL, VP, LL, I, VP, LL, I, I, I, I, F, VP, LL, VP, VP], I),
    "ach_pn2_group_max": ([VP, LL, VP, LL, I, I, I, I, VP], I),
    "ach_pn2_interp3": ([VP, LL, VP, LL, VP, LL, I, I, I, I, VP, LL, VP], I),
    "ach_seg_softmax": ([VP, LL, VP, LL, I, I, I, VP], I),
    "ach_seg_resize_argmax": ([VP, LL, I, I, I, I, I, I, I, I, VP, I, I, VP], I),
    "ach_seg_softmax_resize_argmax": ([VP, LL, I, I, I, I, I, I, I, I, VP, LL, I, I, C.c_uint, VP], I),
    "ach_seg_argmax_u8": ([VP, LL, I, I, I, C.c_uint, VP, LL, VP], I),
    "ach_logsoftmax_argmax_t": ([VP, LL, VP, LL, I, I, I, VP], I),
    "ach_ef_attention": ([VP, LL, VP, LL, VP, LL, VP, VP, VP, VP, LL, VP, LL, I, I, I, I, I, I, F, I, VP], I),
    "ach_upsample2x_hp": ([VP, LL, VP, LL, I, I, I, I, I, VP], I),
    "ach_s2d": ([VP, LL, VP, LL, I, I, I, I, I, VP], I),
    "ach_subsample": ([VP, LL, VP, LL, I, I, I, I, I, VP], I),
    "ach_mhsa": ([VP, LL, VP, LL, I, I, I, I, F, VP], I),
    "ach_dw_convT": ([VP, LL, VP, VP, VP, LL, I, I, I, I, I, VP], I),
    "ach_conv3x3_tc_k": ([I], I),
    "ach_conv3x3_tc": ([C.POINTER(AchConv3x3Tc), VP, VP, VP], I),
    "ach_pre_resize_h": ([VP, LL, I, I, I, I, I, VP, VP, I, VP, LL, VP], I),
    "ach_pre_resize_v_norm": ([VP, LL, I, I, I, VP, VP, I, I, VP, LL, I, I, I, I, VP], I),
    "ach_pre_radar": ([VP, LL, I, I, LL, VP, LL, VP], I),
    "ach_pre_points": ([VP, I, I, VP, I, I, VP, VP], I),
    "ach_decode_outputs": ([C.POINTER(VP), C.POINTER(LL), C.POINTER(I), C.POINTER(I), I, VP, I, I, F, F, VP], I),
    "ach_nms_workspace_bytes": ([I, I], LL),
    "ach_nms": ([VP, I, I, I, F, F, VP, VP, VP, VP, LL, VP], I),
    "ach_nms_rows": ([VP, I, I, I, F, F, VP, LL, I, VP, LL, VP, LL, VP], I),
    "ach_peer_alloc": ([LL, C.POINTER(VP)], I),
    "ach_peer_free": ([VP], I),
    "ach_peer_handle_bytes": ([], I),
    "ach_peer_export": ([VP, C.c_char_p], I),
    "ach_peer_open": ([C.c_char_p, C.POINTER(VP)], I),
    "ach_peer_close": ([VP], I),
    "ach_peer_copy": ([VP, VP, LL, VP], I),
    "ach_peer_signal": ([VP, I, C.c_uint, VP], I),
    "ach_peer_wait": ([VP, I, C.c_uint, VP], I),
}

EXPORTED_SYMBOLS = ["ach_last_error"] + list(_SIGNATURES)

_lib = None


class AchelousKernelError(RuntimeError):
    pass


def load():
    """Loads the shared library; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise AchelousKernelError(
            f"{LIB_PATH} not found: build it with `python -m achelous_b200.build` "
            "(achelous_b200 has no CPU / ATen fallback by design)")
    lib = C.CDLL(LIB_PATH)
    lib.ach_last_error.argtypes = []
    lib.ach_last_error.restype = C.c_char_p
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def check(status, what=""):
    if status != 0:
        msg = load().ach_last_error().decode(errors="replace")
        raise AchelousKernelError(f"{what or 'achelous_b200 kernel'} failed (status {status}): {msg}")
