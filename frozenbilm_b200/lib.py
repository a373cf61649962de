"""ctypes binding of libfrozenbilm_sm100.so (the C-ABI in include/frozenbilm_b200.h).

There is no fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfrozenbilm_sm100.so")
LIB_DBG_PATH = os.path.join(_HERE, "libfrozenbilm_sm100_dbg.so")   # test-only (include/frozenbilm_b200_debug.h)

FB_ACT_NONE, FB_ACT_GELU, FB_ACT_RELU = 0, 1, 2
FB_AUX_NONE, FB_AUX_MUL_GELU_GRAD, FB_AUX_MUL_RELU_MASK, FB_AUX_ADD = 0, 1, 2, 3
FB_DT_BF16, FB_DT_F32 = 0, 1
SCORE_DTYPE = torch.float16   # dtype of the saved attention scores (fb_attn_fwd s_out -> fb_attn_bwd s2)


class FbError(RuntimeError):
    pass


class Dropout(C.Structure):
    _fields_ = [("p", C.c_float), ("seed", C.c_ulonglong), ("offset", C.c_ulonglong)]


def drop(p=0.0, seed=0, offset=0):
    """fb_dropout value (None / p == 0 -> disabled)."""
    return Dropout(float(p), int(seed), int(offset))


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("batch", C.c_int),
        ("A", C.c_void_p), ("lda", C.c_longlong), ("stride_a", C.c_longlong), ("a_mn_major", C.c_int),
        ("B", C.c_void_p), ("ldb", C.c_longlong), ("stride_b", C.c_longlong), ("b_mn_major", C.c_int),
        ("D", C.c_void_p), ("ldd", C.c_longlong), ("stride_d", C.c_longlong), ("d_dtype", C.c_int),
        ("alpha", C.c_float),
        ("bias", C.c_void_p),
        ("D_pre", C.c_void_p), ("ld_pre", C.c_longlong),
        ("act", C.c_int),
        ("aux", C.c_void_p), ("ld_aux", C.c_longlong), ("aux_op", C.c_int),
        ("resid", C.c_void_p), ("ld_resid", C.c_longlong),
        ("accumulate", C.c_int),
        ("split_k", C.c_int),
        ("block_n", C.c_int),
        ("batch_inner", C.c_int), ("stride_a_in", C.c_longlong), ("stride_b_in", C.c_longlong),
        ("stride_d_in", C.c_longlong),
        ("rowmask", C.c_void_p),
        ("drop", Dropout), ("drop_pos", C.c_int),
        ("stride_bias", C.c_longlong),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("L", C.c_int), ("h", C.c_int),
        ("Q", C.c_void_p), ("ldq", C.c_longlong),
        ("K", C.c_void_p), ("ldk", C.c_longlong),
        ("V", C.c_void_p), ("ldv", C.c_longlong),
        ("PKrel", C.c_void_p), ("PQrel", C.c_void_p), ("rel_rows", C.c_int), ("rel_off", C.c_int),
        ("mask", C.c_void_p),
        ("ctx", C.c_void_p), ("ldc", C.c_longlong),
        ("lse", C.c_void_p),
        ("scale", C.c_float),
        ("dropout_p", C.c_float), ("seed", C.c_ulonglong), ("offset", C.c_ulonglong),
        ("s_out", C.c_void_p),
        ("key_mask_only", C.c_int),
    ]


class AttnBwdPrepArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("L", C.c_int), ("h", C.c_int), ("lp", C.c_int), ("rel_rows", C.c_int), ("rel_off", C.c_int),
        ("s2", C.c_void_p), ("lse", C.c_void_p),
        ("dp", C.c_void_p),
        ("ctx", C.c_void_p), ("dctx", C.c_void_p), ("ldc", C.c_longlong),
        ("ds", C.c_void_p), ("pd", C.c_void_p), ("z1", C.c_void_p), ("z2", C.c_void_p),
        ("scale", C.c_float),
        ("dropout_p", C.c_float), ("seed", C.c_ulonglong), ("offset", C.c_ulonglong),
    ]


class AttnBwdArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("L", C.c_int), ("h", C.c_int), ("lp", C.c_int),
        ("Q", C.c_void_p), ("ldq", C.c_longlong),
        ("K", C.c_void_p), ("ldk", C.c_longlong),
        ("V", C.c_void_p), ("ldv", C.c_longlong),
        ("O", C.c_void_p), ("dO", C.c_void_p), ("ldo", C.c_longlong),
        ("s2", C.c_void_p), ("lse", C.c_void_p),
        ("PKrev", C.c_void_p), ("PQrel", C.c_void_p), ("rel_rows", C.c_int), ("rel_off", C.c_int),
        ("dQ", C.c_void_p), ("lddq", C.c_longlong),
        ("dK", C.c_void_p), ("lddk", C.c_longlong),
        ("dV", C.c_void_p), ("lddv", C.c_longlong),
        ("ds_out", C.c_void_p),
        ("scale", C.c_float),
        ("dropout_p", C.c_float), ("seed", C.c_ulonglong), ("offset", C.c_ulonglong),
    ]


class AttnBwdRelArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("L", C.c_int), ("h", C.c_int), ("lp", C.c_int), ("rel_rows", C.c_int), ("rel_off", C.c_int),
        ("ds", C.c_void_p),
        ("Q", C.c_void_p), ("ldq", C.c_longlong),
        ("K", C.c_void_p), ("ldk", C.c_longlong),
        ("gpk", C.c_void_p), ("gpq", C.c_void_p),
        ("units_per_cta", C.c_int),
        ("scale", C.c_float),
    ]


class EpilogueArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("N", C.c_int),
        ("inp", C.c_void_p), ("ld_in", C.c_longlong), ("alpha", C.c_float),
        ("bias", C.c_void_p), ("rowmask", C.c_void_p), ("drop", Dropout), ("drop_pos", C.c_int),
        ("D_pre", C.c_void_p), ("ld_pre", C.c_longlong), ("act", C.c_int),
        ("aux", C.c_void_p), ("ld_aux", C.c_longlong), ("aux_op", C.c_int),
        ("resid", C.c_void_p), ("ld_resid", C.c_longlong), ("accumulate", C.c_int),
        ("D", C.c_void_p), ("ldd", C.c_longlong),
    ]


class AttnF32Args(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("L", C.c_int), ("h", C.c_int),
        ("Q", C.c_void_p), ("ldq", C.c_longlong), ("K", C.c_void_p), ("ldk", C.c_longlong),
        ("V", C.c_void_p), ("ldv", C.c_longlong),
        ("pqk", C.c_void_p), ("ld_pqk", C.c_longlong), ("bucket_of_rel", C.c_void_p),
        ("mask", C.c_void_p), ("key_mask_only", C.c_int),
        ("scale", C.c_float), ("dropout_p", C.c_float), ("seed", C.c_ulonglong), ("offset", C.c_ulonglong),
        ("ctx", C.c_void_p), ("ldc", C.c_longlong), ("lse", C.c_void_p),
        ("dctx", C.c_void_p), ("dQ", C.c_void_p), ("lddq", C.c_longlong), ("dK", C.c_void_p), ("lddk", C.c_longlong),
        ("dV", C.c_void_p), ("lddv", C.c_longlong),
        ("dpqk", C.c_void_p), ("ld_dpqk", C.c_longlong), ("ws_ds", C.c_void_p), ("ws_pd", C.c_void_p),
    ]


class UmmaProbeArgs(C.Structure):
    _fields_ = [
        ("a_image", C.c_void_p), ("a_bytes", C.c_int),
        ("b_image", C.c_void_p), ("b_bytes", C.c_int),
        ("n_mma", C.c_int), ("a_off", C.POINTER(C.c_int)), ("b_off", C.POINTER(C.c_int)),
        ("a_desc", C.c_ulonglong), ("b_desc", C.c_ulonglong),
        ("idesc", C.c_uint),
        ("N", C.c_int),
        ("out", C.c_void_p),
        ("cycles", C.c_void_p),
    ]


_lib = None
_lib_dbg = None


def use_debug_library():
    """scripts/ and tests/ only: route every call of this process through the debug build (same kernels compiled with
    -DFB_DEBUG_SURFACE plus fb_umma_probe / fb_attn_bwd_set_debug).  Must be called before the first load()."""
    global _lib, _lib_dbg
    if _lib is not None and _lib is not _lib_dbg:
        raise FbError("use_debug_library() must be called before the product library is loaded")
    _lib = _lib_dbg = _load(LIB_DBG_PATH, EXPORTS + DEBUG_EXPORTS)
    return _lib


def load() -> C.CDLL:
    """Loads the shared library; raises FbError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    _lib = _load(LIB_PATH, EXPORTS)
    return _lib


def _load(path, exports) -> C.CDLL:
    if not os.path.exists(path):
        raise FbError(f"{path} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no CPU / eager fallback)")
    lib = C.CDLL(path)
    lib.fb_last_error.restype = C.c_char_p
    lib.fb_abi_version.restype = C.c_int
    lib.fb_launch_count.restype = C.c_longlong
    if lib.fb_abi_version() != ABI_VERSION:
        raise FbError(f"{path} has ABI version {lib.fb_abi_version()}, this binding expects {ABI_VERSION}: rebuild it "
                      "(python -c 'import __graft_entry__ as g; g.build()')")
    for name in exports:
        if name in ("fb_last_error", "fb_abi_version", "fb_launch_count"):
            continue
        getattr(lib, name).restype = C.c_int
    vp, ll, i, f = C.c_void_p, C.c_longlong, C.c_int, C.c_float
    dp = C.POINTER(Dropout)
    lib.fb_layernorm_fwd.argtypes = [vp, i, i, vp, vp, f, vp, dp, vp, vp, vp, vp, vp]
    lib.fb_layernorm_bwd.argtypes = [vp, vp, vp, vp, vp, i, i, vp, dp, dp, vp, vp, vp, vp, vp, vp]
    lib.fb_move_rows.argtypes = [vp, ll, i, vp, vp, i, i, vp, ll, i, vp]
    lib.fb_gelu_bwd.argtypes = [vp, vp, i, i, vp, dp, vp, vp]
    lib.fb_add_rowbcast.argtypes = [vp, vp, i, i, i, vp, vp, vp]
    lib.fb_conv_im2col.argtypes = [vp, i, i, i, vp, vp]
    lib.fb_conv_col2im.argtypes = [vp, i, i, i, vp, i, vp]
    lib.fb_colsum_bf16.argtypes = [vp, ll, i, i, vp, vp]
    lib.fb_cross_entropy.argtypes = [vp, ll, i, i, vp, vp, f, vp, ll, f, vp]
    lib.fb_dropout_apply.argtypes = [vp, i, i, dp, vp, i, vp, vp]
    lib.fb_sumsq.argtypes = [vp, ll, vp, vp]
    lib.fb_sumsq_det.argtypes = [vp, ll, vp, vp, i, vp]
    lib.fb_adam_step.argtypes = [vp, vp, vp, vp, ll, f, f, f, f, i, vp, f, f, vp, vp]
    lib.fb_cast_f32_bf16.argtypes = [vp, ll, vp, vp]
    lib.fb_build_rel_tables.argtypes = [vp, ll, i, i, vp, i, i, i, vp, vp, vp, vp]
    lib.fb_rel_to_bucket.argtypes = [vp, vp, i, i, vp, i, i, i, vp, ll, vp]
    lib.fb_build_rel_tables_batched.argtypes = [vp, ll, i, i, vp, i, i, i, vp, vp, vp, i, ll, vp]
    lib.fb_rel_to_bucket_batched.argtypes = [vp, vp, i, i, vp, i, i, i, vp, ll, i, ll, vp]
    lib.fb_dropout_batched.argtypes = [vp, i, i, i, dp, C.c_ulonglong, vp, vp, vp]
    lib.fb_mlm_corrupt.argtypes = [vp, vp, ll, ll, ll, ll, f, C.c_ulonglong, vp, vp, vp]
    lib.fb_gather_frames.argtypes = [vp, vp, vp, i, i, i, vp, vp, vp, vp]
    lib.fb_softmax_topk.argtypes = [vp, ll, i, i, i, vp, ll, vp, vp, vp]
    lib.fb_split3.argtypes = [vp, ll, i, i, i, i, ll, vp, ll, vp]
    lib.fb_colsum_f32.argtypes = [vp, ll, i, i, vp, vp]
    lib.fb_conv_im2col_f32.argtypes = [vp, i, i, i, vp, vp]
    lib.fb_conv_col2im_f32.argtypes = [vp, i, i, i, vp, i, vp]
    lib.fb_cross_entropy_f32.argtypes = [vp, ll, i, i, vp, vp, f, vp, ll, f, vp]
    return lib


ABI_VERSION = 3   # FB_ABI_VERSION of include/frozenbilm_b200.h this binding (struct layouts, argtypes) was written for

# every symbol include/frozenbilm_b200.h declares (tests check the .so exports all of them)
EXPORTS = [
    "fb_last_error", "fb_abi_version", "fb_launch_count", "fb_gemm", "fb_attn_fwd", "fb_build_rel_tables",
    "fb_layernorm_fwd", "fb_layernorm_bwd", "fb_move_rows", "fb_gelu_bwd", "fb_add_rowbcast", "fb_conv_im2col",
    "fb_conv_col2im", "fb_colsum_bf16", "fb_cross_entropy", "fb_dropout_apply", "fb_sumsq", "fb_adam_step",
    "fb_cast_f32_bf16", "fb_attn_bwd_prep", "fb_rel_to_bucket", "fb_attn_bwd", "fb_attn_bwd_rel",
    "fb_mlm_corrupt", "fb_gather_frames", "fb_gemm_cluster_capacity",
    "fb_split3", "fb_epilogue_f32", "fb_colsum_f32", "fb_conv_im2col_f32", "fb_conv_col2im_f32",
    "fb_cross_entropy_f32", "fb_attn_fwd_f32", "fb_attn_bwd_f32", "fb_softmax_topk", "fb_sumsq_det", "fb_set_pdl",
    "fb_build_rel_tables_batched", "fb_rel_to_bucket_batched", "fb_dropout_batched",
]
# include/frozenbilm_b200_debug.h: only in libfrozenbilm_sm100_dbg.so
DEBUG_EXPORTS = ["fb_umma_probe", "fb_attn_bwd_set_debug"]


def check(rc: int, what: str):
    if rc != 0:
        raise FbError(f"{what} failed ({rc}): {load().fb_last_error().decode()}")


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dp(d):
    return C.byref(d) if d is not None else None


def set_pdl(on: bool) -> bool:
    """Programmatic dependent launch for the following launches; returns the previous setting."""
    return bool(load().fb_set_pdl(int(bool(on))))


def launch_count() -> int:
    return int(load().fb_launch_count())


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def gemm(A: torch.Tensor, B: torch.Tensor, D: torch.Tensor, *, a_mn_major=False, b_mn_major=False,
         bias=None, act=FB_ACT_NONE, D_pre=None, aux=None, aux_op=FB_AUX_NONE, resid=None, alpha=1.0,
         accumulate=False, split_k=1, block_n=0, M=None, N=None, K=None, rowmask=None, dropout=None,
         drop_pos=0, batch=None):
    """D = epilogue(alpha * A @ B^T) on the tcgen05 GEMM.

    A: bf16 [M, K] (or [K, M] when a_mn_major);  B: bf16 [N, K] (or [K, N] when b_mn_major);
    2-D (or 3-D batched) views with unit inner stride.  D: bf16 or fp32 [M, N].
    """
    if A.dtype == torch.float32:
        return _gemm_split3(A, B, D, a_mn_major=a_mn_major, b_mn_major=b_mn_major, bias=bias, act=act, D_pre=D_pre,
                            aux=aux, aux_op=aux_op, resid=resid, alpha=alpha, accumulate=accumulate or split_k > 1,
                            M=M, N=N, K=K, rowmask=rowmask, dropout=dropout, drop_pos=drop_pos, batch=batch)
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    assert A.stride(-1) == 1 and B.stride(-1) == 1 and D.stride(-1) == 1
    nbatch = 1
    sa = sb = sd = 0
    bin_, sai, sbi, sdi = 1, 0, 0, 0
    args_stride_bias = 0
    if batch is not None:
        # explicit batching over 2-D base views: batch = dict(outer=(n, sa, sb, sd), inner=(n, sa, sb, sd))
        A2, B2, D2 = A, B, D
        nbatch, sa, sb, sd = batch.get("outer", (1, 0, 0, 0))
        args_stride_bias = int(batch.get("bias_stride", 0))
        bin_, sai, sbi, sdi = batch.get("inner", (1, 0, 0, 0))
    elif A.dim() == 3:
        nbatch = A.shape[0]
        sa, sb, sd = A.stride(0), B.stride(0), D.stride(0)
        A2, B2, D2 = A[0], B[0], D[0]
    else:
        A2, B2, D2 = A, B, D
    if M is None:
        M = A2.shape[1] if a_mn_major else A2.shape[0]
    if K is None:
        K = A2.shape[0] if a_mn_major else A2.shape[1]
    if N is None:
        N = B2.shape[1] if b_mn_major else B2.shape[0]
    args = GemmArgs()
    args.M, args.N, args.K, args.batch = M, N, K, nbatch
    args.batch_inner, args.stride_a_in, args.stride_b_in, args.stride_d_in = bin_, sai, sbi, sdi
    if rowmask is not None:
        assert rowmask.dtype == torch.uint8
        args.rowmask = _ptr(rowmask)
    if dropout is not None and dropout.p > 0:
        args.drop, args.drop_pos = dropout, int(drop_pos)
    args.A, args.lda, args.stride_a, args.a_mn_major = _ptr(A), A2.stride(0), sa, int(a_mn_major)
    args.B, args.ldb, args.stride_b, args.b_mn_major = _ptr(B), B2.stride(0), sb, int(b_mn_major)
    args.D, args.ldd, args.stride_d = _ptr(D), D2.stride(0), sd
    args.d_dtype = FB_DT_F32 if D.dtype == torch.float32 else FB_DT_BF16
    assert D.dtype in (torch.float32, torch.bfloat16)
    args.alpha = float(alpha)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
    args.bias = _ptr(bias)
    args.stride_bias = args_stride_bias
    if D_pre is not None:
        assert D_pre.dtype == torch.bfloat16
        args.D_pre, args.ld_pre = _ptr(D_pre), D_pre.stride(0)
    args.act = act
    if aux is not None:
        assert aux.dtype == torch.bfloat16
        args.aux, args.ld_aux, args.aux_op = _ptr(aux), aux.stride(0), aux_op
    if resid is not None:
        assert resid.dtype == torch.float32
        args.resid, args.ld_resid = _ptr(resid), resid.stride(0)
    args.accumulate = int(accumulate)
    args.split_k = int(split_k)
    args.block_n = int(block_n)
    check(load().fb_gemm(C.byref(args), stream_ptr()), "fb_gemm")
    return D


# --------------------------------------------------------------------------------------------------------------
# fp32-class instantiation (FB_PRECISION=fp32x3): the engine keeps its "bf16" buffers in fp32 and every wrapper
# below routes an fp32-typed operand to the split-bf16 / fp32 kernels of fb_precise.cu.  Same schedule, same call
# sites; only the operand precision differs.  Not tuned (temporaries are allocated per call).
# --------------------------------------------------------------------------------------------------------------
def split3(x: torch.Tensor, rows: int, cols: int, *, along_rows: bool, b_mode: bool):
    """fp32 2-D view [>= rows, >= cols] -> bf16 [rows, 6 Kp] (K = columns) or [6 Kp, ceil8(cols)] (K = rows)."""
    assert x.dtype == torch.float32 and x.stride(-1) == 1
    if along_rows:
        kp = (rows + 7) // 8 * 8
        out = torch.zeros(6 * kp, (cols + 7) // 8 * 8, device=x.device, dtype=torch.bfloat16)
    else:
        kp = (cols + 7) // 8 * 8
        out = torch.zeros(rows, 6 * kp, device=x.device, dtype=torch.bfloat16)
    check(load().fb_split3(_ptr(x), x.stride(0), rows, cols, int(along_rows), int(b_mode), kp, _ptr(out),
                           out.stride(0), stream_ptr()), "fb_split3")
    return out, kp


def epilogue_f32(inp, D, *, alpha=1.0, bias=None, rowmask=None, dropout=None, drop_pos=0, D_pre=None, act=FB_ACT_NONE,
                 aux=None, aux_op=FB_AUX_NONE, resid=None, accumulate=False, M=None, N=None):
    a = EpilogueArgs()
    a.M, a.N = (inp.shape[0] if M is None else M), (inp.shape[1] if N is None else N)
    for t in (inp, D, D_pre, aux, resid):
        assert t is None or (t.dtype == torch.float32 and t.stride(-1) == 1)
    a.inp, a.ld_in, a.alpha = _ptr(inp), inp.stride(0), float(alpha)
    a.bias, a.rowmask = _ptr(bias), _ptr(rowmask)
    if dropout is not None and dropout.p > 0:
        a.drop, a.drop_pos = dropout, int(drop_pos)
    if D_pre is not None:
        a.D_pre, a.ld_pre = _ptr(D_pre), D_pre.stride(0)
    a.act = act
    if aux is not None:
        a.aux, a.ld_aux, a.aux_op = _ptr(aux), aux.stride(0), aux_op
    if resid is not None:
        a.resid, a.ld_resid = _ptr(resid), resid.stride(0)
    a.accumulate = int(accumulate)
    a.D, a.ldd = _ptr(D), D.stride(0)
    check(load().fb_epilogue_f32(C.byref(a), stream_ptr()), "fb_epilogue_f32")
    return D


def _gemm_split3(A, B, D, *, a_mn_major, b_mn_major, bias, act, D_pre, aux, aux_op, resid, alpha, accumulate, M, N, K,
                 rowmask, dropout, drop_pos, batch):
    """fb_gemm at fp32-class precision: each operand split into three bf16 pieces laid out as six blocks along K, ONE
    tcgen05 GEMM over the 6x longer K (all products above 2^-24, fp32 accumulation), then the fp32 epilogue."""
    if batch is not None or A.dim() != 2:
        raise NotImplementedError("split-bf16 GEMM: batched mode is not part of the fp32-class instantiation")
    assert B.dtype == torch.float32 and D.dtype == torch.float32
    if M is None:
        M = A.shape[1] if a_mn_major else A.shape[0]
    if K is None:
        K = A.shape[0] if a_mn_major else A.shape[1]
    if N is None:
        N = B.shape[1] if b_mn_major else B.shape[0]
    A3, kp = split3(A, K, M, along_rows=True, b_mode=False) if a_mn_major else \
        split3(A, M, K, along_rows=False, b_mode=False)
    B3, kp2 = split3(B, K, N, along_rows=True, b_mode=True) if b_mn_major else \
        split3(B, N, K, along_rows=False, b_mode=True)
    assert kp == kp2
    # TMEM accumulation is not round-to-nearest: every tcgen05.mma added into a live accumulator truncates (measured:
    # relative error ~ 2^-24 x the number of accumulating instructions: 1.5e-5 after 414 of them, worse than the split
    # itself).  So each CTA accumulates ONE 64-wide K block (4 instructions) and the partial tiles are summed by
    # fp32 reductions in L2 (split-K with cp.reduce.async.bulk / red.global.add.f32, both round-to-nearest).
    raw = torch.zeros(M, (N + 7) // 8 * 8, device=A.device, dtype=torch.float32)
    gemm(A3, B3, raw, a_mn_major=a_mn_major, b_mn_major=b_mn_major, M=M, N=N, K=6 * kp,
         split_k=(6 * kp + 63) // 64)
    return epilogue_f32(raw, D, alpha=alpha, bias=bias, rowmask=rowmask, dropout=dropout, drop_pos=drop_pos,
                        D_pre=D_pre, act=act, aux=aux, aux_op=aux_op, resid=resid, accumulate=accumulate, M=M, N=N)


def _attn_f32_args(Q, K, V, pqk, bucket_of_rel, mask_u8, ctx, lse, B, L, h, scale, drop, key_mask_only):
    a = AttnF32Args()
    a.B, a.L, a.h = B, L, h
    for t in (Q, K, V, ctx, lse):
        assert t.dtype == torch.float32
    a.Q, a.ldq, a.K, a.ldk, a.V, a.ldv = _ptr(Q), Q.stride(0), _ptr(K), K.stride(0), _ptr(V), V.stride(0)
    if pqk is not None:
        assert pqk.dtype == torch.float32 and bucket_of_rel.dtype == torch.int32
        a.pqk, a.ld_pqk, a.bucket_of_rel = _ptr(pqk), pqk.stride(0), _ptr(bucket_of_rel)
    a.mask, a.key_mask_only = _ptr(mask_u8), int(key_mask_only)
    a.scale = float(scale)
    if drop is not None:
        a.dropout_p, a.seed, a.offset = float(drop.p), int(drop.seed), int(drop.offset)
    a.ctx, a.ldc, a.lse = _ptr(ctx), ctx.stride(0), _ptr(lse)
    return a


def attn_fwd_f32(Q, K, V, pqk, bucket_of_rel, mask_u8, ctx, lse, *, B, L, h, scale, drop=None, key_mask_only=False):
    a = _attn_f32_args(Q, K, V, pqk, bucket_of_rel, mask_u8, ctx, lse, B, L, h, scale, drop, key_mask_only)
    check(load().fb_attn_fwd_f32(C.byref(a), stream_ptr()), "fb_attn_fwd_f32")


def attn_bwd_f32(Q, K, V, pqk, bucket_of_rel, mask_u8, ctx, lse, dctx, dQ, dK, dV, dpqk, *, B, L, h, scale, drop=None,
                 key_mask_only=False):
    a = _attn_f32_args(Q, K, V, pqk, bucket_of_rel, mask_u8, ctx, lse, B, L, h, scale, drop, key_mask_only)
    assert dctx.stride(0) == ctx.stride(0)
    ws = torch.empty(2, B, h, L, L, device=Q.device, dtype=torch.float32)
    a.dctx = _ptr(dctx)
    a.dQ, a.lddq, a.dK, a.lddk, a.dV, a.lddv = _ptr(dQ), dQ.stride(0), _ptr(dK), dK.stride(0), _ptr(dV), dV.stride(0)
    if dpqk is not None:
        a.dpqk, a.ld_dpqk = _ptr(dpqk), dpqk.stride(0)
    a.ws_ds, a.ws_pd = _ptr(ws[0]), _ptr(ws[1])
    check(load().fb_attn_bwd_f32(C.byref(a), stream_ptr()), "fb_attn_bwd_f32")


def rel_table_geometry(L: int):
    """(rel_rows, rel_off) of the per-relative-position tables for sequence length L."""
    lpad = (L + 127) // 128 * 128
    return 2 * lpad + 8, lpad


def build_rel_tables(pqk: torch.Tensor, bucket_of_rel: torch.Tensor, L: int, h: int, out_pq=None, out_pk=None,
                     out_pk_rev=None):
    """pqk: bf16 [2S, 2H] (PQ | PK).  Returns (PQrel, PKrel) bf16 [h, rel_rows, 64]; ``out_pk_rev`` (optional, same
    shape) receives PKrel with its rows reversed (the operand of the fused backward)."""
    H = pqk.shape[1] // 2
    rel_rows, rel_off = rel_table_geometry(L)
    if out_pq is None:
        out_pq = torch.empty(h, rel_rows, 64, device=pqk.device, dtype=torch.bfloat16)
    if out_pk is None:
        out_pk = torch.empty(h, rel_rows, 64, device=pqk.device, dtype=torch.bfloat16)
    assert bucket_of_rel.dtype == torch.int32 and bucket_of_rel.numel() == 1023
    check(load().fb_build_rel_tables(_ptr(pqk), C.c_longlong(pqk.stride(0)), H, h, _ptr(bucket_of_rel), rel_rows,
                                     rel_off, L, _ptr(out_pq), _ptr(out_pk), _ptr(out_pk_rev), stream_ptr()),
          "fb_build_rel_tables")
    return out_pq, out_pk


def build_rel_tables_batched(pqk: torch.Tensor, bucket_of_rel: torch.Tensor, L: int, h: int, out_pq, out_pk,
                             out_pk_rev=None):
    """pqk: bf16 [nb, 2S, 2H]; out_*: bf16 [nb, h, rel_rows, 64] (contiguous): the tables of nb layer passes in ONE
    launch."""
    nb, _, H2 = pqk.shape
    rel_rows, rel_off = rel_table_geometry(L)
    for t in (out_pq, out_pk, out_pk_rev):
        assert t is None or (t.is_contiguous() and tuple(t.shape) == (nb, h, rel_rows, 64) and t.dtype == torch.bfloat16)
    assert pqk.dtype == torch.bfloat16 and pqk.stride(2) == 1
    check(load().fb_build_rel_tables_batched(_ptr(pqk), C.c_longlong(pqk.stride(1)), H2 // 2, h, _ptr(bucket_of_rel),
                                             rel_rows, rel_off, L, _ptr(out_pq), _ptr(out_pk), _ptr(out_pk_rev), nb,
                                             C.c_longlong(pqk.stride(0)), stream_ptr()), "fb_build_rel_tables_batched")


def attn_fwd(Q, K, V, PKrel, PQrel, mask_u8, ctx, lse, *, B, L, h, scale, dropout_p=0.0, seed=0, offset=0,
             s_out=None, key_mask_only=False):
    """Q/K/V: bf16 2-D views [B*L, >= h*64] (row pitch = stride(0)); ctx: bf16 [B*L, H]; lse: fp32 [B, h, L]."""
    rel_rows, rel_off = rel_table_geometry(L)
    assert mask_u8.dtype == torch.uint8 and lse.dtype == torch.float32
    assert PKrel is None or PKrel.shape[1] == rel_rows
    a = AttnArgs()
    a.key_mask_only = int(key_mask_only)
    a.B, a.L, a.h = B, L, h
    a.Q, a.ldq = _ptr(Q), Q.stride(0)
    a.K, a.ldk = _ptr(K), K.stride(0)
    a.V, a.ldv = _ptr(V), V.stride(0)
    a.PKrel, a.PQrel, a.rel_rows, a.rel_off = _ptr(PKrel), _ptr(PQrel), rel_rows, rel_off
    a.mask = _ptr(mask_u8)
    a.ctx, a.ldc = _ptr(ctx), ctx.stride(0)
    a.lse = _ptr(lse)
    a.scale = float(scale)
    a.dropout_p, a.seed, a.offset = float(dropout_p), int(seed), int(offset)
    if s_out is not None:
        assert s_out.dtype == SCORE_DTYPE and s_out.shape[-1] == (L + 127) // 128 * 128
        a.s_out = _ptr(s_out)
    check(load().fb_attn_fwd(C.byref(a), stream_ptr()), "fb_attn_fwd")
    return ctx, lse


def layernorm_fwd(x, gamma, beta, eps, *, rowmask=None, out_drop=None, y32=None, y16=None, mean=None, rstd=None):
    M, H = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    wide16 = y16 if (y16 is not None and y16.dtype == torch.float32) else None   # fp32-class mode: no rounding
    if wide16 is not None:
        y16 = None
        y32 = wide16 if y32 is None else y32
    check(load().fb_layernorm_fwd(_ptr(x), M, H, _ptr(gamma), _ptr(beta), float(eps), _ptr(rowmask), _dp(out_drop),
                                  _ptr(y32), _ptr(y16), _ptr(mean), _ptr(rstd), stream_ptr()), "fb_layernorm_fwd")
    if wide16 is not None and wide16 is not y32:
        wide16.copy_(y32)


def layernorm_bwd(dy, x, mean, rstd, gamma, *, rowmask=None, out_drop=None, in_drop=None, dx32=None, dx16=None,
                  dgamma=None, dbeta=None, dx16_colsum=None):
    M, H = x.shape
    assert dy.dtype == torch.float32 and dy.is_contiguous() and x.is_contiguous()
    if dx16 is not None and dx16.dtype == torch.float32:
        # fp32-class mode: dx "16" = dropout_bwd(dx32) unrounded, its column sums from the fp32 kernel
        tmp = dx32 if dx32 is not None else torch.empty_like(dx16)
        check(load().fb_layernorm_bwd(_ptr(dy), _ptr(x), _ptr(mean), _ptr(rstd), _ptr(gamma), M, H, _ptr(rowmask),
                                      _dp(out_drop), None, _ptr(tmp), None, _ptr(dgamma), _ptr(dbeta), None,
                                      stream_ptr()), "fb_layernorm_bwd")
        dropout_apply(tmp, in_drop, out32=dx16)
        if dx16_colsum is not None:
            colsum_bf16(dx16, dx16_colsum)
        return
    check(load().fb_layernorm_bwd(_ptr(dy), _ptr(x), _ptr(mean), _ptr(rstd), _ptr(gamma), M, H, _ptr(rowmask),
                                  _dp(out_drop), _dp(in_drop), _ptr(dx32), _ptr(dx16), _ptr(dgamma), _ptr(dbeta),
                                  _ptr(dx16_colsum), stream_ptr()), "fb_layernorm_bwd")


def move_rows(src, dst, *, idx_in=None, idx_out=None, R=None, ncols=None):
    """dst[idx_out[r]] = src[idx_in[r]] (int64 index tensors; None = identity)."""
    if R is None:
        R = (idx_in if idx_in is not None else idx_out).numel() if (idx_in is not None or idx_out is not None) \
            else src.shape[0]
    if R == 0:
        return
    ncols = ncols or src.shape[1]
    for t in (idx_in, idx_out):
        assert t is None or (t.dtype == torch.int64 and t.is_contiguous())
    check(load().fb_move_rows(_ptr(src), src.stride(0), int(src.dtype == torch.bfloat16), _ptr(idx_in), _ptr(idx_out),
                              R, ncols, _ptr(dst), dst.stride(0), int(dst.dtype == torch.bfloat16), stream_ptr()),
          "fb_move_rows")


def gelu_bwd(dy, pre16, out16, *, rowmask=None, dropout=None):
    M, N = dy.shape
    if pre16.dtype == torch.float32:
        # same order as gelu_bwd_kernel: gelu'(pre), dropout backward, row mask
        epilogue_f32(dy, out16, aux=pre16, aux_op=FB_AUX_MUL_GELU_GRAD, dropout=dropout, drop_pos=2)
        if rowmask is not None:
            epilogue_f32(out16, out16, rowmask=rowmask)
        return
    check(load().fb_gelu_bwd(_ptr(dy), _ptr(pre16), M, N, _ptr(rowmask), _dp(dropout), _ptr(out16), stream_ptr()),
          "fb_gelu_bwd")


def add_rowbcast(a, b, B, L, H, out32, out16=None):
    if out16 is not None and out16.dtype == torch.float32:
        add_rowbcast(a, b, B, L, H, out32, None)
        out16.copy_(out32)
        return
    check(load().fb_add_rowbcast(_ptr(a), _ptr(b), B, L, H, _ptr(out32), _ptr(out16), stream_ptr()), "fb_add_rowbcast")


def conv_im2col(x16, B, L, H, out16):
    if x16.dtype == torch.float32:
        check(load().fb_conv_im2col_f32(_ptr(x16), B, L, H, _ptr(out16), stream_ptr()), "fb_conv_im2col_f32")
        return
    check(load().fb_conv_im2col(_ptr(x16), B, L, H, _ptr(out16), stream_ptr()), "fb_conv_im2col")


def conv_col2im(dcol16, B, L, H, dx32, accumulate=False):
    if dcol16.dtype == torch.float32:
        check(load().fb_conv_col2im_f32(_ptr(dcol16), B, L, H, _ptr(dx32), int(accumulate), stream_ptr()),
              "fb_conv_col2im_f32")
        return
    check(load().fb_conv_col2im(_ptr(dcol16), B, L, H, _ptr(dx32), int(accumulate), stream_ptr()), "fb_conv_col2im")


def colsum_bf16(x16, out32):
    M, N = x16.shape
    if x16.dtype == torch.float32:
        check(load().fb_colsum_f32(_ptr(x16), x16.stride(0), M, N, _ptr(out32), stream_ptr()), "fb_colsum_f32")
        return
    check(load().fb_colsum_bf16(_ptr(x16), x16.stride(0), M, N, _ptr(out32), stream_ptr()), "fb_colsum_bf16")


def cross_entropy(logits32, V, labels, loss_sum, loss_scale, dlogits16=None, grad_scale=0.0):
    P = logits32.shape[0]
    assert labels.dtype == torch.int64 and labels.is_contiguous()
    if dlogits16 is not None and dlogits16.dtype == torch.float32:
        check(load().fb_cross_entropy_f32(_ptr(logits32), logits32.stride(0), P, V, _ptr(labels), _ptr(loss_sum),
                                          float(loss_scale), _ptr(dlogits16), dlogits16.stride(0), float(grad_scale),
                                          stream_ptr()), "fb_cross_entropy_f32")
        return
    check(load().fb_cross_entropy(_ptr(logits32), logits32.stride(0), P, V, _ptr(labels), _ptr(loss_sum),
                                  float(loss_scale), _ptr(dlogits16),
                                  dlogits16.stride(0) if dlogits16 is not None else 0, float(grad_scale),
                                  stream_ptr()), "fb_cross_entropy")


def softmax_topk(logits32, V, k, want_probs=False):
    """fb_softmax_topk: logits fp32 [P, >= V] -> (probs [P, V] or None, top_p [P, k], top_i [P, k])."""
    assert logits32.dtype == torch.float32 and logits32.stride(-1) == 1
    P = logits32.shape[0]
    probs = torch.empty(P, V, device=logits32.device, dtype=torch.float32) if want_probs else None
    top_p = torch.empty(P, k, device=logits32.device, dtype=torch.float32)
    top_i = torch.empty(P, k, device=logits32.device, dtype=torch.int64)
    check(load().fb_softmax_topk(_ptr(logits32), logits32.stride(0), P, V, k, _ptr(probs), V, _ptr(top_p), _ptr(top_i),
                                 stream_ptr()), "fb_softmax_topk")
    return probs, top_p, top_i


def dropout_apply(x32, dropout, *, out32=None, accumulate=False, out16=None):
    M, N = x32.shape
    if out16 is not None and out16.dtype == torch.float32:
        assert out32 is None
        out32, out16 = out16, None
    check(load().fb_dropout_apply(_ptr(x32), M, N, _dp(dropout), _ptr(out32), int(accumulate), _ptr(out16),
                                  stream_ptr()), "fb_dropout_apply")


def dropout_batched(x32, dropout, offset_stride, *, out16=None, out32_accum=None):
    """nb layer passes of fb_dropout_apply in one launch (pass b: site offset + b * offset_stride).
    out16 [nb, M, N] bf16 = dropout_b(x32 [M, N])   |   out32_accum [M, N] fp32 += sum_b dropout_b(x32 [nb, M, N])."""
    if out16 is not None:
        nb, M, N = out16.shape
        assert out16.is_contiguous() and tuple(x32.shape) == (M, N) and out32_accum is None
    else:
        nb, M, N = x32.shape
        assert tuple(out32_accum.shape) == (M, N) and out32_accum.is_contiguous()
    assert x32.is_contiguous() and x32.dtype == torch.float32
    check(load().fb_dropout_batched(_ptr(x32), M, N, nb, _dp(dropout), int(offset_stride), _ptr(out32_accum),
                                    _ptr(out16), stream_ptr()), "fb_dropout_batched")


def sumsq(g, out):
    check(load().fb_sumsq(_ptr(g), g.numel(), _ptr(out), stream_ptr()), "fb_sumsq")


def sumsq_det(g, out, scratch):
    """Deterministic sum of squares (overwrites out[0]); scratch: fp32 [>= 2], zeroed once by the caller."""
    check(load().fb_sumsq_det(_ptr(g), g.numel(), _ptr(out), _ptr(scratch), scratch.numel(), stream_ptr()),
          "fb_sumsq_det")


def adam_step(p, g, m, v, *, lr, beta1, beta2, eps, step, gnorm_sq=None, max_norm=0.0, grad_scale=1.0, p16=None):
    check(load().fb_adam_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), float(lr), float(beta1), float(beta2),
                              float(eps), int(step), _ptr(gnorm_sq), float(max_norm), float(grad_scale), _ptr(p16),
                              stream_ptr()), "fb_adam_step")


def cast_bf16(x32, out16):
    if out16.dtype == torch.float32:
        out16.copy_(x32.reshape(out16.shape))
        return
    check(load().fb_cast_f32_bf16(_ptr(x32), x32.numel(), _ptr(out16), stream_ptr()), "fb_cast_f32_bf16")


def attn_bwd_prep(s2, lse, dp, ctx, dctx, ds, pd, z1=None, z2=None, *, B, L, h, scale, dropout_p=0.0, seed=0,
                  offset=0):
    rel_rows, rel_off = rel_table_geometry(L)
    assert s2.dtype == SCORE_DTYPE
    a = AttnBwdPrepArgs()
    a.B, a.L, a.h, a.lp, a.rel_rows, a.rel_off = B, L, h, s2.shape[-1], rel_rows, rel_off
    a.s2, a.lse, a.dp = _ptr(s2), _ptr(lse), _ptr(dp)
    a.ctx, a.dctx, a.ldc = _ptr(ctx), _ptr(dctx), ctx.stride(0)
    assert dctx.stride(0) == ctx.stride(0)
    a.ds, a.pd, a.z1, a.z2 = _ptr(ds), _ptr(pd), _ptr(z1), _ptr(z2)
    a.scale = float(scale)
    a.dropout_p, a.seed, a.offset = float(dropout_p), int(seed), int(offset)
    check(load().fb_attn_bwd_prep(C.byref(a), stream_ptr()), "fb_attn_bwd_prep")


def attn_bwd(Q, K, V, O, dO, s2, lse, PKrev, PQrel, dQ, dK, dV, ds_out=None, *, B, L, h, scale, dropout_p=0.0, seed=0,
             offset=0):
    """Fused attention backward (fb_attn_bwd).  Q/K/V/O/dO/dQ/dK/dV: bf16 2-D views [B*L, >= h*64]; s2: fp32
    [B,h,L,lp]; lse: fp32 [B,h,L]; PKrev: row-reversed PKrel (build_rel_tables(out_pk_rev=...)); ds_out: optional
    bf16 [B,h,L,lp], receives the UNSCALED dS'."""
    rel_rows, rel_off = rel_table_geometry(L)
    assert s2.dtype == SCORE_DTYPE and lse.dtype == torch.float32
    assert O.stride(0) == dO.stride(0)
    a = AttnBwdArgs()
    a.B, a.L, a.h, a.lp = B, L, h, s2.shape[-1]
    a.Q, a.ldq = _ptr(Q), Q.stride(0)
    a.K, a.ldk = _ptr(K), K.stride(0)
    a.V, a.ldv = _ptr(V), V.stride(0)
    a.O, a.dO, a.ldo = _ptr(O), _ptr(dO), O.stride(0)
    a.s2, a.lse = _ptr(s2), _ptr(lse)
    a.PKrev, a.PQrel, a.rel_rows, a.rel_off = _ptr(PKrev), _ptr(PQrel), rel_rows, rel_off
    a.dQ, a.lddq = _ptr(dQ), dQ.stride(0)
    a.dK, a.lddk = _ptr(dK), dK.stride(0)
    a.dV, a.lddv = _ptr(dV), dV.stride(0)
    a.ds_out = _ptr(ds_out)
    a.scale = float(scale)
    a.dropout_p, a.seed, a.offset = float(dropout_p), int(seed), int(offset)
    check(load().fb_attn_bwd(C.byref(a), stream_ptr()), "fb_attn_bwd")


def attn_bwd_rel(ds, Q, K, gpk, gpq, *, B, L, h, scale, units_per_cta=0):
    """gpk / gpq (fp32 [h, rel_rows, 64], overwritten) = gradients of the per-relative-position tables."""
    rel_rows, rel_off = rel_table_geometry(L)
    assert gpk.dtype == torch.float32 and gpq.dtype == torch.float32 and gpk.shape[1] == rel_rows
    a = AttnBwdRelArgs()
    a.B, a.L, a.h, a.lp, a.rel_rows, a.rel_off = B, L, h, ds.shape[-1], rel_rows, rel_off
    a.ds = _ptr(ds)
    a.Q, a.ldq = _ptr(Q), Q.stride(0)
    a.K, a.ldk = _ptr(K), K.stride(0)
    a.gpk, a.gpq = _ptr(gpk), _ptr(gpq)
    a.units_per_cta = int(units_per_cta)
    a.scale = float(scale)
    check(load().fb_attn_bwd_rel(C.byref(a), stream_ptr()), "fb_attn_bwd_rel")


def umma_probe(a_image, b_image, a_off, b_off, a_desc, b_desc, idesc, N, cycles=None):
    """Hardware test bench (fb_umma_probe): a_image / b_image are uint8 CUDA tensors (raw shared-memory images);
    returns the fp32 [128, N] accumulator."""
    out = torch.zeros(128, N, device=a_image.device, dtype=torch.float32)
    n = len(a_off)
    a = UmmaProbeArgs()
    a.a_image, a.a_bytes = _ptr(a_image), a_image.numel()
    a.b_image, a.b_bytes = _ptr(b_image), b_image.numel()
    a.n_mma = n
    ao, bo = (C.c_int * n)(*a_off), (C.c_int * n)(*b_off)
    a.a_off, a.b_off = ao, bo
    a.a_desc, a.b_desc, a.idesc, a.N = int(a_desc), int(b_desc), int(idesc), int(N)
    a.out = _ptr(out)
    a.cycles = _ptr(cycles)
    check(use_debug_library().fb_umma_probe(C.byref(a), stream_ptr()), "fb_umma_probe")
    return out


def rel_to_bucket(gpq, gpk, bucket_of_rel, L, h, out):
    """gpq/gpk: fp32 [h, rel_rows, 64]; out: fp32 [2S, 2H], ACCUMULATES (PQ | PK column blocks)."""
    rel_rows, rel_off = rel_table_geometry(L)
    H = h * 64
    check(load().fb_rel_to_bucket(_ptr(gpq), _ptr(gpk), H, h, _ptr(bucket_of_rel), rel_rows, rel_off, L, _ptr(out),
                                  out.stride(0), stream_ptr()), "fb_rel_to_bucket")


def rel_to_bucket_batched(gpq, gpk, bucket_of_rel, L, h, out):
    """gpq/gpk: fp32 [nb, h, rel_rows, 64]; out: fp32 [nb, 2S, 2H], ACCUMULATES."""
    rel_rows, rel_off = rel_table_geometry(L)
    nb = gpq.shape[0]
    assert gpq.is_contiguous() and gpk.is_contiguous() and out.shape[0] == nb
    check(load().fb_rel_to_bucket_batched(_ptr(gpq), _ptr(gpk), h * 64, h, _ptr(bucket_of_rel), rel_rows, rel_off, L,
                                          _ptr(out), out.stride(1), nb, C.c_longlong(out.stride(0)), stream_ptr()),
          "fb_rel_to_bucket_batched")


def mlm_corrupt(ids: torch.Tensor, special_mask, pad_id, mask_id: int, vocab: int, mlm_probability: float, seed: int,
                ids_out=None, labels_out=None):
    """fb_mlm_corrupt: ids int64 (any shape, contiguous, on the GPU); special_mask uint8 or None; returns
    (corrupted ids, labels)."""
    assert ids.dtype == torch.int64 and ids.is_contiguous() and ids.is_cuda
    if special_mask is not None:
        assert special_mask.dtype == torch.uint8 and special_mask.numel() == ids.numel() and special_mask.is_contiguous()
    ids_out = torch.empty_like(ids) if ids_out is None else ids_out
    labels_out = torch.empty_like(ids) if labels_out is None else labels_out
    check(load().fb_mlm_corrupt(_ptr(ids), _ptr(special_mask), ids.numel(), -1 if pad_id is None else int(pad_id),
                                int(mask_id), int(vocab), float(mlm_probability), C.c_ulonglong(int(seed)),
                                _ptr(ids_out), _ptr(labels_out), stream_ptr()), "fb_mlm_corrupt")
    return ids_out, labels_out


def gather_frames(shard: torch.Tensor, offsets: torch.Tensor, nframes: torch.Tensor, max_feats: int, out_f32=None,
                  out_bf16=None, video_len=None):
    """fb_gather_frames: shard fp16 [rows, D] (GPU), offsets int64 [B], nframes int32 [B]."""
    assert shard.dtype == torch.float16 and shard.is_contiguous() and shard.is_cuda
    assert offsets.dtype == torch.int64 and nframes.dtype == torch.int32
    B, D = nframes.numel(), shard.shape[1]
    if out_f32 is None and out_bf16 is None:
        out_f32 = torch.empty(B, max_feats, D, device=shard.device, dtype=torch.float32)
    if video_len is None:
        video_len = torch.empty(B, device=shard.device, dtype=torch.int64)
    check(load().fb_gather_frames(_ptr(shard), _ptr(offsets), _ptr(nframes), B, int(max_feats), D, _ptr(out_f32),
                                  _ptr(out_bf16), _ptr(video_len), stream_ptr()), "fb_gather_frames")
    return (out_f32 if out_f32 is not None else out_bf16), video_len
