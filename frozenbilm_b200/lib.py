"""ctypes binding of libfrozenbilm_sm100.so (the C-ABI in include/frozenbilm_b200.h).

There is no fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfrozenbilm_sm100.so")

FB_ACT_NONE, FB_ACT_GELU, FB_ACT_RELU = 0, 1, 2
FB_AUX_NONE, FB_AUX_MUL_GELU_GRAD, FB_AUX_MUL_RELU_MASK, FB_AUX_ADD = 0, 1, 2, 3
FB_DT_BF16, FB_DT_F32 = 0, 1


class FbError(RuntimeError):
    pass


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
    ]


_lib = None


def load() -> C.CDLL:
    """Loads the shared library; raises FbError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FbError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no CPU / eager fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.fb_last_error.restype = C.c_char_p
    lib.fb_abi_version.restype = C.c_int
    lib.fb_launch_count.restype = C.c_longlong
    for name in EXPORTS:
        if name in ("fb_last_error", "fb_abi_version", "fb_launch_count"):
            continue
        getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


# every symbol include/frozenbilm_b200.h declares (tests check the .so exports all of them)
EXPORTS = [
    "fb_last_error", "fb_abi_version", "fb_launch_count", "fb_gemm", "fb_attn_fwd", "fb_build_rel_tables",
]


def check(rc: int, what: str):
    if rc != 0:
        raise FbError(f"{what} failed ({rc}): {load().fb_last_error().decode()}")


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count() -> int:
    return int(load().fb_launch_count())


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def gemm(A: torch.Tensor, B: torch.Tensor, D: torch.Tensor, *, a_mn_major=False, b_mn_major=False,
         bias=None, act=FB_ACT_NONE, D_pre=None, aux=None, aux_op=FB_AUX_NONE, resid=None, alpha=1.0,
         accumulate=False, split_k=1, block_n=0, M=None, N=None, K=None):
    """D = epilogue(alpha * A @ B^T) on the tcgen05 GEMM.

    A: bf16 [M, K] (or [K, M] when a_mn_major);  B: bf16 [N, K] (or [K, N] when b_mn_major);
    2-D (or 3-D batched) views with unit inner stride.  D: bf16 or fp32 [M, N].
    """
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    assert A.stride(-1) == 1 and B.stride(-1) == 1 and D.stride(-1) == 1
    batch = 1
    sa = sb = sd = 0
    if A.dim() == 3:
        batch = A.shape[0]
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
    args.M, args.N, args.K, args.batch = M, N, K, batch
    args.A, args.lda, args.stride_a, args.a_mn_major = _ptr(A), A2.stride(0), sa, int(a_mn_major)
    args.B, args.ldb, args.stride_b, args.b_mn_major = _ptr(B), B2.stride(0), sb, int(b_mn_major)
    args.D, args.ldd, args.stride_d = _ptr(D), D2.stride(0), sd
    args.d_dtype = FB_DT_F32 if D.dtype == torch.float32 else FB_DT_BF16
    assert D.dtype in (torch.float32, torch.bfloat16)
    args.alpha = float(alpha)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
    args.bias = _ptr(bias)
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


def rel_table_geometry(L: int):
    """(rel_rows, rel_off) of the per-relative-position tables for sequence length L."""
    lpad = (L + 127) // 128 * 128
    return 2 * lpad + 8, lpad


def build_rel_tables(pqk: torch.Tensor, bucket_of_rel: torch.Tensor, L: int, h: int, out_pq=None, out_pk=None):
    """pqk: bf16 [2S, 2H] (PQ | PK).  Returns (PQrel, PKrel) bf16 [h, rel_rows, 64]."""
    H = pqk.shape[1] // 2
    rel_rows, rel_off = rel_table_geometry(L)
    if out_pq is None:
        out_pq = torch.empty(h, rel_rows, 64, device=pqk.device, dtype=torch.bfloat16)
    if out_pk is None:
        out_pk = torch.empty(h, rel_rows, 64, device=pqk.device, dtype=torch.bfloat16)
    assert bucket_of_rel.dtype == torch.int32 and bucket_of_rel.numel() == 1023
    check(load().fb_build_rel_tables(_ptr(pqk), C.c_longlong(pqk.stride(0)), H, h, _ptr(bucket_of_rel), rel_rows,
                                     rel_off, L, _ptr(out_pq), _ptr(out_pk), stream_ptr()), "fb_build_rel_tables")
    return out_pq, out_pk


def attn_fwd(Q, K, V, PKrel, PQrel, mask_u8, ctx, lse, *, B, L, h, scale, dropout_p=0.0, seed=0, offset=0):
    """Q/K/V: bf16 2-D views [B*L, >= h*64] (row pitch = stride(0)); ctx: bf16 [B*L, H]; lse: fp32 [B, h, L]."""
    rel_rows, rel_off = rel_table_geometry(L)
    assert PKrel.shape[1] == rel_rows and mask_u8.dtype == torch.uint8 and lse.dtype == torch.float32
    a = AttnArgs()
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
    check(load().fb_attn_fwd(C.byref(a), stream_ptr()), "fb_attn_fwd")
    return ctx, lse
