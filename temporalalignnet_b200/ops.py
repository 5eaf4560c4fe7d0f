"""Tensor-level wrappers over the C ABI (include/tan_b200.h).

torch is used here only for device memory and the current CUDA stream; every arithmetic step is a
kernel of libtan_b200.so.  All functions launch asynchronously on `torch.cuda.current_stream()`.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import LnArgs, SimGeom, check, lib

_launches = 0          # kernels launched through this module (bench.py's `gpu_launches` claim)
_prof = None           # when a list: (kernel, work, start_event, end_event) per call (bench.py roofline pass)
_only = None           # measurement aid: when set, only this kernel class is launched (see `only_class`)
_work = {}             # per-class (algorithmic work, launches) accumulated while `_only` is set


class only_class:
    """Context manager for bench.py's per-kernel timing: inside it every C-ABI wrapper except `name`
    returns without launching, so a CUDA graph captured around one step holds exactly that kernel
    class's launches (in step order, on the step's real buffers).  Kernels do not branch on data
    values, so skipping the producers changes no launch.  Not used on the product path."""

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        global _only
        _only = self.name
        _work[self.name] = [0.0, 0]
        return _work[self.name]

    def __exit__(self, *a):
        global _only
        _only = None


def _skip(name: str, work: float = 0.0, n: int = 1) -> bool:
    if _only is None:
        return False
    if _only == name:
        _work[name][0] += work
        _work[name][1] += n
        return False
    return True


class profile:
    """Context manager: bracket every C-ABI call with CUDA events on the launching stream."""

    def __enter__(self):
        global _prof
        _prof = []
        return _prof

    def __exit__(self, *a):
        global _prof
        _prof = None


class _timed:
    __slots__ = ("name", "work", "e0")

    def __init__(self, name, work):
        self.name, self.work, self.e0 = name, work, None

    def __enter__(self):
        if _prof is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if _prof is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _prof.append((self.name, self.work, self.e0, e1))


def launches() -> int:
    return _launches


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need(t: torch.Tensor, dtype, name: str):
    if t.dtype != dtype or not t.is_cuda or not t.is_contiguous():
        raise _lib.TanError(f"{name}: expected contiguous CUDA {dtype}, got {t.dtype} {t.device} "
                            f"contiguous={t.is_contiguous()}")


def cast_bf16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 -> bf16 (tan_cast_f32_to_bf16)."""
    global _launches
    _need(x, torch.float32, "cast_bf16.in")
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    if _skip("cast", float(x.numel())):
        return out
    check(lib().tan_cast_f32_to_bf16(x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "tan_cast_f32_to_bf16")
    _launches += 1
    return out


def linear(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
           residual: Optional[torch.Tensor] = None, out_f32: Optional[torch.Tensor] = None,
           out_bf16: Optional[torch.Tensor] = None, act: int = _lib.ACT_NONE, tag: str = "linear") -> None:
    """out = act(a @ w.T + bias) [+ residual]  (tan_linear_bf16).  a [M,K] bf16 (row pitch may exceed
    K), w [N,K] bf16; outputs 2-D with arbitrary row pitch.  `tag` names the kernel class in profiles
    (forward projections: "linear"; backward: "dgrad", "wgrad", "sim_bwd")."""
    global _launches
    M, K = a.shape
    N = w.shape[0]
    if _skip(tag, 2.0 * M * N * K):
        return
    with _timed(tag, 2.0 * M * N * K):
        check(lib().tan_linear_bf16(
            a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), _ptr(bias),
            _ptr(residual), residual.stride(0) if residual is not None else 0,
            _ptr(out_f32), out_f32.stride(0) if out_f32 is not None else 0,
            _ptr(out_bf16), out_bf16.stride(0) if out_bf16 is not None else 0,
            M, N, K, act, _stream()), "tan_linear_bf16")
    _launches += 1


def linear_res_ln(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], x: torch.Tensor,
                  gamma: torch.Tensor, beta: torch.Tensor, out_bf16: torch.Tensor) -> None:
    """x += a @ w.T + bias (fp32, in place); out_bf16 = LayerNorm(x) * gamma + beta  (tan_linear_res_ln_bf16;
    output width 512).  Counted with the linear class (flops of the GEMM)."""
    global _launches
    M, K = a.shape
    N = w.shape[0]
    if _skip("linear", 2.0 * M * N * K):
        return
    with _timed("linear", 2.0 * M * N * K):
        check(lib().tan_linear_res_ln_bf16(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), _ptr(bias),
                                           x.data_ptr(), x.stride(0), gamma.data_ptr(), beta.data_ptr(),
                                           out_bf16.data_ptr(), out_bf16.stride(0), M, N, K, _stream()),
              "tan_linear_res_ln_bf16")
    _launches += 1


def linear_res_ln_stage(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], x: torch.Tensor,
                        gamma: torch.Tensor, beta: torch.Tensor, out_bf16: Optional[torch.Tensor], L: int,
                        l_split: int, nrmA_bf16: Optional[torch.Tensor], strideA: int,
                        nrmB_bf16: Optional[torch.Tensor], strideB: int) -> None:
    """tan_linear_res_ln_stage_bf16: linear_res_ln + L2-normalised bf16 stage features scattered by clip
    (nrm pointers are views whose data_ptr() is the stage's first row)."""
    global _launches
    M, K = a.shape
    N = w.shape[0]
    if _skip("linear", 2.0 * M * N * K):
        return
    with _timed("linear", 2.0 * M * N * K):
        check(lib().tan_linear_res_ln_stage_bf16(
            a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), _ptr(bias), x.data_ptr(), x.stride(0),
            gamma.data_ptr(), beta.data_ptr(), _ptr(out_bf16), out_bf16.stride(0) if out_bf16 is not None else 0,
            M, N, K, L, l_split, _ptr(nrmA_bf16), strideA, _ptr(nrmB_bf16), strideB, _stream()),
            "tan_linear_res_ln_stage_bf16")
    _launches += 1


def layernorm(x: torch.Tensor, rows: int, d: int, gamma=None, beta=None, add=None, add_rows: int = 0,
              L_in: Optional[int] = None, L_out: Optional[int] = None, l_off: int = 0,
              out_f32=None, out_bf16=None, l_split: int = 0, strideA: int = 0, strideB: int = 0,
              rawA=None, rawB=None, nrmA_bf16=None, nrmB_bf16=None, nrmA_f32=None, nrmB_f32=None) -> None:
    """tan_layernorm; pointer-valued keyword arguments are tensors whose data_ptr() already points at
    the first destination row (callers pass views)."""
    global _launches
    L_in = rows if L_in is None else L_in
    L_out = L_in if L_out is None else L_out
    if _skip("layernorm", float(rows) * d):
        return
    a = LnArgs(x.data_ptr(), int(x.dtype == torch.bfloat16), rows, d, _ptr(gamma), _ptr(beta), _ptr(add), add_rows,
               L_in, L_out, l_off, _ptr(out_f32), _ptr(out_bf16), l_split, strideA, strideB,
               _ptr(rawA), _ptr(rawB), _ptr(nrmA_bf16), _ptr(nrmB_bf16), _ptr(nrmA_f32), _ptr(nrmB_f32))
    with _timed("layernorm", float(rows) * d):
        check(lib().tan_layernorm(C.byref(a), _stream()), "tan_layernorm")
    _launches += 1


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, kpm_u8: Optional[torch.Tensor],
              out: torch.Tensor, B: int, H: int, Lq: int, Lk: int, lse: Optional[torch.Tensor] = None) -> None:
    """tan_attention_bf16.  q/k/v/out are 2-D (possibly column-sliced) bf16 views [B*L, H*64]; lse (optional)
    [B, H, pad64(Lq)] fp32 receives the rows' log2-domain log-sum-exp for the backward pass."""
    global _launches
    if _skip("attention", 4.0 * B * H * Lq * Lk * 64):
        return
    with _timed("attention", 4.0 * B * H * Lq * Lk * 64):
        check(lib().tan_attention_bf16(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(),
                                       v.stride(0), _ptr(kpm_u8), out.data_ptr(), out.stride(0), B, H, Lq, Lk,
                                       _ptr(lse), _stream()), "tan_attention_bf16")
    _launches += 1


def sim_geom(B_loc, S, T, C_, N, d, b_off=0) -> SimGeom:
    return SimGeom(B_loc, S, T, C_, N, d, b_off)


def sim_workspace_bytes(g: SimGeom) -> int:
    return int(lib().tan_sim_nce_workspace_bytes(C.byref(g)))


def pos_from_time(start: torch.Tensor, end: torch.Tensor, valid_u8: Optional[torch.Tensor], B: int, T: int, N: int,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """tan_pos_from_time: packed target bits [B, T, ceil(N/32)] (int32 storage of the uint32 words)."""
    global _launches
    W = (N + 31) // 32
    if out is None:
        out = torch.empty(B, T, W, dtype=torch.int32, device=start.device)
    if _skip("glue", float(B * T * W)):
        return out
    _need(start, torch.float32, "pos_from_time.start")
    _need(end, torch.float32, "pos_from_time.end")
    check(lib().tan_pos_from_time(start.data_ptr(), end.data_ptr(), _ptr(valid_u8), B, T, N, out.data_ptr(), _stream()),
          "tan_pos_from_time")
    _launches += 1
    return out


def sim_nce_fwd(vfeat, tfeat, tfeat_stage_stride: int, g: SimGeom, posbits, col_valid, logits_out,
                row_sums, col_sums, workspace, row_kill=None) -> None:
    global _launches
    if _skip("sim_nce_fwd", 2.0 * g.B_loc * g.S * g.T * g.C * g.d, 2):
        return
    with _timed("sim_nce_fwd", 2.0 * g.B_loc * g.S * g.T * g.C * g.d):
        check(lib().tan_sim_nce_fwd(vfeat.data_ptr(), tfeat.data_ptr(), tfeat_stage_stride, C.byref(g),
                                    posbits.data_ptr(), col_valid.data_ptr(), _ptr(row_kill), _ptr(logits_out),
                                    row_sums.data_ptr(), col_sums.data_ptr(), workspace.data_ptr(),
                                    workspace.numel() * workspace.element_size(), _stream()), "tan_sim_nce_fwd")
    _launches += 2


def nce_from_logits(logits, g: SimGeom, posbits, col_valid, row_sums, col_sums, workspace, row_kill=None) -> None:
    global _launches
    is_f32 = int(logits.dtype == torch.float32)
    if not is_f32 and logits.dtype != torch.bfloat16:
        raise _lib.TanError(f"nce_from_logits: logits must be fp32 or bf16, got {logits.dtype}")
    if _skip("nce_from_logits", float(logits.numel()) * logits.element_size(), 2):
        return
    with _timed("nce_from_logits", float(logits.numel()) * logits.element_size()):
        check(lib().tan_nce_from_logits(logits.data_ptr(), is_f32, C.byref(g), posbits.data_ptr(),
                                        col_valid.data_ptr(), _ptr(row_kill), row_sums.data_ptr(),
                                        col_sums.data_ptr(), workspace.data_ptr(),
                                        workspace.numel() * workspace.element_size(), _stream()),
              "tan_nce_from_logits")
    _launches += 2


def nce_reduce(row_sums, col_sums, out4_f64, S: int, T: int, C_: int, row_sel=None, col_sel=None) -> None:
    """tan_nce_reduce.  row_sums [2, B_loc*S*T] or None, col_sums [2, S, C] or None; row_sel [B_loc, T] /
    col_sel [C] uint8 optional selections (thresholded loss)."""
    global _launches
    R = row_sums.numel() // 2 if row_sums is not None else 0
    SC = col_sums.numel() // 2 if col_sums is not None else 0
    if _skip("nce_reduce", float(R + SC)):
        return
    check(lib().tan_nce_reduce(_ptr(row_sums), R, S, T, _ptr(row_sel), _ptr(col_sums), SC, C_, _ptr(col_sel),
                               out4_f64.data_ptr(), _stream()), "tan_nce_reduce")
    _launches += 1


def own_clip_sim(vfeat, tfeat, shared_text: bool, B: int, S: int, T: int, N: int, d: int, s_first: int = 0,
                 s_count: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """tan_own_clip_sim -> fp32 [B, s_count, T, N] own-clip cosine blocks."""
    global _launches
    s_count = S - s_first if s_count is None else s_count
    if out is None:
        out = torch.empty(B, s_count, T, N, dtype=torch.float32, device=vfeat.device)
    if _skip("own_clip_sim", 2.0 * B * s_count * T * N * d):
        return out
    check(lib().tan_own_clip_sim(vfeat.data_ptr(), tfeat.data_ptr(), 0 if shared_text else B * N * d, B, S, T, N, d,
                                 s_first, s_count, out.data_ptr(), _stream()), "tan_own_clip_sim")
    _launches += 1
    return out


def agree_scan(own, posbits, vpm_u8, tpm_u8, B: int, T: int, N: int, fill_max: bool):
    """tan_agree_scan -> (win [B,N,2] int32, mean_logit [B,N], max_logit [B,N])."""
    global _launches
    dev = own.device
    win = torch.empty(B, N, 2, dtype=torch.int32, device=dev)
    mean_logit = torch.empty(B, N, dtype=torch.float32, device=dev)
    max_logit = torch.empty(B, N, dtype=torch.float32, device=dev)
    if _skip("agree", float(B * T * N), 2):
        return win, mean_logit, max_logit
    _need(own, torch.float32, "agree_scan.own")
    ws = torch.empty(int(lib().tan_agree_scan_workspace_bytes(B, T, N)), dtype=torch.uint8, device=dev)
    check(lib().tan_agree_scan(own.data_ptr(), posbits.data_ptr(), _ptr(vpm_u8), tpm_u8.data_ptr(), B, T, N,
                               int(bool(fill_max)), win.data_ptr(), mean_logit.data_ptr(), max_logit.data_ptr(),
                               ws.data_ptr(), ws.numel(), _stream()), "tan_agree_scan")
    _launches += 2
    return win, mean_logit, max_logit


AGREE_KINDS = {"i": 0, "u": 1, "keep": 2, "keep-joint": 3}


def agree_targets(old_posbits, win_joint, win_dual, replace_u8, B: int, T: int, N: int, kind: str) -> torch.Tensor:
    """tan_agree_targets -> new packed target bits [B, T, W]."""
    global _launches
    out = torch.empty_like(old_posbits)
    if _skip("agree", float(B * T)):
        return out
    check(lib().tan_agree_targets(old_posbits.data_ptr(), win_joint.data_ptr(), win_dual.data_ptr(),
                                  replace_u8.data_ptr(), B, T, N, AGREE_KINDS[kind], out.data_ptr(), _stream()),
          "tan_agree_targets")
    _launches += 1
    return out


# ------------------------------------------------------------------------------------------------------
# backward pass (training step)
# ------------------------------------------------------------------------------------------------------
def pad64(n: int) -> int:
    return (n + 63) // 64 * 64


def transpose_bf16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """tan_transpose_bf16: x [R, C] bf16 (row pitch may exceed C) -> [C, pad64(R)] with zero padding."""
    global _launches
    R, Ccols = x.shape
    Rp = pad64(R)
    if out is None:
        out = torch.empty(Ccols, Rp, dtype=torch.bfloat16, device=x.device)
    if _skip("bwd_glue", 2.0 * R * Ccols):
        return out
    with _timed("transpose", 0.0):
        check(lib().tan_transpose_bf16(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), R, Ccols, Rp, _stream()),
              "tan_transpose_bf16")
    _launches += 1
    return out


def transpose_colsum_bf16(x: torch.Tensor, colsum_out: torch.Tensor, accumulate: bool = True) -> torch.Tensor:
    """tan_transpose_colsum_bf16 (experimental): transpose_bf16(x) and colsum_out (+)= column sums of x in one pass."""
    global _launches
    R, Ccols = x.shape
    Rp = pad64(R)
    out = torch.empty(Ccols, Rp, dtype=torch.bfloat16, device=x.device)
    if _skip("bwd_glue", 2.0 * R * Ccols):
        return out
    nbytes = int(lib().tan_transpose_colsum_workspace_bytes(R, Ccols))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    with _timed("transpose", 0.0):
        check(lib().tan_transpose_colsum_bf16(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), R, Ccols, Rp,
                                              colsum_out.data_ptr(), int(bool(accumulate)), ws.data_ptr(), nbytes,
                                              _stream()), "tan_transpose_colsum_bf16")
    _launches += 2
    return out


_tn_ws = {}


def gemm_tn(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, accumulate: bool = True, tag: str = "wgrad") -> None:
    """tan_gemm_tn_bf16: out [P, Q] fp32 (+)= a[R, P]^T @ b[R, Q] (a, b bf16 row-major, row pitch may exceed the
    width): weight gradients and the text-side similarity gradient without transposes."""
    global _launches
    R, P = a.shape
    Q = b.shape[1]
    if _skip(tag, 2.0 * R * P * Q):
        return
    nbytes = int(lib().tan_gemm_tn_workspace_bytes(R, P, Q))
    ws = None
    if nbytes:
        key = (str(a.device), torch.cuda.current_stream().cuda_stream)
        ws = _tn_ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = _tn_ws[key] = torch.empty(nbytes, dtype=torch.uint8, device=a.device)
    with _timed(tag, 2.0 * R * P * Q):
        check(lib().tan_gemm_tn_bf16(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), R, P, Q, out.data_ptr(),
                                     out.stride(0), int(bool(accumulate)), _ptr(ws), nbytes, _stream()),
              "tan_gemm_tn_bf16")
    _launches += 2 if nbytes else 1


def colsum(x: torch.Tensor, out: torch.Tensor, accumulate: bool = True) -> None:
    """tan_colsum: out [N] fp32 (+)= column sums of x [M, N] (bf16 or fp32)."""
    global _launches
    M, N = x.shape
    if _skip("bwd_glue", float(M) * N):
        return
    nbytes = int(lib().tan_colsum_workspace_bytes(M, N))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    with _timed("colsum", 0.0):
        check(lib().tan_colsum(x.data_ptr(), int(x.dtype == torch.bfloat16), x.stride(0), M, N, out.data_ptr(),
                               int(bool(accumulate)), ws.data_ptr(), nbytes, _stream()), "tan_colsum")
    _launches += 2


def quickgelu_fwd(u: torch.Tensor, h: torch.Tensor) -> None:
    global _launches
    if _skip("bwd_glue", float(u.numel())):
        return
    with _timed("gelu", 0.0):
        check(lib().tan_quickgelu_fwd(u.data_ptr(), h.data_ptr(), u.numel(), _stream()), "tan_quickgelu_fwd")
    _launches += 1


def quickgelu_bwd(dh: torch.Tensor, u: torch.Tensor, du: torch.Tensor) -> None:
    global _launches
    if _skip("bwd_glue", float(u.numel())):
        return
    with _timed("gelu", 0.0):
        check(lib().tan_quickgelu_bwd(dh.data_ptr(), u.data_ptr(), du.data_ptr(), u.numel(), _stream()),
              "tan_quickgelu_bwd")
    _launches += 1


def layernorm_bwd(dy: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, dx: torch.Tensor, accumulate_dx: bool,
                  rows: int, d: int, dgamma: Optional[torch.Tensor], dbeta: Optional[torch.Tensor],
                  L_in: Optional[int] = None, L_out: Optional[int] = None, l_off: int = 0) -> None:
    """tan_layernorm_bwd (dy, x, dx fp32; dgamma / dbeta accumulated)."""
    global _launches
    L_in = rows if L_in is None else L_in
    L_out = L_in if L_out is None else L_out
    if _skip("bwd_glue", float(rows) * d):
        return
    nbytes = int(lib().tan_layernorm_bwd_workspace_bytes(rows, d))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    with _timed("ln_bwd", 0.0):
        check(lib().tan_layernorm_bwd(dy.data_ptr(), x.data_ptr(), _ptr(gamma), dx.data_ptr(), int(bool(accumulate_dx)),
                                      rows, d, L_in, L_out, l_off, _ptr(dgamma), _ptr(dbeta), ws.data_ptr(), nbytes,
                                      _stream()), "tan_layernorm_bwd")
    _launches += 2


def l2norm_bwd(x: torch.Tensor, g: torch.Tensor, dst: torch.Tensor, accumulate: bool, rows: int, d: int, L_in: int,
               src_stride: int, L_out: int, l_off: int, g_stride: Optional[int] = None) -> None:
    """tan_l2norm_bwd; x / g are views whose data_ptr() is the stage's first row."""
    global _launches
    if _skip("bwd_glue", float(rows) * d):
        return
    with _timed("l2norm_bwd", 0.0):
        check(lib().tan_l2norm_bwd(x.data_ptr(), g.data_ptr(), dst.data_ptr(), int(bool(accumulate)), rows, d, L_in,
                                   src_stride, src_stride if g_stride is None else g_stride, L_out, l_off, _stream()),
              "tan_l2norm_bwd")
    _launches += 1


def batch_sum(x: torch.Tensor, out: torch.Tensor, B: int, L: int, d: int, L_out: int, l_off: int,
              accumulate: bool) -> None:
    global _launches
    if _skip("bwd_glue", float(B) * L * d):
        return
    with _timed("batch_sum", 0.0):
        check(lib().tan_batch_sum(x.data_ptr(), out.data_ptr(), B, L, d, L_out, l_off, int(bool(accumulate)), _stream()),
              "tan_batch_sum")
    _launches += 1


def sim_grad_tiles(z: torch.Tensor, Rc: int, r0: int, g: SimGeom, posbits, col_valid, row_kill, ra, rap, cb, cbp,
                   G: torch.Tensor, GT: torch.Tensor) -> None:
    """tan_sim_grad_tiles: z [Rc, >=C] fp32 -> G [Rc, Cp] bf16, GT [C, pad64(Rc)] bf16."""
    global _launches
    if _skip("bwd_glue", float(Rc) * g.C):
        return
    with _timed("sim_grad", 0.0):
        check(lib().tan_sim_grad_tiles(z.data_ptr(), z.stride(0), Rc, pad64(Rc), r0, C.byref(g), posbits.data_ptr(),
                                       col_valid.data_ptr(), _ptr(row_kill), ra.data_ptr(), rap.data_ptr(), cb.data_ptr(),
                                       cbp.data_ptr(), G.data_ptr(), G.stride(0), GT.data_ptr(), GT.stride(0), _stream()),
              "tan_sim_grad_tiles")
    _launches += 1


def sim_grad_gemm(a: torch.Tensor, t_pad: torch.Tensor, r0: int, g: SimGeom, posbits, col_valid, row_kill, ra, rap, cb,
                  cbp, G: torch.Tensor, GT: Optional[torch.Tensor] = None) -> None:
    """tan_sim_grad_gemm: a [Rc, d] bf16 video rows, t_pad [Cp, d] bf16 text rows (zero beyond C) -> G [Rc, Cp] bf16.
    GT [Cp, pad64(Rc)] (experimental tan_sim_grad_gemm_gt): also the transposed copy, written by the epilogue."""
    global _launches
    Rc, d = a.shape
    Cp = t_pad.shape[0]
    if _skip("sim_bwd", 2.0 * Rc * Cp * d):
        return
    if GT is not None:
        with _timed("sim_bwd", 2.0 * Rc * Cp * d):
            check(lib().tan_sim_grad_gemm_gt(a.data_ptr(), a.stride(0), t_pad.data_ptr(), t_pad.stride(0), Rc, r0,
                                             C.byref(g), Cp, posbits.data_ptr(), col_valid.data_ptr(), _ptr(row_kill),
                                             ra.data_ptr(), rap.data_ptr(), cb.data_ptr(), cbp.data_ptr(), G.data_ptr(),
                                             G.stride(0), GT.data_ptr(), GT.stride(0), pad64(Rc), _stream()),
                  "tan_sim_grad_gemm_gt")
        _launches += 1
        return
    with _timed("sim_bwd", 2.0 * Rc * Cp * d):
        check(lib().tan_sim_grad_gemm(a.data_ptr(), a.stride(0), t_pad.data_ptr(), t_pad.stride(0), Rc, r0, C.byref(g), Cp,
                                      posbits.data_ptr(), col_valid.data_ptr(), _ptr(row_kill), ra.data_ptr(),
                                      rap.data_ptr(), cb.data_ptr(), cbp.data_ptr(), G.data_ptr(), G.stride(0), _stream()),
              "tan_sim_grad_gemm")
    _launches += 1


def attention_bwd(q, k, v, o, d_out, kpm_u8, dq, dk, dv, lse, delta, B: int, H: int, Lq: int, Lk: int) -> None:
    """tan_attention_bwd_bf16 (2-D, possibly column-sliced bf16 views as in `attention`); lse [B, H, pad64(Lq)] is
    what the forward `attention(..., lse=)` stored, delta a workspace of the same shape."""
    global _launches
    if _skip("attention_bwd", 10.0 * B * H * Lq * Lk * 64):
        return
    with _timed("attention_bwd", 10.0 * B * H * Lq * Lk * 64):
        check(lib().tan_attention_bwd_bf16(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(),
                                           v.stride(0), o.data_ptr(), o.stride(0), d_out.data_ptr(), d_out.stride(0),
                                           _ptr(kpm_u8), dq.data_ptr(), dq.stride(0), dk.data_ptr(), dk.stride(0),
                                           dv.data_ptr(), dv.stride(0), lse.data_ptr(), delta.data_ptr(), B, H, Lq, Lk,
                                           _stream()), "tan_attention_bwd_bf16")
    _launches += 2
