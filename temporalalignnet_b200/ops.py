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
_prof = None           # when a list: (kernel, work, start_event, end_event) per call (`profile` context)


def _launch(cls: str, work: float, n: int, call) -> None:
    """The single choke point of every kernel launch of this module: `call()` performs the C-ABI call (which
    enqueues `n` kernels of class `cls` doing `work` algorithmic flops / elements).  The product always launches;
    measurement tools (bench.py's per-class graphs) wrap THIS function from outside -- there is no switch in the
    product that could make a wrapper return without launching."""
    global _launches
    if _prof is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        call()
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        _prof.append((cls, work, e0, e1))
    else:
        call()
    _launches += n


class profile:
    """Context manager: bracket every C-ABI call with CUDA events on the launching stream."""

    def __enter__(self):
        global _prof
        _prof = []
        return _prof

    def __exit__(self, *a):
        global _prof
        _prof = None


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


def pad64(n: int) -> int:
    return (n + 63) // 64 * 64


def cast_bf16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 -> bf16 (tan_cast_f32_to_bf16)."""
    _need(x, torch.float32, "cast_bf16.in")
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _launch("cast", float(x.numel()), 1, lambda: check(
        lib().tan_cast_f32_to_bf16(x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "tan_cast_f32_to_bf16"))
    return out


def linear(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
           residual: Optional[torch.Tensor] = None, out_f32: Optional[torch.Tensor] = None,
           out_bf16: Optional[torch.Tensor] = None, act: int = _lib.ACT_NONE, tag: str = "linear") -> None:
    """out = act(a @ w.T + bias) [+ residual]  (tan_linear_bf16).  a [M,K] bf16 (row pitch may exceed
    K), w [N,K] bf16; outputs 2-D with arbitrary row pitch.  `tag` names the kernel class in profiles
    (forward projections: "linear"; backward: "dgrad", "sim_bwd")."""
    M, K = a.shape
    N = w.shape[0]
    _launch(tag, 2.0 * M * N * K, 1, lambda: check(lib().tan_linear_bf16(
        a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), _ptr(bias),
        _ptr(residual), residual.stride(0) if residual is not None else 0,
        _ptr(out_f32), out_f32.stride(0) if out_f32 is not None else 0,
        _ptr(out_bf16), out_bf16.stride(0) if out_bf16 is not None else 0,
        M, N, K, act, _stream()), "tan_linear_bf16"))


def linear_dual(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], out_act: torch.Tensor,
                out_pre: torch.Tensor, act: int = _lib.ACT_QUICKGELU) -> None:
    """tan_linear_dual_bf16: out_act = act(a @ w.T + bias) and out_pre = a @ w.T + bias (both bf16) from one GEMM."""
    M, K = a.shape
    N = w.shape[0]
    _launch("linear", 2.0 * M * N * K, 1, lambda: check(lib().tan_linear_dual_bf16(
        a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), _ptr(bias), out_act.data_ptr(), out_act.stride(0),
        out_pre.data_ptr(), out_pre.stride(0), M, N, K, act, _stream()), "tan_linear_dual_bf16"))


def linear_gelu_bwd(a: torch.Tensor, w: torch.Tensor, u: torch.Tensor, out: torch.Tensor) -> None:
    """tan_linear_gelu_bwd_bf16: out = (a @ w.T) * gelu'(u)  (dgrad through c_proj and QuickGELU in one GEMM)."""
    M, K = a.shape
    N = w.shape[0]
    _launch("dgrad", 2.0 * M * N * K, 1, lambda: check(lib().tan_linear_gelu_bwd_bf16(
        a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), u.data_ptr(), u.stride(0), out.data_ptr(), out.stride(0),
        M, N, K, _stream()), "tan_linear_gelu_bwd_bf16"))


def linear_res_ln(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], x: torch.Tensor,
                  gamma: torch.Tensor, beta: torch.Tensor, out_bf16: torch.Tensor) -> None:
    """x += a @ w.T + bias (fp32, in place); out_bf16 = LayerNorm(x) * gamma + beta  (tan_linear_res_ln_bf16;
    output width 512).  Counted with the linear class (flops of the GEMM)."""
    M, K = a.shape
    N = w.shape[0]
    _launch("linear", 2.0 * M * N * K, 1, lambda: check(lib().tan_linear_res_ln_bf16(
        a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), _ptr(bias), x.data_ptr(), x.stride(0), gamma.data_ptr(),
        beta.data_ptr(), out_bf16.data_ptr(), out_bf16.stride(0), M, N, K, _stream()), "tan_linear_res_ln_bf16"))


def linear_res_ln_stage(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], x: torch.Tensor,
                        gamma: torch.Tensor, beta: torch.Tensor, out_bf16: Optional[torch.Tensor], L: int,
                        l_split: int, nrmA_bf16: Optional[torch.Tensor], strideA: int,
                        nrmB_bf16: Optional[torch.Tensor], strideB: int) -> None:
    """tan_linear_res_ln_stage_bf16: linear_res_ln + L2-normalised bf16 stage features scattered by clip
    (nrm pointers are views whose data_ptr() is the stage's first row)."""
    M, K = a.shape
    N = w.shape[0]
    _launch("linear", 2.0 * M * N * K, 1, lambda: check(lib().tan_linear_res_ln_stage_bf16(
        a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), _ptr(bias), x.data_ptr(), x.stride(0),
        gamma.data_ptr(), beta.data_ptr(), _ptr(out_bf16), out_bf16.stride(0) if out_bf16 is not None else 0,
        M, N, K, L, l_split, _ptr(nrmA_bf16), strideA, _ptr(nrmB_bf16), strideB, _stream()),
        "tan_linear_res_ln_stage_bf16"))


def layernorm(x: torch.Tensor, rows: int, d: int, gamma=None, beta=None, add=None, add_rows: int = 0,
              L_in: Optional[int] = None, L_out: Optional[int] = None, l_off: int = 0,
              out_f32=None, out_bf16=None, l_split: int = 0, strideA: int = 0, strideB: int = 0,
              rawA=None, rawB=None, nrmA_bf16=None, nrmB_bf16=None, nrmA_f32=None, nrmB_f32=None,
              raw_strideA: int = 0, raw_strideB: int = 0) -> None:
    """tan_layernorm; pointer-valued keyword arguments are tensors whose data_ptr() already points at
    the first destination row (callers pass views).  raw_stride*: row strides of rawA / rawB when they differ
    from strideA / strideB (0 = the same)."""
    L_in = rows if L_in is None else L_in
    L_out = L_in if L_out is None else L_out
    a = LnArgs(x.data_ptr(), int(x.dtype == torch.bfloat16), rows, d, _ptr(gamma), _ptr(beta), _ptr(add), add_rows,
               L_in, L_out, l_off, _ptr(out_f32), _ptr(out_bf16), l_split, strideA, strideB,
               _ptr(rawA), _ptr(rawB), _ptr(nrmA_bf16), _ptr(nrmB_bf16), _ptr(nrmA_f32), _ptr(nrmB_f32),
               raw_strideA, raw_strideB)
    _launch("layernorm", float(rows) * d, 1, lambda: check(lib().tan_layernorm(C.byref(a), _stream()), "tan_layernorm"))


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, kpm_u8: Optional[torch.Tensor],
              out: torch.Tensor, B: int, H: int, Lq: int, Lk: int, lse: Optional[torch.Tensor] = None) -> None:
    """tan_attention_bf16.  q/k/v/out are 2-D (possibly column-sliced) bf16 views [B*L, H*64]; lse (optional)
    [B, H, pad64(Lq)] fp32 receives the rows' log2-domain log-sum-exp for the backward pass."""
    _launch("attention", 4.0 * B * H * Lq * Lk * 64, 1, lambda: check(lib().tan_attention_bf16(
        q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0), _ptr(kpm_u8), out.data_ptr(),
        out.stride(0), B, H, Lq, Lk, _ptr(lse), _stream()), "tan_attention_bf16"))


def sim_geom(B_loc, S, T, C_, N, d, b_off=0, col_off: Optional[torch.Tensor] = None) -> SimGeom:
    """struct tan_sim_geom; col_off (device int32 [B_glob + 1]) switches to ragged columns (see include/tan_b200.h).
    The returned struct only holds the POINTER: the caller keeps `col_off` alive."""
    return SimGeom(B_loc, S, T, C_, N, d, b_off, _ptr(col_off))


def sim_workspace_bytes(g: SimGeom) -> int:
    return int(lib().tan_sim_nce_workspace_bytes(C.byref(g)))


def pos_from_time(start: torch.Tensor, end: torch.Tensor, valid_u8: Optional[torch.Tensor], B: int, T: int, N: int,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """tan_pos_from_time: packed target bits [B, T, ceil(N/32)] (int32 storage of the uint32 words)."""
    W = (N + 31) // 32
    if out is None:
        out = torch.empty(B, T, W, dtype=torch.int32, device=start.device)
    _need(start, torch.float32, "pos_from_time.start")
    _need(end, torch.float32, "pos_from_time.end")
    _launch("glue", float(B * T * W), 1, lambda: check(lib().tan_pos_from_time(
        start.data_ptr(), end.data_ptr(), _ptr(valid_u8), B, T, N, out.data_ptr(), _stream()), "tan_pos_from_time"))
    return out


def sim_nce_fwd(vfeat, tfeat, tfeat_stage_stride: int, g: SimGeom, posbits, col_valid, logits_out,
                row_sums, col_sums, workspace, row_kill=None) -> None:
    _launch("sim_nce_fwd", 2.0 * g.B_loc * g.S * g.T * g.C * g.d, 2, lambda: check(lib().tan_sim_nce_fwd(
        vfeat.data_ptr(), tfeat.data_ptr(), tfeat_stage_stride, C.byref(g), posbits.data_ptr(), col_valid.data_ptr(),
        _ptr(row_kill), _ptr(logits_out), row_sums.data_ptr(), col_sums.data_ptr(), workspace.data_ptr(),
        workspace.numel() * workspace.element_size(), _stream()), "tan_sim_nce_fwd"))


def nce_from_logits(logits, g: SimGeom, posbits, col_valid, row_sums, col_sums, workspace, row_kill=None) -> None:
    is_f32 = int(logits.dtype == torch.float32)
    if not is_f32 and logits.dtype != torch.bfloat16:
        raise _lib.TanError(f"nce_from_logits: logits must be fp32 or bf16, got {logits.dtype}")
    _launch("nce_from_logits", float(logits.numel()) * logits.element_size(), 2, lambda: check(lib().tan_nce_from_logits(
        logits.data_ptr(), is_f32, C.byref(g), posbits.data_ptr(), col_valid.data_ptr(), _ptr(row_kill),
        row_sums.data_ptr(), col_sums.data_ptr(), workspace.data_ptr(), workspace.numel() * workspace.element_size(),
        _stream()), "tan_nce_from_logits"))


def nce_reduce(row_sums, col_sums, out4_f64, S: int, T: int, C_: int, row_sel=None, col_sel=None) -> None:
    """tan_nce_reduce.  row_sums [2, B_loc*S*T] or None, col_sums [2, S, C] or None; row_sel [B_loc, T] /
    col_sel [C] uint8 optional selections (thresholded loss)."""
    R = row_sums.numel() // 2 if row_sums is not None else 0
    SC = col_sums.numel() // 2 if col_sums is not None else 0
    _launch("nce_reduce", float(R + SC), 1, lambda: check(lib().tan_nce_reduce(
        _ptr(row_sums), R, S, T, _ptr(row_sel), _ptr(col_sums), SC, C_, _ptr(col_sel), out4_f64.data_ptr(), _stream()),
        "tan_nce_reduce"))


def own_clip_sim(vfeat, tfeat, shared_text: bool, B: int, S: int, T: int, N: int, d: int, s_first: int = 0,
                 s_count: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """tan_own_clip_sim -> fp32 [B, s_count, T, N] own-clip cosine blocks."""
    s_count = S - s_first if s_count is None else s_count
    if out is None:
        out = torch.empty(B, s_count, T, N, dtype=torch.float32, device=vfeat.device)
    _launch("own_clip_sim", 2.0 * B * s_count * T * N * d, 1, lambda: check(lib().tan_own_clip_sim(
        vfeat.data_ptr(), tfeat.data_ptr(), 0 if shared_text else B * N * d, B, S, T, N, d, s_first, s_count,
        out.data_ptr(), _stream()), "tan_own_clip_sim"))
    return out


def align_stitch(blk_joint: torch.Tensor, blk_dual: torch.Tensor, windows_i32: torch.Tensor, sim_joint: torch.Tensor,
                 sim_dual: torch.Tensor, cover: torch.Tensor, accumulate: bool, finalize: bool) -> None:
    """tan_align_stitch: blk_* [W, T, N] fp32 own-clip cosines of a batch of windows, windows [W, 4] int32
    (t0, t1, n0, n1); sim_* / cover [n_text, vlen] fp32 running sums -> averages when `finalize`."""
    _need(blk_joint, torch.float32, "align_stitch.blk_joint")
    _need(blk_dual, torch.float32, "align_stitch.blk_dual")
    _need(windows_i32, torch.int32, "align_stitch.windows")
    for t in (sim_joint, sim_dual, cover):
        _need(t, torch.float32, "align_stitch.out")
    W, T, N = blk_joint.shape
    n_text, vlen = sim_joint.shape
    _launch("align", float(n_text * vlen * max(W, 1)), 1, lambda: check(lib().tan_align_stitch(
        blk_joint.data_ptr(), blk_dual.data_ptr(), windows_i32.data_ptr(), W, T, N, sim_joint.data_ptr(),
        sim_dual.data_ptr(), cover.data_ptr(), n_text, vlen, int(bool(accumulate)), int(bool(finalize)), _stream()),
        "tan_align_stitch"))


def align_argmax(sim: torch.Tensor) -> torch.Tensor:
    """tan_align_argmax: per sentence, argmax over time of the soft-max of sim (0 = uncovered -> -6e4) -> int64 [n_text]."""
    _need(sim, torch.float32, "align_argmax.sim")
    n_text, vlen = sim.shape
    out = torch.empty(n_text, dtype=torch.int64, device=sim.device)
    _launch("align", float(n_text * vlen), 1, lambda: check(lib().tan_align_argmax(
        sim.data_ptr(), n_text, vlen, out.data_ptr(), _stream()), "tan_align_argmax"))
    return out


def agree_scan(own, posbits, vpm_u8, tpm_u8, B: int, T: int, N: int, fill_max: bool):
    """tan_agree_scan -> (win [B,N,2] int32, mean_logit [B,N], max_logit [B,N])."""
    dev = own.device
    win = torch.empty(B, N, 2, dtype=torch.int32, device=dev)
    mean_logit = torch.empty(B, N, dtype=torch.float32, device=dev)
    max_logit = torch.empty(B, N, dtype=torch.float32, device=dev)
    _need(own, torch.float32, "agree_scan.own")
    ws = torch.empty(int(lib().tan_agree_scan_workspace_bytes(B, T, N)), dtype=torch.uint8, device=dev)
    _launch("agree", float(B * T * N), 2, lambda: check(lib().tan_agree_scan(
        own.data_ptr(), posbits.data_ptr(), _ptr(vpm_u8), tpm_u8.data_ptr(), B, T, N, int(bool(fill_max)),
        win.data_ptr(), mean_logit.data_ptr(), max_logit.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
        "tan_agree_scan"))
    return win, mean_logit, max_logit


AGREE_KINDS = {"i": 0, "u": 1, "keep": 2, "keep-joint": 3}


def agree_targets(old_posbits, win_joint, win_dual, replace_u8, B: int, T: int, N: int, kind: str) -> torch.Tensor:
    """tan_agree_targets -> new packed target bits [B, T, W]."""
    out = torch.empty_like(old_posbits)
    _launch("agree", float(B * T), 1, lambda: check(lib().tan_agree_targets(
        old_posbits.data_ptr(), win_joint.data_ptr(), win_dual.data_ptr(), replace_u8.data_ptr(), B, T, N,
        AGREE_KINDS[kind], out.data_ptr(), _stream()), "tan_agree_targets"))
    return out


# ------------------------------------------------------------------------------------------------------
# backward pass (training step)
# ------------------------------------------------------------------------------------------------------
def transpose_bf16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """tan_transpose_bf16: x [R, C] bf16 (row pitch may exceed C) -> [C, pad64(R)] with zero padding (weight
    shadows W^T of the dgrad GEMMs and the text operand of dA = G @ text)."""
    R, Ccols = x.shape
    Rp = pad64(R)
    if out is None:
        out = torch.empty(Ccols, Rp, dtype=torch.bfloat16, device=x.device)
    _launch("transpose", 2.0 * R * Ccols, 1, lambda: check(lib().tan_transpose_bf16(
        x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), R, Ccols, Rp, _stream()), "tan_transpose_bf16"))
    return out


_tn_ws = {}


def gemm_tn(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, accumulate: bool = True, tag: str = "wgrad") -> None:
    """tan_gemm_tn_bf16: out [P, Q] fp32 (+)= a[R, P]^T @ b[R, Q] (a, b bf16 row-major, row pitch may exceed the
    width): weight gradients and the text-side similarity gradient without transposes."""
    R, P = a.shape
    Q = b.shape[1]
    nbytes = int(lib().tan_gemm_tn_workspace_bytes(R, P, Q))
    ws = None
    if nbytes:
        key = (str(a.device), torch.cuda.current_stream().cuda_stream)
        ws = _tn_ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = _tn_ws[key] = torch.empty(nbytes, dtype=torch.uint8, device=a.device)
    _launch(tag, 2.0 * R * P * Q, 2 if nbytes else 1, lambda: check(lib().tan_gemm_tn_bf16(
        a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), R, P, Q, out.data_ptr(), out.stride(0),
        int(bool(accumulate)), _ptr(ws), nbytes, _stream()), "tan_gemm_tn_bf16"))


def colsum(x: torch.Tensor, out: torch.Tensor, accumulate: bool = True) -> None:
    """tan_colsum: out [N] fp32 (+)= column sums of x [M, N] (bf16 or fp32)."""
    M, N = x.shape
    nbytes = int(lib().tan_colsum_workspace_bytes(M, N))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    _launch("colsum", float(M) * N, 2, lambda: check(lib().tan_colsum(
        x.data_ptr(), int(x.dtype == torch.bfloat16), x.stride(0), M, N, out.data_ptr(), int(bool(accumulate)),
        ws.data_ptr(), nbytes, _stream()), "tan_colsum"))


def quickgelu_fwd(u: torch.Tensor, h: torch.Tensor) -> None:
    _launch("gelu", float(u.numel()), 1, lambda: check(
        lib().tan_quickgelu_fwd(u.data_ptr(), h.data_ptr(), u.numel(), _stream()), "tan_quickgelu_fwd"))


def quickgelu_bwd(dh: torch.Tensor, u: torch.Tensor, du: torch.Tensor) -> None:
    _launch("gelu", float(u.numel()), 1, lambda: check(
        lib().tan_quickgelu_bwd(dh.data_ptr(), u.data_ptr(), du.data_ptr(), u.numel(), _stream()), "tan_quickgelu_bwd"))


def layernorm_bwd(dy: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, dx: torch.Tensor, accumulate_dx: bool,
                  rows: int, d: int, dgamma: Optional[torch.Tensor], dbeta: Optional[torch.Tensor],
                  L_in: Optional[int] = None, L_out: Optional[int] = None, l_off: int = 0,
                  dx_bf16: Optional[torch.Tensor] = None, dx_colsum: Optional[torch.Tensor] = None) -> None:
    """tan_layernorm_bwd (dy, x, dx fp32; dgamma / dbeta accumulated).  dx_bf16 / dx_colsum: bf16 copy and column
    sums (accumulated) of the updated dx -- the operand and the bias gradient of the next linear's backward."""
    L_in = rows if L_in is None else L_in
    L_out = L_in if L_out is None else L_out
    nbytes = int(lib().tan_layernorm_bwd_workspace_bytes(rows, d))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    _launch("ln_bwd", float(rows) * d, 2, lambda: check(lib().tan_layernorm_bwd(
        dy.data_ptr(), x.data_ptr(), _ptr(gamma), dx.data_ptr(), int(bool(accumulate_dx)), rows, d, L_in, L_out, l_off,
        _ptr(dgamma), _ptr(dbeta), _ptr(dx_bf16), _ptr(dx_colsum), ws.data_ptr(), nbytes, _stream()),
        "tan_layernorm_bwd"))


def l2norm_bwd(x: torch.Tensor, g: torch.Tensor, dst: torch.Tensor, accumulate: bool, rows: int, d: int, L_in: int,
               src_stride: int, L_out: int, l_off: int, g_stride: Optional[int] = None) -> None:
    """tan_l2norm_bwd; x / g are views whose data_ptr() is the stage's first row."""
    _launch("l2norm_bwd", float(rows) * d, 1, lambda: check(lib().tan_l2norm_bwd(
        x.data_ptr(), g.data_ptr(), dst.data_ptr(), int(bool(accumulate)), rows, d, L_in, src_stride,
        src_stride if g_stride is None else g_stride, L_out, l_off, _stream()), "tan_l2norm_bwd"))


def batch_sum(x: torch.Tensor, out: torch.Tensor, B: int, L: int, d: int, L_out: int, l_off: int,
              accumulate: bool) -> None:
    _launch("batch_sum", float(B) * L * d, 1, lambda: check(lib().tan_batch_sum(
        x.data_ptr(), out.data_ptr(), B, L, d, L_out, l_off, int(bool(accumulate)), _stream()), "tan_batch_sum"))


def sim_grad_tiles(z: torch.Tensor, Rc: int, r0: int, g: SimGeom, posbits, col_valid, row_kill, ra, rap, cb, cbp,
                   G: torch.Tensor, GT: torch.Tensor) -> None:
    """tan_sim_grad_tiles: z [Rc, >=C] fp32 -> G [Rc, Cp] bf16, GT [C, pad64(Rc)] bf16."""
    _launch("sim_grad", float(Rc) * g.C, 1, lambda: check(lib().tan_sim_grad_tiles(
        z.data_ptr(), z.stride(0), Rc, pad64(Rc), r0, C.byref(g), posbits.data_ptr(), col_valid.data_ptr(),
        _ptr(row_kill), ra.data_ptr(), rap.data_ptr(), cb.data_ptr(), cbp.data_ptr(), G.data_ptr(), G.stride(0),
        GT.data_ptr(), GT.stride(0), _stream()), "tan_sim_grad_tiles"))


def sim_grad_gemm(a: torch.Tensor, t_pad: torch.Tensor, r0: int, g: SimGeom, posbits, col_valid, row_kill, ra, rap, cb,
                  cbp, G: torch.Tensor) -> None:
    """tan_sim_grad_gemm: a [Rc, d] bf16 video rows, t_pad [Cp, d] bf16 text rows (zero beyond C) -> G [Rc, Cp] bf16."""
    Rc, d = a.shape
    Cp = t_pad.shape[0]
    _launch("sim_bwd", 2.0 * Rc * Cp * d, 1, lambda: check(lib().tan_sim_grad_gemm(
        a.data_ptr(), a.stride(0), t_pad.data_ptr(), t_pad.stride(0), Rc, r0, C.byref(g), Cp, posbits.data_ptr(),
        col_valid.data_ptr(), _ptr(row_kill), ra.data_ptr(), rap.data_ptr(), cb.data_ptr(), cbp.data_ptr(), G.data_ptr(),
        G.stride(0), _stream()), "tan_sim_grad_gemm"))


def attention_bwd(q, k, v, o, d_out, kpm_u8, dq, dk, dv, lse, delta, B: int, H: int, Lq: int, Lk: int) -> None:
    """tan_attention_bwd_bf16 (2-D, possibly column-sliced bf16 views as in `attention`); lse [B, H, pad64(Lq)] is
    what the forward `attention(..., lse=)` stored, delta a workspace of the same shape."""
    _launch("attention_bwd", 10.0 * B * H * Lq * Lk * 64, 3, lambda: check(lib().tan_attention_bwd_bf16(
        q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0), o.data_ptr(), o.stride(0),
        d_out.data_ptr(), d_out.stride(0), _ptr(kpm_u8), dq.data_ptr(), dq.stride(0), dk.data_ptr(), dk.stride(0),
        dv.data_ptr(), dv.stride(0), lse.data_ptr(), delta.data_ptr(), B, H, Lq, Lk, _stream()),
        "tan_attention_bwd_bf16"))


# ------------------------------------------------------------------------------------------------------
# text embedder (model/word2vec_model.py)
# ------------------------------------------------------------------------------------------------------
def embed_gather(ids: torch.Tensor, table_bf16: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """tan_embed_gather_bf16: ids [n] int64 -> rows of table_bf16 [V, ld] -> [n, ld] bf16."""
    _need(ids, torch.int64, "embed_gather.ids")
    _need(table_bf16, torch.bfloat16, "embed_gather.table")
    n, (V, ld) = ids.numel(), table_bf16.shape
    if out is None:
        out = torch.empty(n, ld, dtype=torch.bfloat16, device=ids.device)
    _launch("text_embed", float(n) * ld, 1, lambda: check(lib().tan_embed_gather_bf16(
        ids.data_ptr(), table_bf16.data_ptr(), ld, V, n, out.data_ptr(), _stream()), "tan_embed_gather_bf16"))
    return out


def text_pool_fc1(x: torch.Tensor, w1: torch.Tensor, b1: Optional[torch.Tensor], keep_u8: Optional[torch.Tensor],
                  S: int, want_argmax: bool = False):
    """tan_text_pool_fc1: x [S*32, K] bf16, w1 [F, K] bf16 -> (pooled [S, F] bf16, argmax [S, F] uint8 | None)."""
    K = x.shape[1]
    F = w1.shape[0]
    pooled = torch.empty(S, F, dtype=torch.bfloat16, device=x.device)
    arg = torch.empty(S, F, dtype=torch.uint8, device=x.device) if want_argmax else None
    _launch("text_embed", 2.0 * S * 32 * F * K, 1, lambda: check(lib().tan_text_pool_fc1(
        x.data_ptr(), x.stride(0), w1.data_ptr(), w1.stride(0), _ptr(b1), _ptr(keep_u8), S, F, K, pooled.data_ptr(),
        _ptr(arg), _stream()), "tan_text_pool_fc1"))
    return pooled, arg


def text_pool_bwd(dpool: torch.Tensor, pooled: torch.Tensor, argmax: torch.Tensor) -> torch.Tensor:
    """tan_text_pool_bwd -> dH [S*32, F] bf16."""
    S, F = pooled.shape
    dH = torch.empty(S * 32, F, dtype=torch.bfloat16, device=pooled.device)
    _launch("text_embed", float(S) * 32 * F, 1, lambda: check(lib().tan_text_pool_bwd(
        dpool.data_ptr(), pooled.data_ptr(), argmax.data_ptr(), S, F, dH.data_ptr(), _stream()), "tan_text_pool_bwd"))
    return dH
