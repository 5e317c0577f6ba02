"""Batched tensor-level operators over the C ABI (include/dmm_b200.h), with autograd.

PyTorch is only the plumbing here: device memory, the current CUDA stream, autograd bookkeeping.
Every function takes CUDA fp32 tensors and launches hand-written sm_100a kernels from libdmm_b200.so;
nothing falls back to torch ops or to the CPU.

Batch convention: B independent (video, frame) problems of P proposals x O templates over H*W pixels;
optional int32 ``n_prop[B]`` / ``n_tmpl[B]`` give the number of real rows per problem (rest is padding).
"""
from __future__ import annotations

import ctypes
import functools
from typing import Optional, Sequence, Tuple

import torch

from . import _lib

_VP = ctypes.c_void_p


# ----------------------------------------------------------------------------------------------------------
# device guard + optional per-op CUDA-event timing
# ----------------------------------------------------------------------------------------------------------
class KernelTimer:
    """Collects (op name, start event, end event) for every public op while installed with ``set_kernel_timer``
    (bench.py's clip legs report each op's share of the frame from it).  Events are recorded on the stream the op
    launches on; nothing synchronises until ``totals()``."""

    def __init__(self):
        self.spans = []

    def totals(self):
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1 in self.spans:
            out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
        return out


_TIMER: Optional[KernelTimer] = None
_TIMER_DEPTH = 0


def set_kernel_timer(timer: Optional[KernelTimer]) -> None:
    global _TIMER
    _TIMER = timer


def _first_cuda_device(args, kwargs):
    for a in list(args) + list(kwargs.values()):
        if isinstance(a, torch.Tensor):
            if a.is_cuda:
                return a.device
        elif isinstance(a, RaggedMasks):
            return a.device
        elif isinstance(a, (list, tuple)) and a and isinstance(a[0], torch.Tensor) and a[0].is_cuda:
            return a[0].device
    return None


def _op(name: str):
    """Every public op runs with the CUDA device of its tensors current (the C ABI launches on the current device's
    stream: without the guard, tensors of a non-current device would be handed to kernels of another GPU), and is
    timed when a KernelTimer is installed (outermost op only)."""

    def deco(fn):
        @functools.wraps(fn)
        def wrap(*args, **kwargs):
            global _TIMER_DEPTH
            dev = _first_cuda_device(args, kwargs)
            if dev is not None and dev.index is not None and dev.index != torch.cuda.current_device():
                with torch.cuda.device(dev):
                    return wrap(*args, **kwargs)
            t = _TIMER
            if t is None or _TIMER_DEPTH > 0:
                return fn(*args, **kwargs)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            _TIMER_DEPTH += 1
            try:
                e0.record()
                out = fn(*args, **kwargs)
                e1.record()
            finally:
                _TIMER_DEPTH -= 1
            t.spans.append((name, e0, e1))
            return out

        return wrap

    return deco


def _p(t: Optional[torch.Tensor]):
    return None if t is None else _VP(t.data_ptr())


def _stream() -> ctypes.c_void_p:
    return _VP(torch.cuda.current_stream().cuda_stream)


def _cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"dmm_net_b200: `{name}` must be a CUDA tensor (there is no CPU fallback)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _counts(t: Optional[torch.Tensor], B: int, dev) -> Optional[torch.Tensor]:
    if t is None:
        return None
    t = torch.as_tensor(t, device=dev).to(torch.int32).contiguous()
    assert t.shape == (B,), t.shape
    return t


_LIMITS = None


def limits() -> dict:
    """Shape envelope of the solver kernel (dmm_b200_limits): the whole [O x max(P, O+1)] problem lives in the registers
    of one CTA, so O (template slots = the reference's ``-maxseqlen``) and P (``sort_max_num``) are bounded."""
    global _LIMITS
    if _LIMITS is None:
        out = (ctypes.c_int * 4)()
        _lib.check(_lib.load().dmm_b200_limits(out), "dmm_b200_limits")
        _LIMITS = {"max_templates": int(out[0]), "max_solver_cols": int(out[1]), "max_batch": int(out[2])}
    return _LIMITS


def check_solver_shape(P: int, O: int) -> None:
    lim = limits()
    if O > lim["max_templates"] or pad_cols(P, O) > lim["max_solver_cols"]:
        raise RuntimeError(
            f"dmm_net_b200: the relaxed-matching kernel holds one problem per CTA and supports at most "
            f"{lim['max_templates']} template slots (-maxseqlen) and {lim['max_solver_cols']} proposals (sort_max_num); "
            f"got O={O}, P={P}. Lower -maxseqlen / sort_max_num, or pass only the valid templates.")


def pad_cols(P: int, O: int) -> int:
    """Column stride of the [O x m] solver matrices: P <= O problems are padded to O+1 (match_model.py:109-113)."""
    return max(P, O + 1)


class RaggedMasks:
    """Per-problem proposal mask tensors used IN PLACE through a device pointer table (no torch.stack copy):
    tensors[b] is [P_b, H, W] (or [P_b, HW]) fp32 on the GPU; P = max P_b; n_prop[b] = P_b."""

    def __init__(self, tensors):
        assert len(tensors) > 0
        self.tensors = [_cuda_f32(t, "prop_mask[b]") for t in tensors]       # kept alive with this object
        t0 = self.tensors[0]
        self.device = t0.device
        self.shape_hw = tuple(t0.shape[1:])
        self.HW = 1
        for dsz in self.shape_hw:
            self.HW *= int(dsz)
        for t in self.tensors:
            assert tuple(t.shape[1:]) == self.shape_hw, (t.shape, self.shape_hw)
        self.B = len(self.tensors)
        counts = [int(t.shape[0]) for t in self.tensors]
        self.P = max(max(counts), 1)
        # an empty proposal set still needs a valid address: point it at any other tensor (n_prop = 0 masks it)
        anyptr = next((t.data_ptr() for t in self.tensors if t.numel() > 0), 0)
        ptrs = [t.data_ptr() if t.numel() > 0 else anyptr for t in self.tensors]
        self.aligned16 = int(all(q % 16 == 0 for q in ptrs) and self.HW % 4 == 0)
        self.ptrs = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
        self.n_prop = torch.tensor(counts, dtype=torch.int32, device=self.device)


# ----------------------------------------------------------------------------------------------------------
# K1  mask IoU
# ----------------------------------------------------------------------------------------------------------
@_op("K1 mask_iou")
def mask_iou_pairwise(prop: torch.Tensor, tmpl: torch.Tensor, tmpl2: Optional[torch.Tensor] = None,
                      n_prop=None, n_tmpl=None, cos: Optional[torch.Tensor] = None, w_cos: float = 0.0,
                      w_iou: float = 0.0, want_counts: bool = False):
    """prop [B,P,...], tmpl [B,O,...] (trailing dims flattened to HW) -> dict(iou [B,O,P], iou2, sim, counts)."""
    lib = _lib.load()
    if isinstance(prop, (list, tuple)):
        prop = RaggedMasks(prop)
    ragged = prop if isinstance(prop, RaggedMasks) else None
    tmpl = _cuda_f32(tmpl, "tmpl")
    O = tmpl.shape[1]
    if ragged is not None:
        B, P, HW = ragged.B, ragged.P, ragged.HW
        n_prop = ragged.n_prop if n_prop is None else n_prop
    else:
        prop = _cuda_f32(prop, "prop")
        B, P = prop.shape[:2]
        HW = 1
        for dsz in prop.shape[2:]:
            HW *= int(dsz)
        prop = prop.reshape(B, P, HW)
    assert tmpl.shape[0] == B and tmpl.numel() == B * O * HW, (B, P, HW, tmpl.shape)
    tmpl = tmpl.reshape(B, O, HW)
    dev = tmpl.device
    if tmpl2 is not None:
        tmpl2 = _cuda_f32(tmpl2, "tmpl2")
        assert tmpl2.numel() == tmpl.numel(), (tmpl2.shape, tmpl.shape)
        tmpl2 = tmpl2.reshape(B, O, HW)
    n_prop, n_tmpl = _counts(n_prop, B, dev), _counts(n_tmpl, B, dev)
    iou = torch.empty(B, O, P, device=dev)
    iou2 = torch.empty(B, O, P, device=dev) if tmpl2 is not None else None
    sim = None
    if cos is not None:
        cos = _cuda_f32(cos, "cos")
        assert cos.shape == (B, O, P)
        sim = torch.empty(B, O, P, device=dev)
    counts = torch.empty(B, O * P + O + P, device=dev, dtype=torch.int32) if want_counts else None
    out = {"iou": iou, "iou2": iou2, "sim": sim, "counts": counts}
    if B == 0 or P == 0 or O == 0:
        return out
    one_pass_wide = ragged is None and HW % 4 == 0 and P <= 64 and 2 * O <= 32 and P + 2 * O <= 96    # K1's 96-row TMA tile
    if tmpl2 is not None and not one_pass_wide and (P + 2 * O > 64 or 2 * O > 16) and P + O <= 64 and O <= 16:
        # both template sets in one pass would need several tiles (rows re-read, LDG path); two single-tile passes
        # keep the TMA ring and read the proposals twice instead of up to four times
        first = mask_iou_pairwise(ragged if ragged is not None else prop, tmpl, None, n_prop, n_tmpl, cos, w_cos, w_iou,
                                  want_counts)
        second = mask_iou_pairwise(ragged if ragged is not None else prop, tmpl2, None, n_prop, n_tmpl)
        first["iou2"] = second["iou"]
        return first
    step = 65535  # grid.y limit
    for s in range(0, B, step):
        e = min(B, s + step)
        nb = e - s
        ws_bytes = lib.dmm_mask_iou_workspace_bytes(nb, P, O, max(HW, 1), int(tmpl2 is not None))
        ws = torch.empty(max(ws_bytes, 256), device=dev, dtype=torch.uint8)
        sl = lambda t: None if t is None else t[s:e]
        if ragged is not None:
            rc = lib.dmm_mask_iou_pairwise_ptrs(_p(ragged.ptrs[s:e]), ragged.aligned16, _p(tmpl[s:e]), O * HW,
                                                _p(sl(tmpl2)), O * HW, nb, P, O, HW, _p(sl(n_prop)), _p(sl(n_tmpl)),
                                                _p(iou[s:e]), _p(sl(iou2)), _p(sl(cos)), float(w_cos), float(w_iou),
                                                _p(sl(sim)), _p(sl(counts)), _p(ws), ws.numel(), _stream())
        else:
            rc = lib.dmm_mask_iou_pairwise(_p(prop[s:e]), P * HW, _p(tmpl[s:e]), O * HW, _p(sl(tmpl2)), O * HW, nb, P, O,
                                           HW, _p(sl(n_prop)), _p(sl(n_tmpl)), _p(iou[s:e]), _p(sl(iou2)), _p(sl(cos)),
                                           float(w_cos), float(w_iou), _p(sl(sim)), _p(sl(counts)), _p(ws), ws.numel(),
                                           _stream())
        _lib.check(rc, "dmm_mask_iou_pairwise")
    return out


@_op("K1 mask_iou_rowwise")
def mask_iou_rowwise(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a [N,M], b [N,M] -> iou [N]  (compute_iou_binary_mask_2D, match_helper.py:9-28)."""
    lib = _lib.load()
    a = _cuda_f32(a, "annotation")
    b = _cuda_f32(b, "segmentation")
    N, M = a.shape
    out = torch.empty(N, device=a.device)
    if N == 0:
        return out
    ws_bytes = lib.dmm_mask_iou_rowwise_workspace_bytes(N, max(M, 1))
    ws = torch.empty(max(ws_bytes, 256), device=a.device, dtype=torch.uint8)
    rc = lib.dmm_mask_iou_rowwise(_p(a), _p(b), N, M, _p(out), _p(ws), ws.numel(), _stream())
    _lib.check(rc, "dmm_mask_iou_rowwise")
    return out


# ----------------------------------------------------------------------------------------------------------
# bit-packed masks
# ----------------------------------------------------------------------------------------------------------
def packed_words(HW: int) -> int:
    return (int(HW) + 31) // 32


def pack_masks_host(masks: torch.Tensor, mask_dims: int = 2, threads: int = 0, out: Optional[torch.Tensor] = None):
    """HOST fp32 masks [..., H, W] (mask_dims=2) or [..., HW] (mask_dims=1) -> HOST int32 bit planes [..., words]
    (bit i of word j = pixel 32j+i > 0.5), packed by the host cores (std::thread + AVX-512/AVX2 inside libdmm_b200).  Pass a pinned
    ``out`` to make the following H2D copy asynchronous."""
    lib = _lib.load()
    assert not masks.is_cuda and masks.dtype == torch.float32
    masks = masks.contiguous()
    lead = tuple(masks.shape[:masks.dim() - mask_dims])
    HW = 1
    for dsz in masks.shape[masks.dim() - mask_dims:]:
        HW *= int(dsz)
    rows = 1
    for dsz in lead:
        rows *= int(dsz)
    words = packed_words(HW)
    if out is None:
        out = torch.empty(lead + (words,), dtype=torch.int32)
    assert out.dtype == torch.int32 and out.is_contiguous() and out.numel() == rows * words
    rc = lib.dmm_host_pack_masks(_VP(masks.data_ptr()), rows, HW, _VP(out.data_ptr()), int(threads))
    _lib.check(rc, "dmm_host_pack_masks")
    return out


@_op("mask_pack_bits")
def pack_masks(masks: torch.Tensor, mask_dims: int = 2) -> torch.Tensor:
    """DEVICE fp32 masks [..., H, W] -> DEVICE int32 bit planes [..., words]."""
    lib = _lib.load()
    masks = _cuda_f32(masks, "masks")
    lead = tuple(masks.shape[:masks.dim() - mask_dims])
    HW = 1
    for dsz in masks.shape[masks.dim() - mask_dims:]:
        HW *= int(dsz)
    rows = 1
    for dsz in lead:
        rows *= int(dsz)
    out = torch.empty(lead + (packed_words(HW),), dtype=torch.int32, device=masks.device)
    if rows * HW > 0:
        rc = lib.dmm_mask_pack_bits(_p(masks), rows, HW, _p(out), _stream())
        _lib.check(rc, "dmm_mask_pack_bits")
    else:
        out.zero_()
    return out


@_op("K1 mask_iou_packed")
def mask_iou_pairwise_packed(prop_bits: torch.Tensor, tmpl_bits: torch.Tensor, tmpl2_bits: Optional[torch.Tensor] = None,
                             n_prop=None, n_tmpl=None, cos: Optional[torch.Tensor] = None, w_cos: float = 0.0,
                             w_iou: float = 0.0, want_counts: bool = False):
    """K1 on bit-packed rows: prop_bits [B,P,words], tmpl_bits [B,O,words] (int32, CUDA) -> the dict of mask_iou_pairwise."""
    lib = _lib.load()
    for t in (prop_bits, tmpl_bits):
        assert t.is_cuda and t.dtype == torch.int32 and t.dim() == 3, "packed masks are CUDA int32 [B, rows, words]"
    prop_bits, tmpl_bits = prop_bits.contiguous(), tmpl_bits.contiguous()
    B, P, words = prop_bits.shape
    O = tmpl_bits.shape[1]
    assert tmpl_bits.shape == (B, O, words)
    dev = prop_bits.device
    if tmpl2_bits is not None:
        tmpl2_bits = tmpl2_bits.contiguous()
        assert tmpl2_bits.shape == (B, O, words)
    n_prop, n_tmpl = _counts(n_prop, B, dev), _counts(n_tmpl, B, dev)
    iou = torch.empty(B, O, P, device=dev)
    iou2 = torch.empty(B, O, P, device=dev) if tmpl2_bits is not None else None
    sim = None
    if cos is not None:
        cos = _cuda_f32(cos, "cos")
        sim = torch.empty(B, O, P, device=dev)
    counts = torch.empty(B, O * P + O + P, device=dev, dtype=torch.int32) if want_counts else None
    out = {"iou": iou, "iou2": iou2, "sim": sim, "counts": counts}
    if B * P * O == 0:
        return out
    step = 65535
    for s in range(0, B, step):
        e = min(B, s + step)
        nb = e - s
        ws = torch.empty(max(lib.dmm_mask_iou_packed_workspace_bytes(nb, P, O, words, int(tmpl2_bits is not None)), 256),
                         device=dev, dtype=torch.uint8)
        sl = lambda t: None if t is None else t[s:e]
        rc = lib.dmm_mask_iou_pairwise_packed(_p(prop_bits[s:e]), P * words, _p(tmpl_bits[s:e]), O * words, _p(sl(tmpl2_bits)),
                                              O * words, nb, P, O, words, _p(sl(n_prop)), _p(sl(n_tmpl)), _p(iou[s:e]),
                                              _p(sl(iou2)), _p(sl(cos)), float(w_cos), float(w_iou), _p(sl(sim)),
                                              _p(sl(counts)), _p(ws), ws.numel(), _stream())
        _lib.check(rc, "dmm_mask_iou_pairwise_packed")
    return out


# ----------------------------------------------------------------------------------------------------------
# K2  cosine
# ----------------------------------------------------------------------------------------------------------
class _CosineFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tmpl_feat, prop_feat, n_prop, n_tmpl, eps, impl):
        lib = _lib.load()
        B, T, O, D = tmpl_feat.shape
        P = prop_feat.shape[1]
        cos = torch.empty(B, O, P, device=prop_feat.device)
        if B * O * P > 0:
            if impl is None:
                rc = lib.dmm_cosine_pairwise(_p(tmpl_feat), _p(prop_feat), B, T, P, O, D, _p(n_prop), _p(n_tmpl),
                                             float(eps), _p(cos), _stream())
            else:
                rc = lib.dmm_cosine_pairwise_impl(_p(tmpl_feat), _p(prop_feat), B, T, P, O, D, _p(n_prop), _p(n_tmpl),
                                                  float(eps), _p(cos), COSINE_IMPLS[impl], _stream())
            _lib.check(rc, "dmm_cosine_pairwise")
        ctx.save_for_backward(tmpl_feat, prop_feat, n_prop, n_tmpl, cos)
        ctx.eps = eps
        return cos

    @staticmethod
    def backward(ctx, g_cos):
        lib = _lib.load()
        tmpl_feat, prop_feat, n_prop, n_tmpl, cos = ctx.saved_tensors
        B, T, O, D = tmpl_feat.shape
        P = prop_feat.shape[1]
        g_cos = g_cos.contiguous().float()
        gq = torch.zeros_like(tmpl_feat)
        gk = torch.zeros_like(prop_feat)
        if B * O * P * D > 0:
            rc = lib.dmm_cosine_pairwise_bwd(_p(g_cos), _p(cos), _p(tmpl_feat), _p(prop_feat), B, T, P, O, D, _p(n_prop),
                                             _p(n_tmpl), float(ctx.eps), _p(gq), _p(gk), _stream())
            _lib.check(rc, "dmm_cosine_pairwise_bwd")
        return gq, gk, None, None, None, None


COSINE_IMPLS = {"auto": 0, "simt": 1, "tc": 2}


@_op("K2 cosine")
def cosine_pairwise(tmpl_feat: torch.Tensor, prop_feat: torch.Tensor, n_prop=None, n_tmpl=None, eps: float = 1e-8,
                    impl: Optional[str] = None):
    """tmpl_feat [B,T,O,D] (T template-feature sets), prop_feat [B,P,D] -> mean_t cos [B,O,P]; differentiable.
    ``impl``: None (library default: the tcgen05 3xTF32 kernel inside its envelope, else fp32 FFMA), "tc", "simt"."""
    tmpl_feat = _cuda_f32(tmpl_feat, "tmpl_feat")
    prop_feat = _cuda_f32(prop_feat, "prop_feat")
    B = prop_feat.shape[0]
    assert tmpl_feat.dim() == 4 and prop_feat.dim() == 3 and tmpl_feat.shape[0] == B
    assert tmpl_feat.shape[3] == prop_feat.shape[2], (tmpl_feat.shape, prop_feat.shape)
    return _CosineFn.apply(tmpl_feat, prop_feat, _counts(n_prop, B, prop_feat.device),
                           _counts(n_tmpl, B, prop_feat.device), eps, impl)


# ----------------------------------------------------------------------------------------------------------
# K3  solver + head
# ----------------------------------------------------------------------------------------------------------
class _SolveFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mat, score, n_prop, n_tmpl, max_iter, proj_iter, lr, negate, pad_rule, is_test, want_xlist):
        lib = _lib.load()
        B, O, P = mat.shape
        MS = pad_cols(P, O) if pad_rule else P
        dev = mat.device
        new = lambda *s: torch.empty(*s, device=dev)
        R, Xf, Bm, logic = new(B, O, MS), new(B, O, MS), new(B, O, MS), new(B, O, MS)
        ms, ds = new(B, O), new(B, O)
        n_list = torch.empty(B, device=dev, dtype=torch.int32)
        xlist = new(B, max_iter + 1, O, MS) if want_xlist else None
        cost = new(B, max_iter + 1) if want_xlist else None
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        saved = None
        if need_grad:
            saved = torch.empty(lib.dmm_relax_saved_bytes(B, max_iter, proj_iter), device=dev, dtype=torch.uint8)
        if B * O > 0:
            rc = lib.dmm_relax_solve(_p(mat), _p(score), B, P, O, _p(n_prop), _p(n_tmpl), max_iter, proj_iter, float(lr),
                                     int(negate), int(pad_rule), int(is_test), _p(R), _p(Xf), _p(Bm), _p(logic), _p(ms),
                                     _p(ds), _p(n_list), _p(xlist), _p(cost), _p(saved), _stream())
            _lib.check(rc, "dmm_relax_solve")
        ctx.save_for_backward(mat, score, R, logic, n_list, saved, n_prop, n_tmpl)
        ctx.cfg = (max_iter, proj_iter, float(lr), int(negate), int(pad_rule))
        ctx.mark_non_differentiable(logic, n_list)
        if want_xlist:
            ctx.mark_non_differentiable(xlist, cost)
            return R, Bm, ms, ds, Xf, logic, n_list, xlist, cost
        return R, Bm, ms, ds, Xf, logic, n_list

    @staticmethod
    def backward(ctx, gR, gBm, gms, gds, gXf, *unused):
        lib = _lib.load()
        mat, score, R, logic, n_list, saved, n_prop, n_tmpl = ctx.saved_tensors
        max_iter, proj_iter, lr, negate, pad_rule = ctx.cfg
        B, O, P = mat.shape
        c = lambda g: None if g is None else g.contiguous().float()
        gR, gBm, gms, gds, gXf = c(gR), c(gBm), c(gms), c(gds), c(gXf)
        g_mat = torch.zeros_like(mat)
        g_score = torch.zeros_like(score) if score is not None else None
        if saved is None:
            raise RuntimeError("dmm_relax_solve: backward requested but the forward ran without grad inputs")
        if B * O * P > 0:
            rc = lib.dmm_relax_solve_bwd(_p(gR), _p(gXf), _p(gBm), _p(gms), _p(gds), _p(mat), _p(score), _p(R), _p(logic),
                                         _p(n_list), _p(saved), B, P, O, _p(n_prop), _p(n_tmpl), max_iter, proj_iter,
                                         lr, negate, pad_rule, _p(g_mat), _p(g_score), _stream())
            _lib.check(rc, "dmm_relax_solve_bwd")
        return g_mat, g_score, None, None, None, None, None, None, None, None, None


@_op("K3 relax_solve")
def relax_solve(mat: torch.Tensor, score: Optional[torch.Tensor] = None, n_prop=None, n_tmpl=None,
                max_iter: int = 20, proj_iter: int = 5, lr: float = 0.1, negate: bool = True, pad_rule: bool = True,
                is_test: bool = True, want_xlist: bool = False):
    """mat [B,O,P] (similarity when negate else cost) -> (R, Bmat, match_score, det_score, X_final, logic, n_list[, xlist, cost])."""
    mat = _cuda_f32(mat, "mat")
    B = mat.shape[0]
    check_solver_shape(mat.shape[2], mat.shape[1])
    if score is not None:
        score = _cuda_f32(score, "prop_score")
        assert score.shape == (B, mat.shape[2]), (score.shape, mat.shape)
    dev = mat.device
    return _SolveFn.apply(mat, score, _counts(n_prop, B, dev), _counts(n_tmpl, B, dev), int(max_iter), int(proj_iter),
                          float(lr), bool(negate), bool(pad_rule), bool(is_test), bool(want_xlist))


# ----------------------------------------------------------------------------------------------------------
# K4  assignment apply
# ----------------------------------------------------------------------------------------------------------
class _ApplyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Bm, prop, logic, n_prop, n_tmpl, row_map, O_out, zero_fill):
        lib = _lib.load()
        B, O, MS = Bm.shape
        P, HW = prop.shape[1], prop.shape[2]
        out = torch.empty(B, O_out, HW, device=prop.device)
        if B * O_out * HW > 0:
            rc = lib.dmm_assign_apply(_p(Bm), _p(prop), P * HW, B, P, O, MS, HW, _p(n_prop), _p(n_tmpl), _p(row_map),
                                      O_out, int(zero_fill), _p(out), O_out * HW, _stream())
            _lib.check(rc, "dmm_assign_apply")
        ctx.save_for_backward(Bm, prop, logic, n_prop, n_tmpl, row_map)
        return out

    @staticmethod
    def backward(ctx, g_out):
        lib = _lib.load()
        Bm, prop, logic, n_prop, n_tmpl, row_map = ctx.saved_tensors
        B, O, MS = Bm.shape
        P, HW = prop.shape[1], prop.shape[2]
        g_out = g_out.contiguous().float()
        O_out = g_out.shape[1]
        gB = torch.zeros_like(Bm) if ctx.needs_input_grad[0] else None
        gprop = torch.zeros_like(prop) if ctx.needs_input_grad[1] else None
        if B * O * P * HW > 0 and (gB is not None or gprop is not None):
            ws = torch.empty(lib.dmm_assign_apply_bwd_workspace_bytes(B, P, O, HW), device=prop.device, dtype=torch.uint8)
            sel = logic if logic is not None else (Bm != 0).float()
            rc = lib.dmm_assign_apply_bwd(_p(g_out), O_out * HW, _p(prop), P * HW, _p(Bm), _p(sel), B, P, O, MS, HW,
                                          _p(n_prop), _p(n_tmpl), _p(row_map), _p(gB), _p(gprop), _p(ws), ws.numel(),
                                          _stream())
            _lib.check(rc, "dmm_assign_apply_bwd")
        return gB, gprop, None, None, None, None, None, None


class _ApplyRaggedFn(torch.autograd.Function):
    """assign_apply over a RaggedMasks pointer table (the proposal masks carry no gradient on this path)."""

    @staticmethod
    def forward(ctx, Bm, logic, n_tmpl, row_map, ragged, n_prop, O_out, zero_fill):
        lib = _lib.load()
        B, O, MS = Bm.shape
        out = torch.empty(B, O_out, ragged.HW, device=Bm.device)
        if B * O_out * ragged.HW > 0:
            rc = lib.dmm_assign_apply_ptrs(_p(Bm), _p(ragged.ptrs), ragged.aligned16, B, ragged.P, O, MS, ragged.HW,
                                           _p(n_prop), _p(n_tmpl), _p(row_map), O_out, int(zero_fill), _p(out),
                                           O_out * ragged.HW, _stream())
            _lib.check(rc, "dmm_assign_apply_ptrs")
        ctx.save_for_backward(Bm, logic, n_tmpl, row_map, n_prop)
        ctx.ragged = ragged
        return out

    @staticmethod
    def backward(ctx, g_out):
        lib = _lib.load()
        Bm, logic, n_tmpl, row_map, n_prop = ctx.saved_tensors
        rg = ctx.ragged
        B, O, MS = Bm.shape
        g_out = g_out.contiguous().float()
        O_out = g_out.shape[1]
        gB = torch.zeros_like(Bm)
        if B * O * rg.P * rg.HW > 0:
            ws = torch.empty(lib.dmm_assign_apply_bwd_workspace_bytes(B, rg.P, O, rg.HW), device=Bm.device, dtype=torch.uint8)
            sel = logic if logic is not None else (Bm != 0).float()
            rc = lib.dmm_assign_apply_bwd_ptrs(_p(g_out), O_out * rg.HW, _p(rg.ptrs), rg.aligned16, _p(Bm), _p(sel), B, rg.P,
                                               O, MS, rg.HW, _p(n_prop), _p(n_tmpl), _p(row_map), _p(gB), _p(ws), ws.numel(),
                                               _stream())
            _lib.check(rc, "dmm_assign_apply_bwd_ptrs")
        return gB, None, None, None, None, None, None, None


@_op("K4 assign_apply")
def assign_apply(Bm: torch.Tensor, prop: torch.Tensor, logic: Optional[torch.Tensor] = None, n_prop=None, n_tmpl=None,
                 row_map: Optional[torch.Tensor] = None, O_out: Optional[int] = None, zero_fill: bool = True):
    """Bmat [B,O,MS] x prop [B,P,HW] -> out [B,O_out,HW]; row o of problem b lands in row row_map[b,o].
    ``logic`` (the solver's selection mask) restricts the gradient w.r.t. Bmat to the selected entries, exactly the
    entries through which the reference's ``R * logic_mask`` lets gradient flow."""
    Bm = _cuda_f32(Bm, "Bmat")
    B, O, MS = Bm.shape
    if isinstance(prop, (list, tuple)):
        prop = RaggedMasks(prop)
    if isinstance(prop, RaggedMasks):
        dev = Bm.device
        if row_map is not None:
            row_map = torch.as_tensor(row_map, device=dev).to(torch.int32).contiguous()
            assert row_map.shape == (B, O)
        if logic is not None:
            logic = _cuda_f32(logic, "logic")
        n_prop = prop.n_prop if n_prop is None else _counts(n_prop, B, dev)
        return _ApplyRaggedFn.apply(Bm, logic, _counts(n_tmpl, B, dev), row_map, prop, n_prop,
                                    int(O if O_out is None else O_out), bool(zero_fill))
    prop = _cuda_f32(prop, "prop")
    prop = prop.reshape(B, prop.shape[1], -1)
    dev = prop.device
    if row_map is not None:
        row_map = torch.as_tensor(row_map, device=dev).to(torch.int32).contiguous()
        assert row_map.shape == (B, O)
    if O_out is None:
        O_out = O
    if logic is not None:
        logic = _cuda_f32(logic, "logic")
    return _ApplyFn.apply(Bm, prop, logic, _counts(n_prop, B, dev), _counts(n_tmpl, B, dev), row_map, int(O_out),
                          bool(zero_fill))


# ----------------------------------------------------------------------------------------------------------
# K5  ROI mean pooling
# ----------------------------------------------------------------------------------------------------------
class _RoiPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rois, impl, *feats):
        lib = _lib.load()
        N, C = feats[0].shape[:2]
        R = rois.shape[0]
        out = torch.empty(R, 4 * C, device=rois.device)
        Hl = (ctypes.c_int * 4)(*[f.shape[2] for f in feats])
        Wl = (ctypes.c_int * 4)(*[f.shape[3] for f in feats])
        ptrs = (ctypes.c_void_p * 4)(*[f.data_ptr() for f in feats])
        if R * C > 0:
            ws_bytes = 0 if impl == "simt" else lib.dmm_roi_mean_pool_workspace_bytes(Hl, Wl, N, C, R)
            ws = torch.empty(ws_bytes, device=rois.device, dtype=torch.uint8) if ws_bytes else None
            rc = lib.dmm_roi_mean_pool(ptrs, Hl, Wl, N, C, _p(rois), R, _p(out), _p(ws), ws_bytes, ROI_POOL_IMPLS[impl],
                                       _stream())
            _lib.check(rc, "dmm_roi_mean_pool")
        ctx.save_for_backward(rois)
        ctx.shapes = [tuple(f.shape) for f in feats]
        return out

    @staticmethod
    def backward(ctx, g_out):
        lib = _lib.load()
        (rois,) = ctx.saved_tensors
        shapes = ctx.shapes
        N, C = shapes[0][:2]
        R = rois.shape[0]
        g_out = g_out.contiguous().float()
        Hl = (ctypes.c_int * 4)(*[s[2] for s in shapes])
        Wl = (ctypes.c_int * 4)(*[s[3] for s in shapes])
        # deterministic gather (overwrites every element: plain empty buffers) when the library has a plan for these
        # shapes, else the atomic scatter into zeroed buffers
        ws_bytes = lib.dmm_roi_mean_pool_bwd_workspace_bytes(Hl, Wl, N, C, R) if (R * C > 0 and ROI_POOL_BWD_IMPL != "atomic") else 0
        gf = [(torch.empty if ws_bytes else torch.zeros)(s, device=rois.device) for s in shapes]
        ptrs = (ctypes.c_void_p * 4)(*[g.data_ptr() for g in gf])
        if R * C > 0:
            ws = torch.empty(ws_bytes, device=rois.device, dtype=torch.uint8) if ws_bytes else None
            wrote = ctypes.c_int(0)
            rc = lib.dmm_roi_mean_pool_bwd(_p(g_out), Hl, Wl, N, C, _p(rois), R, ptrs, _p(ws), ws_bytes, 2 if ws_bytes else 1,
                                           ctypes.byref(wrote), _stream())
            _lib.check(rc, "dmm_roi_mean_pool_bwd")
        else:
            for g in gf:
                g.zero_()
        return (None, None, *gf)


ROI_POOL_IMPLS = {"auto": 0, "simt": 1, "tc": 2}
ROI_POOL_BWD_IMPL = "auto"          # "atomic" forces the legacy scatter kernel (tests compare the two)


@_op("K5 roi_mean_pool")
def roi_mean_pool(features: Sequence[torch.Tensor], rois: torch.Tensor, impl: str = "auto") -> torch.Tensor:
    """4 levels [N,C,Hl,Wl] at strides 4/8/16/32, rois [R,5] = (batch idx, x1, y1, x2, y2) -> [R, 4*C].
    ``impl``: "auto" (tensor-core contraction per frame where the shapes allow -- C == 128, Wl % 4 == 0 or a tiny level --
    and the SIMT gather kernel for the rest), "simt", "tc" (raise if no level can take the tensor-core path)."""
    assert len(features) == 4, "FeatureExtractor pools 4 levels (feature_extractor.py:13)"
    feats = [_cuda_f32(f, "feature") for f in features]
    N, C = feats[0].shape[:2]
    for f in feats:
        assert f.dim() == 4 and f.shape[0] == N and f.shape[1] == C, [tuple(g.shape) for g in feats]
    rois = _cuda_f32(rois, "rois")
    assert rois.dim() == 2 and rois.shape[1] == 5, rois.shape
    return _RoiPoolFn.apply(rois, impl, *feats)


# ----------------------------------------------------------------------------------------------------------
# K6 / K7  rows after the layer: decoder mask-input pyramid, merged label map, hard-IoU metric
# ----------------------------------------------------------------------------------------------------------
def _bohw(t: torch.Tensor, name: str, B: int, O: int, H: int, W: int) -> torch.Tensor:
    """[B,O,H,W] or [B,O,HW] fp32 CUDA whose (O,H,W) block is dense; the batch stride may be larger (a view)."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"dmm_net_b200: `{name}` must be a CUDA tensor (there is no CPU fallback)")
    assert t.shape[0] == B and t.shape[1] == O and t.numel() == B * O * H * W, (name, tuple(t.shape), (B, O, H, W))
    t = t.float().reshape(B, O, H * W)
    if t.stride(2) != 1 or t.stride(1) != H * W:
        t = t.contiguous()
    return t


class _PyramidFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prev, ref, init, H, W, L):
        lib = _lib.load()
        B, O = init.shape[:2]
        sizes = []
        for k in range(L):
            hk, wk = ctypes.c_int(), ctypes.c_int()
            _lib.check(lib.dmm_mask_pyramid_level_size(H, W, k, ctypes.byref(hk), ctypes.byref(wk)), "dmm_mask_pyramid_level_size")
            sizes.append((hk.value, wk.value))
        outs = [torch.empty(O, B, 3, hk, wk, device=init.device) for hk, wk in sizes]
        if B * O * H * W > 0 and L > 0:
            ptrs = (ctypes.c_void_p * L)(*[o.data_ptr() for o in outs])
            rc = lib.dmm_mask_pyramid(_p(prev), prev.stride(0), _p(ref), ref.stride(0), _p(init), init.stride(0), B, O, H, W,
                                      L, ptrs, _stream())
            _lib.check(rc, "dmm_mask_pyramid")
        ctx.save_for_backward(prev, ref, init)
        ctx.dims = (H, W, L)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *g_outs):
        lib = _lib.load()
        prev, ref, init = ctx.saved_tensors
        H, W, L = ctx.dims
        B, O = init.shape[:2]
        gs = [None if g is None else g.contiguous().float() for g in g_outs]
        grads = [torch.empty(B, O, H * W, device=init.device) if ctx.needs_input_grad[i] else None for i in range(3)]
        if B * O * H * W > 0 and any(g is not None for g in grads):
            ptrs = (ctypes.c_void_p * max(L, 1))(*[None if g is None else g.data_ptr() for g in gs])
            rc = lib.dmm_mask_pyramid_bwd(ptrs, _p(prev), prev.stride(0), _p(ref), ref.stride(0), _p(init), init.stride(0),
                                          B, O, H, W, L, _p(grads[0]), _p(grads[1]), _p(grads[2]), _stream())
            _lib.check(rc, "dmm_mask_pyramid_bwd")
        return grads[0], grads[1], grads[2], None, None, None


@_op("K6 mask_pyramid")
def mask_pyramid(prev_mask: torch.Tensor, ref_mask: torch.Tensor, init_pred: torch.Tensor, n_levels: int = 4):
    """The decoder's mask inputs for EVERY object in one pass (trainer.py:256-263, evaluator.py:187-194).

    prev_mask / ref_mask / init_pred: [B,O,H,W] (or [B,O,HW] for the first two, as the reference holds them).
    Returns ``n_levels`` tensors [O,B,3,hk,wk], finest first (window 4, 8, ...): ``levels[k][t]`` is what the reference
    calls ``mask_lstm`` (before its ``reversed``) entry k of object t.  Differentiable w.r.t. all three inputs."""
    assert init_pred.dim() == 4, "init_pred_inst is [B,O,H,W]"
    B, O, H, W = init_pred.shape
    init = _bohw(init_pred, "init_pred_inst", B, O, H, W)
    prev = _bohw(prev_mask, "prev_mask", B, O, H, W)
    ref = _bohw(ref_mask, "ref_mask", B, O, H, W)
    return list(_PyramidFn.apply(prev, ref, init, int(H), int(W), int(n_levels)))


@_op("K7 merge_labels")
def merge_labels(outs: torch.Tensor, n_valid=None) -> torch.Tensor:
    """outs [B,O,HW] (or [B,O,H,W]) sigmoid masks -> uint8 label map [B,HW]: 0 = background, t+1 = object t
    (evaluator.py:139-145, for all videos of the batch at once; ``n_valid[b]`` = tplt_valid_batch[b].sum())."""
    lib = _lib.load()
    if not outs.is_cuda:
        raise RuntimeError("dmm_net_b200: `outs` must be a CUDA tensor (there is no CPU fallback)")
    B, O = outs.shape[:2]
    o3 = outs.float().reshape(B, O, -1)
    if o3.stride(2) != 1 or (O > 1 and o3.stride(1) != o3.shape[2]):
        o3 = o3.contiguous()
    HW = o3.shape[2]
    label = torch.empty(B, HW, dtype=torch.uint8, device=outs.device)
    if B * HW > 0:
        rc = lib.dmm_merge_labels(_p(o3), o3.stride(0), B, O, HW, _p(_counts(n_valid, B, outs.device)), _p(label), _stream())
        _lib.check(rc, "dmm_merge_labels")
    return label


def hard_iou_mean(y_mask: torch.Tensor, pred: torch.Tensor, valid: torch.Tensor) -> torch.Tensor:
    """trainer.py:189-196 / :296-300: mean hard IoU over the valid templates.  The [B*O] row-paired IoU is K1's row-wise
    entry; the [B,O] masking and the scalar mean are left to torch (50 numbers)."""
    B, O = valid.shape
    iou = mask_iou_rowwise(y_mask.reshape(B * O, -1), pred.reshape(B * O, -1)).view(B, O) * valid.float()
    n = valid.sum()
    return torch.where(n > 0, iou.sum() / (n + 1e-6), iou.sum() * 0)


# ----------------------------------------------------------------------------------------------------------
# K8 / K9  rows before the layer: proposal paste (+ bit rows + tight boxes), box NMS
# ----------------------------------------------------------------------------------------------------------
@_op("K8 paste_masks")
def paste_masks(masks: torch.Tensor, boxes: torch.Tensor, im_h: int, im_w: int, thresh: float = 0.5, padding: int = 1,
                want_pasted: bool = True, want_bits: bool = False, want_tight: bool = True):
    """masks [N,1,M,M] or [N,M,M] soft, boxes [N,4] xyxy -> dict(pasted [N,im_h,im_w], bits [N,words] int32, tight [N,4]
    int64), every proposal of a batch of frames in one launch (masker.py:91-206).  ``pasted`` rows are what K1 / K4 read;
    ``bits`` rows are the packed-K1 format."""
    lib = _lib.load()
    masks, boxes = _cuda_f32(masks, "masks"), _cuda_f32(boxes, "boxes")
    N, M = masks.shape[0], masks.shape[-1]
    assert masks.numel() == N * M * M and boxes.shape == (N, 4), (tuple(masks.shape), tuple(boxes.shape))
    dev = masks.device
    pasted = torch.empty(N, im_h, im_w, device=dev) if want_pasted else None
    bits = torch.empty(N, packed_words(im_h * im_w), dtype=torch.int32, device=dev) if want_bits else None
    tight = torch.empty(N, 4, dtype=torch.int64, device=dev) if want_tight else None
    out = {"pasted": pasted, "bits": bits, "tight": tight}
    for s in range(0, N, 65535):
        e = min(N, s + 65535)
        ws = torch.empty(max(lib.dmm_paste_masks_workspace_bytes(e - s), 256), device=dev, dtype=torch.uint8)
        sl = lambda t: None if t is None else t[s:e]
        rc = lib.dmm_paste_masks(_p(masks[s:e]), _p(boxes[s:e]), e - s, M, int(padding), int(im_h), int(im_w), float(thresh),
                                 _p(sl(pasted)), _p(sl(bits)), _p(sl(tight)), _p(ws), ws.numel(), _stream())
        _lib.check(rc, "dmm_paste_masks")
    return out


class _PasteApplyFn(torch.autograd.Function):
    """K10 with a gradient for the assignment (the mask-head outputs are constants: offline proposals in the reference)."""

    @staticmethod
    def forward(ctx, Bm, logic, masks, boxes, src_index, n_prop, n_tmpl, row_map, dims):
        lib = _lib.load()
        im_h, im_w, O_out, zero_fill, padding = dims
        B, O, MS = Bm.shape
        P, M = src_index.shape[1], masks.shape[-1]
        out = torch.empty(B, O_out, im_h, im_w, device=Bm.device)
        step = max(1, 65535 // max(O_out, 1))
        for s in range(0, B, step):
            e = min(B, s + step)
            sl = lambda t: None if t is None else t[s:e]
            rc = lib.dmm_paste_apply(_p(Bm[s:e]), _p(masks), _p(boxes), _p(src_index[s:e]), e - s, P, O, MS, M, int(padding),
                                     int(im_h), int(im_w), _p(sl(n_prop)), _p(sl(n_tmpl)), _p(sl(row_map)), O_out,
                                     int(zero_fill), _p(out[s:e]), O_out * im_h * im_w, _stream())
            _lib.check(rc, "dmm_paste_apply")
        ctx.save_for_backward(Bm, logic, masks, boxes, src_index, n_prop, n_tmpl, row_map)
        ctx.dims = dims
        return out

    @staticmethod
    def backward(ctx, g_out):
        lib = _lib.load()
        Bm, logic, masks, boxes, src_index, n_prop, n_tmpl, row_map = ctx.saved_tensors
        im_h, im_w, O_out, zero_fill, padding = ctx.dims
        B, O, MS = Bm.shape
        P, M = src_index.shape[1], masks.shape[-1]
        g_out = g_out.contiguous().float()
        sel = logic if logic is not None else (Bm != 0).float()
        gB = torch.empty_like(Bm)
        if B * O > 0:
            rc = lib.dmm_paste_apply_bwd(_p(g_out), O_out * im_h * im_w, _p(sel), _p(masks), _p(boxes), _p(src_index), B, P, O, MS, M,
                                         int(padding), int(im_h), int(im_w), _p(n_prop), _p(n_tmpl), _p(row_map), O_out, _p(gB),
                                         _stream())
            _lib.check(rc, "dmm_paste_apply_bwd")
        return gB, None, None, None, None, None, None, None, None


@_op("K10 paste_apply")
def paste_apply(Bm: torch.Tensor, masks: torch.Tensor, boxes: torch.Tensor, src_index: torch.Tensor, im_h: int, im_w: int,
                n_prop=None, n_tmpl=None, row_map: Optional[torch.Tensor] = None, O_out: Optional[int] = None,
                zero_fill: bool = True, padding: int = 1, logic: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused paste + assignment apply ("lazy paste", K10): out[b, row(o)] = sum_p Bm[b,o,p] * paste(masks[src_index[b,p]]).

    Bm [B,O,MS]; masks [Nsrc,1,M,M] / [Nsrc,M,M] mask-head outputs and boxes [Nsrc,4] of ALL detections; src_index [B,P]
    int32 = which detection sits behind column p of problem b (-1: none) -- the NMS keep list.  Returns [B,O_out,im_h,im_w],
    bit-identical to ``assign_apply(Bm, paste_masks(...)["pasted"] gathered by src_index)`` without ever writing or reading
    the P pasted masks.  Differentiable w.r.t. ``Bm`` (``logic`` = the solver's selection mask restricts the gradient to the
    selected entries, as in ``assign_apply``); the mask-head outputs are treated as constants."""
    Bm = _cuda_f32(Bm, "Bmat")
    masks, boxes = _cuda_f32(masks, "masks"), _cuda_f32(boxes, "boxes")
    B, O, MS = Bm.shape
    dev = Bm.device
    src_index = torch.as_tensor(src_index, device=dev).to(torch.int32).contiguous()
    P = src_index.shape[1]
    assert src_index.shape == (B, P) and MS >= P and boxes.shape == (masks.shape[0], 4), (src_index.shape, Bm.shape, boxes.shape)
    if row_map is not None:
        row_map = torch.as_tensor(row_map, device=dev).to(torch.int32).contiguous()
        assert row_map.shape == (B, O)
    if logic is not None:
        logic = _cuda_f32(logic, "logic")
    O_out = int(O if O_out is None else O_out)
    return _PasteApplyFn.apply(Bm, logic, masks.detach(), boxes.detach(), src_index, _counts(n_prop, B, dev), _counts(n_tmpl, B, dev),
                               row_map, (int(im_h), int(im_w), O_out, bool(zero_fill), int(padding)))


@_op("K9 box_nms")
def box_nms(boxes: torch.Tensor, scores: torch.Tensor, thresh: float, max_keep: int = 0, n_boxes=None):
    """boxes [F,n,4] (or [n,4]), scores [F,n] (or [n]) -> (keep [F,n] int64 kept indices in score order, -1 padded; n_keep
    [F] int32).  Greedy NMS with the legacy +1 widths, one CTA per frame (boxlist_ops.py:15-29)."""
    lib = _lib.load()
    single = boxes.dim() == 2
    boxes = _cuda_f32(boxes, "boxes")
    scores = _cuda_f32(scores, "scores")
    if single:
        boxes, scores = boxes[None], scores[None]
    F_, n = scores.shape
    assert boxes.shape == (F_, n, 4), (tuple(boxes.shape), tuple(scores.shape))
    keep = torch.empty(F_, n, dtype=torch.int64, device=boxes.device)
    n_keep = torch.zeros(F_, dtype=torch.int32, device=boxes.device)
    if F_ > 0:
        rc = lib.dmm_box_nms(_p(boxes), _p(scores), _p(_counts(n_boxes, F_, boxes.device)), F_, n, float(thresh),
                             int(max_keep), _p(keep), _p(n_keep), _stream())
        _lib.check(rc, "dmm_box_nms")
    return keep, n_keep


# ----------------------------------------------------------------------------------------------------------
# the fused layer over a batch of problems
# ----------------------------------------------------------------------------------------------------------
_SIDE_STREAMS = {}


def _side_streams(dev: torch.device):
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
    return _SIDE_STREAMS[key]


@_op("cost_and_solve (K2+K1+K3)")
def cost_and_solve(prop_feat, prop_mask, tmpl_feat, tmpl_mask, prop_score, *, max_iter: int, proj_iter: int, lr: float,
                   score_weight: float, is_test: bool, n_prop=None, n_tmpl=None, chunks: Optional[int] = None,
                   k1_events: Optional[list] = None):
    """Inference path (no autograd) of cost-build + solve for a dense batch: K2 -> K1(+finalize/mix) -> K3.

    Large batches are cut into chunks that run on two side streams, staggered so that the HBM-bound K1 launches stay
    back to back while the latency-bound cosine of the next chunk and the solver of the previous chunk run underneath
    them (the TMA K1 kernel leaves registers and shared memory for one solver CTA per SM).  Every output is allocated
    on the caller's stream, which waits for both side streams before returning, so results and memory lifetimes are
    exactly those of the single-stream order.  Off by default (it measured slower, see below); ``chunks=N`` or
    ``DMM_PIPELINE=1`` enables it.
    ``k1_events``: optional list that receives one (start, end) CUDA-event pair per K1 launch, recorded on the stream
    the kernel runs on (bench.py's roofline timing).
    Returns dict(sim, cos, iou, R, Bmat, logic, X_final, match_score, det_score, n_list)."""
    import os
    lib = _lib.load()
    if tmpl_feat.dim() == 3:
        tmpl_feat = tmpl_feat.unsqueeze(1)
    prop_feat, tmpl_feat = _cuda_f32(prop_feat, "prop_feat"), _cuda_f32(tmpl_feat, "tmpl_feat")
    prop_mask, tmpl_mask = _cuda_f32(prop_mask, "prop_mask"), _cuda_f32(tmpl_mask, "tmpl_mask")
    prop_score = _cuda_f32(prop_score, "prop_score")
    B, P = prop_mask.shape[:2]
    O = tmpl_mask.shape[1]
    T, D = tmpl_feat.shape[1], tmpl_feat.shape[3]
    HW = 1
    for dsz in prop_mask.shape[2:]:
        HW *= int(dsz)
    dev = prop_mask.device
    MS = pad_cols(P, O)
    check_solver_shape(P, O)
    n_prop, n_tmpl = _counts(n_prop, B, dev), _counts(n_tmpl, B, dev)
    w = float(score_weight)
    new = lambda *shape: torch.empty(*shape, device=dev)
    out = {"cos": new(B, O, P), "iou": new(B, O, P), "sim": new(B, O, P), "R": new(B, O, MS), "Bmat": new(B, O, MS),
           "logic": new(B, O, MS), "X_final": new(B, O, MS), "match_score": new(B, O), "det_score": new(B, O),
           "n_list": torch.empty(B, device=dev, dtype=torch.int32)}
    if B * O == 0 or P == 0:
        for v in out.values():
            v.zero_()
        return out
    if chunks is None:
        # measured on B200 (profiles/README.md): 4 staggered chunks of 256 problems run 4.34 ms vs 4.14 ms for one launch
        # of 1024 -- smaller K1 launches pay more tail, and the co-running solver slows K1 more than it hides.  So the
        # overlap is opt-in (DMM_PIPELINE=1); the default is one chunk per 65535 problems.
        chunks = (4 if B >= 512 else (2 if B >= 128 else 1)) if os.environ.get("DMM_PIPELINE", "0") == "1" else 1
    chunks = max(1, min(int(chunks), B))
    chunks = max(chunks, (B + 65534) // 65535)
    bounds = [(B * i // chunks, B * (i + 1) // chunks) for i in range(chunks)]
    ws = [torch.empty(max(lib.dmm_mask_iou_workspace_bytes(e - s, P, O, max(HW, 1), 0), 256), device=dev, dtype=torch.uint8)
          for s, e in bounds]
    sl = lambda t, s, e: None if t is None else t[s:e]

    def run_chunk(i, k1_after=None):
        s, e = bounds[i]
        nb = e - s
        st = _stream()
        rc = lib.dmm_cosine_pairwise(_p(tmpl_feat[s:e]), _p(prop_feat[s:e]), nb, T, P, O, D, _p(sl(n_prop, s, e)),
                                     _p(sl(n_tmpl, s, e)), 1e-8, _p(out["cos"][s:e]), st)
        _lib.check(rc, "dmm_cosine_pairwise")
        if k1_after is not None:
            torch.cuda.current_stream().wait_event(k1_after)        # keep the HBM-bound kernels back to back, not concurrent
        if k1_events is not None:
            t0 = torch.cuda.Event(enable_timing=True)
            t0.record()
        rc = lib.dmm_mask_iou_pairwise(_p(prop_mask[s:e]), P * HW, _p(tmpl_mask[s:e]), O * HW, None, 0, nb, P, O, HW,
                                       _p(sl(n_prop, s, e)), _p(sl(n_tmpl, s, e)), _p(out["iou"][s:e]), None,
                                       _p(out["cos"][s:e]), float(1 - w), w, _p(out["sim"][s:e]), None, _p(ws[i]),
                                       ws[i].numel(), st)
        _lib.check(rc, "dmm_mask_iou_pairwise")
        ev = torch.cuda.Event(enable_timing=k1_events is not None)
        ev.record()
        if k1_events is not None:
            k1_events.append((t0, ev))
        rc = lib.dmm_relax_solve(_p(out["sim"][s:e]), _p(prop_score[s:e]), nb, P, O, _p(sl(n_prop, s, e)),
                                 _p(sl(n_tmpl, s, e)), int(max_iter), int(proj_iter), float(lr), 1, 1, int(bool(is_test)),
                                 _p(out["R"][s:e]), _p(out["X_final"][s:e]), _p(out["Bmat"][s:e]), _p(out["logic"][s:e]),
                                 _p(out["match_score"][s:e]), _p(out["det_score"][s:e]), _p(out["n_list"][s:e]), None, None,
                                 None, st)
        _lib.check(rc, "dmm_relax_solve")
        return ev

    if chunks == 1:
        run_chunk(0)
        return out
    main = torch.cuda.current_stream()
    start = torch.cuda.Event()
    start.record(main)
    side = _side_streams(dev)
    prev = None
    for i in range(chunks):
        st = side[i & 1]
        if i < 2:
            st.wait_event(start)
        with torch.cuda.stream(st):
            prev = run_chunk(i, prev)
    for st in side:
        done = torch.cuda.Event()
        done.record(st)
        main.wait_event(done)
    return out


def match_batch(prop_feat: torch.Tensor, prop_mask: torch.Tensor, tmpl_feat: torch.Tensor, tmpl_mask: torch.Tensor,
                prop_score: torch.Tensor, targets: Optional[torch.Tensor] = None, *, max_iter: int, proj_iter: int,
                lr: float, score_weight: float, is_test: bool, n_prop=None, n_tmpl=None, row_map=None,
                O_out: Optional[int] = None, apply: bool = True):
    """MatchModel.forward for B problems in five launches (cosine, IoU, IoU-finalize+mix, solve+head, apply).

    prop_feat [B,P,D], prop_mask [B,P,H,W], tmpl_feat [B,T,O,D] (or [B,O,D]), tmpl_mask [B,O,H,W], prop_score [B,P],
    targets [B,O,H,W] or None.  Returns dict(full_outmask [B,O_out,H,W], match_score, det_score [B,O], sim, R, Bmat,
    logic, n_list, cost_loss [B] or None).
    """
    if tmpl_feat.dim() == 3:
        tmpl_feat = tmpl_feat.unsqueeze(1)
    if isinstance(prop_mask, (list, tuple)):
        prop_mask = RaggedMasks(prop_mask)                      # per-video tensors used in place (pointer table)
    if isinstance(prop_mask, RaggedMasks):
        B, P = prop_mask.B, prop_mask.P
        n_prop = prop_mask.n_prop if n_prop is None else n_prop
        assert prop_feat.shape[1] == P and prop_score.shape[1] == P, "features / scores must be padded to max P_b"
    else:
        B, P = prop_mask.shape[:2]
    O = tmpl_mask.shape[1]
    H, W = tmpl_mask.shape[-2:]
    dev = tmpl_mask.device
    n_prop, n_tmpl = _counts(n_prop, B, dev), _counts(n_tmpl, B, dev)
    needs_grad = torch.is_grad_enabled() and any(t is not None and torch.is_tensor(t) and t.requires_grad
                                                 for t in (prop_feat, tmpl_feat, prop_score, prop_mask))
    if not needs_grad and targets is None and not isinstance(prop_mask, RaggedMasks):
        out = cost_and_solve(prop_feat, prop_mask, tmpl_feat, tmpl_mask, prop_score, max_iter=max_iter, proj_iter=proj_iter,
                             lr=lr, score_weight=score_weight, is_test=is_test, n_prop=n_prop, n_tmpl=n_tmpl)
        full = None
        if apply:
            full = assign_apply(out["Bmat"], prop_mask, out["logic"], n_prop, n_tmpl, row_map, O_out).view(B, -1, H, W)
        out.update(full_outmask=full, cost_loss=None)
        return out
    cos = cosine_pairwise(tmpl_feat, prop_feat, n_prop, n_tmpl)                       # K2
    w = float(score_weight)
    if torch.is_grad_enabled() and cos.requires_grad:
        r = mask_iou_pairwise(prop_mask, tmpl_mask, targets, n_prop, n_tmpl)           # K1 (both template sets in one pass)
        # sim = cos*(1-w) + iou*w with separate fp32 roundings (match_model.py:90); kept in torch here so that
        # autograd sees cos.  IoU carries no gradient (match_helper.py:20 runs under no_grad).
        sim = cos * (1 - w) + r["iou"] * w
    else:
        r = mask_iou_pairwise(prop_mask, tmpl_mask, targets, n_prop, n_tmpl, cos=cos, w_cos=1 - w, w_iou=w)
        sim = r["sim"]                                                                 # mixed in K1's finalize kernel
    cost_loss = None
    if targets is not None:
        gt = relax_solve(r["iou2"], None, n_prop, n_tmpl, 0, 0, 0.0, True, False, True)[0]   # greedy one-hot (match_helper.py:44)
        d2 = (cos - gt) ** 2
        if n_prop is None and n_tmpl is None:
            cost_loss = d2.mean(dim=(1, 2))                                           # F.mse_loss per problem
        else:
            npv = n_prop if n_prop is not None else torch.full((B,), P, device=dev, dtype=torch.int32)
            ntv = n_tmpl if n_tmpl is not None else torch.full((B,), O, device=dev, dtype=torch.int32)
            valid = (torch.arange(P, device=dev)[None, None, :] < npv[:, None, None]) & \
                    (torch.arange(O, device=dev)[None, :, None] < ntv[:, None, None])
            cost_loss = (d2 * valid).sum(dim=(1, 2)) / (npv * ntv).clamp(min=1).float()
    R, Bm, ms, ds, Xf, logic, n_list = relax_solve(sim, prop_score, n_prop, n_tmpl, max_iter, proj_iter, lr, True, True,
                                                   is_test)                            # K3
    full = None
    if apply:
        full = assign_apply(Bm, prop_mask, logic, n_prop, n_tmpl, row_map, O_out).view(B, -1, H, W)   # K4
    return {"full_outmask": full, "match_score": ms, "det_score": ds, "sim": sim, "R": R, "Bmat": Bm, "logic": logic,
            "n_list": n_list, "cost_loss": cost_loss, "cos": cos, "iou": r["iou"], "X_final": Xf}


# ----------------------------------------------------------------------------------------------------------
# host-buffer entry: masks held in HOST memory cross PCIe as bits
# ----------------------------------------------------------------------------------------------------------
_PINNED = {}


def host_threads() -> int:
    """Threads for host-side packing: the cgroup CPU quota when there is one, else all cores.
    (Round-1 GPU box: 128 logical CPUs under cpu.max = 16 CPUs.  Short bursts pack at 340 GB/s with 64 threads, but
    sustained the quota throttles: e2e 3.8k matches/s with 16 threads, 3.0k with 32, 1.5k with 64, 0.8k with 96.)"""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(2, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, n)


def _pinned(key, shape, dtype):
    t = _PINNED.get(key)
    if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
        from . import hostmem
        with hostmem.prefer_node(hostmem.gpu_numa_node(torch.cuda.current_device())):   # staging next to the GPU's socket
            t = torch.empty(shape, dtype=dtype).pin_memory()
        _PINNED[key] = t
    return t


_HOST_SPLIT = {}          # (P, O, HW) -> best measured seconds per problem of the two routes (pack, dma)
_HOST_PENDING = {}        # (P, O, HW) -> measurements of the previous call, not yet folded in
_STAGE_TURN, _STAGE_EVENT = {}, {}   # double-buffered pinned staging: next slot, and the H2D-done event of each slot
_RAW_TURN, _RAW_STAGE = {}, {}       # raw route: next device slot; (key, slot) -> [prop buffer, tmpl buffer, last-read event]


def _fold_route_measurement(key):
    """Exponential average of the measured seconds per problem of each route (under several ranks per host the packing
    rate depends on what the other ranks are doing, so the estimate has to follow it); a route's measurement only counts
    when it carried enough problems for its fixed costs not to dominate."""
    pend = _HOST_PENDING.pop(key, None)
    if pend is None:
        return
    e0, ev_dma, t_host, nraw, npk, B = pend
    if not ev_dma.query():
        _HOST_PENDING[key] = pend
        return
    old = _HOST_SPLIT.get(key)
    new_pack = t_host / npk if npk >= max(1, B // 8) else None
    new_dma = e0.elapsed_time(ev_dma) * 1e-3 / nraw if nraw >= max(1, B // 8) else None
    if old is None:
        if new_pack is not None and new_dma is not None:
            _HOST_SPLIT[key] = (new_pack, new_dma)
        return
    # one slow sample (a copy queued behind another rank's traffic, a throttled packing burst) must not swing the split:
    # a sample counts at most as 3x the current estimate
    mix = lambda o, n: o if n is None else 0.5 * (o + min(n, 3.0 * o))
    _HOST_SPLIT[key] = (mix(old[0], new_pack), mix(old[1], new_dma))



def match_batch_host(prop_feat, prop_mask, tmpl_feat, tmpl_mask, prop_score, *, device="cuda", **kw):
    """``_match_batch_host`` with ``device`` made current for the whole call (streams, events, allocations and the
    C-ABI launches all follow the current device)."""
    dev = torch.device(device)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    with torch.cuda.device(dev):
        return _match_batch_host(prop_feat, prop_mask, tmpl_feat, tmpl_mask, prop_score, device=dev, **kw)


def _match_batch_host(prop_feat, prop_mask, tmpl_feat, tmpl_mask, prop_score, *, max_iter: int, proj_iter: int, lr: float,
                      score_weight: float, is_test: bool, device="cuda", threads: Optional[int] = None, n_prop=None,
                      n_tmpl=None, raw_fraction: Optional[float] = None):
    """cost-build + solve for a batch whose inputs live in HOST memory (CPU fp32 tensors, pinned or not).

    Two routes feed the device at the same time:
    * **packed**: the IoU needs only the thresholded bits, so the host cores pack the masks (std::thread + AVX-512/AVX2, memory speed)
      and only bits cross PCIe: 0.86 MB instead of 27.5 MB of masks per match at the headline size (K1 on packed rows);
    * **raw**: while the cores are packing, the copy engine -- otherwise idle -- DMAs the fp32 masks of the first
      ``raw_fraction`` of the problems straight from pinned memory (K1 on fp32 rows, the TMA kernel).
    The split is chosen so that both routes finish together: ``raw_fraction=None`` tracks the measured seconds per problem
    of each route (host wall clock for the packing, CUDA events for the DMA) across calls; 0 disables the raw route
    (also when the masks are not pinned: the copy would not be asynchronous).  Both routes produce the same integer
    counts, so ``iou`` / ``sim`` are bit-identical whichever route a problem takes.
    Device side: K2 cosine -> K1 (fp32 rows | packed rows, +finalize/mix) -> K3 solver+head.  Returns the dict of
    ``cost_and_solve`` (device tensors) plus ``inputs_consumed``: a CUDA event that completes when the LAST host->device
    copy reading the caller's buffers has finished.  **Buffer-reuse contract:** the call returns while those copies may
    still be in flight (that is what lets consecutive calls overlap); a producer that refills the same pinned buffers in
    place must ``out["inputs_consumed"].synchronize()`` first.  The soft masks stay on the host: the assignment-apply
    (which needs <= O selected rows per problem) is left to the caller / ``match_batch``."""
    import time
    lib = _lib.load()
    for t in (prop_feat, prop_mask, tmpl_feat, tmpl_mask, prop_score):
        assert not t.is_cuda and t.dtype == torch.float32, "match_batch_host takes CPU fp32 tensors"
    if tmpl_feat.dim() == 3:
        tmpl_feat = tmpl_feat.unsqueeze(1)
    dev = torch.device(device)
    B, P = prop_mask.shape[:2]
    O = tmpl_mask.shape[1]
    HW = 1
    for dsz in prop_mask.shape[2:]:
        HW *= int(dsz)
    words = packed_words(HW)
    th = host_threads() if threads is None else int(threads)
    prop_mask = prop_mask.reshape(B, P, HW)
    tmpl_mask = tmpl_mask.reshape(B, O, HW)
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), P, O, HW)   # per device: one thread per GPU
    can_raw = prop_mask.is_pinned() and tmpl_mask.is_pinned() and prop_mask.is_contiguous() and tmpl_mask.is_contiguous()
    _fold_route_measurement(key)
    auto_split = raw_fraction is None
    if raw_fraction is None:
        est = _HOST_SPLIT.get(key)
        # first call: nominal 100 GB/s of packing against 50 GB/s of PCIe; afterwards the measured rates
        t_pack, t_dma = est if est else (1.0, 2.0)
        # The raw route should finish a little BEFORE the packing does: the bit rows cross PCIe behind it on the same copy
        # engine and the device work of the call starts only when both have landed -> weigh the packing time by 0.85.
        # Both routes always carry at least an eighth of the batch, so that both keep being measured.
        raw_fraction = min(0.875, max(0.125, 0.85 * t_pack / (0.85 * t_pack + t_dma)))
    nraw = int(round(B * float(raw_fraction))) if can_raw and B >= 4 else 0
    if auto_split and B >= 32:
        nraw = (nraw + 2) // 4 * 4                               # few distinct staging sizes: the allocator reuses its blocks
    nraw = max(0, min(nraw, B))
    main = torch.cuda.current_stream(dev)
    to_dev = lambda t: t.to(dev, non_blocking=True)
    t_call = time.perf_counter()
    pm_raw = tm_raw = None
    ev_dma = None
    raw_slot = None
    if nraw > 0:
        # Device staging of the raw route: two persistent slots of full-batch size, alternated call by call.  (Fresh
        # tensors per call on the side stream went through the caching allocator with a size that follows the moving
        # split: every new size is a cudaMalloc -- milliseconds, and a device-wide synchronisation -- inside the call.)
        # Slot k is overwritten only after the K1 launch that read it two calls ago has finished (event on `main`).
        raw_slot = _RAW_TURN.get(key, 0)
        _RAW_TURN[key] = raw_slot ^ 1
        st = _RAW_STAGE.get((key, raw_slot))
        copy_stream = _side_streams(dev)[0]
        if st is None or st[0].shape[0] < B:
            copy_stream.synchronize()                          # a larger batch than before: no DMA may still target the old slot
            st = [torch.empty((B, P, HW), device=dev), torch.empty((B, O, HW), device=dev), None]
            _RAW_STAGE[(key, raw_slot)] = st
        # copy_stream is not ordered after `main`: the DMA may start while the previous
        with torch.cuda.stream(copy_stream):                   # call's kernels still run (the other slot, read-only source)
            if st[2] is not None:
                copy_stream.wait_event(st[2])
            pm_raw, tm_raw = st[0][:nraw], st[1][:nraw]
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
            pm_raw.copy_(prop_mask[:nraw], non_blocking=True)
            tm_raw.copy_(tmpl_mask[:nraw], non_blocking=True)
            ev_dma = torch.cuda.Event(enable_timing=True)
            ev_dma.record()
    npk = B - nraw
    t_host = 0.0
    t_raw_enqueued = time.perf_counter() - t_call
    if npk > 0:
        # Pinned staging buffers, sized for the whole batch (the split moves from call to call; re-pinning would cost more)
        # and double-buffered: callers may issue the next call before this call's H2D has drained, and the packer must
        # not overwrite bits that are still being copied.  The event of a slot is waited for before the slot is reused.
        slot = _STAGE_TURN.get(key, 0)
        _STAGE_TURN[key] = slot ^ 1
        pb_full = _pinned(("pb", key[0], slot), (B, P, words), torch.int32)
        tb_full = _pinned(("tb", key[0], slot), (B, O, words), torch.int32)
        busy = _STAGE_EVENT.get((key, slot))
        if busy is not None:
            busy.synchronize()
        t0 = time.perf_counter()
        pb, tb = pb_full[:npk], tb_full[:npk]
        rc = lib.dmm_host_pack_masks2(_VP(prop_mask[nraw:].data_ptr()), npk * P, _VP(pb.data_ptr()),
                                      _VP(tmpl_mask[nraw:].data_ptr()), npk * O, _VP(tb.data_ptr()), HW, th)   # one thread team
        _lib.check(rc, "dmm_host_pack_masks2")
        t_host = time.perf_counter() - t0
        pbd, tbd = to_dev(pb), to_dev(tb)
        done = torch.cuda.Event()
        done.record(main)
        _STAGE_EVENT[(key, slot)] = done
    pf, tf, sc = to_dev(prop_feat.contiguous()), to_dev(tmpl_feat.contiguous()), to_dev(prop_score.contiguous())
    n_prop, n_tmpl = _counts(n_prop, B, dev), _counts(n_tmpl, B, dev)
    sl = lambda t, a, b: None if t is None else t[a:b]
    consumed = torch.cuda.Event()
    consumed.record(main)                                        # every H2D copy on main is queued (re-recorded below when
    with torch.no_grad():                                        #   the raw route is in use)
        cos = cosine_pairwise(tf, pf, n_prop, n_tmpl)
        w = float(score_weight)
        parts = [None, None]
        if npk > 0:                                              # packed rows first: they do not wait for the raw DMA
            parts[1] = mask_iou_pairwise_packed(pbd, tbd, None, sl(n_prop, nraw, B), sl(n_tmpl, nraw, B), cos=cos[nraw:],
                                                w_cos=1 - w, w_iou=w)
        if nraw > 0:
            main.wait_event(ev_dma)
            consumed = torch.cuda.Event()
            consumed.record(main)                                # after the wait on the raw DMA and every copy queued on main
            parts[0] = mask_iou_pairwise(pm_raw, tm_raw, None, sl(n_prop, 0, nraw), sl(n_tmpl, 0, nraw), cos=cos[:nraw],
                                         w_cos=1 - w, w_iou=w)
            read_done = torch.cuda.Event()
            read_done.record(main)                               # the slot may be overwritten once this K1 has run
            _RAW_STAGE[(key, raw_slot)][2] = read_done
        parts = [q for q in parts if q is not None]
        if len(parts) == 1:
            iou, sim = parts[0]["iou"], parts[0]["sim"]
        else:
            iou = torch.cat([q["iou"] for q in parts], 0)
            sim = torch.cat([q["sim"] for q in parts], 0)
        R, Bm, ms, ds, Xf, logic, n_list = relax_solve(sim, sc, n_prop, n_tmpl, max_iter, proj_iter, lr, True, True,
                                                       bool(is_test))
    if nraw > 0 and npk > 0:
        # route measurements of this call; folded into the estimate at the start of the next call (the DMA may still be
        # running here when it is the longer route -- never block on it)
        _HOST_PENDING[key] = (e0, ev_dma, t_host, nraw, npk, B)
    h2d = 4 * (prop_feat.numel() + tmpl_feat.numel() + prop_score.numel()) + 4 * npk * (P + O) * words + \
        4 * nraw * (P + O) * HW
    return {"cos": cos, "iou": iou, "sim": sim, "R": R, "Bmat": Bm, "logic": logic, "X_final": Xf,
            "match_score": ms, "det_score": ds, "n_list": n_list, "h2d_bytes": h2d, "inputs_consumed": consumed,
            "host_packed_bytes": 4 * npk * (P + O) * HW, "host_threads": th, "raw_problems": nraw, "packed_problems": npk,
            "host_pack_seconds": t_host, "route_estimate": _HOST_SPLIT.get(key),
            "host_seconds": {"enqueue_raw_route": t_raw_enqueued, "pack": t_host, "whole_call": time.perf_counter() - t_call}}
