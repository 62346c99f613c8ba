"""Array-level calls into libb2q.so: every function takes framework arrays (torch tensors, MXNet NDArrays,
anything exporting DLPack), views them zero-copy and enqueues CUDA kernels on the arrays' stream.

Host (CPU) arrays are accepted by the operators that have a host-buffer entry point (the "e2e" path of
bench.py): they are staged through device memory by the library; nothing is ever computed on the CPU.
"""
import functools

import numpy as np

from . import _lib
from .dlpack import as_buffer, current_stream

REQ = _lib.REQ


def _req(r):
    if isinstance(r, int):
        return r
    try:
        return REQ[r]
    except KeyError:
        raise ValueError("unknown req %r (expected one of %s)" % (r, sorted(REQ)))


@functools.lru_cache(maxsize=256)
def _f32(v):
    """Python double -> float32 the way MXNet applies scalar operands."""
    return float(np.float32(v))


def _same_place(*bufs):
    first = bufs[0]
    for b in bufs:
        if b.on_device != first.on_device or (b.on_device and b.device_id != first.device_id):
            raise ValueError("all tensors of one operator call must live on the same device")
    return first.on_device, first.device_id


def _rows_cols(shape):
    rows = int(shape[0]) if len(shape) else 1
    n = 1
    for s in shape:
        n *= int(s)
    return rows, (n // rows if rows else 0), n


def _host_device():
    """Device whose staging buffers serve host-buffer calls: the process's current CUDA device."""
    try:
        import torch
        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except ImportError:  # pragma: no cover
        pass
    return 0


def host_sync():
    """Wait for every host-buffer call issued so far on the current device (they are asynchronous)."""
    _lib.context(_host_device()).host_sync()


def require_device(on_device, what):
    if not on_device:
        raise _lib.B2QError("%s has no host-buffer entry point; pass CUDA tensors (there is no CPU fallback)" % what)


# ---------------------------------------------------------------------------------------------------------
def assign(dst, req, src):
    """CustomOp.assign."""
    r = _req(req)
    if r == 0:
        return
    d, s = as_buffer(dst, write=True), as_buffer(src)
    if d.numel != s.numel:
        raise ValueError("assign: size mismatch %s vs %s" % (d.shape, s.shape))
    on_dev, dev = _same_place(d, s)
    if on_dev:
        _lib.context(dev).call("b2q_ste_bwd_f32", s.ptr, d.ptr, d.numel, r, current_stream(d))
    else:
        if r == 3:
            raise _lib.B2QError("assign(add) on host buffers is not supported")
        _lib.context(_host_device()).call("b2q_ste_bwd_host_f32", s.ptr, d.ptr, d.numel)


def zero_(dst):
    """in_grad[i][:] = 0 (fold_bn_v1_gdrq.py:124-125): a memset on the array's stream."""
    d = as_buffer(dst, write=True)
    if d.on_device:
        _lib.context(d.device_id).call("b2q_zero_f32", d.ptr, d.numel, current_stream(d))
    else:
        dst[:] = 0


def scratch_like(ref, n=1, dtype_name="float32"):
    """Small device scratch array next to ``ref``, allocated by whichever framework owns ``ref``: torch for torch
    tensors, ``mx.nd.empty(ctx=ref.context)`` for MXNet NDArrays (they have no ``.device``), and for any other DLPack
    exporter a torch tensor on the same CUDA device (torch only supplies the memory)."""
    try:
        import torch
    except ImportError:  # pragma: no cover
        torch = None
    if torch is not None and isinstance(ref, torch.Tensor):
        return torch.empty(n, dtype=getattr(torch, dtype_name), device=ref.device)
    if hasattr(ref, "context") and hasattr(ref, "to_dlpack_for_write"):   # mx.nd.NDArray
        import mxnet as mx
        return mx.nd.empty((n,), ctx=ref.context, dtype=dtype_name)
    b = as_buffer(ref)
    if torch is None:  # pragma: no cover
        raise _lib.B2QError("cannot allocate scratch memory for %r without torch or mxnet" % type(ref))
    dev = torch.device("cuda", b.device_id) if b.on_device else torch.device("cpu")
    return torch.empty(n, dtype=getattr(torch, dtype_name), device=dev)


def minmax_quant_fwd(variant, x, y, aux, is_weight, per_channel, is_train, init, ema_decay, req):
    xb, yb, ab = as_buffer(x), as_buffer(y, write=True), as_buffer(aux, write=True)
    rows, cols, n = _rows_cols(xb.shape)
    if yb.numel != n:
        raise ValueError("output shape %s does not match input %s" % (yb.shape, xb.shape))
    want_aux = rows if (per_channel and is_weight) else 1
    if ab.numel != want_aux:
        raise ValueError("aux has %d elements, expected %d" % (ab.numel, want_aux))
    on_dev, dev = _same_place(xb, yb, ab)
    d, omd = _f32(ema_decay), _f32(1 - ema_decay)
    if on_dev:
        _lib.context(dev).call("b2q_minmax_quant_fwd_f32", int(variant), xb.ptr, yb.ptr, ab.ptr, rows, cols,
                               int(bool(is_weight)), int(bool(per_channel)), int(bool(is_train)), int(bool(init)),
                               d, omd, _req(req), current_stream(xb))
    else:
        if _req(req) not in (1, 2):
            raise _lib.B2QError("host-buffer forward supports req=write only")
        _lib.context(_host_device()).call("b2q_minmax_quant_fwd_host_f32", int(variant), xb.ptr, yb.ptr, ab.ptr, rows, cols,
                             int(bool(is_weight)), int(bool(per_channel)), int(bool(is_train)), int(bool(init)),
                             d, omd)


def minmax_quant_stat(x, stat, per_channel):
    xb, sb = as_buffer(x), as_buffer(stat, write=True)
    rows, cols, _ = _rows_cols(xb.shape)
    on_dev, dev = _same_place(xb, sb)
    require_device(on_dev, "minmax_quant_stat")
    _lib.context(dev).call("b2q_minmax_quant_stat_f32", xb.ptr, rows, cols, int(bool(per_channel)), sb.ptr,
                           current_stream(xb))


def minmax_quant_finish(variant, x, y, aux, stat, is_weight, per_channel, is_train, init, ema_decay, req):
    xb, yb, ab, sb = as_buffer(x), as_buffer(y, write=True), as_buffer(aux, write=True), as_buffer(stat)
    rows, cols, _ = _rows_cols(xb.shape)
    on_dev, dev = _same_place(xb, yb, ab, sb)
    require_device(on_dev, "minmax_quant_finish")
    _lib.context(dev).call("b2q_minmax_quant_finish_f32", int(variant), xb.ptr, yb.ptr, ab.ptr, sb.ptr, rows, cols,
                           int(bool(is_weight)), int(bool(per_channel)), int(bool(is_train)), int(bool(init)),
                           _f32(ema_decay), _f32(1 - ema_decay), _req(req), current_stream(xb))


def ste_bwd(dy, dx, req):
    assign(dx, req, dy)


def clipgrad_bwd(x, dy, dx, aux):
    xb, gb, ob, ab = as_buffer(x), as_buffer(dy), as_buffer(dx, write=True), as_buffer(aux)
    on_dev, dev = _same_place(xb, gb, ob, ab)
    if on_dev:
        _lib.context(dev).call("b2q_clipgrad_bwd_f32", xb.ptr, gb.ptr, ob.ptr, ab.ptr, xb.numel, current_stream(xb))
    else:
        _lib.context(_host_device()).call("b2q_clipgrad_bwd_host_f32", xb.ptr, gb.ptr, ob.ptr, ab.ptr, xb.numel)


def _gdrq_view(shape, group_size, is_weight):
    """(outer, groups, inner) view of GDRQ's grouping (GDRQ.py:88-94) without transposing anything."""
    rows, cols, n = _rows_cols(shape)
    if group_size == -1:
        return 1, 1, n
    if is_weight:
        channels = int(shape[0])
        assert channels % group_size == 0
        return 1, channels // group_size, group_size * cols
    channels = int(shape[1])
    assert channels % group_size == 0
    inner = group_size
    for s in shape[2:]:
        inner *= int(s)
    return int(shape[0]), channels // group_size, inner


def gdrq_fwd(x, y, alpha, group_size, is_weight, fix_alpha, do_round, qlevel, ktimes, lamda, req):
    xb, yb, ab = as_buffer(x), as_buffer(y, write=True), as_buffer(alpha, write=True)
    outer, groups, inner = _gdrq_view(xb.shape, group_size, is_weight)
    if ab.numel != groups:
        raise ValueError("alpha has %d elements, expected %d" % (ab.numel, groups))
    on_dev, dev = _same_place(xb, yb, ab)
    require_device(on_dev, "GDRQ_PY")
    _lib.context(dev).call("b2q_gdrq_fwd_f32", xb.ptr, yb.ptr, ab.ptr, outer, groups, inner, int(bool(is_weight)),
                           int(bool(fix_alpha)), int(bool(do_round)), _f32(qlevel), _f32(ktimes), _f32(lamda),
                           _req(req), current_stream(xb))


def gdrq_bwd(x, dy, dx, alpha, group_size, req):
    xb, gb, ob, ab = as_buffer(x), as_buffer(dy), as_buffer(dx, write=True), as_buffer(alpha)
    outer, groups, inner = _gdrq_view(xb.shape, group_size, False)
    on_dev, dev = _same_place(xb, gb, ob, ab)
    require_device(on_dev, "GDRQ_PY")
    _lib.context(dev).call("b2q_gdrq_bwd_f32", xb.ptr, gb.ptr, ob.ptr, ab.ptr, outer, groups, inner, _req(req),
                           current_stream(xb))


def foldbn_data_fwd(x, y, aux_data, init, ema_decay):
    xb, yb, ab = as_buffer(x), as_buffer(y, write=True), as_buffer(aux_data, write=True)
    on_dev, dev = _same_place(xb, yb, ab)
    require_device(on_dev, "GDRQ_Fold_BN")
    _lib.context(dev).call("b2q_foldbn_data_fwd_f32", xb.ptr, yb.ptr, ab.ptr, xb.numel, int(bool(init)),
                           _f32(ema_decay), _f32(1 - ema_decay), current_stream(xb))


def foldbn_weight_fwd(w, w_q, bias, aux_weight, gamma, beta, mean, var, eps, per_channel, quantize, is_train):
    wb, qb, bb = as_buffer(w), as_buffer(w_q, write=True), as_buffer(bias, write=True)
    ab = as_buffer(aux_weight, write=True)
    g, b, m, v = as_buffer(gamma), as_buffer(beta), as_buffer(mean), as_buffer(var)
    cout, cols, _ = _rows_cols(wb.shape)
    for t in (g, b, m, v, bb):
        if t.numel != cout:
            raise ValueError("BN parameter has %d elements, expected num_filter=%d" % (t.numel, cout))
    on_dev, dev = _same_place(wb, qb, bb, ab, g, b, m, v)
    require_device(on_dev, "GDRQ_Fold_BN")
    _lib.context(dev).call("b2q_foldbn_weight_fwd_f32", wb.ptr, qb.ptr, bb.ptr, ab.ptr, g.ptr, b.ptr, m.ptr, v.ptr,
                           _f32(eps), cout, cols, int(bool(per_channel)), int(bool(quantize)), int(bool(is_train)),
                           current_stream(wb))


def _nchw(shape):
    n, c = int(shape[0]), int(shape[1])
    hw = 1
    for d in shape[2:]:
        hw *= int(d)
    return n, c, hw


def bn_batch_stats(y, mean, var):
    """BatchNorm_v1(output_mean_var=True) statistics of an (N, C, ...) tensor (fold_bn_v1_gdrq.py:271-272)."""
    yb, mb, vb = as_buffer(y), as_buffer(mean, write=True), as_buffer(var, write=True)
    n, c, hw = _nchw(yb.shape)
    if mb.numel != c or vb.numel != c:
        raise ValueError("mean / var must have %d elements" % c)
    on_dev, dev = _same_place(yb, mb, vb)
    require_device(on_dev, "bn_batch_stats")
    _lib.context(dev).call("b2q_bn_batch_stats_f32", yb.ptr, n, c, hw, mb.ptr, vb.ptr, current_stream(yb))


def bnstat_foldbn_weight_fwd(conv_out, mean, var, w, w_q, bias, aux_weight, gamma, beta, eps, per_channel, quantize,
                             is_train):
    """batch statistics of ``conv_out`` -> ``mean`` / ``var`` AND the fold-BN weight path (w_q, bias, aux_weight) that
    consumes them, in one launch for per-channel weights (fold_bn_v1_gdrq.py:268-287 -> :70-96,113)."""
    yb, mb, vb = as_buffer(conv_out), as_buffer(mean, write=True), as_buffer(var, write=True)
    wb, qb, bb = as_buffer(w), as_buffer(w_q, write=True), as_buffer(bias, write=True)
    ab, g, b = as_buffer(aux_weight, write=True), as_buffer(gamma), as_buffer(beta)
    n, c, hw = _nchw(yb.shape)
    cout, cols, _ = _rows_cols(wb.shape)
    if cout != c:
        raise ValueError("the convolution output has %d channels, the weight %d rows" % (c, cout))
    for t in (mb, vb, bb, g, b):
        if t.numel != c:
            raise ValueError("per-channel array has %d elements, expected %d" % (t.numel, c))
    on_dev, dev = _same_place(yb, mb, vb, wb, qb, bb, ab, g, b)
    require_device(on_dev, "GDRQ_Fold_BN")
    _lib.context(dev).call("b2q_bnstat_foldbn_weight_fwd_f32", yb.ptr, n, c, hw, mb.ptr, vb.ptr, wb.ptr, qb.ptr, bb.ptr,
                           ab.ptr, g.ptr, b.ptr, _f32(eps), cols, int(bool(per_channel)), int(bool(quantize)),
                           int(bool(is_train)), current_stream(yb))


def clip_relu_fwd(x, y, threshold, qlevel, req):
    xb, yb = as_buffer(x), as_buffer(y, write=True)
    on_dev, dev = _same_place(xb, yb)
    require_device(on_dev, "CLIP_RELU_PY")
    q = _f32(threshold / qlevel)   # python double division, then float32 (GDRQ.py:203)
    _lib.context(dev).call("b2q_clip_relu_fwd_f32", xb.ptr, yb.ptr, xb.numel, _f32(threshold), q, _req(req),
                           current_stream(xb))


def mask_bwd(x, dy, dx, thr, thr_imm, mask_mode, req, view=None):
    xb, gb, ob = as_buffer(x), as_buffer(dy), as_buffer(dx, write=True)
    tb = as_buffer(thr) if thr is not None else None
    outer, groups, inner = view if view is not None else (1, 1, xb.numel)
    on_dev, dev = _same_place(xb, gb, ob)
    require_device(on_dev, "masked backward")
    _lib.context(dev).call("b2q_mask_bwd_f32", xb.ptr, gb.ptr, ob.ptr, outer, groups, inner,
                           tb.ptr if tb else None, _f32(thr_imm), int(mask_mode), _req(req), current_stream(xb))


def absmax(x, stat, view=None):
    xb, sb = as_buffer(x), as_buffer(stat, write=True)
    outer, groups, inner = view if view is not None else (1, 1, xb.numel)
    on_dev, dev = _same_place(xb, sb)
    require_device(on_dev, "absmax")
    _lib.context(dev).call("b2q_absmax_f32", xb.ptr, outer, groups, inner, sb.ptr, current_stream(xb))


def meanabs(x, stat, view=None):
    xb, sb = as_buffer(x), as_buffer(stat, write=True)
    outer, groups, inner = view if view is not None else (1, 1, xb.numel)
    on_dev, dev = _same_place(xb, sb)
    require_device(on_dev, "meanabs")
    _lib.context(dev).call("b2q_meanabs_f32", xb.ptr, outer, groups, inner, sb.ptr, current_stream(xb))


def threshold_update(mode, stat, aux, p0, p1, clip_out=None):
    sb, ab = as_buffer(stat), as_buffer(aux, write=True)
    cb = as_buffer(clip_out, write=True) if clip_out is not None else None
    on_dev, dev = _same_place(sb, ab)
    require_device(on_dev, "threshold_update")
    _lib.context(dev).call("b2q_threshold_update_f32", int(mode), sb.ptr, ab.ptr, cb.ptr if cb else None, ab.numel,
                           _f32(p0), _f32(p1), current_stream(sb))


def qdq(x, y, thr, qlevel, clip_mode, req, view=None, clip_thr=None, do_round=True, codes=None, prescale=None):
    xb, yb, tb = as_buffer(x), as_buffer(y, write=True), as_buffer(thr)
    outer, groups, inner = view if view is not None else (1, 1, xb.numel)
    cb = as_buffer(clip_thr) if clip_thr is not None else None
    on_dev, dev = _same_place(xb, yb, tb)
    require_device(on_dev, "qdq")
    codes_ptr = None
    if codes is not None:
        import torch
        assert isinstance(codes, torch.Tensor) and codes.dtype == torch.int32 and codes.is_contiguous()
        codes_ptr = codes.data_ptr()
    g_ptr = v_ptr = None
    eps = 0.0
    if prescale is not None:
        gam, var, eps = prescale
        g_ptr, v_ptr = as_buffer(gam).ptr, as_buffer(var).ptr
    _lib.context(dev).call("b2q_qdq_f32", xb.ptr, yb.ptr, outer, groups, inner, tb.ptr, cb.ptr if cb else None,
                           _f32(qlevel), int(clip_mode), int(bool(do_round)), _req(req), codes_ptr, g_ptr, v_ptr,
                           _f32(eps), current_stream(xb))


def wnq_fwd(x, y, per_channel, qlevel, req):
    xb, yb = as_buffer(x), as_buffer(y, write=True)
    rows, cols, _ = _rows_cols(xb.shape)
    on_dev, dev = _same_place(xb, yb)
    require_device(on_dev, "WNQ_PY")
    _lib.context(dev).call("b2q_wnq_fwd_f32", xb.ptr, yb.ptr, rows, cols, int(bool(per_channel)), _f32(qlevel),
                           _req(req), current_stream(xb))


def wnq_bwd(x, dy, dx, per_channel, req):
    xb, gb, ob = as_buffer(x), as_buffer(dy), as_buffer(dx, write=True)
    rows, cols, _ = _rows_cols(xb.shape)
    on_dev, dev = _same_place(xb, gb, ob)
    require_device(on_dev, "WNQ_PY")
    _lib.context(dev).call("b2q_wnq_bwd_f32", xb.ptr, gb.ptr, ob.ptr, rows, cols, int(bool(per_channel)), _req(req),
                           current_stream(xb))


def pact_bwd(x, dy, dx, dgamma, gamma, two_sided, req, req_gamma):
    xb, gb, ob = as_buffer(x), as_buffer(dy), as_buffer(dx, write=True)
    dgb, gmb = as_buffer(dgamma, write=True), as_buffer(gamma)
    on_dev, dev = _same_place(xb, gb, ob, dgb, gmb)
    require_device(on_dev, "PACT_PY")
    _lib.context(dev).call("b2q_pact_bwd_f32", xb.ptr, gb.ptr, ob.ptr, dgb.ptr, gmb.ptr, xb.numel,
                           int(bool(two_sided)), _req(req), _req(req_gamma), current_stream(xb))


def dorefa_fwd(x, y, vmax, qlevel, req):
    xb, yb, vb = as_buffer(x), as_buffer(y, write=True), as_buffer(vmax, write=True)
    on_dev, dev = _same_place(xb, yb, vb)
    require_device(on_dev, "DoReFa_PY")
    _lib.context(dev).call("b2q_dorefa_fwd_f32", xb.ptr, yb.ptr, vb.ptr, xb.numel, _f32(qlevel), _req(req),
                           current_stream(xb))


def dorefa_bwd(x, dy, dx, vmax, req):
    xb, gb, ob, vb = as_buffer(x), as_buffer(dy), as_buffer(dx, write=True), as_buffer(vmax)
    on_dev, dev = _same_place(xb, gb, ob, vb)
    require_device(on_dev, "DoReFa_PY")
    _lib.context(dev).call("b2q_dorefa_bwd_f32", xb.ptr, gb.ptr, ob.ptr, vb.ptr, xb.numel, _req(req),
                           current_stream(xb))


def qil_fwd(variant, x, y, p0, p1, qlevel, req):
    xb, yb = as_buffer(x), as_buffer(y, write=True)
    b0, b1 = as_buffer(p0, write=True), as_buffer(p1, write=True)
    on_dev, dev = _same_place(xb, yb, b0, b1)
    require_device(on_dev, "QIL")
    _lib.context(dev).call("b2q_qil_fwd_f32", int(variant), xb.ptr, yb.ptr, b0.ptr, b1.ptr, xb.numel, _f32(qlevel),
                           _req(req), current_stream(xb))


def qil_bwd(variant, x, dy, dx, p0, p1, dp0, dp1, req, req_p0, req_p1):
    xb, gb, ob = as_buffer(x), as_buffer(dy), as_buffer(dx, write=True)
    b0, b1 = as_buffer(p0), as_buffer(p1)
    d0, d1 = as_buffer(dp0, write=True), as_buffer(dp1, write=True)
    on_dev, dev = _same_place(xb, gb, ob, b0, b1, d0, d1)
    require_device(on_dev, "QIL")
    _lib.context(dev).call("b2q_qil_bwd_f32", int(variant), xb.ptr, gb.ptr, ob.ptr, b0.ptr, b1.ptr, d0.ptr, d1.ptr,
                           xb.numel, _req(req), _req(req_p0), _req(req_p1), current_stream(xb))


def export_int8(x, thr, qlevel, clip_mode, view=None):
    """(int8 codes, float32 steps): the integers and the per-group step the fake-quant output is made of."""
    import torch
    xb, tb = as_buffer(x), as_buffer(thr)
    outer, groups, inner = view if view is not None else (1, 1, xb.numel)
    on_dev, dev = _same_place(xb, tb)
    require_device(on_dev, "export_int8")
    tdev = x.device if isinstance(x, torch.Tensor) else torch.device("cuda", xb.device_id)
    codes = torch.empty(xb.shape, dtype=torch.int8, device=tdev)
    steps = torch.empty(groups, dtype=torch.float32, device=tdev)
    _lib.context(dev).call("b2q_export_int8_f32", xb.ptr, codes.data_ptr(), steps.data_ptr(), outer, groups, inner,
                           tb.ptr, _f32(qlevel), int(clip_mode), current_stream(xb))
    return codes, steps
