"""ctypes binding of the C ABI declared in ``include/sphk.h`` (libsphk.so, sm_100a CUDA).

This is the ONLY compute path of the package.  There is no CPU or eager-PyTorch fallback: a
missing library raises at import of this module, a non-CUDA tensor raises at the call."""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libsphk.so")
_PROBE_LIB = os.environ.get("SPHK_PROBE_LIB")      # tools/ only: an instrumented build of the same sources

KIND = {"sph2pob_efficient": 0, "sph2pob_standard": 1, "sph": 2, "fov": 3, "naive": 4, "unbiased": 5,
        "sph2pob_legacy": 6}
MODE = {"iou": 0, "iof": 1}
EDGE = {"arc": 0, "chord": 1, "tangent": 2}
ANGLE = {"equator": 0, "project": 1}

SPHK_OK = 0
ABI_VERSION = 8

_c_float_p = ctypes.c_void_p  # raw device addresses
_i64 = ctypes.c_int64
_i32 = ctypes.c_int32
_int = ctypes.c_int

# name -> (restype, argtypes); must list every symbol of include/sphk.h (tests check this)
SIGNATURES = {
    "sphk_abi_version": (_int, []),
    "sphk_last_error_string": (ctypes.c_char_p, []),
    "sphk_device_info": (_int, [ctypes.POINTER(_int)] * 3),
    "sphk_iou_aligned": (_int, [_int, _c_float_p, _c_float_p, _i64, _int, _int, _int, _int, _c_float_p, ctypes.c_void_p]),
    "sphk_iou_pairwise_workspace_bytes": (_i64, [_i64, _i64]),
    "sphk_iou_pairwise": (_int, [_int, _c_float_p, _i64, _c_float_p, _i64, _int, _int, _int, _int, _c_float_p, _i64,
                                 _c_float_p, ctypes.c_void_p, _c_float_p, ctypes.c_void_p, _i32, _i32,
                                 ctypes.c_void_p, ctypes.c_void_p]),
    "sphk_iou_pairwise_keys": (_int, [_int, _c_float_p, _i64, _c_float_p, _i64, _int, _int, _int, ctypes.c_void_p, ctypes.c_void_p,
                                      _i32, _i32, _int, ctypes.c_void_p, ctypes.c_void_p]),
    "sphk_iou_pairwise_ties": (_int, [_int, _c_float_p, _i64, _c_float_p, _i64, _int, _int, _int, _c_float_p, ctypes.c_void_p,
                                      _i32, ctypes.c_void_p, ctypes.c_void_p]),
    "sphk_unpack_gathered_keys": (_int, [ctypes.c_void_p, _i32, _i64, _i64, _i64, _c_float_p, ctypes.c_void_p, _c_float_p,
                                         ctypes.c_void_p, ctypes.c_void_p]),
    "sphk_key_push_parts": (_i32, [_i64]),
    "sphk_iou_pairwise_keys_push": (_int, [_int, _c_float_p, _i64, _c_float_p, _i64, _int, _int, _int, ctypes.c_void_p, _i32, _i32,
                                           ctypes.c_void_p, _i32, _i64, _i64, ctypes.c_void_p, ctypes.c_void_p]),
    "sphk_unpack_peer_keys": (_int, [ctypes.c_void_p, _i32, _i32, ctypes.c_uint64, _i64, _i64, _i64, _i64, _i64, _i32, _i32, _c_float_p,
                                     ctypes.c_void_p, _c_float_p, ctypes.c_void_p, ctypes.c_void_p]),
    "sphk_max_iou_assign_workspace_bytes": (_i64, [_i64, _i64, _i32]),
    "sphk_max_iou_assign": (_int, [_int, _c_float_p, ctypes.POINTER(_i32), _i32, _c_float_p, _i64, _int, ctypes.c_float,
                                   ctypes.c_float, ctypes.c_float, ctypes.c_float, _int, _int, ctypes.c_void_p, ctypes.c_void_p,
                                   _c_float_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "sphk_anchor_targets": (_int, [ctypes.c_void_p, _i32, _i64, _int, _c_float_p, _c_float_p, ctypes.c_void_p, ctypes.c_void_p, _i64,
                                   ctypes.c_float, _int, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), ctypes.c_void_p,
                                   _c_float_p, _c_float_p, _c_float_p, ctypes.c_void_p, ctypes.c_void_p]),
    "sphk_loss_fwd_bwd": (_int, [_c_float_p, _c_float_p, _i64, _int, _c_float_p, _c_float_p, _c_float_p, _c_float_p,
                                 ctypes.c_void_p]),
    "sphk_loss_reduce_partials": (_i64, [_i64]),
    "sphk_loss_reduce": (_int, [_c_float_p, _c_float_p, _c_float_p, _i64, _int, ctypes.c_float, _c_float_p, _c_float_p, _c_float_p,
                                ctypes.c_void_p]),
    "sphk_loss_total_scratch_bytes": (_i64, [_i64]),
    "sphk_loss_reduce_total": (_int, [_c_float_p, _c_float_p, _c_float_p, _i64, _int, ctypes.c_float, _c_float_p, ctypes.c_void_p,
                                      _c_float_p, _c_float_p, ctypes.c_void_p]),
    "sphk_obb_fwd": (_int, [_int, _c_float_p, _c_float_p, _i64, _int, _int, _c_float_p, _c_float_p, ctypes.c_void_p]),
    "sphk_obb_bwd": (_int, [_int, _c_float_p, _c_float_p, _i64, _int, _int, _c_float_p, _c_float_p, _c_float_p,
                            _c_float_p, ctypes.c_void_p]),
    "sphk_riou_fwd_bwd": (_int, [_c_float_p, _c_float_p, _i64, _c_float_p, _c_float_p, _c_float_p, _c_float_p,
                                 ctypes.c_void_p]),
    "sphk_obb_loss": (_int, [_int, _int, _int, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, _int, _c_float_p,
                             _c_float_p, _i64, _int, _c_float_p, _int, ctypes.c_float, _c_float_p, _c_float_p, _c_float_p,
                             _c_float_p, ctypes.c_void_p]),
    "sphk_obb_loss_total": (_int, [_int, _int, _int, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, _int, _c_float_p,
                                   _c_float_p, _i64, _int, _c_float_p, _int, ctypes.c_float, _c_float_p, ctypes.c_void_p, _c_float_p,
                                   _c_float_p, ctypes.c_void_p]),
    "sphk_coder_decode": (_int, [_c_float_p, _c_float_p, _i64, _int, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float),
                                 ctypes.c_float, _int, _int, ctypes.c_float, _c_float_p, ctypes.c_void_p]),
    "sphk_coder_decode_bwd": (_int, [_c_float_p, _c_float_p, _c_float_p, _i64, _int, ctypes.POINTER(ctypes.c_float),
                                     ctypes.POINTER(ctypes.c_float), ctypes.c_float, _int, _int, ctypes.c_float, _c_float_p,
                                     ctypes.c_void_p]),
    "sphk_coder_encode": (_int, [_c_float_p, _c_float_p, _i64, _int, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float),
                                 _c_float_p, ctypes.c_void_p]),
    "sphk_box_format": (_int, [_int, _c_float_p, _i64, _int, _int, ctypes.c_float, ctypes.c_float, _c_float_p, ctypes.c_void_p]),
    "sphk_decode_loss_partials": (_i64, [_i64]),
    "sphk_decode_loss_reduce": (_int, [_c_float_p, _c_float_p, _c_float_p, _c_float_p, _int, _i64, _int,
                                       ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), ctypes.c_float, _int, _int,
                                       ctypes.c_float, ctypes.c_float, _c_float_p, _c_float_p, ctypes.c_void_p]),
    "sphk_nms_batched": (_int, [_c_float_p, ctypes.c_void_p, ctypes.c_void_p, _i32, _i32, _i32, _int, _int, ctypes.c_float,
                                ctypes.c_void_p, ctypes.c_void_p]),
    "sphk_nms_images_workspace_bytes": (_i64, [_i32, _i32, _i32]),
    "sphk_nms_images": (_int, [_c_float_p, _c_float_p, ctypes.c_void_p, ctypes.c_void_p, _i32, _i32, _i32, _int, _int, ctypes.c_float,
                               _i32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "sphk_probe_fp32": (_int, [_i32, _i32, _c_float_p, ctypes.c_void_p]),
    "sphk_set_dense": (_int, [_int]),
    "sphk_prefilter_count": (_int, [_c_float_p, _i64, _c_float_p, _i64, _int, _int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
}


class SphkError(RuntimeError):
    pass


def _load():
    # (re)build in-tree when the library is missing or older than its sources; nothing else can serve this path
    try:
        from . import build as _build
        if _build.is_stale():
            _build.build()
    except Exception as e:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                "sph_retina_b200: %s is missing and could not be built (%s). Build it with "
                "`python -m sph_retina_b200.build` (nvcc, sm_100a). There is no CPU fallback." % (LIB_PATH, e))
    lib = ctypes.CDLL(_PROBE_LIB or LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    got = lib.sphk_abi_version()
    if got != ABI_VERSION:
        raise ImportError("libsphk.so ABI version %d, binding expects %d: rebuild" % (got, ABI_VERSION))
    return lib


lib = _load()
launches = 0  # number of C-ABI compute calls issued by this process (bench.py reports kernel counts from it)


def _check(status: int):
    if status != SPHK_OK:
        raise SphkError("libsphk: %s (status %d)" % (lib.sphk_last_error_string().decode(), status))


def _ptr(t):
    return None if t is None else t.data_ptr()


try:
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:  # pragma: no cover
    def _raw_stream(index):
        return torch.cuda.current_stream(index).cuda_stream


def _stream(t):
    return _raw_stream(t.device.index)


class _on_device:
    """`with torch.cuda.device(dev)` only when dev is not already current (the context manager costs ~10 us)."""
    __slots__ = ("ctx",)

    def __init__(self, dev):
        self.ctx = None if dev.index == torch.cuda.current_device() else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


_workspaces = {}


def _workspace(dev, nbytes):
    """Per (device, stream) scratch buffer, grown geometrically.  Re-use across calls is safe because every
    consumer is enqueued on that same stream."""
    key = (dev.index, _raw_stream(dev.index))
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes + nbytes // 2, 1 << 16), dtype=torch.uint8, device=dev)
        _workspaces[key] = ws
    return ws


def _boxes(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise SphkError("%s is on %s: the sm_100a kernels are the only implementation of this path "
                        "(no CPU fallback); move the boxes to a CUDA device" % (name, t.device))
    if t.dim() != 2 or t.size(1) not in (4, 5):
        raise SphkError("%s must have shape [n, 4] (BFoV) or [n, 5] (RBFoV), got %s" % (name, tuple(t.shape)))
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def device_info():
    sm, major, minor = _int(), _int(), _int()
    _check(lib.sphk_device_info(ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor)))
    return sm.value, major.value, minor.value


def iou_aligned(kind: str, b1, b2, mode="iou", edge="arc", angle="equator") -> torch.Tensor:
    global launches
    b1, b2 = _boxes(b1, "bboxes1"), _boxes(b2, "bboxes2")
    if b1.shape != b2.shape:
        raise SphkError("aligned IoU needs equal shapes, got %s and %s" % (tuple(b1.shape), tuple(b2.shape)))
    out = torch.empty(b1.size(0), dtype=torch.float32, device=b1.device)
    with _on_device(b1.device):
        _check(lib.sphk_iou_aligned(KIND[kind], _ptr(b1), _ptr(b2), b1.size(0), b1.size(1), MODE[mode], EDGE[edge],
                                    ANGLE[angle], _ptr(out), _stream(b1)))
    launches += 1
    return out


def iou_pairwise(kind: str, rows, cols, mode="iou", edge="arc", want_matrix=True, want_row_max=False,
                 want_col_max=False, row_base=0, col_base=0, out=None, angle="equator"):
    """Returns (matrix|None, (row_max,row_arg)|None, (col_max,col_arg)|None)."""
    global launches
    rows, cols = _boxes(rows, "bboxes1"), _boxes(cols, "bboxes2")
    if rows.size(1) != cols.size(1):
        raise SphkError("box widths differ: %d vs %d" % (rows.size(1), cols.size(1)))
    R, C, dev = rows.size(0), cols.size(0), rows.device
    mat = None
    if want_matrix:
        mat = out if out is not None else torch.empty((R, C), dtype=torch.float32, device=dev)
        assert mat.is_cuda and mat.dtype == torch.float32 and mat.shape == (R, C) and mat.stride(1) == 1
    ld = mat.stride(0) if (mat is not None and R > 1) else C
    rmax = rarg = cmax = carg = ws = None
    if want_row_max:
        rmax = torch.empty(R, dtype=torch.float32, device=dev)
        rarg = torch.empty(R, dtype=torch.int32, device=dev)
    if want_col_max:
        cmax = torch.empty(C, dtype=torch.float32, device=dev)
        carg = torch.empty(C, dtype=torch.int32, device=dev)
    if want_row_max or want_col_max or kind in ("sph2pob_efficient", "sph2pob_standard"):
        ws = _workspace(dev, 136 * (R + C) + 32)     # >= sphk_iou_pairwise_workspace_bytes(R, C)
    with _on_device(dev):
        _check(lib.sphk_iou_pairwise(KIND[kind], _ptr(rows), R, _ptr(cols), C, rows.size(1), MODE[mode], EDGE[edge],
                                     ANGLE[angle], _ptr(mat), ld, _ptr(rmax), _ptr(rarg), _ptr(cmax), _ptr(carg), row_base, col_base,
                                     _ptr(ws), _stream(rows)))
    if kind in ("sph2pob_efficient", "sph2pob_standard"):
        if mat is not None and R <= 32 and not (want_row_max or want_col_max):
            launches += 1                                        # k_iou_rows32: records computed inside the CTAs
        else:
            launches += 2 + int(want_row_max or want_col_max)   # k_box_pre (zeroes the keys), k_iou_pairwise2[, k_unpack_keys2]
    else:
        launches += 1 + 2 * int(want_row_max) + 2 * int(want_col_max) - int(want_row_max and want_col_max)
    return mat, ((rmax, rarg) if want_row_max else None), ((cmax, carg) if want_col_max else None)


def iou_pairwise_keys(kind: str, rows, cols, mode="iou", edge="arc", row_base=0, col_base=0, row_keys_out=None,
                      col_keys_out=None, keep_row_keys=False, keep_col_keys=False):
    """Fused max/argmax of the N x M overlaps as the kernel's packed keys, without unpacking:
    (row_keys[R], col_keys[C]) int64 = float32 bits << 32 | (0xFFFFFFFF - index); 0 = no positive overlap.
    Integer MAX over shards of such keys = (max value, lowest index): what the multi-GPU sweep reduces.
    ``row_keys_out`` / ``col_keys_out``: contiguous int64 CUDA tensors to write into (slices of a communication buffer);
    ``keep_*_keys``: raise the keys already there instead of resetting them first (row-chunked sweeps)."""
    global launches
    rows, cols = _boxes(rows, "bboxes1"), _boxes(cols, "bboxes2")
    R, C, dev = rows.size(0), cols.size(0), rows.device
    rk = row_keys_out if row_keys_out is not None else torch.empty(R, dtype=torch.int64, device=dev)
    ck = col_keys_out if col_keys_out is not None else torch.empty(C, dtype=torch.int64, device=dev)
    for k, n in ((rk, R), (ck, C)):
        if not (k.is_cuda and k.dtype == torch.int64 and k.is_contiguous() and k.numel() == n and k.device == dev):
            raise SphkError("key outputs must be contiguous int64 CUDA tensors of %d elements on %s" % (n, dev))
    ws = _workspace(dev, 136 * (R + C) + 32)
    with _on_device(dev):
        _check(lib.sphk_iou_pairwise_keys(KIND[kind], _ptr(rows), R, _ptr(cols), C, rows.size(1), MODE[mode], EDGE[edge],
                                          _ptr(rk), _ptr(ck), row_base, col_base, (1 if keep_row_keys else 0) | (2 if keep_col_keys else 0),
                                          _ptr(ws), _stream(rows)))
    launches += 2
    return rk, ck


def key_push_parts(C: int) -> int:
    """Partial arrays per row key that iou_pairwise_keys_push writes: one per column tile of the compute kernel."""
    return int(lib.sphk_key_push_parts(C))


def iou_pairwise_keys_push(kind: str, rows, cols, col_keys_out, peer_bufs_dev: int, world: int, push_offset: int, part_stride: int,
                           mode="iou", edge="arc", row_base=0, col_base=0):
    """The fused max / argmax sweep with the row keys stored into EVERY rank's symmetric buffer from inside the launch
    (include/sphk.h: sphk_iou_pairwise_keys_push): partial array t (one per column tile, key_push_parts(C) of them) at
    element ``push_offset + t * part_stride`` of each buffer; the column keys go to ``col_keys_out``.
    ``peer_bufs_dev``: device address of the array of the ranks' buffer base pointers."""
    global launches
    rows, cols = _boxes(rows, "bboxes1"), _boxes(cols, "bboxes2")
    R, C, dev = rows.size(0), cols.size(0), rows.device
    k = col_keys_out
    if not (k.is_cuda and k.dtype == torch.int64 and k.is_contiguous() and k.numel() == C and k.device == dev):
        raise SphkError("col_keys_out must be a contiguous int64 CUDA tensor of %d elements on %s" % (C, dev))
    ws = _workspace(dev, 136 * (R + C) + 32)
    with _on_device(dev):
        _check(lib.sphk_iou_pairwise_keys_push(KIND[kind], _ptr(rows), R, _ptr(cols), C, rows.size(1), MODE[mode], EDGE[edge],
                                               _ptr(k), row_base, col_base, peer_bufs_dev, world, push_offset, part_stride,
                                               _ptr(ws), _stream(rows)))
    launches += 2
    return k


def unpack_gathered_keys(gathered, world: int, n_long: int, n_short: int, cap: int, out=None):
    """Row-sharded N x M: (long_max[n_long], long_arg[n_long] int64, short_max[n_short], short_arg[n_short] int64) from
    the all-gathered key blocks [world, cap + n_short] (include/sphk.h: sphk_unpack_gathered_keys), one launch.
    ``out``: optional 4-tuple of preallocated CUDA tensors to write into."""
    global launches
    if not (gathered.is_cuda and gathered.dtype == torch.int64 and gathered.is_contiguous()
            and gathered.numel() == world * (cap + n_short)):
        raise SphkError("gathered must be a contiguous int64 CUDA tensor of world * (cap + n_short) keys")
    dev = gathered.device
    if out is None:
        out = (torch.empty(n_long, dtype=torch.float32, device=dev), torch.empty(n_long, dtype=torch.int64, device=dev),
               torch.empty(n_short, dtype=torch.float32, device=dev), torch.empty(n_short, dtype=torch.int64, device=dev))
    lmax, larg, smax, sarg = out
    for t, n, dt in ((lmax, n_long, torch.float32), (larg, n_long, torch.int64), (smax, n_short, torch.float32), (sarg, n_short, torch.int64)):
        if not (t.is_cuda and t.dtype == dt and t.is_contiguous() and t.numel() == n):
            raise SphkError("unpack_gathered_keys: output tensors must be contiguous CUDA tensors of the result shapes")
    with _on_device(dev):
        _check(lib.sphk_unpack_gathered_keys(_ptr(gathered), world, n_long, n_short, cap, _ptr(lmax), _ptr(larg), _ptr(smax),
                                             _ptr(sarg), _stream(gathered)))
    launches += 1
    return lmax, larg, smax, sarg


def unpack_peer_keys(peer_bufs_dev: int, rank: int, world: int, step: int, block_offset: int, flag_offset: int, n_long: int,
                     n_short: int, cap: int, device, out=None, long_parts=1, long_pushed=False):
    """The exchange step of the row-sharded N x M without a collective (include/sphk.h: sphk_unpack_peer_keys): flag
    handshake over the ranks' symmetric buffers, then the keys of the long operand from the local slots (``long_pushed``:
    every rank's compute kernel stored ``long_parts`` partial arrays there) or from their owners' buffers over NVLink, the
    short operand's from their owners.  ``peer_bufs_dev``: device address of the array of the ranks' buffer base pointers
    (``_SymmetricMemory.buffer_ptrs_dev``); ``out``: optional preallocated outputs."""
    global launches
    if out is None:
        out = (torch.empty(n_long, dtype=torch.float32, device=device), torch.empty(n_long, dtype=torch.int64, device=device),
               torch.empty(n_short, dtype=torch.float32, device=device), torch.empty(n_short, dtype=torch.int64, device=device))
    lmax, larg, smax, sarg = out
    with _on_device(device):
        _check(lib.sphk_unpack_peer_keys(peer_bufs_dev, rank, world, step, block_offset, flag_offset, n_long, n_short, cap,
                                         long_parts, 1 if long_pushed else 0, _ptr(lmax), _ptr(larg), _ptr(smax), _ptr(sarg),
                                         _raw_stream(device.index)))
    launches += 1
    return lmax, larg, smax, sarg


def iou_pairwise_ties(kind: str, rows, cols, row_target, mode="iou", edge="arc", row_base=0):
    """col_tie[j] = max over rows i with IoU(rows[i], cols[j]) == row_target[i] > 0 of (row_base + i + 1), else 0."""
    global launches
    rows, cols = _boxes(rows, "bboxes1"), _boxes(cols, "bboxes2")
    R, C, dev = rows.size(0), cols.size(0), rows.device
    row_target = row_target.to(device=dev, dtype=torch.float32).contiguous()
    assert row_target.numel() == R
    tie = torch.empty(C, dtype=torch.int32, device=dev)
    ws = _workspace(dev, 136 * (R + C) + 32)
    with _on_device(dev):
        _check(lib.sphk_iou_pairwise_ties(KIND[kind], _ptr(rows), R, _ptr(cols), C, rows.size(1), MODE[mode], EDGE[edge],
                                          _ptr(row_target), _ptr(tie), row_base, _ptr(ws), _stream(rows)))
    launches += 2
    return tie


def max_iou_assign(kind: str, gts, gt_offsets, boxes, pos_iou_thr, neg_lo, neg_hi, min_pos_iou, gt_max_assign_all,
                   match_low_quality, gt_labels=None):
    """Batched MaxIoUAssigner.  gts [sumK, D] (all images), gt_offsets: python list of batch+1 ints, boxes [N, D] shared.
    Returns (gt_inds [B, N] int64, max_overlaps [B, N] float32, labels [B, N] int64 | None)."""
    global launches
    boxes = _boxes(boxes, "bboxes")
    dev, N, B = boxes.device, boxes.size(0), len(gt_offsets) - 1
    sumK = int(gt_offsets[-1])
    if sumK > 0:
        gts = _boxes(gts, "gt_bboxes")
        assert gts.size(0) == sumK and gts.size(1) == boxes.size(1)
    gt_inds = torch.empty((B, N), dtype=torch.int64, device=dev)
    max_overlaps = torch.empty((B, N), dtype=torch.float32, device=dev)
    labels = None
    if gt_labels is not None:
        gt_labels = gt_labels.to(device=dev, dtype=torch.int64).contiguous()
        assert gt_labels.numel() == sumK
        labels = torch.empty((B, N), dtype=torch.int64, device=dev)
    if B == 0 or N == 0:
        return gt_inds, max_overlaps, labels
    offs = (_i32 * (B + 1))(*[int(v) for v in gt_offsets])
    ws = _workspace(dev, lib.sphk_max_iou_assign_workspace_bytes(sumK, N, B))
    with _on_device(dev):
        _check(lib.sphk_max_iou_assign(KIND[kind], _ptr(gts) if sumK > 0 else None, offs, B, _ptr(boxes), N, boxes.size(1),
                                       float(pos_iou_thr), float(neg_lo), float(neg_hi), float(min_pos_iou),
                                       int(bool(gt_max_assign_all)), int(bool(match_low_quality)), _ptr(gt_labels), _ptr(gt_inds),
                                       _ptr(max_overlaps), _ptr(labels), _ptr(ws), _stream(boxes)))
    launches += 6
    return gt_inds, max_overlaps, labels


def anchor_targets(gt_inds, anchors, gts, gt_labels, gt_offsets, num_classes, pos_weight=-1.0, reg_decoded_bbox=True, means=None,
                   stds=None):
    """labels, label_weights, bbox_targets, bbox_weights, counts[B, 2] of the anchor head from an assignment
    (sphk_anchor_targets).  gt_inds [B, N] int64 as returned by max_iou_assign; gt_offsets: python list of B + 1 ints."""
    global launches
    anchors = _boxes(anchors, "anchors")
    dev, N, D = anchors.device, anchors.size(0), anchors.size(1)
    B = len(gt_offsets) - 1
    sumK = int(gt_offsets[-1])
    assert gt_inds.shape == (B, N) and gt_inds.dtype == torch.int64 and gt_inds.is_cuda and gt_inds.is_contiguous()
    if sumK > 0:
        gts = _boxes(gts, "gt_bboxes")
        assert gts.shape == (sumK, D)
        if gt_labels is not None:
            gt_labels = gt_labels.to(device=dev, dtype=torch.int64).contiguous()
            assert gt_labels.numel() == sumK
    else:
        gts, gt_labels = None, None
    offs = torch.tensor([int(v) for v in gt_offsets], dtype=torch.int32).to(dev, non_blocking=True)
    labels = torch.empty((B, N), dtype=torch.int64, device=dev)
    label_weights = torch.empty((B, N), dtype=torch.float32, device=dev)
    bbox_targets = torch.empty((B, N, D), dtype=torch.float32, device=dev)
    bbox_weights = torch.empty((B, N, D), dtype=torch.float32, device=dev)
    counts = torch.empty((B, 2), dtype=torch.int32, device=dev)
    if B > 0:
        with _on_device(dev):
            _check(lib.sphk_anchor_targets(_ptr(gt_inds), B, N, D, _ptr(anchors), _ptr(gts), _ptr(gt_labels), _ptr(offs), int(num_classes),
                                           float(pos_weight), int(bool(reg_decoded_bbox)), _host5(means, D, 0.0), _host5(stds, D, 1.0),
                                           _ptr(labels), _ptr(label_weights), _ptr(bbox_targets), _ptr(bbox_weights), _ptr(counts),
                                           _stream(anchors)))
        launches += 1
    return labels, label_weights, bbox_targets, bbox_weights, counts


def loss_fwd_bwd(pred, target, grad_iou=None, want_grad_pred=False, want_grad_target=False):
    """Fused Sph2Pob (standard transform) IoU of aligned pairs and its gradients w.r.t. pred / target
    (scaled by grad_iou if given, else d(iou)/d(box) itself)."""
    global launches
    pred, target = _boxes(pred, "pred"), _boxes(target, "target")
    if pred.shape != target.shape:
        raise SphkError("pred/target shapes differ: %s vs %s" % (tuple(pred.shape), tuple(target.shape)))
    n, dev = pred.size(0), pred.device
    iou = torch.empty(n, dtype=torch.float32, device=dev)
    gp = torch.empty_like(pred) if want_grad_pred else None
    gt = torch.empty_like(target) if want_grad_target else None
    if grad_iou is not None:
        grad_iou = grad_iou.to(device=dev, dtype=torch.float32).contiguous()
        assert grad_iou.numel() == n
    with _on_device(dev):
        _check(lib.sphk_loss_fwd_bwd(_ptr(pred), _ptr(target), n, pred.size(1), _ptr(iou), _ptr(grad_iou), _ptr(gp),
                                     _ptr(gt), _stream(pred)))
    launches += 1
    return iou, gp, gt


def loss_reduce(pred, target, weight, scale, want_grad_pred=False, want_grad_target=False):
    """partial sums of weight * (1 - iou) (loss = scale * partial.sum()) and the gradients of that loss."""
    global launches
    pred, target = _boxes(pred, "pred"), _boxes(target, "target")
    if pred.shape != target.shape:
        raise SphkError("pred/target shapes differ: %s vs %s" % (tuple(pred.shape), tuple(target.shape)))
    n, dev = pred.size(0), pred.device
    if weight is not None:
        weight = weight.to(device=dev, dtype=torch.float32).contiguous()
        assert weight.numel() == n
    partial = torch.empty(max(1, int(lib.sphk_loss_reduce_partials(n))), dtype=torch.float32, device=dev)
    if n == 0:
        partial.zero_()
    gp = torch.empty_like(pred) if want_grad_pred else None
    gt = torch.empty_like(target) if want_grad_target else None
    with _on_device(dev):
        _check(lib.sphk_loss_reduce(_ptr(pred), _ptr(target), _ptr(weight), n, pred.size(1), float(scale), _ptr(partial), _ptr(gp),
                                    _ptr(gt), _stream(pred)))
    launches += 1
    return partial, gp, gt


_loss_scratch = {}


def loss_reduce_total(pred, target, weight, scale, want_grad_pred=False, want_grad_target=False):
    """(loss, grad_pred, grad_target): loss = scale * sum_i weight_i (1 - iou_i) as a 0-dim tensor written by the kernel
    itself (ONE launch, no reduction op afterwards) and the gradients of that loss.  The scratch buffer (per-block sums
    + ticket counter) is kept per (device, stream): every user of it is ordered on that stream, and the kernel hands the
    counter back at zero."""
    global launches
    pred, target = _boxes(pred, "pred"), _boxes(target, "target")
    if pred.shape != target.shape:
        raise SphkError("pred/target shapes differ: %s vs %s" % (tuple(pred.shape), tuple(target.shape)))
    n, dev = pred.size(0), pred.device
    if weight is not None:
        weight = weight.to(device=dev, dtype=torch.float32).contiguous()
        assert weight.numel() == n
    need = int(lib.sphk_loss_total_scratch_bytes(n))
    key = (dev.index, _raw_stream(dev.index))
    scratch = _loss_scratch.get(key)
    if scratch is None or scratch.numel() < need:
        scratch = torch.zeros(max(need * 2, 1 << 12), dtype=torch.uint8, device=dev)
        _loss_scratch[key] = scratch
    region = scratch                       # ticket counter in its first 16 bytes, zero between calls
    total = torch.empty((), dtype=torch.float32, device=dev)
    gp = torch.empty_like(pred) if want_grad_pred else None
    gt = torch.empty_like(target) if want_grad_target else None
    with _on_device(dev):
        _check(lib.sphk_loss_reduce_total(_ptr(pred), _ptr(target), _ptr(weight), n, pred.size(1), float(scale), _ptr(total),
                                          _ptr(region), _ptr(gp), _ptr(gt), _stream(pred)))
    launches += 1
    return total, gp, gt


LOSS_KIND = {"gwd": 0, "kld": 1, "jd": 2, "kld_symmax": 3, "kld_symmin": 4, "kfiou": 5, "l1": 6}


def obb_loss(loss_kind, pred, target, upstream=None, scale=1.0, fun=0, flags=0, tau=0.0, alpha=1.0, beta=1.0 / 9.0, eps=1e-6,
             transform="sph2pob_standard", want_loss=True, want_partial=False, want_grad_pred=False, want_grad_target=False):
    """GD / KF / L1 loss rows on the Sph2Pob OBBs of aligned (pred, target), forward + backward in one launch
    (sphk_obb_loss).  Returns (loss [n] or [n, 5] | None, partial | None, grad_pred | None, grad_target | None)."""
    global launches
    pred, target = _boxes(pred, "pred"), _boxes(target, "target")
    if pred.shape != target.shape:
        raise SphkError("pred/target shapes differ: %s vs %s" % (tuple(pred.shape), tuple(target.shape)))
    n, dev = pred.size(0), pred.device
    kind = LOSS_KIND[loss_kind]
    L = 5 if kind == 6 else 1
    up_cols = 0
    if upstream is not None:
        upstream = upstream.to(device=dev, dtype=torch.float32).contiguous()
        if upstream.numel() == n * L and L > 1:
            up_cols = L
        elif upstream.numel() == n:
            up_cols = 1
        else:
            raise SphkError("upstream has %d elements for %d rows of %d loss columns" % (upstream.numel(), n, L))
    loss = torch.empty((n, L) if L > 1 else (n,), dtype=torch.float32, device=dev) if want_loss else None
    partial = None
    if want_partial:
        partial = torch.empty(max(1, int(lib.sphk_loss_reduce_partials(n))), dtype=torch.float32, device=dev)
        if n == 0:
            partial.zero_()
    gp = torch.empty_like(pred) if want_grad_pred else None
    gt = torch.empty_like(target) if want_grad_target else None
    with _on_device(dev):
        _check(lib.sphk_obb_loss(kind, int(fun), int(flags), float(tau), float(alpha), float(beta), float(eps), KIND[transform],
                                 _ptr(pred), _ptr(target), n, pred.size(1), _ptr(upstream), up_cols, float(scale), _ptr(loss),
                                 _ptr(partial), _ptr(gp), _ptr(gt), _stream(pred)))
    launches += 1
    return loss, partial, gp, gt


def obb_loss_total(loss_kind, pred, target, upstream=None, scale=1.0, fun=0, flags=0, tau=0.0, alpha=1.0, beta=1.0 / 9.0, eps=1e-6,
                   transform="sph2pob_standard", want_grad_pred=False, want_grad_target=False):
    """(total, grad_pred, grad_target): total = scale * sum(upstream * loss rows) as a 0-dim tensor written by the launch
    itself (sphk_obb_loss_total: one launch, no reduction op afterwards) and the gradients of that total.  Scratch buffer as
    loss_reduce_total: per (device, stream), the kernel hands the ticket counter back at zero."""
    global launches
    pred, target = _boxes(pred, "pred"), _boxes(target, "target")
    if pred.shape != target.shape:
        raise SphkError("pred/target shapes differ: %s vs %s" % (tuple(pred.shape), tuple(target.shape)))
    n, dev = pred.size(0), pred.device
    kind = LOSS_KIND[loss_kind]
    L = 5 if kind == 6 else 1
    up_cols = 0
    if upstream is not None:
        upstream = upstream.to(device=dev, dtype=torch.float32).contiguous()
        if upstream.numel() == n * L and L > 1:
            up_cols = L
        elif upstream.numel() == n:
            up_cols = 1
        else:
            raise SphkError("upstream has %d elements for %d rows of %d loss columns" % (upstream.numel(), n, L))
    need = int(lib.sphk_loss_total_scratch_bytes(n))
    key = (dev.index, _raw_stream(dev.index))
    scratch = _loss_scratch.get(key)
    if scratch is None or scratch.numel() < need:
        scratch = torch.zeros(max(need * 2, 1 << 12), dtype=torch.uint8, device=dev)
        _loss_scratch[key] = scratch
    total = torch.empty((), dtype=torch.float32, device=dev)
    gp = torch.empty_like(pred) if want_grad_pred else None
    gt = torch.empty_like(target) if want_grad_target else None
    with _on_device(dev):
        _check(lib.sphk_obb_loss_total(kind, int(fun), int(flags), float(tau), float(alpha), float(beta), float(eps), KIND[transform],
                                       _ptr(pred), _ptr(target), n, pred.size(1), _ptr(upstream), up_cols, float(scale), _ptr(total),
                                       _ptr(scratch), _ptr(gp), _ptr(gt), _stream(pred)))
    launches += 1
    return total, gp, gt


def _host5(values, D, default):
    v = list(values) if values is not None else [default] * D
    if len(v) < D:
        raise SphkError("coder: %d means/stds given for %d-column boxes" % (len(v), D))
    return (ctypes.c_float * 5)(*([float(x) for x in v[:D]] + [default] * (5 - D)))


def coder_decode(rois, deltas, means=None, stds=None, wh_ratio_clip=16 / 1000, clip_border=True, add_ctr_clamp=False,
                 ctr_clamp=32, grad_out=None):
    """delta2bbox(rois, deltas) [n, D]; with grad_out = d(total)/d(decoded) it returns d(total)/d(deltas) instead."""
    global launches
    rois, deltas = _boxes(rois, "rois"), _boxes(deltas, "deltas")
    if rois.shape != deltas.shape:
        raise SphkError("coder: rois %s and deltas %s differ in shape" % (tuple(rois.shape), tuple(deltas.shape)))
    n, D, dev = rois.size(0), rois.size(1), rois.device
    out = torch.empty_like(rois)
    m, sd = _host5(means, D, 0.0), _host5(stds, D, 1.0)
    with _on_device(dev):
        if grad_out is None:
            _check(lib.sphk_coder_decode(_ptr(rois), _ptr(deltas), n, D, m, sd, float(wh_ratio_clip), int(bool(clip_border)),
                                         int(bool(add_ctr_clamp)), float(ctr_clamp), _ptr(out), _stream(rois)))
        else:
            grad_out = grad_out.to(device=dev, dtype=torch.float32).contiguous()
            assert grad_out.shape == rois.shape
            _check(lib.sphk_coder_decode_bwd(_ptr(rois), _ptr(deltas), _ptr(grad_out), n, D, m, sd, float(wh_ratio_clip),
                                             int(bool(clip_border)), int(bool(add_ctr_clamp)), float(ctr_clamp), _ptr(out),
                                             _stream(rois)))
    launches += 1
    return out


def coder_encode(proposals, gt, means=None, stds=None):
    """bbox2delta(proposals, gt) [n, D]."""
    global launches
    proposals, gt = _boxes(proposals, "proposals"), _boxes(gt, "gt")
    if proposals.shape != gt.shape:
        raise SphkError("coder: proposals %s and gt %s differ in shape" % (tuple(proposals.shape), tuple(gt.shape)))
    n, D, dev = proposals.size(0), proposals.size(1), proposals.device
    out = torch.empty_like(proposals)
    with _on_device(dev):
        _check(lib.sphk_coder_encode(_ptr(proposals), _ptr(gt), n, D, _host5(means, D, 0.0), _host5(stds, D, 1.0), _ptr(out),
                                     _stream(proposals)))
    launches += 1
    return out


BOX_FORMAT = {"xyxy2xywh": 0, "xywh2xyxy": 1, "obb2hbb_xywh": 2, "obb2hbb_xyxy": 3, "bfov2rbfov": 4, "geo2sph": 5, "sph2geo": 6,
              "sph2pix": 7, "pix2sph": 8, "sph2tan": 9, "tan2sph": 10, "sph2planar_pix": 11, "sph2planar_tan": 12,
              "planar2sph_pix": 13, "planar2sph_tan": 14}


def box_format(fmt, boxes, d_out, img_size=(512, 1024)):
    """One launch of k_box_format: boxes [n, d_in] float32 -> [n, d_out] (sphdet/bbox/box_formator.py conversions)."""
    global launches
    if not boxes.is_cuda:
        raise SphkError("box_format: boxes must be a CUDA tensor (no CPU path)")
    if boxes.dim() != 2 or boxes.size(1) not in (4, 5):
        raise SphkError("box_format: boxes must be [n, 4] or [n, 5], got %s" % (tuple(boxes.shape),))
    b = boxes.detach()
    if b.dtype != torch.float32 or not b.is_contiguous():
        b = b.float().contiguous()
    out = torch.empty((b.size(0), d_out), dtype=torch.float32, device=b.device)
    with _on_device(b.device):
        _check(lib.sphk_box_format(BOX_FORMAT[fmt], _ptr(b), b.size(0), b.size(1), int(d_out), float(img_size[0]), float(img_size[1]),
                                   _ptr(out), _stream(b)))
    launches += 1
    return out if out.dtype == boxes.dtype else out.to(boxes.dtype)


def decode_loss_reduce(anchors, deltas, target, weight, scale, want_grad=True, means=None, stds=None, wh_ratio_clip=16 / 1000,
                       clip_border=True, add_ctr_clamp=False, ctr_clamp=32):
    """One launch for the head's regression loss with decoded boxes: per-CTA partial sums of weight * (1 - iou)
    (loss = scale * partial.sum()) and d(loss)/d(deltas).  weight: None, [n] or [n, k] (row mean)."""
    global launches
    anchors, deltas, target = _boxes(anchors, "anchors"), _boxes(deltas, "deltas"), _boxes(target, "target")
    if not (anchors.shape == deltas.shape == target.shape):
        raise SphkError("decode+loss: anchors %s, deltas %s, target %s differ in shape" %
                        (tuple(anchors.shape), tuple(deltas.shape), tuple(target.shape)))
    n, D, dev = anchors.size(0), anchors.size(1), anchors.device
    wcols = 0
    if weight is not None:
        weight = weight.to(device=dev, dtype=torch.float32).contiguous()
        wcols = 1 if weight.dim() == 1 else weight.size(1)
        if weight.size(0) != n or weight.dim() > 2:
            raise SphkError("decode+loss: weight %s does not match %d rows" % (tuple(weight.shape), n))
    partial = torch.empty(max(1, lib.sphk_decode_loss_partials(n)), dtype=torch.float32, device=dev)
    if n == 0:
        partial.zero_()
    grad = torch.empty_like(deltas) if want_grad else None
    with _on_device(dev):
        _check(lib.sphk_decode_loss_reduce(_ptr(anchors), _ptr(deltas), _ptr(target), _ptr(weight), wcols, n, D,
                                           _host5(means, D, 0.0), _host5(stds, D, 1.0), float(wh_ratio_clip),
                                           int(bool(clip_border)), int(bool(add_ctr_clamp)), float(ctr_clamp), float(scale),
                                           _ptr(partial), _ptr(grad), _stream(anchors)))
    launches += 1
    return partial, grad


def obb_fwd(kind: str, b1, b2, edge="arc"):
    global launches
    b1, b2 = _boxes(b1, "bboxes1"), _boxes(b2, "bboxes2")
    assert b1.shape == b2.shape
    n, dev = b1.size(0), b1.device
    o1 = torch.empty((n, 5), dtype=torch.float32, device=dev)
    o2 = torch.empty((n, 5), dtype=torch.float32, device=dev)
    with _on_device(dev):
        _check(lib.sphk_obb_fwd(KIND[kind], _ptr(b1), _ptr(b2), n, b1.size(1), EDGE[edge], _ptr(o1), _ptr(o2), _stream(b1)))
    launches += 1
    return o1, o2


def obb_bwd(kind: str, b1, b2, g1, g2, edge="arc", want1=True, want2=True):
    global launches
    b1, b2 = _boxes(b1, "bboxes1"), _boxes(b2, "bboxes2")
    n, dev = b1.size(0), b1.device
    g1 = None if g1 is None else g1.to(device=dev, dtype=torch.float32).contiguous()
    g2 = None if g2 is None else g2.to(device=dev, dtype=torch.float32).contiguous()
    gb1 = torch.empty_like(b1) if want1 else None
    gb2 = torch.empty_like(b2) if want2 else None
    with _on_device(dev):
        _check(lib.sphk_obb_bwd(KIND[kind], _ptr(b1), _ptr(b2), n, b1.size(1), EDGE[edge], _ptr(g1), _ptr(g2), _ptr(gb1),
                                _ptr(gb2), _stream(b1)))
    launches += 1
    return gb1, gb2


def riou_fwd_bwd(o1, o2, grad_iou=None, want1=False, want2=False):
    global launches
    if not (o1.is_cuda and o2.is_cuda):
        raise SphkError("rotated IoU: OBB tensors must be CUDA tensors (no CPU fallback)")
    o1, o2 = o1.float().contiguous(), o2.float().contiguous()
    assert o1.shape == o2.shape and o1.dim() == 2 and o1.size(1) == 5
    n, dev = o1.size(0), o1.device
    iou = torch.empty(n, dtype=torch.float32, device=dev)
    g1 = torch.empty_like(o1) if want1 else None
    g2 = torch.empty_like(o2) if want2 else None
    if grad_iou is not None:
        grad_iou = grad_iou.to(device=dev, dtype=torch.float32).contiguous()
    with _on_device(dev):
        _check(lib.sphk_riou_fwd_bwd(_ptr(o1), _ptr(o2), n, _ptr(iou), _ptr(grad_iou), _ptr(g1), _ptr(g2), _stream(o1)))
    launches += 1
    return iou, g1, g2


# SphNMS's iou_calculator names (sph_nms.py:8-16); 'planar' = naive_iou with mmcv nms's rule (SPHK_NMS_RULE_GT: a NaN IoU keeps)
NMS_KIND = {"sph2pob_efficient": 0, "naive_iou": 4, "unbiased_iou": 5, "planar": 4 | 0x100}


def nms_batched(boxes, order, seg_offsets, max_seg_len: int, iou_threshold: float, typical_seg_len: int = 0,
                iou_calculator: str = "sph2pob_efficient") -> torch.Tensor:
    """keep flags (uint8, aligned with `order`) of the greedy per-segment spherical NMS."""
    global launches
    boxes = _boxes(boxes, "boxes")
    dev = boxes.device
    order = order.to(device=dev, dtype=torch.int32).contiguous()
    seg_offsets = seg_offsets.to(device=dev, dtype=torch.int32).contiguous()
    S = seg_offsets.numel() - 1
    keep = torch.zeros(order.numel(), dtype=torch.uint8, device=dev)
    if S <= 0 or order.numel() == 0:
        return keep
    with _on_device(dev):
        _check(lib.sphk_nms_batched(_ptr(boxes), _ptr(order), _ptr(seg_offsets), S, int(max_seg_len), int(typical_seg_len), boxes.size(1),
                                    NMS_KIND[iou_calculator], float(iou_threshold), _ptr(keep), _stream(boxes)))
    launches += 1
    return keep


def nms_images(boxes, scores, labels, num_images: int, num_classes: int, iou_threshold: float, max_out: int, valid=None,
               iou_calculator: str = "sph2pob_efficient"):
    """Greedy per-(image, class) NMS of a batch laid out as `num_images` equal blocks of candidates, entirely on the
    device (three launches, no sort on the host side, no synchronisation).  Returns (idx [num_images, max_out] int32
    into boxes, score-descending per image, -1 padded; count [num_images] int32)."""
    global launches
    boxes = _boxes(boxes, "boxes")
    dev = boxes.device
    M = boxes.size(0)
    if num_images <= 0 or M % num_images != 0:
        raise SphkError("nms_images: %d boxes are not %d equal blocks" % (M, num_images))
    per_image = M // num_images
    scores = scores.to(device=dev, dtype=torch.float32).contiguous()
    labels = labels.to(device=dev, dtype=torch.int64).contiguous()
    if scores.numel() != M or labels.numel() != M:
        raise SphkError("nms_images: scores / labels do not match the boxes")
    if valid is not None:
        valid = valid.to(device=dev, dtype=torch.uint8).contiguous()
        assert valid.numel() == M
    out_idx = torch.empty((num_images, max_out), dtype=torch.int32, device=dev)
    out_count = torch.empty(num_images, dtype=torch.int32, device=dev)
    ws = _workspace(dev, lib.sphk_nms_images_workspace_bytes(num_images, per_image, num_classes))
    with _on_device(dev):
        _check(lib.sphk_nms_images(_ptr(boxes), _ptr(scores), _ptr(labels), _ptr(valid), num_images, per_image, num_classes,
                                   boxes.size(1), NMS_KIND[iou_calculator], float(iou_threshold), int(max_out), _ptr(out_idx),
                                   _ptr(out_count), _ptr(ws), _stream(boxes)))
    launches += 3
    return out_idx, out_count


def probe_fp32(blocks: int, iters: int, device) -> float:
    """Runs the FMA-chain probe once; returns the flop count of the launch (time it with CUDA events)."""
    sink = torch.empty(blocks * 256, dtype=torch.float32, device=device)
    with _on_device(sink.device):
        _check(lib.sphk_probe_fp32(blocks, iters, _ptr(sink), _stream(sink)))
    return 2.0 * 8 * iters * 256 * blocks


def prefilter_live_pairs(rows, cols, edge="arc") -> int:
    """Measurement helper: how many of the R x C pairs survive the prefilter of the N x M kernels (host int; syncs)."""
    rows, cols = _boxes(rows, "bboxes1"), _boxes(cols, "bboxes2")
    R, C, dev = rows.size(0), cols.size(0), rows.device
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = _workspace(dev, 136 * (R + C) + 32)
    with _on_device(dev):
        _check(lib.sphk_prefilter_count(_ptr(rows), R, _ptr(cols), C, rows.size(1), EDGE[edge], _ptr(cnt), _ptr(ws), _stream(rows)))
    return int(cnt.item())


def set_dense(on: bool) -> bool:
    """Disable (True) / enable (False) the disjoint-pair early-outs; returns the previous setting."""
    return bool(lib.sphk_set_dense(1 if on else 0))
