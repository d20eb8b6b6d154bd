"""The C-ABI library loads on a machine without a GPU and exports every symbol include/sphk.h
declares; the ctypes binding covers all of them.  No compute call is made here."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sphk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sphk_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built_lib():
    from sph_retina_b200 import build
    return build.build()


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for must in ("sphk_iou_aligned", "sphk_iou_pairwise", "sphk_loss_fwd_bwd", "sphk_nms_batched", "sphk_obb_fwd",
                 "sphk_obb_bwd", "sphk_riou_fwd_bwd", "sphk_coder_decode", "sphk_coder_encode", "sphk_decode_loss_reduce",
                 "sphk_last_error_string", "sphk_abi_version"):
        assert must in names


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for name in declared_symbols():
        assert hasattr(lib, name), "libsphk.so does not export %s" % name
    lib.sphk_abi_version.restype = ctypes.c_int
    assert lib.sphk_abi_version() == 8


def test_binding_covers_the_header(built_lib):
    from sph_retina_b200 import _native
    assert sorted(_native.SIGNATURES) == declared_symbols()


def test_argument_validation_needs_no_gpu(built_lib):
    """Invalid arguments are rejected before any CUDA call, with a message."""
    from sph_retina_b200 import _native
    lib = _native.lib
    assert lib.sphk_iou_aligned(0, None, None, 10, 3, 0, 0, 0, None, None) == -1          # D = 3
    assert b"D not in" in lib.sphk_last_error_string()
    assert lib.sphk_iou_aligned(2, None, None, 10, 5, 0, 0, 0, None, None) == -3          # sph_iou on RBFoV
    assert lib.sphk_iou_aligned(7, None, None, 10, 4, 0, 0, 0, None, None) == -1          # unknown kind
    assert lib.sphk_iou_aligned(0, None, None, 0, 4, 0, 0, 0, None, None) == 0            # empty is fine
    assert lib.sphk_iou_pairwise_workspace_bytes(10, 20) == 240 + 30 * 128
    assert lib.sphk_nms_batched(None, None, None, 0, 0, 0, 4, 0, 0.5, None, None) == 0
    assert lib.sphk_unpack_gathered_keys(None, 0, 10, 4, 10, None, None, None, None, None) == -1    # world < 1
    assert lib.sphk_unpack_gathered_keys(None, 2, 10, 4, 4, None, None, None, None, None) == -1     # cap < ceil(n / world)
    assert lib.sphk_unpack_gathered_keys(None, 2, 0, 0, 0, None, None, None, None, None) == 0       # nothing to do
    assert lib.sphk_unpack_peer_keys(None, 0, 17, 1, 0, 0, 8, 2, 8, 1, 0, None, None, None, None, None) == -1   # more than 16 ranks
    assert lib.sphk_unpack_peer_keys(None, 0, 2, 0, 0, 0, 8, 2, 4, 1, 0, None, None, None, None, None) == -1    # steps count from 1
    assert lib.sphk_unpack_peer_keys(None, 2, 2, 1, 0, 0, 8, 2, 4, 1, 0, None, None, None, None, None) == -1    # rank >= world
    assert lib.sphk_unpack_peer_keys(None, 0, 2, 1, 0, 0, 8, 2, 4, 4, 0, None, None, None, None, None) == -1    # parts > 1 without push
    assert lib.sphk_iou_pairwise_keys_push(0, None, 8, None, 8, 5, 0, 0, None, 0, 0, None, 2, 0, 8, None, None) == -1   # null pointers
    assert lib.sphk_key_push_parts(1024) == 4 and lib.sphk_key_push_parts(1025) == 5 and lib.sphk_key_push_parts(0) == 1
    assert lib.sphk_nms_batched(None, None, None, 3, 9, 0, 4, 1, 0.5, None, None) == -3    # NMS calculator: efficient, naive or unbiased
    assert lib.sphk_iou_aligned(4, None, None, 10, 5, 1, 0, 0, None, None) == -3          # naive_iou: mode 'iou' only
    assert lib.sphk_iou_aligned(5, None, None, 10, 4, 1, 0, 0, None, None) == -3          # unbiased_iou: mode 'iou' only
    assert lib.sphk_iou_aligned(6, None, None, 10, 5, 0, 0, 0, None, None) == -3          # sph2pob_legacy_iou: BFoV only
    assert lib.sphk_iou_aligned(7, None, None, 10, 4, 0, 0, 0, None, None) == -1          # unknown kind
    assert lib.sphk_loss_fwd_bwd(None, None, 5, 4, None, None, None, None, None) == -1  # null boxes
    assert lib.sphk_box_format(15, None, 0, 4, 4, 512.0, 1024.0, None, None) == -1        # unknown format
    assert lib.sphk_box_format(2, None, 0, 4, 4, 512.0, 1024.0, None, None) == -1         # obb2hbb takes 5 columns
    assert lib.sphk_box_format(11, None, 0, 5, 5, 512.0, 1024.0, None, None) == 0         # empty is fine
    assert lib.sphk_box_format(13, None, 3, 4, 5, 512.0, 1024.0, None, None) == -1        # null pointers


def test_sass_is_sm100a(built_lib):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out
