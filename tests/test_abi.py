"""The drop-in boundary: libb200vision.so loads, exports every symbol include/b200vision.h
declares (and the legacy process_frame with the reference's signature), and fails loudly when no
CUDA device is usable.  No compute calls here."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200vision.h")


@pytest.fixture(scope="module")
def libpath():
    from cuauv_vision_pipeline_b200 import build
    return build.build()


def declared_functions():
    txt = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(bv_[a-z0-9_]+|process_frame)\s*\(", txt)))


def test_header_declares_the_path():
    names = declared_functions()
    for must in ("bv_create", "bv_color_balance", "bv_cvt_color", "bv_in_range", "bv_morph", "bv_label",
                 "bv_letterbox", "bv_resize_linear", "bv_stage", "bv_stage_host", "process_frame"):
        assert must in names


def test_library_exports_every_declared_symbol(libpath):
    lib = ctypes.CDLL(libpath)
    for name in declared_functions():
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", libpath], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (\w+)", out))
    assert set(declared_functions()) <= exported


def test_legacy_alias_exports_process_frame(libpath):
    legacy = os.path.join(os.path.dirname(libpath), "libauv-color-balance.so")
    assert os.path.exists(legacy)
    assert hasattr(ctypes.CDLL(legacy), "process_frame")


def test_cffi_binding_matches_header():
    from cuauv_vision_pipeline_b200 import _ffi
    assert _ffi.exported_symbols() == declared_functions()
    assert _ffi.lib.bv_version() == 100
    assert _ffi.ffi.sizeof("bv_blob") == 96
    p = _ffi.ffi.new("bv_balance_params *")
    _ffi.lib.bv_balance_default(p)
    # defaults of balance(), modules/color_balance.py:93-96
    assert (p.equalize_rgb, p.rgb_contrast_correct, p.hsv_contrast_correct, p.hsi_contrast_correct,
            p.rgb_extrema_clipping, p.adaptive_cast_correction, p.horizontal_blocks, p.vertical_blocks) == \
        (1, 0, 1, 0, 1, 0, 1, 1)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import cuauv_vision_pipeline_b200 as bv
    with pytest.raises(bv.BVError) as e:
        bv.Context(0)
    assert "no CPU path" in str(e.value)
    # the legacy symbol reports failure through its return code instead of aborting
    import numpy as np
    lib = ctypes.CDLL(os.path.join(ROOT, "cuauv_vision_pipeline_b200", "lib", "libauv-color-balance.so"))
    buf = np.zeros((4, 4, 3), np.uint8)
    assert lib.process_frame(buf.ctypes.data_as(ctypes.c_void_p), 4, 4, 3, True, False, True, False, True, False, 1, 1) != 0


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "cuauv_vision_pipeline_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                if f == "modules.py":
                    # the drop-in modules draw the accepted rectangles into the image they post, as modules/bins.py:71-74
                    # does, with the reference's own cv2 (imported lazily inside _draw_box): drawing only, no pixel path
                    code = "\n".join(ln.split("#")[0] for ln in src.splitlines() if not ln.lstrip().startswith(("#", '"""', "cv2.")))
                    code = re.sub(r'"""[\s\S]*?"""', "", code)
                    assert set(re.findall(r"\bcv2\.(\w+)\(", code)) <= {"drawContours", "boxPoints"}, f
                    assert src.count("import cv2") == 1, f
                else:
                    assert "import cv2" not in src, f


CABI_SRC = os.path.join(ROOT, "tests", "cabi", "example.c")


def _build_c_example(libpath, out):
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), CABI_SRC,
                           "-L", os.path.dirname(libpath), "-lb200vision", "-Wl,-rpath," + os.path.dirname(libpath), "-o", out])


def test_header_is_plain_c99_and_a_c_caller_links(libpath, tmp_path):
    """The boundary is a C ABI: the header compiles as C99 and a C program that uses the legacy symbol and the fused
    host-buffer entry point links against the library (no C++ types, no CUDA headers on the caller's side)."""
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), CABI_SRC])
    _build_c_example(libpath, str(tmp_path / "example"))


@pytest.mark.gpu
def test_c_caller_gets_the_reference_results(libpath, tmp_path):
    """tests/cabi/example.c run as a process: process_frame (modules/color_balance.py:105-107) and the bins stage
    (modules/bins.py:13-27) called from plain C give the reference's bytes."""
    import cv2
    import numpy as np
    from oracle import synth, ref_balance, color_balance_np, cv_ops
    exe = str(tmp_path / "example")
    _build_c_example(libpath, exe)
    img = synth.gen_underwater(480, 640, 4242)
    raw, bal, mask = (str(tmp_path / n) for n in ("frame.raw", "bal.raw", "mask.raw"))
    img.tofile(raw)
    out = subprocess.run([exe, raw, "480", "640", bal, mask], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    want = ref_balance.balance(img) if ref_balance.available() else color_balance_np.process_frame_np(img)
    assert np.array_equal(np.fromfile(bal, np.uint8).reshape(img.shape), want)
    _, cleaned = cv_ops.bins_mask(img)
    assert np.array_equal(np.fromfile(mask, np.uint8).reshape(480, 640), cleaned)
    n_ref = cv2.connectedComponents(cleaned, connectivity=8)[0] - 1
    assert out.stdout.split()[1] == str(n_ref)
