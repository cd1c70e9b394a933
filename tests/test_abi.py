"""The drop-in boundary without a GPU: the shared library loads, exports exactly the symbols include/slideo_b200.h
declares, struct layouts agree between C and the binding, and compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "slideo_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(slideo_b200_[a-z0-9_]+)\s*\(", src)))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    from slideo_b200 import ffi
    lib = ffi.load()
    names = _declared()
    assert len(names) >= 28
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(ffi.SYMBOLS) == names, "ffi.SYMBOLS and the header disagree"
    out = subprocess.check_output(["nm", "-D", "--defined-only", ffi.LIB_PATH], text=True)
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert exported == names, "the library exports symbols the header does not declare (or vice versa)"


def test_struct_layouts_match_c(tmp_path):
    from slideo_b200 import ffi
    c = tmp_path / "sz.c"
    c.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "slideo_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %d\\n",'
                 "sizeof(slideo_b200_config), sizeof(slideo_b200_frame_result), sizeof(slideo_b200_match), sizeof(slideo_b200_timings),"
                 "offsetof(slideo_b200_config, vote_ratio), offsetof(slideo_b200_timings, knn_pairs), SLIDEO_B200_ABI_VERSION);return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    vals = [int(v) for v in subprocess.check_output([str(exe)], text=True).split()]
    assert vals == [ctypes.sizeof(ffi.Config), ctypes.sizeof(ffi.FrameResult), ctypes.sizeof(ffi.Match), ctypes.sizeof(ffi.Timings),
                    ffi.Config.vote_ratio.offset, ffi.Timings.knn_pairs.offset, ffi.ABI_VERSION]


def test_default_config_is_the_reference_literals():
    import slideo_b200
    cfg = slideo_b200.default_config()
    # feature_extractor.rs:13-23, lib.rs:266, lib.rs:275
    assert (cfg.nfeatures, cfg.nlevels, cfg.edge_threshold, cfg.patch_size, cfg.fast_threshold, cfg.knn_k) == (2000, 8, 62, 62, 20, 30)
    assert cfg.scale_factor == np.float32(1.2) and cfg.vote_ratio == np.float32(1.05)
    assert cfg.descriptor_kind == slideo_b200.ffi.DESC_ORB256
    with pytest.raises(TypeError):
        slideo_b200.default_config(no_such_field=1)


def test_version_and_null_handling():
    from slideo_b200 import ffi
    lib = ffi.load()
    assert b"sm_100a" in lib.slideo_b200_version()
    assert lib.slideo_b200_default_config(None) == ffi.E_INVALID_ARG
    assert lib.slideo_b200_destroy(None) == ffi.OK
    assert lib.slideo_b200_finalize_pool(None) == ffi.E_INVALID_ARG
    assert lib.slideo_b200_match_frames_bgr8(None, None, 0, 0, 0, 0, 0, None) == ffi.E_INVALID_ARG


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    import slideo_b200
    with pytest.raises(slideo_b200.SlideoError) as e:
        slideo_b200.Context()
    assert e.value.status == slideo_b200.ffi.E_CUDA
    assert "no CPU fallback" in e.value.message


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under slideo_b200/ may import or link it."""
    pkg = os.path.join(ROOT, "slideo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt and "orb_oracle" not in txt.replace("oracle/orb_oracle.c", ""), f
    code = "import sys; sys.path.insert(0, %r); import slideo_b200; assert 'oracle' not in sys.modules and 'cv2' not in sys.modules" % ROOT
    subprocess.check_call([sys.executable, "-c", code])


def test_tools_and_bench_compile():
    """bench.py, __graft_entry__.py and every developer tool under tools/ at least parse (they only run on the GPU box)."""
    import glob
    import py_compile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = [os.path.join(root, "bench.py"), os.path.join(root, "__graft_entry__.py")] + sorted(glob.glob(os.path.join(root, "tools", "*.py")))
    assert len(files) > 5
    for f in files:
        py_compile.compile(f, doraise=True)
