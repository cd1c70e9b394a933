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
    c.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "slideo_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %d %zu %zu %zu\\n",'
                 "sizeof(slideo_b200_config), sizeof(slideo_b200_frame_result), sizeof(slideo_b200_match), sizeof(slideo_b200_timings),"
                 "offsetof(slideo_b200_config, vote_ratio), offsetof(slideo_b200_timings, knn_pairs), SLIDEO_B200_ABI_VERSION,"
                 "sizeof(slideo_b200_verify_result), sizeof(slideo_b200_decision), offsetof(slideo_b200_decision, refined_matrix));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    vals = [int(v) for v in subprocess.check_output([str(exe)], text=True).split()]
    assert vals == [ctypes.sizeof(ffi.Config), ctypes.sizeof(ffi.FrameResult), ctypes.sizeof(ffi.Match), ctypes.sizeof(ffi.Timings),
                    ffi.Config.vote_ratio.offset, ffi.Timings.knn_pairs.offset, ffi.ABI_VERSION,
                    ctypes.sizeof(ffi.VerifyResult), ctypes.sizeof(ffi.Decision), ffi.Decision.refined_matrix.offset]


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


# ---- the Rust binding source (integration/matching-b200/src/ffi.rs) against the header ----------------------------------------------
FFI_RS = os.path.join(ROOT, "integration", "matching-b200", "src", "ffi.rs")


def _c_functions():
    """name -> number of parameters, from the header."""
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int32_t|const char\*)\s+(slideo_b200_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def _c_structs():
    """name -> [(field, C type, array dims)] from the header."""
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    out = {}
    for m in re.finditer(r"typedef struct (slideo_b200_[a-z_]+) \{(.*?)\} \1;", src, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            fm = re.match(r"([a-z0-9_]+)\s+([a-z0-9_]+)((?:\[[A-Z0-9_]+\])*)$", decl)
            assert fm, decl
            fields.append((fm.group(2), fm.group(1), re.findall(r"\[([A-Z0-9_]+)\]", fm.group(3))))
        out[m.group(1)] = fields
    return out


def _rs_functions():
    src = re.sub(r"//[^\n]*", "", open(FFI_RS).read())
    out = {}
    for m in re.finditer(r"pub fn (slideo_b200_[a-z0-9_]+)\s*\((.*?)\)\s*->", src, flags=re.S):
        args = [a for a in m.group(2).split(",") if a.strip()]
        out[m.group(1)] = len(args)
    return out


def _rs_structs():
    src = re.sub(r"//[^\n]*", "", open(FFI_RS).read())
    out = {}
    for m in re.finditer(r"#\[repr\(C\)\]\s*(?:#\[derive\([^)]*\)\]\s*)?pub struct (slideo_b200_[a-z_]+) \{(.*?)\n\}", src, flags=re.S):
        fields = []
        for fm in re.finditer(r"pub ([a-z0-9_]+):\s*([^,\n]+),", m.group(2)):
            fields.append((fm.group(1), fm.group(2).strip()))
        out[m.group(1)] = fields
    return out


def test_rust_binding_declares_every_function_with_the_same_arity():
    c, rs = _c_functions(), _rs_functions()
    assert sorted(c) == _declared()
    assert sorted(rs) == sorted(c), "ffi.rs and the header declare different functions"
    for name, n in c.items():
        assert rs[name] == n, f"{name}: {rs[name]} parameters in ffi.rs, {n} in the header"


def test_rust_binding_structs_match_the_header():
    consts = {"SLIDEO_B200_TOP_SLIDES": 40, "SLIDEO_B200_TOP_RATED": 10}
    ctype = {"int32_t": "i32", "float": "f32", "double": "f64", "int64_t": "i64"}
    c, rs = _c_structs(), _rs_structs()
    assert sorted(c) == sorted(rs) and len(c) == 6
    for name, fields in c.items():
        assert [f for f, _, _ in fields] == [f for f, _ in rs[name]], name
        for (fname, t, dims), (_, rtype) in zip(fields, rs[name]):
            want = ctype[t]
            for d in reversed(dims):                      # C a[X][Y] == Rust [[T; Y]; X]
                want = f"[{want}; {d if d in consts else int(d)}]"
            got = rtype
            for k, v in consts.items():
                want, got = want.replace(k, str(v)), got.replace(k, str(v))
            assert got.replace(" ", "") == want.replace(" ", ""), f"{name}.{fname}: {rtype} vs {t}{dims}"
    src = open(FFI_RS).read()
    for k, v in consts.items():
        assert re.search(rf"pub const {k}: usize = {v};", src)
    # status codes and kinds
    hdr = open(HEADER).read()
    for name, val in re.findall(r"(SLIDEO_B200_(?:OK|E_[A-Z_]+|DESC_[A-Z0-9]+)) = (-?\d+)", hdr):
        assert re.search(rf"pub const {name}: i32 = {val};", src), name


def test_rust_crate_uses_the_decision_tail():
    """the crate must decide like the reference (gate chain), not by min_votes."""
    lib_rs = open(os.path.join(ROOT, "integration", "matching-b200", "src", "lib.rs")).read()
    assert "geometric_verification = 2" in lib_rs and "slideo_b200_get_decisions" in lib_rs and "slideo_b200_mark_changed_bgr8" in lib_rs
    assert "min_votes" not in lib_rs
    for mod in re.findall(r"^mod ([a-z_]+);", lib_rs, flags=re.M):
        assert os.path.exists(os.path.join(ROOT, "integration", "matching-b200", "src", mod + ".rs")), f"mod {mod} has no source file"
    assert os.path.exists(os.path.join(ROOT, "integration", "matching-b200", "build.rs"))


def test_committed_bench_lines_follow_the_contract():
    """The bench lines kept under profiles/ (what bench.py printed on the GPU box) carry every key the driver's contract names, and
    their derived figures are consistent (frac = achieved / peak, value = frames per step / step time)."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for name in ("r2_bench_r.json", "r2_bench_q.json", "r2_bench_p.json", "r2_bench_o_8gpu.json"):
        d = json.loads(open(os.path.join(root, "profiles", name)).read().strip().splitlines()[-1])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                  "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
            assert k in d, (name, k)
        assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["warmup"] >= 3 and d["gpu_launches"] > 0
        assert "workload" in d["config"] and "model" not in d["config"]
        for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
            assert k in d["e2e"], (name, k)
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and 0 < d["e2e"]["value"] <= d["value"] * 1.001
        r = d["roofline"]
        for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
            assert k in r, (name, k)
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if d["n_gpus"] == 1:
            c = d["cpu_baseline"]
            for k in ("value", "unit", "cores", "kind", "sample"):
                assert k in c, (name, k)
            assert c["kind"] in ("reference", "port") and "flann_lsh" in c


def test_reference_arm_runs_on_the_host_and_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU path of the reference through cv2, no GPU, no repo kernels) at a tiny size: one JSON line
    with the reference arm's keys (tier contract: impl, cpu_baseline describing this run, e2e with zero copied bytes)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--frames", "4", "--pages", "2"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "1080p frames matched/s" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert "workload" in d["config"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and abs(c["value"] - d["value"]) < 1e-9 * max(1.0, d["value"])
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["unit"] == d["unit"]
    assert "libslideo_b200" not in out.stderr

