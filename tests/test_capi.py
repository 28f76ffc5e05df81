"""CPU: the C-ABI library loads and exports every symbol include/orbx_b200.h declares; host-side error behaviour
that needs no GPU.  (No compute calls here: compute needs a B200 and lives in the -m gpu tests.)"""
import ctypes
import os
import re
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    syms = set()
    for name in sorted(os.listdir(os.path.join(ROOT, "include"))):           # the product ABI and the test-tap header
        if name.endswith(".h"):
            hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", name)).read(), flags=re.S)
            syms |= set(re.findall(r"\b(orbx_[a-z0-9_]+)\s*\(", hdr))
    return sorted(syms)


def test_product_header_has_no_test_taps():
    hdr = open(os.path.join(ROOT, "include", "orbx_b200.h")).read()
    assert "orbx_debug_" not in hdr


def test_library_exports_every_declared_symbol(orbx):
    assert os.path.exists(orbx.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(orbx.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_keypoint_layout_matches_cv_keypoint(orbx):
    assert orbx.KP_DTYPE.itemsize == 28
    assert [orbx.KP_DTYPE.fields[n][1] for n in ("x", "y", "size", "angle", "response", "octave", "class_id")] == [0, 4, 8, 12, 16, 20, 24]


def test_no_cpu_fallback(orbx):
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(orbx.OrbxError) as e:
        orbx.ORBextractor(1000, 1.2, 8, 20, 7)
    assert e.value.code == orbx.E_CUDA
    with pytest.raises(orbx.OrbxError):
        orbx.ORBmatcher(0.9, True)


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU implementation)."""
    pkg = os.path.join(ROOT, "amos-slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inl", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace("the oracle", ""), f
                assert "cvlite" not in src or f.endswith(".hpp"), f


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (no C++-isms, no torch / CUDA types) and a C program must link against the library."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "orbx_b200.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    src = tmp_path / "link.c"
    src.write_text('#include "orbx_b200.h"\n#include <stdio.h>\nint main(void) { orbx_extractor* h = 0; int rc = orbx_create(1000, 1.2f, 8, 20, 7, 0, &h);\n'
                   '  printf("%d %s\\n", rc, rc ? orbx_last_error() : "ok"); if (!rc) orbx_destroy(h); return 0; }\n')
    lib_dir = os.path.join(ROOT, "amos-slam_b200")
    exe = str(tmp_path / "link")
    subprocess.check_call(["gcc", "-std=c99", "-I" + os.path.join(ROOT, "include"), str(src), "-L" + lib_dir, "-lorbx_b200", "-Wl,-rpath," + lib_dir, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip()           # without a GPU: a negative status and a message, never a crash or a CPU result
