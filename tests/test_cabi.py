"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/eth3d_b200.h declares;
without a GPU the product path fails loudly (no CPU fallback). No compute calls here."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from dataset_pipeline_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()


def test_header_symbols_exported(L):
    hdr = open(os.path.join(ROOT, "include", "eth3d_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", hdr)) - {"b2_allreduce_fn"})
    assert len(declared) >= 15
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, "header declares symbols the library does not export: %s" % missing
    from dataset_pipeline_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared


def test_abi_version(L):
    assert L.b2_abi_version() == 2


def test_product_path_does_not_touch_oracle():
    """The oracle is test infrastructure: nothing under dataset_pipeline_b200/ may import, include or load it."""
    pkg = os.path.join(ROOT, "dataset_pipeline_b200")
    for dp, _, files in os.walk(pkg):
        if "_build" in dp:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".cc", ".c")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "liboracle" not in txt and "orc_api" not in txt and "from oracle" not in txt and "import oracle" not in txt, \
                    "%s references the oracle" % os.path.join(dp, f)


def test_fails_loudly_without_gpu(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dataset_pipeline_b200 import PointToPlaneICP
    from dataset_pipeline_b200._lib import B2Error
    with pytest.raises(B2Error) as ei:
        PointToPlaneICP()
    assert "NO_DEVICE" in str(ei.value) or "CUDA" in str(ei.value)


def test_cpp_shims_compile_and_fail_loudly(L, tmp_path):
    """The header-only C++ shims (reference class shapes over the C ABI) compile as C++14 — the reference's standard — and, without
    a GPU, surface the library's error as an exception instead of aborting."""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    from dataset_pipeline_b200 import _lib
    exe = str(tmp_path / "shim_check")
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["g++", "-std=c++14", "-Wall", "-Werror", "-I", ROOT, "-o", exe, os.path.join(ROOT, "tests", "cpp", "shim_check.cc"),
           "-L", libdir, "-leth3d_b200", "-Wl,-rpath," + libdir]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
