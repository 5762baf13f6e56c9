"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads without a GPU, exports exactly
the symbols include/vistracker_b200.h declares, and rejects bad arguments with an error string instead of crashing."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from vistracker_b200.build import build
    build()
    from vistracker_b200 import _lib
    return _lib.load()


def _declared():
    src = open(os.path.join(ROOT, "include", "vistracker_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vt_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from vistracker_b200 import _lib
    declared = _declared()
    assert len(declared) >= 14
    assert sorted(_lib.SIGNATURES) == declared, "ctypes table and header disagree"
    for name in declared:
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r"\bT (vt_[a-z0-9_]+)", out)))
    assert exported == declared


def test_library_is_sm100a_only(lib):
    from vistracker_b200 import _lib
    assert lib.vt_version() == 1 and lib.vt_compiled_arch() == 100
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_bad_arguments_are_rejected_without_a_gpu(lib):
    # argument validation happens before any CUDA call, so these run on the CPU-only box
    rc = lib.vt_conv_mma(None, None, 1, 12, 12, 64, 1, 3, None, None, 64, None, None, 0, None, 64, None, 0, None)
    assert rc < 0 and b"W=12" in lib.vt_last_error()
    rc = lib.vt_conv_ffma(None, 64, None, None, 0, 1, 8, 8, 64, 5, None, 64, None, None, 0, None, 64, None, 0, None)
    assert rc < 0 and b"kernel size" in lib.vt_last_error()
    rc = lib.vt_gn_finalize(None, 0, None, None, 1, 48, 32, 1, ctypes.c_float(1e-5), None, None, None)
    assert rc < 0 and b"groups" in lib.vt_last_error()
    cam = (ctypes.c_float * 7)()
    rc = lib.vt_query_fwd(None, None, None, 1, 1, None, None, None, None, 4, 4, 8, 8, 128, 64, 32, 64, cam, None, None, None, None, None)
    assert rc < 0 and b"611" in lib.vt_last_error()


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under vistracker_b200/ may import or execute it."""
    pkg = os.path.join(ROOT, "vistracker_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "oracle/" not in text and "import_module(\"oracle" not in text, f


def test_tools_do_not_touch_the_oracle_either():
    """tools/ are timing / integration drivers of the product path: no oracle imports, no test-helper imports that pull it in."""
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith(".py"):
            text = open(os.path.join(ROOT, "tools", f)).read()
            assert not re.search(r"^\s*(from|import)\s+(oracle|fit_problem|recon_problem)\b", text, flags=re.M), f


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from vistracker_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_model_refuses_cpu_device():
    import torch
    from vistracker_b200 import CHORETriplaneVisibility, default_options
    net = CHORETriplaneVisibility(default_options(), device="cpu")
    with pytest.raises(RuntimeError, match="no CPU path"):
        net.load_state_dict({k: torch.zeros(s) for k, s in net._expected.items()})
