"""CPU tests: the C-ABI library loads and exports every symbol include/bshark.h declares (no compute calls)."""
import ctypes
import os
import re

import baby_shark_b200 as B
from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "bshark.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(B.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "libbshark_cuda.so does not export %s" % s
    assert sorted(B.EXPORTS) == syms, "baby_shark_b200.EXPORTS out of sync with include/bshark.h"


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        return
    lib = B.load_library()
    h = ctypes.c_void_p()
    assert lib.bs_context_create(-1, ctypes.byref(h)) == B.BS_ERR_NO_DEVICE
    try:
        B.Context()
    except B.BsharkError as e:
        assert e.status == B.BS_ERR_NO_DEVICE
    else:
        raise AssertionError("Context() must fail loudly without a GPU")


def test_product_does_not_touch_oracle():
    # the product path may not import, link or execute anything under oracle/
    pkg = os.path.join(ROOT, "baby_shark_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower() or f == "bs_stubs.cu" and False, "%s mentions the oracle" % f


def test_rust_ffi_declarations_match_the_header():
    # no Rust toolchain here: rust/src/ffi.rs is generated from include/bshark.h (tools/gen_ffi_rs.py) and must be current --
    # every export declared, same order, same arity and pointer constness; lib.rs may only call what ffi.rs declares
    import re
    import subprocess
    import sys
    assert subprocess.call([sys.executable, os.path.join(ROOT, "tools", "gen_ffi_rs.py"), "--check"]) == 0, "run tools/gen_ffi_rs.py"
    ffi = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    declared = {m.group(1): m.group(2).count(":") for m in re.finditer(r"pub fn (bs_\w+)\((.*?)\)", ffi)}
    assert set(declared) == set(B.EXPORTS)
    lib = open(os.path.join(ROOT, "rust", "src", "lib.rs")).read()
    for name, args in re.findall(r"ffi::(bs_[a-z_0-9]+)\((.*?)\) \}", lib):
        assert name in declared, name
