"""The C-ABI library builds, loads on a GPU-less host, and exports every symbol include/usot_b200.h declares."""
import ctypes
import os

import pytest

from usot_b200 import _lib


def test_header_symbols_exported_and_bound(lib):
    syms = _lib.header_symbols()
    assert len(syms) >= 15 and len(set(syms)) == len(syms)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/usot_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in usot_b200/_lib.py"
    for s in _lib.SIGNATURES:
        assert s in syms, f"{s} bound in _lib.py but not declared in the header"


def test_abi_version_and_error_channel(lib):
    assert lib.usot_abi_version() == 3
    assert lib.usot_feature_size(255) == 31 and lib.usot_feature_size(271) == 33 and lib.usot_feature_size(127) == 15
    rc = lib.usot_set_tunable(b"no_such_knob", 1)
    assert rc != 0 and b"unknown tunable" in lib.usot_last_error()


def test_no_gpu_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = lib.usot_engine_create(ctypes.byref(h), 0, 0)
    assert rc != 0 and len(lib.usot_last_error()) > 0


def test_header_is_plain_c_and_a_c_host_links(lib, tmp_path):
    """include/usot_b200.h must be consumable by a C compiler (no C++ / torch types) and a C host must link against the shared
    library: examples/c_abi_demo.c is compiled with -std=c99 -Wall -Wextra -Werror, linked, and run in its ABI-probe mode."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    src, exe = os.path.join(root, "examples", "c_abi_demo.c"), str(tmp_path / "c_abi_demo")
    libdir = os.path.join(root, "usot_b200")
    cudart = next((d for d in ("/usr/local/cuda/lib64", "/usr/local/cuda/targets/x86_64-linux/lib") if os.path.exists(os.path.join(d, "libcudart.so"))), None)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"), "-c", src, "-o", str(tmp_path / "demo.o")], check=True)
    if cudart is None:
        pytest.skip("libcudart.so not found: compiled, not linked")
    subprocess.run([gcc, str(tmp_path / "demo.o"), "-o", exe, "-L", libdir, "-lusot_b200", "-L", cudart, "-lcudart", f"-Wl,-rpath,{libdir}",
                    f"-Wl,-rpath,{cudart}"], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    assert "ABI version 3" in out and "255 crop: 31" in out


def test_no_oracle_import_in_product():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in list(os.walk(os.path.join(root, "usot_b200"))) + list(os.walk(os.path.join(root, "lib"))):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "usot_oracle" not in src and "import oracle" not in src, f"{f} references the oracle"
