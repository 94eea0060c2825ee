"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/collider_b200.h declares.
No compute calls here (no GPU in this container)."""
import ctypes
import importlib
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build = importlib.import_module("collider-rs_b200.build")
    build.build()
    cb = importlib.import_module("collider-rs_b200")
    return cb, cb.load_library()


def declared_symbols():
    with open(os.path.join(ROOT, "include", "collider_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cb200_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    cb, L = lib
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/collider_b200.h but not exported"
    assert set(cb.EXPORTS) == set(names)
    assert L.cb200_abi_version() == 1


def test_struct_layouts_match_header(lib):
    cb, _ = lib
    assert ctypes.sizeof(cb.Event) == 32
    assert ctypes.sizeof(cb.StepResult) == 8 + 32 + 7 * 8 + 8 + 8 + 8
    assert ctypes.sizeof(cb.StateView) == 16 * 8
    assert ctypes.sizeof(cb.VelView) == 5 * 8


def test_built_for_sm100a_and_fma_free(lib, tmp_path):
    """The f64 contract: sm_100a only, and no fused multiply-add anywhere in the PTX of the engine —
    every product and sum is rounded separately, like the reference's SSE2 code (div/sqrt stay IEEE `div.rn`/`sqrt.rn`)."""
    cb, _ = lib
    out = subprocess.run(["cuobjdump", "-lelf", cb.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    build = importlib.import_module("collider-rs_b200.build")
    ptx = tmp_path / "engine.ptx"
    subprocess.run([build.nvcc(), "-gencode", "arch=compute_100a,code=compute_100a", "-O3", "-std=c++17", "-fmad=false", "-ptx", "-o", str(ptx), build.SRC],
                   check=True, capture_output=True)
    text = ptx.read_text()
    assert "fma.rn.f64" not in text and "mad.rn.f64" not in text
    assert "div.rn.f64" in text and "sqrt.rn.f64" in text and "mul.rn.f64" in text and "add.rn.f64" in text
    assert "-fmad=false" in build.NVCC_FLAGS


def test_create_fails_loudly_without_a_gpu(lib):
    cb, L = lib
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(cb.ColliderError) as e:
        cb.Engine(4.0, 0.01)
    assert e.value.code == cb.E_CUDA and "no CPU fallback" in e.value.msg
    # argument validation happens before any device work: Collider::new's asserts (collider.rs:60-61)
    with pytest.raises(cb.ColliderError) as e:
        cb.Engine(1.0, 2.0)
    assert "cell_width > padding" in e.value.msg


def test_cpp_wrapper_compiles_and_links(lib, tmp_path):
    """The header-only C++ `Collider<P>` (include/collider_b200.hpp) and its test program build against the library."""
    cb, _ = lib
    libdir = os.path.dirname(cb.LIB_PATH)
    exe = str(tmp_path / "host_collider_check")
    res = subprocess.run(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "host_collider_check.cpp"), "-L" + libdir,
                          "-lcollider_b200", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
