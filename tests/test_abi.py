"""The C-ABI shared library loads and exports every symbol include/firework_b200.h declares;
POD layouts of the Python binding match the compiled structs; creating a context without a GPU
fails loudly (no CPU fallback). No compute is called here -- runs without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

from bevy_firework_b200 import _abi
from bevy_firework_b200.build import LIB_PATH, build_native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "firework_b200.h")


@pytest.fixture(scope="module")
def lib():
    build_native()
    from bevy_firework_b200._native import load_library

    return load_library()


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fw_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_what_binding_expects():
    assert _declared_functions() == sorted(_abi.EXPORTS.keys())


def test_library_exports_every_declared_symbol(lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only", LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (fw_[a-z0-9_]+)", out))
    for name in _declared_functions():
        assert name in exported, f"{name} declared in the header but not exported"
        getattr(lib, name)


def test_library_is_sm100a_only():
    out = subprocess.check_output(["cuobjdump", "-lelf", LIB_PATH], text=True)
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_pod_layouts_match(lib):
    assert lib.fw_abi_version() == _abi.FW_ABI_VERSION
    for name, cls in _abi.POD_TYPES.items():
        assert lib.fw_abi_sizeof(name.encode()) == C.sizeof(cls), name
    assert lib.fw_abi_sizeof(b"nope") == 0
    assert C.sizeof(_abi.fw_particle_instance) == 64      # reference src/render.rs:95-103
    assert _abi.particle_instance_dtype().itemsize == 64
    assert _abi.particle_data_dtype().itemsize == C.sizeof(_abi.fw_particle_data) == 104
    # vertex attribute offsets 0/16/32/48 (reference src/render.rs:737-766)
    d = _abi.particle_instance_dtype()
    assert [d.fields[f][1] for f in ("position", "scale", "rotation", "base_color", "emissive_color")] == [0, 12, 16, 32, 48]


def test_no_cpu_fallback(lib):
    """Without a CUDA device fw_create must fail with FW_ERR_NO_DEVICE, never compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path is checked on the CPU box")
    cfg = _abi.fw_config(_abi.FW_ABI_VERSION, 0, 1, None, 0, 0)
    ctx = C.c_void_p()
    rc = lib.fw_create(C.byref(cfg), C.byref(ctx))
    assert rc == _abi.FW_ERR_NO_DEVICE
    assert not ctx.value
    assert b"no CPU fallback" in lib.fw_last_global_error()
    from bevy_firework_b200._native import Engine, FireworkError

    with pytest.raises(FireworkError):
        Engine()


def test_create_rejects_bad_arguments(lib):
    ctx = C.c_void_p()
    assert lib.fw_create(None, C.byref(ctx)) == _abi.FW_ERR_INVALID_ARGUMENT
    cfg = _abi.fw_config(_abi.FW_ABI_VERSION + 7, 0, 1, None, 0, 0)
    assert lib.fw_create(C.byref(cfg), C.byref(ctx)) == _abi.FW_ERR_INVALID_ARGUMENT
    assert b"ABI version" in lib.fw_last_global_error()


def test_product_does_not_reference_oracle():
    """Nothing under the package may import, link or call oracle/."""
    pkg = os.path.join(ROOT, "bevy_firework_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "fwo_" not in text and "fw_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f
    out = subprocess.check_output(["nm", "-D", LIB_PATH], text=True)
    assert "fwo_" not in out


def test_header_is_plain_c_and_cpp_host_compiles(tmp_path):
    """include/firework_b200.h compiles as C99; the C++ host mirror compiles against it."""
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", HEADER])
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsyntax-only",
                           os.path.join(ROOT, "host", "examples", "sparks.cpp")])
