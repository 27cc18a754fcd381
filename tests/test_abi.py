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


# ---------------------------------------------------------------- header <-> Rust <-> compiled library
RUST_SYS = os.path.join(ROOT, "rust", "firework_b200_sys.rs")
_C_SCALARS = {"float": ("f32", 4, 4), "uint32_t": ("u32", 4, 4), "int32_t": ("i32", 4, 4), "uint64_t": ("u64", 8, 8),
              "uint8_t": ("u8", 1, 1), "void*": ("*mut c_void", 8, 8)}


def _parse_c_structs():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    consts = {m.group(1): int(m.group(2).rstrip("u"), 0) for m in re.finditer(r"#define\s+(FW_\w+)\s+(0x[0-9A-Fa-f]+u?|\d+u?)\s*$", text, flags=re.M)}
    structs = {}
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            fm = re.match(r"(.+?)\s*(\*?)\s*(\w+)((?:\[[^\]]+\])*)$", decl)
            ctype = fm.group(1).strip() + fm.group(2)
            dims = [consts[d] if d in consts else int(d) for d in re.findall(r"\[([^\]]+)\]", fm.group(4))]
            fields.append((fm.group(3), ctype, dims))
        structs[m.group(3)] = fields
    return structs, consts


def _parse_rust():
    text = re.sub(r"//.*", "", open(RUST_SYS).read())
    consts = {m.group(1): int(m.group(2).replace("_", ""), 0) for m in re.finditer(r"pub const (FW_\w+): \w+ = (0x[0-9A-Fa-f_]+|\d+);", text)}
    structs = {}
    for m in re.finditer(r"#\[repr\(C\)\][^{]*?pub struct (\w+)\s*\{(.*?)\n\}", text, flags=re.S):
        fields = []
        for fm in re.finditer(r"pub (\w+):\s*([^,\n]+),", m.group(2)):
            fields.append((fm.group(1), fm.group(2).strip()))
        structs[m.group(1)] = fields
    ext = text[text.index('extern "C" {'):]
    fns = {}
    for m in re.finditer(r"pub fn (fw_\w+)\s*\((.*?)\)\s*(?:->\s*([^;]+))?;", ext, flags=re.S):
        args = [a.split(":", 1)[1].strip() for a in m.group(2).replace("\n", " ").split(",") if ":" in a]
        fns[m.group(1)] = (args, (m.group(3) or "()").strip())
    return structs, consts, fns


def _rust_type_of(ctype, dims, consts_by_value_ok=True):
    base = _C_SCALARS[ctype][0] if ctype in _C_SCALARS else ctype
    for d in reversed(dims):
        base = f"[{base}; {d}]"
    return base


def _layout(structs, name, memo):
    """(size, align, {field: offset}) by the C layout rules"""
    if name in memo:
        return memo[name]
    off, align, offsets = 0, 1, {}
    for fname, ctype, dims in structs[name]:
        if ctype in _C_SCALARS:
            _, sz, al = _C_SCALARS[ctype]
        else:
            sz, al, _ = _layout(structs, ctype, memo)
        n = 1
        for d in dims:
            n *= d
        off = (off + al - 1) // al * al
        offsets[fname] = off
        off += sz * n
        align = max(align, al)
    size = (off + align - 1) // align * align
    memo[name] = (size, align, offsets)
    return memo[name]


def test_rust_binding_matches_header_and_library(lib):
    """rust/firework_b200_sys.rs cannot be compiled here (no Rust toolchain), so it is checked
    mechanically: same constants, same structs (field names, order, element types, array lengths),
    offsets by the C layout rules == fw_abi_offsetof of the compiled library == the ctypes binding,
    and every export of the header declared with the same arity."""
    c_structs, c_consts = _parse_c_structs()
    r_structs, r_consts, r_fns = _parse_rust()
    for k in ("FW_ABI_VERSION", "FW_MAX_KNOTS", "FW_MAX_EXCLUDED", "FW_NO_KEY", "FW_GATHER_MAX_RANKS", "FW_FLAG_PROFILE", "FW_FLAG_NO_GRAPHS",
              "FW_FLAG_NO_CONCURRENT_SPAWN", "FW_LAYOUT_COMPACTING", "FW_LAYOUT_COLLIDES", "FW_LAYOUT_ROTATES", "FW_STORE_BASE_COLOR",
              "FW_STORE_EMISSIVE_COLOR", "FW_STORE_SCALE", "FW_STORE_LIFETIME"):
        assert r_consts[k] == c_consts[k], k
    resolve = lambda t: re.sub(r"FW_\w+", lambda m: str(r_consts[m.group(0)]), t)
    memo = {}
    for name, fields in c_structs.items():
        assert name in r_structs, f"struct {name} missing in the Rust binding"
        r_fields = r_structs[name]
        assert [f for f, _ in r_fields] == [f for f, _, _ in fields], name
        for (fname, ctype, dims), (_, rtype) in zip(fields, r_fields):
            assert resolve(rtype) == _rust_type_of(ctype, dims), f"{name}.{fname}: {rtype} vs {ctype}{dims}"
        size, _, offsets = _layout(c_structs, name, memo)
        assert lib.fw_abi_sizeof(name.encode()) == size, name
        for fname, off in offsets.items():
            assert lib.fw_abi_offsetof(name.encode(), fname.encode()) == off, f"{name}.{fname}"
        if name in _abi.POD_TYPES:
            cls = _abi.POD_TYPES[name]
            assert C.sizeof(cls) == size and {f: getattr(cls, f).offset for f, _ in cls._fields_} == offsets, name
    assert lib.fw_abi_offsetof(b"fw_collider", b"nope") == 0xFFFFFFFF
    # exports: every function of the header, same number of arguments
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for m in re.finditer(r"\b(fw_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        name, args = m.group(1), m.group(2).strip()
        n_args = 0 if args in ("", "void") else len(args.split(","))
        assert name in r_fns, f"{name} missing in the Rust binding"
        assert len(r_fns[name][0]) == n_args, f"{name}: {len(r_fns[name][0])} arguments in Rust, {n_args} in the header"
        assert len(_abi.EXPORTS[name][1]) == n_args, name
    assert set(r_fns) == set(_abi.EXPORTS), set(r_fns) ^ set(_abi.EXPORTS)


def test_c99_example_compiles_and_links(tmp_path):
    """host/examples/sparks_c99.c drives sparks through the bare C ABI (no C++, no Python)"""
    exe = tmp_path / "sparks_c99"
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-O1", "-o", str(exe),
                           os.path.join(ROOT, "host", "examples", "sparks_c99.c"), "-L" + os.path.dirname(LIB_PATH),
                           "-lfirework_b200", "-lm", "-Wl,-rpath," + os.path.dirname(LIB_PATH)])
    import torch

    if not torch.cuda.is_available():  # without a device the program must fail loudly, not compute
        r = subprocess.run([str(exe)], capture_output=True, text=True)
        assert r.returncode != 0 and "no CPU fallback" in (r.stdout + r.stderr)
