"""CPU checks of the drop-in boundary: the shared library builds, loads and exports every symbol that
include/polus_b200.h declares (no compute calls -- there is no GPU here), and the product package never
imports the oracle."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "polus_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(polus_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "polus_b200", "libpolus_b200.so"))
    syms = header_symbols()
    assert len(syms) > 60
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_ctypes_binding_covers_the_header():
    from polus_b200 import _lib
    assert not _lib.missing_exports()
    declared = set(header_symbols())
    bound = set(_lib.EXPORTS)
    assert declared == bound, (sorted(declared - bound), sorted(bound - declared))
    assert _lib.call("polus_version") >= 100


def test_error_convention_without_gpu():
    """Entry points return a negative status + message instead of crashing (SURVEY §8b error convention)."""
    import ctypes as C
    from polus_b200 import _lib
    n = C.c_int(-1)
    rc = _lib.load().polus_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        return  # running on a GPU box: nothing to check here
    rc = _lib.load().polus_init(0)
    assert rc < 0
    assert _lib.last_error() != ""


def test_struct_layout_matches_header(tmp_path):
    """ctypes mirrors of polus_gemm_t / polus_operand_t / polus_adam_cfg_t against what the C compiler makes of
    include/polus_b200.h: sizes and the offsets of every field a caller sets."""
    import ctypes as C
    import subprocess
    from polus_b200 import _lib
    fields = ["M", "K", "batch1", "A", "B", "C", "ldc", "c_dtype", "C2", "bias", "alpha", "act", "accumulate", "split_k",
              "c2_kind", "Emul", "colsum"]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "polus_b200.h"\nint main(void){\n'
                   'printf("%zu %zu %zu\\n", sizeof(polus_operand_t), sizeof(polus_gemm_t), sizeof(polus_adam_cfg_t));\n'
                   + "".join(f'printf("%zu\\n", offsetof(polus_gemm_t, {f}));\n' for f in fields) + "return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert [C.sizeof(_lib.Operand), C.sizeof(_lib.Gemm), C.sizeof(_lib.AdamCfg)] == [int(x) for x in out[:3]]
    assert [getattr(_lib.Gemm, f).offset for f in fields] == [int(x) for x in out[3:]]


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "polus_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(dirpath, f)


def test_product_path_never_imports_torch():
    """north_star: "no PyTorch" -- device memory, streams, graphs and NCCL are driven from libpolus_b200.so and the
    host-side rendezvous is a stdlib socket store (polus_b200/comm.py).  Neither the package sources nor a process that
    has imported every module of it may pull torch in."""
    import subprocess
    import sys
    pkg = os.path.join(ROOT, "polus_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "import torch" not in text and "from torch" not in text, os.path.join(dirpath, f)
    code = ("import sys, pkgutil, importlib, polus_b200\n"
            "for m in pkgutil.walk_packages(polus_b200.__path__, 'polus_b200.'):\n"
            "    if 'libpolus' not in m.name: importlib.import_module(m.name)\n"
            "import polus\n"
            "assert 'torch' not in sys.modules, 'torch was imported'\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
