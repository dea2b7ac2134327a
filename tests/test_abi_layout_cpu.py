"""include/nrf_b200.h against its ctypes mirror (_lib.py), without a GPU: the header is compiled as plain C by gcc and
every struct's size and field offsets are compared with the ctypes Structures the Python host hands to the library, and
every function the header declares is compared with the binding's signature table (name and argument count).  A field
added on one side only would otherwise show up as a silently shifted pointer."""
import ctypes as C
import os
import re
import shutil
import subprocess

import pytest

from tests.conftest import load_pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nrf_b200.h")
STRUCTS = ["NrfCompositeReuse", "NrfGemm", "NrfMlpParams", "NrfMlpGrads", "NrfMlpSizes"]


def test_header_structs_match_the_ctypes_mirror(tmp_path):
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    lib = load_pkg("_lib")
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void) {"]
    for name in STRUCTS:
        st = getattr(lib, name)
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for field, _ in st._fields_:
            lines.append(f'  printf("{name}.{field} %zu\\n", offsetof({name}, {field}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run([gcc, "-std=c11", "-Wall", "-Werror", "-o", str(exe), str(src)], check=True)   # the header is plain C
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for name in STRUCTS:
        st = getattr(lib, name)
        assert int(got[name]) == C.sizeof(st), (name, got[name], C.sizeof(st))
        for field, _ in st._fields_:
            assert int(got[f"{name}.{field}"]) == getattr(st, field).offset, (name, field)
    assert lib.NRF_MAX_BLOCKS == int(re.search(r"#define NRF_MAX_BLOCKS (\d+)", open(HEADER).read()).group(1))


def test_header_functions_match_the_binding_table():
    lib = load_pkg("_lib")
    text = re.sub(r"/\*.*?\*/", " ", open(HEADER).read(), flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|int64_t|const char\*|void)\s+(nrf_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        decls[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    for name, sig in lib._SIGNATURES.items():
        assert name in decls, f"{name} is bound by _lib.py but not declared in the header"
        assert decls[name] == len(sig), (name, decls[name], len(sig))
    for name in lib.EXPORTS:
        assert name in decls, f"{name} is exported but not declared in the header"
    for name in decls:
        assert name in lib.EXPORTS or name.startswith("nrf_debug_"), f"{name} is declared but not in _lib.EXPORTS"
