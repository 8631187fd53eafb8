"""The OCaml shim (ocaml-hnsw_b200/ocaml/) cannot be built here: the image has no OCaml toolchain.
What CAN be checked on this machine:
  * the C stub file compiles (-fsyntax-only, warnings as errors) against include/hnsw_b200.h and a MOCK of the
    few <caml/*.h> declarations it uses (tests/mock_caml/, written for this test) - so every call into the C ABI
    has the argument count and types the header declares;
  * every `external` in hnsw_b200.ml names stubs that the C file defines, with the arity OCaml will call them
    with (native stub: one `value` per argument; more than 5 arguments need a bytecode twin taking (value*, int)).
"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OCAML = os.path.join(ROOT, "ocaml-hnsw_b200", "ocaml")


def test_stub_file_typechecks_against_the_c_abi():
    r = subprocess.run(["gcc", "-fsyntax-only", "-std=c11", "-Wall", "-Wextra", "-Werror",
                        "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "tests", "mock_caml"),
                        os.path.join(OCAML, "hnsw_b200_stubs.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def _arity(ml_type):
    """Arguments of an OCaml arrow type written on one line (tuples / parenthesised types are one argument)."""
    depth, parts, cur = 0, [], ""
    i = 0
    while i < len(ml_type):
        c = ml_type[i]
        if c in "([":
            depth += 1
        elif c in ")]":
            depth -= 1
        if depth == 0 and ml_type.startswith("->", i):
            parts.append(cur)
            cur = ""
            i += 2
            continue
        cur += c
        i += 1
    parts.append(cur)
    return len(parts) - 1


def test_every_external_has_a_stub_of_the_right_arity():
    ml = open(os.path.join(OCAML, "hnsw_b200.ml")).read() + open(os.path.join(OCAML, "hnsw_b200_graph.ml")).read()
    ml = re.sub(r"^(\s*external [^\n=]*)\n\s*(= )", r"\1 \2", ml, flags=re.M)      # externals written over two lines
    ml = re.sub(r"^\s+external ", "external ", ml, flags=re.M)                        # externals inside a module
    c = open(os.path.join(OCAML, "hnsw_b200_stubs.c")).read()
    stubs = {m.group(1): m.group(2) for m in re.finditer(r"CAMLprim value (\w+)\(([^)]*)\)", c)}
    externals = re.findall(r"^external (\w+) : (.+?) = ((?:\"\w+\"\s*)+)$", ml, re.M)
    assert len(externals) >= 22
    names = {n for n, _, _ in externals}
    for needed in ("import_graph_", "export_layer_", "export_levels_", "stats_", "create_", "search_"):
        assert needed in names, needed
    for name, ml_type, syms in externals:
        syms = re.findall(r"\"(\w+)\"", syms)
        n = _arity(ml_type)
        native = syms[-1]
        assert native in stubs, f"external {name}: no stub {native}"
        assert stubs[native].count("value ") == n, f"external {name}: {native} takes {stubs[native]!r}, OCaml passes {n}"
        if n > 5:
            assert len(syms) == 2, f"external {name}: {n} arguments need a bytecode stub too"
            assert re.fullmatch(r"value\s*\*\s*\w+,\s*int\s+\w+", stubs[syms[0]].strip()), stubs[syms[0]]
        else:
            assert len(syms) == 1
