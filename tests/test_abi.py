"""CPU: the C-ABI library loads and exports every symbol include/sqfa_b200.h declares, and the
ctypes prototypes in sqfa_b200/_lib.py cover exactly that set (no compute calls: no GPU here)."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sqfa_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sqfa_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge

    ge.build()
    from sqfa_b200 import _lib

    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sqfa_b200.h but not exported"
    assert lib.sqfa_missing_symbols == ()


def test_ctypes_prototypes_match_header(lib):
    from sqfa_b200 import _lib

    assert sorted(_lib.SIGNATURES) == header_functions()


def test_version_and_pure_queries(lib):
    assert lib.sqfa_version() >= 100
    assert lib.sqfa_bucket_workspace_bytes(1000, 10) > 0
    assert lib.sqfa_class_gram_workspace_bytes(50000, 3072, 10) >= 10 * 78 * 16
    assert lib.sqfa_class_factor_floats(5, 0) == 50
    assert lib.sqfa_class_factor_floats(5, 2) == 60
    assert isinstance(lib.sqfa_last_error(), bytes)


def test_argument_validation_without_gpu(lib):
    # invalid arguments are rejected before any CUDA call
    assert lib.sqfa_label_max(None, -1, None, None) == -1
    assert lib.sqfa_bucket_labels(None, 5, 3, None, None, None, None, 0, None) == -1
    assert b"sqfa_bucket_labels" in lib.sqfa_last_error()
    assert lib.sqfa_class_factor(None, 1, 100, 0, None, None, None) == -1


def test_sass_has_blackwell_tensor_core_instructions():
    so = os.path.join(ROOT, "sqfa_b200", "libsqfa_b200.so")
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", so], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCMMA" in sass  # tcgen05.mma
    assert "LDTM" in sass                          # tcgen05.ld
    assert "HMMA.16" not in sass                   # no legacy mma.sync path


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: the header must compile as C99 (and as C++) on its own."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    src = tmp_path / "t.c"
    src.write_text('#include "sqfa_b200.h"\nint main(void) { return 0; }\n')
    for args in (["-std=c99", "-pedantic"], ["-x", "c++", "-std=c++17"]):
        res = subprocess.run([gcc, *args, "-Wall", "-Wextra", "-Werror", "-I", inc, "-fsyntax-only", str(src)],
                             capture_output=True, text=True)
        assert res.returncode == 0, res.stderr


def test_integration_stub_matches_the_abi():
    """The ctypes stub a maintainer is told to add declares the same number of
    arguments as the header for every entry point it binds, and its calls pass that many."""
    import re

    from sqfa_b200 import _lib

    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = text[text.index("# sqfa/_native.py"):]
    block = block[: block.index("```")]
    decls = re.findall(r'\("(sqfa_\w+)",\s*ctypes\.\w+,\s*\[(.*?)\]\)', block, flags=re.S)
    assert len(decls) >= 8
    for name, args in decls:
        n_args = len([a for a in args.replace("\n", " ").split(",") if a.strip()])
        assert n_args == len(_lib.SIGNATURES[name][1]), (name, n_args, len(_lib.SIGNATURES[name][1]))
    for name, _ in decls:
        for hit in re.finditer(r"_lib\." + name + r"\(", block):
            depth, n, i = 1, 1, hit.end()
            while depth:
                ch = block[i]
                depth += ch in "(["
                depth -= ch in ")]"
                n += ch == "," and depth == 1
                i += 1
            assert n == len(_lib.SIGNATURES[name][1]), (name, n, block[hit.start():i])


def header_prototypes():
    """name -> (return C type, [argument C types]) parsed from include/sqfa_b200.h."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    protos = {}
    for ret, name, args in re.findall(r"([A-Za-z_][\w \*]*?)\s*\b(sqfa_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        args = " ".join(args.split())
        types = []
        if args != "void":
            for a in args.split(","):
                a = a.strip()
                # drop the parameter name: everything up to the last '*' or the last space
                t = a[: a.rindex("*") + 1] if "*" in a else a[: a.rindex(" ")]
                types.append(" ".join(t.replace("*", " * ").split()))
        protos[name] = (" ".join(ret.replace("*", " * ").split()), types)
    return protos


def _ctype_of(c_type):
    """The ctypes class _lib.py must use for a C type of the header."""
    from sqfa_b200 import _lib

    if c_type == "const char *":
        return ctypes.c_char_p
    if c_type.endswith("*") or c_type == "sqfa_stream_t":
        return _lib.c_ptr
    return {"int": _lib.c_int, "int32_t": _lib.c_i32, "int64_t": _lib.c_i64, "float": _lib.c_f32,
            "size_t": _lib.c_size}[c_type]


def test_ctypes_argument_types_match_header_prototypes():
    """Every argument of every entry point: the ctypes declaration in sqfa_b200/_lib.py has the width and
    kind (pointer / int32 / int64 / float / size_t) of the C prototype -- a drifted binding would pass
    garbage across the boundary without any error on the Python side."""
    from sqfa_b200 import _lib

    protos = header_prototypes()
    assert sorted(protos) == sorted(_lib.SIGNATURES)
    for name, (ret, args) in protos.items():
        restype, argtypes = _lib.SIGNATURES[name]
        assert restype is _ctype_of(ret), (name, ret, restype)
        assert len(argtypes) == len(args), (name, len(argtypes), len(args))
        for i, (c_type, ct) in enumerate(zip(args, argtypes)):
            assert ct is _ctype_of(c_type), (name, i, c_type, ct)


def test_pure_size_queries_are_consistent(lib):
    """The host-only queries of the ABI (no GPU needed): executed tile area and packed Gram size follow the
    tile shapes DESIGN.md section 4 states, workspaces grow with the problem, the exchange spans of the
    sharded closure lie inside its workspace, 16-byte aligned and disjoint."""
    # D <= 128: one 128 x 128 tile; larger: 256 x 256 upper-triangular tiles
    assert lib.sqfa_gram_executed_tile_area(104) == 128 * 128
    assert lib.sqfa_gram_executed_tile_area(128) == 128 * 128
    for d in (129, 256, 784, 1024, 3072):
        t = -(-d // 256)
        assert lib.sqfa_gram_executed_tile_area(d) == t * (t + 1) // 2 * 256 * 256, d
        assert lib.sqfa_gram_packed_floats(d, 7) == 7 * t * (t + 1) // 2 * 256 * 256, d
    assert lib.sqfa_gram_executed_tile_area(0) == 0 and lib.sqfa_gram_packed_floats(0, 3) == 0
    # the single-call workspace covers its parts
    n, d, c = 50000, 3072, 10
    total = lib.sqfa_class_statistics_workspace_bytes(n, d, c)
    parts = (lib.sqfa_bucket_workspace_bytes(n, c) + lib.sqfa_class_sums_workspace_bytes(n, d, c) + c * d * 4
             + lib.sqfa_class_gram_workspace_bytes(n, d, c) + lib.sqfa_stats_epilogue_workspace_bytes(c))
    assert parts <= total <= parts + 5 * 256
    # closure workspace: larger pair ranges need at least as much; spans inside, aligned, disjoint
    C, D, k, dist = 1000, 512, 16, 1
    P = C * (C - 1) // 2
    full = lib.sqfa_fused_loss_workspace_bytes(C, D, k, dist, 0, P)
    half = lib.sqfa_fused_loss_workspace_bytes(C, D, k, dist, 0, P // 2)
    assert 0 < half <= full
    spans = []
    for which in (0, 1):
        off, nb = ctypes.c_size_t(0), ctypes.c_size_t(0)
        assert lib.sqfa_fused_loss_exchange_span(C, D, k, dist, 0, P // 2, which, ctypes.byref(off), ctypes.byref(nb)) == 0
        assert off.value % 16 == 0 and nb.value % 4 == 0 and nb.value > 0
        assert off.value + nb.value <= half
        spans.append((off.value, off.value + nb.value))
    assert spans[0][1] <= spans[1][0] or spans[1][1] <= spans[0][0]
    assert lib.sqfa_fused_loss_exchange_span(C, D, k, dist, 0, P, 2, None, None) == -1
    # completion signals of the class groups add up to those of one group holding every class
    n_groups = 4
    per_group = [lib.sqfa_class_gram_group_signals(n, d, c, n_groups, g) for g in range(n_groups)]
    assert all(s > 0 for s in per_group)
    assert sum(per_group) == lib.sqfa_class_gram_group_signals(n, d, c, 1, 0)
    assert lib.sqfa_lbfgs_max_n() >= 32 * 3072 and lib.sqfa_lbfgs_max_history() >= 100
