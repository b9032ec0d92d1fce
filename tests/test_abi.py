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
