"""Build libsqfa_b200.so (hand-written sm_100a CUDA kernels + the C ABI) with nvcc, in-tree.

Usage: python -m sqfa_b200.build [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU; the resulting .so is git-ignored but travels to the
GPU box with the repo snapshot. An object is rebuilt when the CONTENT of its source or of any header
changed (hashes kept in _build/hashes.json; modification times are not trusted). The hash of all
sources is compiled into the library (`sqfa_build_id()`); `sqfa_b200._lib.load()` compares it with the
sources next to it and refuses a stale library.
"""

import hashlib
import json
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
BUILD_DIR = os.path.join(HERE, "_build")
LIB_PATH = os.path.join(HERE, "libsqfa_b200.so")
# test-only object (tcgen05 operand-layout probe): its own library, not part of the product or its header
PROBE_SRC = os.path.join(ROOT, "tests", "native", "umma_probe.cu")
PROBE_LIB = os.path.join(ROOT, "tests", "native", "libsqfa_probe.so")

NVCC_FLAGS = [
    "-std=c++17",
    "-O3",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas",
    "-v",
]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libsqfa_b200.so")
    return nvcc


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f.endswith(".h")]
    return sorted(hs)


def _sha(paths):
    h = hashlib.sha256()
    for p in paths:
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def source_build_id():
    """Hash of every source and header of the library (16 hex digits), or None without the sources."""
    try:
        return _sha(_sources() + _headers())[:16]
    except OSError:
        return None


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a and link libsqfa_b200.so. Returns the .so path."""
    nvcc = _nvcc()
    os.makedirs(BUILD_DIR, exist_ok=True)
    hash_file = os.path.join(BUILD_DIR, "hashes.json")
    try:
        with open(hash_file) as f:
            old = json.load(f)
    except (OSError, ValueError):
        old = {}
    hdr_hash = _sha(_headers())
    build_id = source_build_id()
    new, jobs, objs = {"flags": " ".join(NVCC_FLAGS)}, [], []
    for src in _sources():
        name = os.path.basename(src)
        obj = os.path.join(BUILD_DIR, name[:-3] + ".o")
        objs.append(obj)
        # capi.cu carries the build id: it is recompiled whenever anything changed
        key = _sha([src]) + hdr_hash + (build_id if name == "capi.cu" else "")
        new[name] = key
        if force or not os.path.exists(obj) or old.get(name) != key or old.get("flags") != new["flags"]:
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc, *NVCC_FLAGS, f'-DSQFA_BUILD_ID="{build_id}"', "-I", INCLUDE, "-c", src, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(BUILD_DIR, os.path.basename(src)[:-3] + ".ptxas.log")
        with open(log, "w") as f:
            f.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            print(res.stderr, file=sys.stderr)
        return obj

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    stale_objs = set(os.path.basename(o) for o in objs)
    for f in os.listdir(BUILD_DIR):  # objects of sources that no longer exist must not be linked
        if f.endswith(".o") and f not in stale_objs:
            os.remove(os.path.join(BUILD_DIR, f))

    if jobs or not os.path.exists(LIB_PATH):
        cmd = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    with open(hash_file, "w") as f:
        json.dump(new, f, indent=1)
    return LIB_PATH


def build_probe(force=False):
    """The test-only tcgen05 layout probe (tests/native/libsqfa_probe.so)."""
    if not os.path.exists(PROBE_SRC):
        return None
    deps = [PROBE_SRC] + _headers()
    if not force and os.path.exists(PROBE_LIB) and os.path.getmtime(PROBE_LIB) >= max(os.path.getmtime(d) for d in deps):
        return PROBE_LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-shared", "-I", INCLUDE, "-I", CSRC, PROBE_SRC, "-o", PROBE_LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {PROBE_SRC}:\n{res.stdout}\n{res.stderr}")
    return PROBE_LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    build_probe(force="--force" in sys.argv)
    print(path)
