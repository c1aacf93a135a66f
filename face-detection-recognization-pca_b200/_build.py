"""Build libeigenfaces_b200.so in-tree with nvcc for sm_100a (the only supported target).

Every .cu file is compiled to its own object under build/obj (in parallel, only when the content hash of the source,
the headers or the flags changed) and the objects are linked into the shared library.  Staleness is decided by content
hashes kept in a manifest beside the library, not by mtimes (a snapshot copied to another box keeps contents, not
times).  The whole build runs under a file lock and writes to process-unique temporary names, so the ranks of a
torchrun launch that load the library at the same time cannot interleave their outputs."""
import fcntl
import hashlib
import json
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libeigenfaces_b200.so")
OBJ_DIR = os.path.join(HERE, "build", "obj")
MANIFEST = os.path.join(HERE, "build", "manifest.json")
HEADER = os.path.join(HERE, "..", "include", "eigenfaces_b200.h")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-Xcompiler", "-fPIC"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libeigenfaces_b200.so cannot be built (set NVCC=/path/to/nvcc)")


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha1(f.read()).hexdigest()


def _wanted():
    """{object name: hash of everything it is built from}; headers and flags go into every entry."""
    h = hashlib.sha1(" ".join(NVCC_FLAGS + LINK_FLAGS).encode())
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cuh", ".h")):
            h.update(_sha(os.path.join(CSRC, f)).encode())
    h.update(_sha(HEADER).encode())
    common = h.hexdigest()
    return {s[:-3] + ".o": hashlib.sha1((common + _sha(os.path.join(CSRC, s))).encode()).hexdigest() for s in sources()}


def _have():
    try:
        with open(MANIFEST) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def needs_build():
    if not os.path.exists(LIB):
        return True
    have = _have()
    return have.get("objects") != _wanted() or not have.get("linked")


def _compile_one(nvcc, src, obj, verbose):
    tmp = obj + f".{os.getpid()}.tmp"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", tmp]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {os.path.basename(src)}:\n" + res.stdout)
    os.replace(tmp, obj)
    return res.stdout


def build(force=False, verbose=False):
    """Compile every CUDA source into one shared library. Returns the library path."""
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    with open(os.path.join(OBJ_DIR, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():          # another process finished the build while we waited
                return LIB
            want, have = _wanted(), _have().get("objects", {})
            jobs = [(os.path.join(CSRC, o[:-2] + ".cu"), os.path.join(OBJ_DIR, o)) for o in want
                    if force or have.get(o) != want[o] or not os.path.exists(os.path.join(OBJ_DIR, o))]
            workers = max(1, min(len(jobs), int(os.environ.get("EF_BUILD_JOBS", os.cpu_count() or 4))))
            with ThreadPoolExecutor(workers) as pool:
                logs = list(pool.map(lambda j: _compile_one(nvcc, j[0], j[1], verbose), jobs))
            tmp = LIB + f".{os.getpid()}.tmp"
            res = subprocess.run([nvcc] + LINK_FLAGS + [os.path.join(OBJ_DIR, o) for o in sorted(want)] + ["-o", tmp],
                                 stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if res.returncode != 0:
                raise RuntimeError("nvcc link failed:\n" + res.stdout)
            os.replace(tmp, LIB)
            with open(MANIFEST + f".{os.getpid()}.tmp", "w") as f:
                json.dump({"objects": want, "linked": True}, f)
            os.replace(MANIFEST + f".{os.getpid()}.tmp", MANIFEST)
            if verbose:
                print("\n".join(logs))
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB
