"""Build libeigenfaces_b200.so in-tree with nvcc for sm_100a (the only supported target)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libeigenfaces_b200.so")
SOURCES = ["ef_abi.cu", "ef_preprocess.cu", "ef_project.cu", "ef_project_tc.cu", "ef_recognize_cluster.cu", "ef_recognize_pipe.cu", "ef_epilogue.cu", "ef_match.cu", "ef_match_small.cu", "ef_match_tc.cu", "ef_linalg.cu",
           "ef_model.cu", "ef_fit.cu", "ef_gram_tc.cu", "ef_template.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libeigenfaces_b200.so cannot be built (set NVCC=/path/to/nvcc)")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "eigenfaces_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile every CUDA source into one shared library. Returns the library path."""
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB + ".tmp"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    os.replace(LIB + ".tmp", LIB)
    if verbose:
        print(res.stdout)
    return LIB
