"""Build libmag2d_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

    python -m mag2d_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The host compiler is pinned to /usr/bin/g++: the image's default
CXX (/opt/gcc/bin/g++) links libstdc++ statically, which breaks shared objects loaded into python.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmag2d_b200.so")
OBJ = os.path.join(HERE, "build")
SOURCES = ["abi.cu", "push.cu", "sort.cu", "poisson.cu", "poisson_direct.cu", "push3d.cu", "push3d_brick.cu", "poisson3d.cu", "comm.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def _ccbin():
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".hpp"))]
    out.append(os.path.join(os.path.dirname(HERE), "include", "mag2d_b200.h"))
    return out


def build(force=False, verbose=False, extra=(), out=None):
    """extra: additional nvcc flags (tuning experiments), out: alternative output path"""
    global LIB, OBJ
    if out:
        LIB = out
        OBJ = out + ".build"
    newest = max(os.path.getmtime(p) for p in _deps())
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest:
        sys.stderr.write("mag2d_b200.build: reused %s (newer than every source; --force recompiles)\n" % LIB)
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    common = [_nvcc(), "-O3", "-std=c++17", "-lineinfo", "-ccbin", _ccbin(), "-Xcompiler", "-fPIC",
              "--cudart", "shared"] + ARCH + list(extra)
    if verbose:
        common += ["-Xptxas", "-v"]

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = common + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    link = [_nvcc(), "-shared", "-ccbin", _ccbin(), "--cudart", "shared", "-o", LIB] + objs + ["-ldl"] + ARCH
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    sys.stderr.write("mag2d_b200.build: compiled %d translation units for sm_100a -> %s\n" % (len(SOURCES), LIB))
    return LIB


if __name__ == "__main__":
    extra = [a for a in sys.argv[1:] if a.startswith("-D")]
    out = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")), None)
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, extra=extra, out=out))
