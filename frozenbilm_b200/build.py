"""In-tree build of libfrozenbilm_sm100.so (nvcc, sm_100a only).

The .so lives next to this file so that it travels to the GPU box with the gpurun snapshot
(git-ignored, not gpurun-ignored).  nvcc cross-compiles without a GPU.
"""
import glob
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfrozenbilm_sm100.so")
STAMP = LIB + ".stamp"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-cudart", "static"]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _digest():
    h = hashlib.sha256()
    for f in _sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + \
            sorted(glob.glob(os.path.join(HERE, "..", "include", "*.h"))):
        h.update(f.encode())
        h.update(open(f, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read() == dig:
        return LIB
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(CSRC, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        procs.append((src, subprocess.Popen([NVCC, *FLAGS, "-c", src, "-o", obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, pr in procs:
        out, _ = pr.communicate()
        log.append(f"== {os.path.basename(src)}\n{out}")
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{out}")
    link = subprocess.run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
                           "-o", LIB, *objs], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if link.returncode != 0:
        raise RuntimeError("link failed:\n" + link.stdout)
    open(STAMP, "w").write(dig)
    open(os.path.join(HERE, "build.log"), "w").write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose=True))
