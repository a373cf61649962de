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
# test-only debug build: same sources + fb_probe.cu with -DFB_DEBUG_SURFACE (phase stamps, tcgen05 layout probe);
# loaded by scripts/ and tests/ only (include/frozenbilm_b200_debug.h)
LIB_DBG = os.path.join(HERE, "libfrozenbilm_sm100_dbg.so")
DEBUG_ONLY = ("fb_probe.cu",)
STAMP = LIB + ".stamp"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-cudart", "static"]


def _sources(debug=False):
    src = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    return src if debug else [f for f in src if os.path.basename(f) not in DEBUG_ONLY]


def _digest():
    h = hashlib.sha256()
    for f in _sources(True) + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + \
            sorted(glob.glob(os.path.join(HERE, "..", "include", "*.h"))):
        h.update(f.encode())
        h.update(open(f, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read() == dig:
        return LIB
    objs, objs_dbg = [], []
    procs = []
    for debug in (False, True):
        for src in _sources(debug):
            obj = os.path.join(CSRC, os.path.basename(src)[:-3] + (".dbg.o" if debug else ".o"))
            (objs_dbg if debug else objs).append(obj)
            extra = ["-DFB_DEBUG_SURFACE"] if debug else []
            procs.append((src, debug, subprocess.Popen([NVCC, *FLAGS, *extra, "-c", src, "-o", obj],
                                                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, debug, pr in procs:
        out, _ = pr.communicate()
        if not debug:
            log.append(f"== {os.path.basename(src)}\n{out}")
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{out}")
    for lib, ob in ((LIB, objs), (LIB_DBG, objs_dbg)):
        link = subprocess.run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
                               "-o", lib, *ob], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
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
