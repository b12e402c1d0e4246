"""In-tree build of the C-ABI library: plain nvcc, sm_100a only, no torch headers.

    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --shared -Xcompiler -fPIC
         csrc/*.cu -o csrc/libbevpool_b200.so

The .so stays next to the sources (git-ignored, shipped to the GPU box by gpurun).
Precision-critical code uses explicit __fmul_rn/__fadd_rn/__fdiv_rn, and -use_fast_math
is never passed.
"""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libbevpool_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "--shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(HERE, "..", "include", "bevpool_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force=False, verbose=False):
    """Compile if sources are newer than the library. Returns the library path."""
    if not force and not _stale():
        return LIB
    if not os.path.exists(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC}; cannot build {LIB}")
    objs = []
    procs = []
    for src in sources():
        obj = src[:-3] + ".o"
        cmd = [NVCC] + [f for f in FLAGS if f != "--shared"] + ["-c", src, "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        log.append(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        objs.append(obj)
    subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", LIB] + objs)
    with open(os.path.join(CSRC, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print(f"built {LIB}")
    return LIB
