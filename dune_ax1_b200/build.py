"""Build libax1gpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libax1gpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"] + ARCH

# (source, extra flags); boundary_fd.cu must not fuse multiply-adds (see its header comment)
UNITS = [
    ("assembly.cu", []),
    ("boundary_fd.cu", ["--fmad=false"]),
    ("solver.cu", []),
    ("ilu1.cu", []),
    ("amg.cu", []),
    ("driver.cu", []),
    ("timeloop.cu", []),
    ("mori.cu", []),
    ("capi.cu", []),
    ("host_model.cpp", []),
]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "ax1gpu.h"))
    objs = []
    procs = []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [NVCC] + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
        if verbose:
            sys.stderr.write(out)
    if force or procs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ARCH + ["-lcudart", "-ldl"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
