"""Build libpandora_b200.so in-tree with nvcc for sm_100a (no torch extension machinery).

`python -m pandora_b200.build` or `build_library()`; used by `__graft_entry__.build()`.
Objects are rebuilt only when their source (or a header) is newer.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(ROOT, "csrc")
LIBDIR = os.path.join(ROOT, "lib")
LIB = os.path.join(LIBDIR, "libpandora_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
    "-DPB_PDL",        # programmatic dependent launch of the tcgen05 GEMM kernels (csrc/common.cuh)
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _newer(src, dst):
    return (not os.path.exists(dst)) or os.path.getmtime(src) > os.path.getmtime(dst)


def build_library(verbose=False, force=False, defines=(), variant=None):
    """`variant` + `defines`: a profiling build (e.g. -DPB_ATT_TRACE) into lib/<variant>/, next to the product
    library, selected at load time with PANDORA_B200_LIB (tools only)."""
    libdir = LIBDIR if variant is None else os.path.join(LIBDIR, variant)
    lib_path = os.path.join(libdir, "libpandora_b200.so")
    flags = [*NVCC_FLAGS, *(f"-D{d}" for d in defines)]
    os.makedirs(libdir, exist_ok=True)
    objdir = os.path.join(libdir, "obj")
    os.makedirs(objdir, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(ROOT), "include", "pandora_b200.h"))
    hdr_m = max(os.path.getmtime(h) for h in hdrs)
    nvcc = _nvcc()
    jobs = []
    for s in srcs:
        src, obj = os.path.join(CSRC, s), os.path.join(objdir, s[:-3] + ".o")
        if force or _newer(src, obj) or hdr_m > os.path.getmtime(obj):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        r = subprocess.run([nvcc, *flags, "-c", src, "-o", obj], capture_output=True, text=True)
        return src, r

    logs = []
    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for src, r in ex.map(compile_one, jobs):
            logs.append((src, r.stderr))
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    objs = [os.path.join(objdir, s[:-3] + ".o") for s in srcs]
    if jobs or not os.path.exists(lib_path):
        r = subprocess.run([nvcc, "-shared", "-o", lib_path, *objs, "-lcuda", "-Xcompiler", "-fPIC"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        for src, log in logs:
            print(f"== {os.path.basename(src)}\n{log}")
    return lib_path


if __name__ == "__main__":
    print(build_library(verbose="-v" in sys.argv, force="-f" in sys.argv))
