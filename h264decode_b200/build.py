"""Build libh264b200.so in-tree with nvcc for sm_100a (no torch extension machinery: the product is a plain C-ABI
shared library).  `python -m h264decode_b200.build` or build() from __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libh264b200.so")
SOURCES = ["api.cu", "annexb_scan.cu", "cabac_engine.cu", "ctx_init.cu", "slice_header.cu", "param_sets.cu", "ctx_glue.cu", "batch.cu", "mb_type.cu"]
HEADERS = ["common.cuh", "annexb_local.cuh", "cabac_lane.cuh", "slice_header.cuh", "param_sets.cuh", "ctx_glue.cuh", "tables.inc", "../../include/h264b200.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.exists(c) or c == "nvcc"):
            return c
    return "nvcc"


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False, extra=(), out=None):
    """out: alternative output path (experiment builds with extra -D flags; always rebuilt)"""
    if out is None and not force and up_to_date():
        return LIB
    objs = []
    procs = []
    objdir = os.path.join(HERE, "build" if out is None else "build_" + os.path.basename(out))
    os.makedirs(objdir, exist_ok=True)
    for s in SOURCES:
        o = os.path.join(objdir, s.replace(".cu", ".o"))
        cmd = [nvcc()] + NVCC_FLAGS + list(extra) + ["-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            print(" ".join(cmd))
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        log, _ = p.communicate()
        if log.strip() and (verbose or p.returncode):
            print(log)
        if p.returncode:
            raise RuntimeError("nvcc failed on %s" % s)
    target = LIB if out is None else out
    # -cudart shared: the statically linked runtime would carry the symbol names of driver entry points this library
    # never calls (every copy here is a plain cudaMemcpyAsync)
    cmd = [nvcc(), "-shared", "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64", "-o", target] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lpthread"]
    subprocess.check_call(cmd)
    return target


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True,
          extra=(["-Xptxas", "-v"] if "--ptxas" in sys.argv else []))
    print(LIB)
