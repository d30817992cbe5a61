"""Build ucd_b200/libucd_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles).

    python -m ucd_b200.build [--force] [--verbose] [--debug]

--debug builds ucd_b200/libucd_b200_debug.so instead: the same sources with -DUCD_DEBUG_KNOBS plus selftest.cu
(per-role cycle tracing, environment tuning knobs, tcgen05 probes: include/ucd_b200_debug.h).  The product library
carries none of that.

The library is a plain C-ABI shared object (include/ucd_b200.h); it links the CUDA runtime statically
and has no torch or libcuda link-time dependency, so it loads on a CPU-only box too (symbols can be
enumerated; any compute call needs a B200).
"""
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libucd_b200.so")
LIB_DEBUG = os.path.join(HERE, "libucd_b200_debug.so")
SOURCES = ["capi.cu", "ce_kd.cu", "upsample.cu", "prep.cu", "contrast.cu", "seg_fused.cu"]
DEBUG_SOURCES = SOURCES + ["selftest.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libucd_b200.so")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, debug=False):
    nvcc = _nvcc()
    obj_dir = OBJ + ("_debug" if debug else "")
    lib, sources = (LIB_DEBUG, DEBUG_SOURCES) if debug else (LIB, SOURCES)
    flags = NVCC_FLAGS + (["-DUCD_DEBUG_KNOBS"] if debug else [])
    os.makedirs(obj_dir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    inc = os.path.join(os.path.dirname(HERE), "include")
    headers += [os.path.join(inc, "ucd_b200.h"), os.path.join(inc, "ucd_b200_debug.h")]
    jobs = []
    for src in sources:
        s = os.path.join(CSRC, src)
        o = os.path.join(obj_dir, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for cmd, res in zip(jobs, ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs)):
                if verbose and res.stderr:
                    print(res.stderr, file=sys.stderr)
                if res.returncode != 0:
                    raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), res.stdout, res.stderr))
    objs = [os.path.join(obj_dir, s.replace(".cu", ".o")) for s in sources]
    if force or jobs or _stale(lib, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed: %s\n%s" % (" ".join(cmd), res.stderr))
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, debug="--debug" in sys.argv))
