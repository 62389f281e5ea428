"""Builds the sm_100a shared library (C ABI of include/ripcurrents_b200.h) in-tree with nvcc.

    python -m ripcurrents_b200.build [--force]

Output: ripcurrents_b200/lib/libripcurrents_b200.so (git-ignored; travels to the GPU box with gpurun).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
SO = os.path.join(LIBDIR, "libripcurrents_b200.so")
SOURCES = ["farneback.cu", "aggregate.cu", "advect.cu", "compat.cu", "fields.cu", "diag.cu", "comm.cu", "api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O2,-Wall", "-Xptxas", "-v"]
# farneback.cu holds both arithmetic modes: its strict code is written with __fmul_rn/__fadd_rn intrinsics (never
# contracted) and its fast code wants FMA contraction; the other files restate reference arithmetic in which every
# product and sum rounds separately, so they are compiled with contraction off.
EXTRA = {"farneback.cu": [], "aggregate.cu": ["-fmad=false"], "advect.cu": ["-fmad=false"], "compat.cu": ["-fmad=false"], "fields.cu": ["-fmad=false"], "diag.cu": ["-fmad=false"], "comm.cu": ["-fmad=false"], "api.cu": ["-fmad=false"]}


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _deps(src):
    return [src, os.path.join(CSRC, "rc_internal.h"), os.path.join(CSRC, "atan2f_ref.h"),
            os.path.join(HERE, "..", "include", "ripcurrents_b200.h"), os.path.abspath(__file__)]


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    objs, rebuilt = [], False
    procs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(LIBDIR, s[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or any(os.path.getmtime(d) > os.path.getmtime(obj) for d in _deps(src)):
            cmd = [_nvcc()] + NVCC_FLAGS + EXTRA.get(s, []) + ["-c", src, "-o", obj]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
            rebuilt = True
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        with open(os.path.join(LIBDIR, s[:-3] + ".ptxas.log"), "w") as f:
            f.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % s)
    if rebuilt or not os.path.exists(SO):
        # --cudart shared: the library carries only the runtime symbols it uses (libcudart.so.12 resolves through the
        # rpath below, or is already loaded by torch in the bench / tests)
        cmd = [_nvcc(), "-shared", "--cudart", "shared", "-o", SO] + objs + [
            "-gencode", "arch=compute_100a,code=sm_100a", "-Xlinker", "-rpath=/usr/local/cuda/lib64", "-ldl"]
        subprocess.check_call(cmd)
    return SO


CPP = os.path.join(HERE, "cpp")
CPP_LIB = os.path.join(LIBDIR, "libripcurrents_cpp.so")
DEMO = os.path.join(LIBDIR, "demo_main")
DEMO_MULTI = os.path.join(LIBDIR, "demo_multi_gpu")


def build_cpp(force=False):
    """Header-compatible C++ wrappers (ripcurrents.hpp / Streakline.hpp / pathlines.h) + the drop-in demo loop."""
    so = build(force)
    srcs = [os.path.join(CPP, f) for f in ("ripcurrents_b200.cpp", "demo_main.cpp", "ripcurrents.hpp", "Streakline.hpp",
                                           "pathlines.h", "cv_compat.hpp")]
    newest = max(os.path.getmtime(f) for f in srcs + [so])
    if force or not os.path.exists(CPP_LIB) or os.path.getmtime(CPP_LIB) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++14", "-fPIC", "-Wall", "-ffp-contract=off", "-shared", "-o", CPP_LIB,
                               os.path.join(CPP, "ripcurrents_b200.cpp"), "-L" + LIBDIR, "-lripcurrents_b200",
                               "-Wl,-rpath,$ORIGIN"])
    if force or not os.path.exists(DEMO) or os.path.getmtime(DEMO) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++14", "-Wall", "-o", DEMO, os.path.join(CPP, "demo_main.cpp"),
                               "-L" + LIBDIR, "-lripcurrents_cpp", "-lripcurrents_b200", "-Wl,-rpath,$ORIGIN"])
    multi_src = os.path.join(CPP, "demo_multi_gpu.cpp")          # plain C ABI + std::thread: one stream sharded over the GPUs
    if force or not os.path.exists(DEMO_MULTI) or os.path.getmtime(DEMO_MULTI) < max(os.path.getmtime(multi_src), os.path.getmtime(so)):
        subprocess.check_call(["g++", "-O2", "-std=c++14", "-Wall", "-pthread", "-o", DEMO_MULTI, multi_src, "-L" + LIBDIR,
                               "-lripcurrents_b200", "-Wl,-rpath,$ORIGIN"])
    return CPP_LIB, DEMO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_cpp(force="--force" in sys.argv))
