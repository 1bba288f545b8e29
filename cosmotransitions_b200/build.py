"""Builds cosmotransitions_b200/libctbounce.so (in-tree) with nvcc for sm_100a.

Usage: python -m cosmotransitions_b200.build [--verbose]
The .so is git-ignored but travels with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libctbounce.so")
SOURCES = ["ctb_api.cu", "ctb_bounce_poly4.cu", "ctb_bounce_spline.cu", "ctb_bounce_model1_line.cu",
           "ctb_bounce_model1f.cu", "ctb_bounce_thermal1f.cu", "ctb_phases.cu"]
HEADER = os.path.join(PKG, "..", "include", "ctbounce.h")
OBJDIR = os.path.join(PKG, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libctbounce.so")


def needs_build():
    if not os.path.exists(LIB):
        return True
    # every kernel source / header in csrc/ plus the public header: a stale .so must never be loaded silently
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))] + [HEADER]
    return any(os.path.getmtime(f) > t for f in deps)


def _compile_one(args):
    nvcc, src, obj, verbose, extra = args
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    res = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return src, res.returncode, res.stdout


def build(force=False, verbose=False, extra_flags=(), lib=None, sources=None):
    """Compile every translation unit (in parallel) and link libctbounce.so in-tree.
    extra_flags / lib / sources exist for tuning experiments (e.g. -maxrregcount variants)."""
    if lib is None and not force and not needs_build():
        return LIB
    lib = lib or LIB
    sources = sources or SOURCES
    tag = os.path.basename(lib).replace(".so", "")
    from concurrent.futures import ThreadPoolExecutor
    nvcc = find_nvcc()
    os.makedirs(OBJDIR, exist_ok=True)
    jobs = [(nvcc, src, os.path.join(OBJDIR, tag + "_" + src.replace(".cu", ".o")), verbose, list(extra_flags))
            for src in sources]
    with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
        results = list(ex.map(_compile_one, jobs))
    for src, rc, out in results:
        if verbose or rc:
            sys.stdout.write("== %s ==\n%s" % (src, out))
        if rc:
            raise RuntimeError("nvcc failed on %s (%d)" % (src, rc))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + [j[2] for j in jobs]
    res = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode:
        sys.stdout.write(res.stdout)
        raise RuntimeError("link failed (%d)" % res.returncode)
    return lib


if __name__ == "__main__":
    build(force=True, verbose="--verbose" in sys.argv)
    print(LIB)
