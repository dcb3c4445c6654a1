"""In-tree build of libmaua_b200.so with nvcc for sm_100a (no torch dependency in the library).

`python -m maua_stylegan2_b200.build` or `__graft_entry__.build()`.  Objects are rebuilt only when their source
(or a header) is newer, so repeated calls are cheap.  The .so is git-ignored but travels to the GPU box with gpurun.
"""
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libmaua_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unused-function",
    "-I", os.path.join(ROOT, "include"),
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newest_header():
    t = 0.0
    for d in (CSRC, os.path.join(ROOT, "include")):
        for f in os.listdir(d):
            if f.endswith((".h", ".cuh")):
                t = max(t, os.path.getmtime(os.path.join(d, f)))
    return t


def build(verbose=False, force=False, extra_flags=(), tag=""):
    """tag / extra_flags: an experimental variant (e.g. tag="epi4", extra_flags=["-DMAUA_EPI_GROUPS=4"]) is built into
    build_<tag>/ and libmaua_b200_<tag>.so; MAUA_LIB_PATH selects it at run time (tools/ A/B measurements)."""
    global OBJ, LIB
    obj_dir, lib_path = (OBJ, LIB) if not tag else (os.path.join(PKG, f"build_{tag}"), os.path.join(PKG, f"libmaua_b200_{tag}.so"))
    return _build(verbose, force, list(extra_flags), obj_dir, lib_path)


def _build(verbose, force, extra_flags, OBJ, LIB):
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdr_t = _newest_header()
    jobs = []
    objs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc] + NVCC_FLAGS + extra_flags + ["-Xptxas", "-v" if verbose else "-warn-spills", "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r.returncode, r.stdout + r.stderr

    failed = False
    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for src, rc, out in ex.map(compile_one, jobs):
            if rc != 0 or verbose:
                print(f"[nvcc] {os.path.basename(src)} rc={rc}\n{out}", file=sys.stderr)
            failed |= rc != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if jobs or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcufft"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stdout + r.stderr, file=sys.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    extra = [a for a in sys.argv[1:] if a.startswith("-D")]
    tag = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--tag=")), "")
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv, extra_flags=extra, tag=tag))
