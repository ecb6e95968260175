"""Builds libar_slam_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libar_slam_b200.so")
# source -> (headers it includes, extra nvcc flags).  detect.cu is compiled without FMA contraction: its
# polygon-approximation and Otsu comparisons must round like the CPU restatement they are tested against.
SOURCES = {
    "arslam.cu": (["model.cuh", "kernels.cuh", "accum_pipe.cuh", "schur.cuh", "schur_local.cuh", "cholesky.cuh",
                   "localize.cuh", "pcg.cuh"], []),
    "detect.cu": (["dict_4x4_50.inc", "aruco_dictionaries.inc"], ["-fmad=false"]),
}
COMMON_HEADERS = [os.path.join("..", "..", "include", "ar_slam_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def nvcc_path():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _obj(src):
    return os.path.join(LIB_DIR, "obj", src.replace(".cu", ".o"))


def _stale(target, src):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    deps = [src] + SOURCES[src][0] + COMMON_HEADERS
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in deps)


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(not _stale(_obj(s), s) and os.path.getmtime(_obj(s)) <= t for s in SOURCES)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    os.makedirs(os.path.join(LIB_DIR, "obj"), exist_ok=True)
    env = {k: v for k, v in os.environ.items() if k not in ("CXX", "CC")}
    procs = []
    for src, (_, extra) in SOURCES.items():
        if force or _stale(_obj(src), src):
            tmp = _obj(src) + ".tmp%d" % os.getpid()
            cmd = [nvcc_path()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
                  ["-c", "-o", tmp, os.path.join(CSRC, src)]
            procs.append((subprocess.Popen(cmd, cwd=CSRC, env=env), cmd, tmp, _obj(src)))
    for pr, cmd, tmp, obj in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, cmd)
        os.replace(tmp, obj)
    tmp = LIB + ".tmp%d" % os.getpid()   # never leave a half-written library where a snapshot could pick it up
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp] + \
          [_obj(s) for s in SOURCES] + ["-ldl"]
    subprocess.check_call(cmd, cwd=CSRC, env=env)
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
