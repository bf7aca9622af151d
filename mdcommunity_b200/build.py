"""In-tree build of libmdb200.so (hand-written CUDA for sm_100a behind a C ABI).

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box.
"""
import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmdb200.so")
SOURCES = ["md_env.cu", "md_env_large.cu", "md_embed.cu", "md_embed_iter.cu", "md_embed_iter_tc.cu", "md_agg.cu", "md_embed_fused.cu", "md_qhead.cu", "md_qhead_tc.cu", "md_tail.cu", "md_heuristic.cu", "md_spmm.cu", "md_train.cu", "md_api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
]
if os.environ.get("MDB200_PROFILE"):  # in-kernel phase clocks (md_debug_read_env / _fused*); never set for release builds
    NVCC_FLAGS.append("-DMD_PROFILE")


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _stamp():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ["../../include/mdcommunity_b200.h"]:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    # (the -I paths are absolute and differ between checkouts of the same tree, e.g. the GPU box's scratch copy: a stamp
    # that included them made every box rebuild the library -- and N ranks of one job race on the output file)
    h.update(" ".join(f for f in NVCC_FLAGS if not f.startswith("-I")).encode())
    return h.hexdigest()


def up_to_date():
    """True when libmdb200.so exists and its stamp matches the current sources + flags."""
    stamp_file = LIB + ".stamp"
    return os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == _stamp()


def build(force=False, verbose=False):
    """Compile every .cu for sm_100a and link libmdb200.so; no-op when sources are unchanged."""
    stamp_file = LIB + ".stamp"
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    # one builder at a time (several ranks of one job may get here together); the others wait and re-check the stamp
    import fcntl
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
            return LIB
        return _build_locked(stamp, stamp_file, verbose)


def _build_locked(stamp, stamp_file, verbose):
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out))
    tmp = LIB + ".tmp.%d" % os.getpid()
    cmd = [_nvcc(), "-shared", "-o", tmp] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    subprocess.run(cmd, check=True)
    os.replace(tmp, LIB)  # atomic: a concurrent dlopen sees the old or the new file, never a partial one
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
