"""Builds the C-ABI shared library (hand-written sm_100a kernels + engines) in-tree with nvcc."""
import fcntl
import glob
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# LS_LIB (development aid): load / build an alternative library, e.g. one compiled with LS_BUILD_DEFINES="-DTBLOCK_CLUSTER=1"
LIB = os.environ.get("LS_LIB") or os.path.join(HERE, "libls_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--threads", "0"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _defines():
    return os.environ.get("LS_BUILD_DEFINES", "").split()


def source_hash():
    """Content hash of everything the library is compiled from (sources, headers, flags).  The stamp file written next
    to the .so holds the hash it was built from: staleness does not depend on file times, which a snapshot copy to
    another machine does not preserve."""
    h = hashlib.sha256()
    deps = sources() + sorted(glob.glob(os.path.join(CSRC, "*.h"))) + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + \
        [os.path.join(os.path.dirname(HERE), "include", "ls_b200.h")]
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS + _defines()).encode())
    return h.hexdigest()


def stamp_path():
    return LIB + ".srchash"


def is_stale():
    if not os.path.exists(LIB):
        return True
    try:
        with open(stamp_path()) as f:
            return f.read().strip() != source_hash()
    except OSError:
        return True


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into libls_b200.so (cross-compiles without a GPU).  Serialised across processes
    (torchrun ranks importing at the same time) by a lock file; a no-op when the stamp matches the sources."""
    if not force and not is_stale():
        return LIB
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not is_stale():  # another process built it while we waited
            return LIB
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        tmp = LIB + ".tmp%d" % os.getpid()
        cmd = [nvcc] + NVCC_FLAGS + _defines() + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + sources()
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            if os.path.exists(tmp):
                os.remove(tmp)
            raise RuntimeError("nvcc failed building libls_b200.so")
        os.replace(tmp, LIB)
        with open(stamp_path(), "w") as f:
            f.write(source_hash())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
