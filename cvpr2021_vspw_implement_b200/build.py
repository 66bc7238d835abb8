"""Build the C-ABI CUDA library in-tree (nvcc, sm_100a only).

    python -m cvpr2021_vspw_implement_b200.build [--force]

The resulting ``csrc/libvspw_b200.so`` is git-ignored but travels to the GPU box with the repo
snapshot.  No torch types cross this boundary; the library links only against cudart.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libvspw_b200.so")
SOURCES = ["layout.cu", "conv_simt.cu", "conv_tc.cu", "bn.cu", "pool.cu", "loss.cu", "ocr.cu", "ocr_tc.cu", "optim.cu", "peer.cu", "ppm.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
] + os.environ.get("VSPW_NVCC_EXTRA", "").split()  # e.g. -DVSPW_BN_UNROLL=8 for tuning experiments


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _deps():
    return sorted([os.path.join(CSRC, s) for s in SOURCES] + [
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))
    ] + [os.path.join(HERE, "..", "include", "vspw_b200.h")])


def _source_hash():
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in _deps():
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


STAMP = os.path.join(CSRC, "libvspw_b200.stamp")


def _stale():
    """Content-based (a snapshot copy to the GPU box does not preserve mtimes reliably): the library is current when the
    hash of its sources and flags equals the stamp written next to it at build time."""
    if not os.path.exists(LIB):
        return True
    if os.path.exists(STAMP):
        return open(STAMP).read().strip() != _source_hash()
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build_library(force=False, verbose=False):
    """Compile every .cu into objects (in parallel) and link libvspw_b200.so."""
    if not force and not _stale():
        return LIB
    import fcntl
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    with open(os.path.join(objdir, ".lock"), "w") as lock:  # ranks of one job must not rebuild the same file concurrently
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not _stale():
            return LIB
        return _build_locked(objdir, verbose)


def _build_locked(objdir, verbose):
    nvcc = _nvcc()
    procs = []
    for s in SOURCES:
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, s), "-o", obj]
        procs.append((s, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    log = []
    for s, obj, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {s}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
        objs.append(obj)
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    tmp = LIB + f".tmp{os.getpid()}"
    cmd = [nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(tmp, LIB)  # atomic: a process that already mapped the old file keeps it
    with open(STAMP, "w") as f:
        f.write(_source_hash())
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
