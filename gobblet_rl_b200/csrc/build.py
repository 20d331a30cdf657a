"""Build libgobblet_b200.so for sm_100a, in-tree (the .so is git-ignored but travels with gpurun)."""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["gobblet_engine.cu", "gobblet_greedy.cu"]
HOST_SOURCES = ["gobblet_host.c"]      # plain C (gcc): thread pool + expander of the packed wire format
HEADERS = ["gobblet_core.cuh", os.path.join("..", "..", "include", "gobblet_b200.h")]
GCC_FLAGS = ["-O3", "-std=gnu11", "-fPIC", "-fvisibility=hidden", "-pthread", "-Wall"]
# tuning builds: GBL_EXTRA_NVCC_FLAGS="-DGBL_BLOCK=128" GBL_LIB_SUFFIX=_b128 python build.py
OUT = os.path.join(HERE, "libgobblet_b200" + os.environ.get("GBL_LIB_SUFFIX", "") + ".so")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--cudart", "static"]


def _fingerprint():
    """Content hash of sources + flags: mtimes do not survive a snapshot copy to another box."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS + GCC_FLAGS + os.environ.get("GBL_EXTRA_NVCC_FLAGS", "").split()).encode())
    for f in SOURCES + HOST_SOURCES + HEADERS:
        with open(os.path.join(HERE, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def stale():
    stamp = OUT + ".srchash"
    if not (os.path.exists(OUT) and os.path.exists(stamp)):
        return True
    with open(stamp) as fh:
        return fh.read().strip() != _fingerprint()


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    tmp = OUT + f".tmp{os.getpid()}"
    extra = os.environ.get("GBL_EXTRA_NVCC_FLAGS", "").split()
    objs = []
    for src in HOST_SOURCES:
        obj = os.path.join(HERE, src + f".tmp{os.getpid()}.o")
        subprocess.check_call([os.environ.get("CC", "gcc"), *GCC_FLAGS, "-c", os.path.join(HERE, src), "-o", obj])
        objs.append(obj)
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-shared", "-o", tmp, *[os.path.join(HERE, s) for s in SOURCES], *objs,
           "-Xlinker", "-lpthread"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    try:
        subprocess.check_call(cmd)
    finally:
        for obj in objs:
            if os.path.exists(obj):
                os.remove(obj)
    os.replace(tmp, OUT)
    with open(OUT + ".srchash.tmp%d" % os.getpid(), "w") as fh:
        fh.write(_fingerprint())
    os.replace(fh.name, OUT + ".srchash")
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
