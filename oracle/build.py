"""Build recipe for the C restatement: gcc -> oracle/_build/libgobblet_oracle.so.  TEST INFRASTRUCTURE."""
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "gobblet_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libgobblet_oracle.so")


def _fingerprint():
    with open(SRC, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    stamp = OUT + ".srchash"
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read().strip() == _fingerprint():
        return OUT
    tmp = OUT + f".tmp{os.getpid()}"
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-fvisibility=hidden", "-Wall",
                           "-o", tmp, SRC])
    os.replace(tmp, OUT)
    with open(stamp + f".tmp{os.getpid()}", "w") as fh:
        fh.write(_fingerprint())
    os.replace(fh.name, stamp)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
