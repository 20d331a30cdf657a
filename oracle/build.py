"""Build recipe for the C restatement: gcc -> oracle/_build/libgobblet_oracle.so.  TEST INFRASTRUCTURE."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "gobblet_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libgobblet_oracle.so")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    tmp = OUT + f".tmp{os.getpid()}"
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-fvisibility=hidden", "-Wall",
                           "-o", tmp, SRC])
    os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
