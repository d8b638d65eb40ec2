"""ORACLE build recipe (test infrastructure only): compiles oracle/codecs.c with gcc into
oracle/_build/liborc_oracle.so.  The reference itself is Rust and cannot be compiled in this image
(no cargo/rustc), so there is no oracle/_ref; the oracle is a "port"."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liborc_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "codecs.c")
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return LIB
    cmd = ["gcc", "-O2", "-std=gnu11", "-shared", "-fPIC", "-Wall", "-o", LIB, src, "-lz", "-ldl"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
