"""Compile the C restatement of the BPE path (oracle/bpe_oracle.c) with gcc into
oracle/_build/libbpe_oracle.so.  The reference itself is pure Python (nothing to compile into
oracle/_ref); its BPE engine is the third-party `tokenizers` wheel."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "bpe_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libbpe_oracle.so")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-shared", "-fPIC", "-o", LIB, SRC])
    return LIB


if __name__ == "__main__":
    print(build())
