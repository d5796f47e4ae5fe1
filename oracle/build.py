"""Compile oracle/march_ref.c -> oracle/_build/libangio_oracle.so (test infrastructure).

oracle/ is test infrastructure: see oracle/__init__.py.  ``-ffp-contract=off`` keeps every
fp32 operation individually rounded so the C restatement and the CUDA kernels (which use
__fmul_rn/__fadd_rn) agree bit for bit.
"""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "march_ref.c")
OUT_DIR = os.path.join(_HERE, "_build")
OUT = os.path.join(OUT_DIR, "libangio_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
           SRC, "-o", OUT, "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
