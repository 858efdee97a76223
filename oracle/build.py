"""Build the oracle's C restatement (oracle/svb_oracle.c) with gcc into oracle/_build/libsvb_oracle.so.

TEST INFRASTRUCTURE.  -ffp-contract=off keeps one IEEE rounding per source operation.
The reference itself is pure Python (no C/C++ sources to compile), so there is no oracle/_ref.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'svb_oracle.c')
OUT_DIR = os.path.join(HERE, '_build')
LIB = os.path.join(OUT_DIR, 'libsvb_oracle.so')


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    subprocess.run(['gcc', '-O2', '-std=c99', '-ffp-contract=off', '-fno-fast-math', '-fPIC', '-shared', '-Wall',
                    '-o', LIB, SRC, '-lm'], check=True)
    return LIB


if __name__ == '__main__':
    print(build(force=True))
