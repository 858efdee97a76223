"""Build libsvb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m supervillain_b200.build            # rebuild if sources are newer than the .so
    python -m supervillain_b200.build --force

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
INCLUDE = os.path.join(ROOT, 'include')
LIB = os.path.join(HERE, 'libsvb200.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-lineinfo', '-O3', '-std=c++17',
    '-Xcompiler', '-fPIC',
    '-I', INCLUDE, '-I', CSRC,
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps():
    out = sources()
    out += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    out += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f.endswith('.h')]
    return out


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(p) <= t for p in _deps())


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ and link them into libsvb200.so.  Returns the library path."""
    if not force and up_to_date():
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    cmd = [nvcc, '-Wno-deprecated-gpu-targets', '-shared', '-o', LIB] + objs + ['-lcudart_static', '-lpthread', '-ldl', '-lrt']
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
