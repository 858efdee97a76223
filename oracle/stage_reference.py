"""Stage the UNMODIFIED reference package under oracle/_ref/ so that it travels to the GPU box.

TEST / BASELINE INFRASTRUCTURE.  The reference (evanberkowitz/supervillain v1.0.0) is pure Python --
there is nothing to compile; "building" it is placing its package directory where the GPU box can
import it.  /root/reference exists only in the build container, oracle/_ref/ is git-ignored (the
sources never enter this repository's history) but NOT gpurun-ignored, so the staged copy ships
with the snapshot like the built .so files.  __graft_entry__.build() calls stage() whenever
/root/reference is present.

    python -m oracle.stage_reference          # copy /root/reference/supervillain -> oracle/_ref/supervillain

What uses it (and nothing else may): bench.py's `cpu_baseline` / `--impl reference` legs (the
reference's own generators timed on the host cores), the `-m gpu` tests that let the reference's
own `Ensemble` drive the GPU generators, and tests/golden/make_*.py.  The product package never
imports it (tests/test_host_logic.py enforces that).
"""
import filecmp
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = os.environ.get('SVB_REFERENCE_ROOT', '/root/reference')
DEST = os.path.join(HERE, '_ref')
PACKAGE = 'supervillain'


def staged():
    return os.path.isfile(os.path.join(DEST, PACKAGE, '__init__.py'))


def _python_files(root):
    out = []
    for base, dirs, files in os.walk(root):
        dirs[:] = sorted(d for d in dirs if d != '__pycache__')
        for f in sorted(files):
            if f.endswith('.py'):
                out.append(os.path.relpath(os.path.join(base, f), root))
    return out


def stage(force=False):
    """Copy the reference's package tree (its .py files, byte for byte) to oracle/_ref/supervillain.  Returns the staged
    root (the directory to put on sys.path), or None when neither the source nor an earlier staging exists."""
    src = os.path.join(SOURCE, PACKAGE)
    if not os.path.isdir(src):
        return DEST if staged() else None
    dst = os.path.join(DEST, PACKAGE)
    files = _python_files(src)
    if not force and staged() and all(
            os.path.isfile(os.path.join(dst, f)) and filecmp.cmp(os.path.join(src, f), os.path.join(dst, f), shallow=False)
            for f in files):
        return DEST
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    for f in files:
        os.makedirs(os.path.dirname(os.path.join(dst, f)), exist_ok=True)
        shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
    with open(os.path.join(DEST, 'STAGED_FROM'), 'w') as fh:
        fh.write(f'{src}\n{len(files)} python files, unmodified (oracle/stage_reference.py)\n')
    return DEST


if __name__ == '__main__':
    print(stage(force=True))
