"""
TEST / BENCH INFRASTRUCTURE ONLY -- never imported by the product path.

Recipe that makes the UNMODIFIED reference travel to the GPU box ("import, don't
re-implement ... or a vendored copy under the git-ignored baseline/_ref/ if the GPU box lacks the mount").

    python -m oracle.vendor_ref            # /root/reference -> baseline/_ref/

Copies the reference's python packages byte for byte (`shallow_encoders/`, `tools/`, `configs/`, LICENSE) into
`baseline/_ref/`, which is listed in .gitignore (so no reference source ever enters this repository's history) but NOT in
.gpurunignore (so `bench.py --impl reference` and the `cpu_baseline` leg can import the real reference on the GPU box,
where /root/reference does not exist).  `__graft_entry__.build()` runs it whenever /root/reference is present.
`oracle/ref_import.py` resolves the reference root to /root/reference first and to this copy second.
"""
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get('SE_REFERENCE_SRC', '/root/reference')
DST = os.path.join(ROOT, 'baseline', '_ref')
TREES = ('shallow_encoders', 'tools', 'configs')
FILES = ('LICENSE', 'requirements.txt')
SUFFIXES = ('.py', '.yaml', '.sh')


def vendor(src: str = SRC, dst: str = DST) -> int:
    """Returns the number of files copied (0 when the reference is not mounted)."""
    if not os.path.isdir(os.path.join(src, 'shallow_encoders')):
        return 0
    n = 0
    for tree in TREES:
        for dirpath, dirnames, filenames in os.walk(os.path.join(src, tree)):
            dirnames[:] = [d for d in dirnames if d != '__pycache__']
            rel = os.path.relpath(dirpath, src)
            os.makedirs(os.path.join(dst, rel), exist_ok=True)
            for f in filenames:
                if f.endswith(SUFFIXES):
                    shutil.copyfile(os.path.join(dirpath, f), os.path.join(dst, rel, f))
                    n += 1
    for f in FILES:
        if os.path.exists(os.path.join(src, f)):
            shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
            n += 1
    with open(os.path.join(dst, 'VENDORED.txt'), 'w') as fh:
        fh.write(f'byte-for-byte copy of {src} made by oracle/vendor_ref.py; git-ignored; do not edit\n')
    return n


def verify(src: str = SRC, dst: str = DST) -> bool:
    """True when every vendored python file is identical to the mounted reference (or the reference is not mounted)."""
    if not os.path.isdir(os.path.join(src, 'shallow_encoders')):
        return True
    for tree in TREES:
        for dirpath, _dirs, filenames in os.walk(os.path.join(dst, tree)):
            rel = os.path.relpath(dirpath, dst)
            for f in filenames:
                if f.endswith(SUFFIXES) and not filecmp.cmp(os.path.join(dirpath, f), os.path.join(src, rel, f), shallow=False):
                    return False
    return True


if __name__ == '__main__':
    count = vendor()
    print(f'vendored {count} files into {DST}' if count else f'reference not mounted at {SRC}: nothing to do')
    sys.exit(0 if verify() else 1)
