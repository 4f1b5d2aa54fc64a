"""Recipe for oracle/_ref/: the UNMODIFIED reference, byte-compiled from where it lies.

    python oracle/build_ref.py           (also run by __graft_entry__.build() when /root/reference exists)

TEST INFRASTRUCTURE ONLY.  The reference is pure Python, so its "binary" is CPython bytecode:
/root/reference/apply_hrtf.py and sphere.py are compiled with py_compile straight into
oracle/_ref/apply_hrtf.bytecode and oracle/_ref/sphere.bytecode (sourceless modules).  No reference source is
copied into the repository; oracle/_ref/ is git-ignored but travels to the GPU box with the snapshot
(same image, same interpreter), where /root/reference does not exist.  Users: bench.py's cpu_baseline
and `--impl reference` legs (kind "reference") and tests that re-check the oracle against it.
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = '/root/reference'
OUT = os.path.join(HERE, '_ref')
MODULES = ('apply_hrtf', 'sphere')


def build(verbose=True) -> bool:
    """Compile the reference's two modules into oracle/_ref/; False when /root/reference is absent."""
    if not all(os.path.exists(os.path.join(REFERENCE, m + '.py')) for m in MODULES):
        return False
    os.makedirs(OUT, exist_ok=True)
    for m in MODULES:
        py_compile.compile(os.path.join(REFERENCE, m + '.py'), cfile=os.path.join(OUT, m + '.bytecode'), doraise=True)
    with open(os.path.join(OUT, 'BUILT_FROM'), 'w') as f:
        f.write('%s (python %s)\n' % (REFERENCE, sys.version.split()[0]))
    if verbose:
        print('oracle/_ref: compiled %s from %s' % (', '.join(m + '.bytecode' for m in MODULES), REFERENCE))
    return True


if __name__ == '__main__':
    sys.exit(0 if build() else 1)
