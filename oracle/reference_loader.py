"""Import the unmodified reference from oracle/_ref/ (bytecode made by oracle/build_ref.py).
TEST INFRASTRUCTURE ONLY - never imported by the product package.

matplotlib is absent from the image and is imported, but never used, on the render path
(apply_hrtf.py:17-18, sphere.py:4-5): empty stub modules are registered first.  The reference's
module names (`apply_hrtf`, `sphere`) would shadow nothing here: they are loaded under private names
and returned, not left importable.
"""
import importlib.machinery
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, m + '.bytecode')) for m in ('apply_hrtf', 'sphere'))


def load():
    """(apply_hrtf, sphere) modules of the reference, or raises ImportError when oracle/_ref is not built."""
    if not available():
        raise ImportError('oracle/_ref is not built: run `python oracle/build_ref.py` where /root/reference exists')
    for name in ('matplotlib', 'matplotlib.pyplot', 'mpl_toolkits', 'mpl_toolkits.mplot3d'):
        sys.modules.setdefault(name, types.ModuleType(name))
    if not hasattr(sys.modules['mpl_toolkits.mplot3d'], 'Axes3D'):
        sys.modules['mpl_toolkits.mplot3d'].Axes3D = object
    saved = {k: sys.modules.get(k) for k in ('sphere', 'apply_hrtf')}
    mods = {}
    try:
        for name in ('sphere', 'apply_hrtf'):          # apply_hrtf does `import sphere`
            loader = importlib.machinery.SourcelessFileLoader(name, os.path.join(REF_DIR, name + '.bytecode'))
            spec = importlib.util.spec_from_loader(name, loader)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            loader.exec_module(mod)
            mods[name] = mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mods['apply_hrtf'], mods['sphere']
