"""Test infrastructure: the REFERENCE's own Python class megaverse.megaverse_env.MegaverseEnv, imported from /root/reference and running
on the reference's own pybind module (bindings/megaverse.cpp compiled in place on the Bullet stand-in with null renderers,
oracle/_ref/pyref).  A throw-away package directory is assembled from symlinks -- nothing of the reference is copied -- plus a minimal
stand-in for the absent `gym` package (the three space classes and the Env base the reference file touches)."""
import glob
import importlib
import importlib.util
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PY = "/root/reference/megaverse/megaverse_env.py"

_GYM_STUB = '''
import numpy as np
class Env:
    pass
class _Space:
    pass
class Discrete(_Space):
    def __init__(self, n): self.n = n
    def sample(self): return int(np.random.randint(self.n))
class Tuple(_Space):
    def __init__(self, spaces): self.spaces = tuple(spaces)
    def sample(self): return tuple(s.sample() for s in self.spaces)
class Box(_Space):
    def __init__(self, low, high, shape, dtype=np.float32): self.low, self.high, self.shape, self.dtype = low, high, shape, dtype
'''


def available():
    return os.path.exists(REF_PY) and os.path.isdir("/root/reference/src/libs/bindings")


def reference_env_class():
    """returns the reference's MegaverseEnv class (cached in sys.modules under its own package name `megaverse`)"""
    if "megaverse.megaverse_env" in sys.modules:
        return sys.modules["megaverse.megaverse_env"].MegaverseEnv
    ext = glob.glob(os.path.join(ROOT, "oracle", "_ref", "pyref", "megaverse*.so"))
    if not ext:
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "pyref"])
        ext = glob.glob(os.path.join(ROOT, "oracle", "_ref", "pyref", "megaverse*.so"))
    d = tempfile.mkdtemp(prefix="refpy_")
    os.makedirs(os.path.join(d, "megaverse", "extension"))
    os.makedirs(os.path.join(d, "gym"))
    open(os.path.join(d, "megaverse", "__init__.py"), "w").close()
    open(os.path.join(d, "megaverse", "extension", "__init__.py"), "w").close()
    os.symlink(REF_PY, os.path.join(d, "megaverse", "megaverse_env.py"))
    os.symlink(ext[0], os.path.join(d, "megaverse", "extension", os.path.basename(ext[0])))
    with open(os.path.join(d, "gym", "__init__.py"), "w") as f:
        f.write("from . import spaces\nfrom .spaces import Env\n")
    with open(os.path.join(d, "gym", "spaces.py"), "w") as f:
        f.write(_GYM_STUB)
    if importlib.util.find_spec("gym") is not None:  # a real gym wins: drop the stand-in
        import shutil

        shutil.rmtree(os.path.join(d, "gym"))
    sys.path.insert(0, d)
    return importlib.import_module("megaverse.megaverse_env").MegaverseEnv
