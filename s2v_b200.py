"""Import alias: `import s2v_b200` == the package in ./disentangled-subject-to-vid_b200/ (whose directory name, fixed by
the project layout, is not a valid Python identifier).  Every submodule is aliased to the SAME module object."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_REAL = "disentangled-subject-to-vid_b200"
_pkg = importlib.import_module(_REAL)
for _k, _v in list(sys.modules.items()):
    if _k.startswith(_REAL + "."):
        sys.modules["s2v_b200" + _k[len(_REAL):]] = _v
sys.modules[__name__] = _pkg
