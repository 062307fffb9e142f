"""Import the UNMODIFIED reference (carpedkm/disentangled-subject-to-vid) from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  Used by
make_golden.py to generate the committed fixtures and by nothing else.

Three shims are needed: stub `imageio`, stub `matplotlib(.pyplot)`,
and `transformers.utils.FLAX_WEIGHTS_NAME`.
"""
import os
import sys
import types

REF_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "diffusers", "src", "diffusers"))


def install():
    if not available():
        raise RuntimeError("reference tree not present at /root/reference")
    for name in ("imageio", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                import importlib.machinery

                mod = types.ModuleType(name)
                mod.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
                if name == "matplotlib":
                    mod.__path__ = []
                sys.modules[name] = mod
    if "matplotlib.pyplot" in sys.modules and isinstance(sys.modules["matplotlib"], types.ModuleType):
        setattr(sys.modules["matplotlib"], "pyplot", sys.modules["matplotlib.pyplot"])
    import transformers.utils as tu

    if not hasattr(tu, "FLAX_WEIGHTS_NAME"):
        tu.FLAX_WEIGHTS_NAME = "flax_model.msgpack"
    for p in (os.path.join(REF_ROOT, "diffusers", "src"), os.path.join(REF_ROOT, "src")):
        if p not in sys.path:
            sys.path.insert(0, p)
