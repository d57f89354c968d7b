"""Import the UNMODIFIED reference modules from /root/reference/src (dev container only).

TEST INFRASTRUCTURE -- not part of the product path.  Only tests/, bench.py's
cpu_baseline leg, the golden-vector generator and __graft_entry__.smoke() may
use anything under oracle/.

/root/reference does not exist on the GPU box, so nothing that runs there may
call `load_reference()`; it is used by `oracle/gen_golden.py` to produce the
committed fixtures in tests/golden/ and by the CPU-side pin tests (skipped when
the reference tree is absent).

Recipe: the reference imports four packages that are not
installed here (IPython, matplotlib, matplotlib.pyplot, path); empty stubs are
registered before the import.  Reference files touched:
  src/AE_model_unet.py:1-16   (imports IPython.display, matplotlib)
  src/utils.py:1-15           (imports path.Path, matplotlib)
  src/calculate_error.py:1-7  (torch, cv2, numpy only)
"""
import os
import sys
import types
import pathlib

REF_SRC = "/root/reference/src"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "AE_model_unet.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_reference():
    """Return (AE_model_unet, calculate_error, utils) reference modules."""
    if not reference_available():
        raise RuntimeError("reference tree not present at " + REF_SRC)
    _stub("IPython", display=types.SimpleNamespace(clear_output=lambda *a, **k: None,
                                                   display=lambda *a, **k: None))
    mpl = _stub("matplotlib", use=lambda *a, **k: None)
    plt = _stub("matplotlib.pyplot")
    mpl.pyplot = plt
    _stub("path", Path=pathlib.Path)
    import importlib.util

    mods = []
    for name in ("AE_model_unet", "calculate_error", "utils"):
        key = "_gdn_reference_" + name
        if key in sys.modules:
            mods.append(sys.modules[key])
            continue
        spec = importlib.util.spec_from_file_location(key, os.path.join(REF_SRC, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[key] = mod
        spec.loader.exec_module(mod)
        mods.append(mod)
    return tuple(mods)
