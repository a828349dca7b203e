"""Resolve the `hiccup` namespace the `.hic` wire format depends on.

A `.hic` Huffman-table payload is `pickle.dumps({"type": TupP, "data": [...]})` (reference
hiccup/hicimage.py:117-121): the pickle stores a *global reference* to the class
`hiccup.hicimage.TupP`.  To write byte-identical files, and to read files written by the reference,
that dotted name must resolve in this process.

  * If the reference package is importable (drop-in deployment: hiccup_b200 installed beside
    hiccup), its own modules are used as they are.
  * Otherwise (standalone deployment, e.g. the GPU box) a minimal alias package `hiccup` is
    installed into `sys.modules`, whose `hiccup.hicimage` / `hiccup.model` are this package's own
    container and model modules.  The alias is marked `__hiccup_b200_alias__ = True`.
"""
import importlib
import importlib.util
import sys
import types

REAL = False          #: True when the reference's own `hiccup` package backs the namespace


def _try_real():
    mod = sys.modules.get("hiccup")
    if mod is not None and getattr(mod, "__hiccup_b200_alias__", False):
        return False
    try:
        if importlib.util.find_spec("hiccup") is None:
            return False
        importlib.import_module("hiccup.model")
        importlib.import_module("hiccup.hicimage")
        return True
    except Exception:
        return False


def _install_alias():
    from hiccup_b200 import hicimage as own_hicimage
    from hiccup_b200 import model as own_model
    # pickle stores classes by (__module__, __qualname__): make ours spell the reference's name
    own_hicimage.TupP.__module__ = "hiccup.hicimage"
    pkg = types.ModuleType("hiccup")
    pkg.__path__ = []
    pkg.__hiccup_b200_alias__ = True
    pkg.hicimage = own_hicimage
    pkg.model = own_model
    sys.modules["hiccup"] = pkg
    sys.modules["hiccup.hicimage"] = own_hicimage
    sys.modules["hiccup.model"] = own_model


def resolve():
    """Idempotent; called once from hiccup_b200/__init__.py after the own modules exist."""
    global REAL
    if _try_real():
        REAL = True
        return
    REAL = False
    existing = sys.modules.get("hiccup.hicimage")
    if existing is None or not hasattr(existing, "TupP"):
        _install_alias()


def wire_tuple_class():
    """The class object pickled as the table payload's "type" entry."""
    mod = sys.modules.get("hiccup.hicimage")
    if mod is None or not hasattr(mod, "TupP"):
        resolve()
        mod = sys.modules["hiccup.hicimage"]
    return mod.TupP
