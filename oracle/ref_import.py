"""Import the UNMODIFIED reference from /root/reference (this container only).

Used by oracle/make_golden.py and by the `needs_reference` tests to pin the oracle.
/root/reference does not exist on the GPU box; nothing on the gpu/bench path imports this.
Missing third-party imports of the reference are replaced by empty stub modules
(SURVEY.md section 8c); none of the stubbed symbols is executed on the hot path.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("AMP_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pointNet"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(_stub(parent), child, mod)
    return mod


def _install_stubs():
    class _Missing:  # instantiating one of these is a bug in the harness
        def __init__(self, *a, **k):
            raise RuntimeError("stubbed third-party symbol was executed")

    _stub("pointNet_2")
    _stub("pointNet_2.models")
    _stub("pointNet_2.models.pointnet2_utils",
          PointNetSetAbstraction=_Missing, PointNetFeaturePropagation=_Missing)
    try:
        import matplotlib  # noqa: F401
    except Exception:
        _stub("matplotlib")
        _stub("matplotlib.pyplot")
        _stub("matplotlib.colors")
        _stub("matplotlib.lines")
    _stub("k_means_constrained", KMeansConstrained=_Missing)
    _stub("progressbar", progressbar=lambda it, **k: it)
    _stub("laspy")


def load():
    """Returns (pointnetAtt module, utils.utils module, collate_fns module)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the reference has a top-level package called `utils`; make sure ours is not shadowing
    for name in ("utils", "utils.utils", "pointNet"):
        m = sys.modules.get(name)
        if m is not None and not str(getattr(m, "__file__", "")).startswith(REFERENCE_ROOT):
            del sys.modules[name]
    model = importlib.import_module("pointNet.model.pointnetAtt")
    uu = importlib.import_module("utils.utils")
    coll = importlib.import_module("pointNet.collate_fns")
    return model, uu, coll
