"""TEST INFRASTRUCTURE ONLY -- import shim for the real reference under /root/reference.

Only usable in the dev container (the GPU box has no /root/reference).  Used by
oracle/make_golden.py to generate the committed fixtures in tests/golden/ and by the
`-m "not gpu"` tests to re-check the restated oracle (oracle/torch_ref.py) against the
live reference when it happens to be present.  The product package never imports this.

The reference model files import matplotlib / imageio / tensorflow / scipy.misc.imsave
(utils.py:7-10, logger.py:2) which are absent from this image; they are replaced by empty
stub modules (with a ModuleSpec, torch probes `find_spec('tensorflow')`) *before* import.
Reference files are never modified.
"""
import importlib
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SRB_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "base_networks.py"))


def _stub(name, **attrs):
    if name in sys.modules and not getattr(sys.modules[name], "__srb_stub__", False):
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__srb_stub__ = True
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = {}


def load():
    """Return dict name -> reference module (base_networks, srcnn, espcn, fsrcnn, vdsr, edsr, srgan, utils)."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    _stub("imageio")
    _stub("tensorflow")
    import scipy  # noqa
    try:
        import scipy.misc as _sm  # scipy.misc was removed upstream; stub what utils.py:10 asks for
        if not hasattr(_sm, "imsave"):
            _sm.imsave = lambda *a, **k: None
    except Exception:
        scipy.misc = _stub("scipy.misc", imsave=lambda *a, **k: None)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import torchvision.transforms as _T
        if not hasattr(_T, "Scale"):
            # utils.py:255,266 call transforms.Scale, torchvision's pre-0.2 name of Resize (renamed, then removed upstream);
            # aliasing it is the only way to run the reference's img_interp under the installed torchvision.  Modern Resize
            # reads a 2-tuple as (h, w) where the reference passes (w, h): identical for the square crops it trains on.
            _T.Scale = _T.Resize
        for name in ("base_networks", "utils", "srcnn", "espcn", "fsrcnn", "vdsr", "edsr", "srgan"):
            _loaded[name] = importlib.import_module(name)
    return _loaded
