"""Import the *live* reference `features` package from /root/reference (CPU container only).

Test infrastructure (see oracle/ref_features.py header).  Recipe from SURVEY.md §8c:
stub matplotlib (plotter.py:8-10 imports it), run from a scratch cwd because
config.py:42-47 creates ./log/<time>.txt at import, put the reference on sys.path.
On the GPU box /root/reference does not exist; bench.py's CPU legs find the copy under baseline/_ref/ instead.
"""
import contextlib
import io
import os
import sys
import tempfile
import types
import warnings

def _find():
    """$DSP_REF_DIR, else /root/reference (CPU container), else baseline/_ref/ (oracle/install_reference.py's copy: the
    only one that exists on the GPU box, used by bench.py's CPU legs alone)."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for d in (os.environ.get("DSP_REF_DIR"), "/root/reference", os.path.join(here, "baseline", "_ref")):
        if d and os.path.isdir(os.path.join(d, "features")):
            return d
    return "/root/reference"


REF_DIR = _find()


def available():
    return os.path.isdir(os.path.join(REF_DIR, "features"))


def load():
    """Returns the reference `features` module (imported under its own name)."""
    if "features" in sys.modules and getattr(sys.modules["features"], "__file__", "").startswith(REF_DIR):
        return sys.modules["features"]
    if "features" in sys.modules:
        raise RuntimeError("a different `features` package is already imported in this process")
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.setdefault("get_cmap", lambda *a, **k: None)
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    scratch = tempfile.mkdtemp(prefix="dsp_ref_")
    cwd = os.getcwd()
    os.chdir(scratch)
    sys.path.insert(0, REF_DIR)
    try:
        with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
            warnings.simplefilter("ignore")
            import features  # noqa: F401  (the reference package)
    finally:
        os.chdir(cwd)
    return sys.modules["features"]


@contextlib.contextmanager
def quiet():
    """Swallow the reference's stdout prints (pitch.py:42,45) and numpy warnings."""
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()), \
            contextlib.redirect_stderr(io.StringIO()):
        warnings.simplefilter("ignore")
        yield
