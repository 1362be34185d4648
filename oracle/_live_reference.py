"""Import the *live* reference `features` package from /root/reference (CPU container only).

Test infrastructure (see oracle/ref_features.py header).  Recipe from SURVEY.md §8c:
stub matplotlib (plotter.py:8-10 imports it), run from a scratch cwd because
config.py:42-47 creates ./log/<time>.txt at import, put the reference on sys.path.
Never used on the GPU box: /root/reference does not exist there.
"""
import contextlib
import io
import os
import sys
import tempfile
import types
import warnings

REF_DIR = os.environ.get("DSP_REF_DIR", "/root/reference")


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
