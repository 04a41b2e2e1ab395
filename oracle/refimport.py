"""Imports the UNMODIFIED reference (tdunnlab/scrubvae) — test / baseline infrastructure, never the product path.

Search order: `oracle/_ref/` (staged by oracle/make_ref.sh; travels to the GPU box), then `/root/reference/src`
(the build container).  The reference's optional third-party imports (neuroposelib, line_profiler, h5py,
matplotlib, seaborn, colorcet, wandb) are not installed in this image and carry no hot-path arithmetic: they are
replaced by inert stub modules before the import.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")
REF_ROOT = os.environ.get("SCV_REFERENCE_ROOT", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "src")


def source_dir():
    """Directory holding the `scrubvae` package of the reference, or None."""
    for d in (STAGED, REF_SRC):
        if os.path.isfile(os.path.join(d, "scrubvae", "__init__.py")):
            return d
    return None


def available() -> bool:
    return source_dir() is not None


class _Anything(types.ModuleType):
    """Module stub: any attribute resolves to another stub / no-op callable."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Anything(self.__name__ + "." + name)
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        return None


def _stub(name):
    parts = name.split(".")
    for i in range(1, len(parts) + 1):
        full = ".".join(parts[:i])
        if full not in sys.modules:
            m = _Anything(full)
            m.__path__ = []  # behave as a package
            sys.modules[full] = m
            if i > 1:
                setattr(sys.modules[".".join(parts[: i - 1])], parts[i - 1], m)
    return sys.modules[name]


def import_reference():
    """Returns the imported `scrubvae` package of the reference."""
    src = source_dir()
    if src is None:
        raise RuntimeError("reference not staged: run `bash oracle/make_ref.sh` in the build container")
    if "scrubvae" in sys.modules and hasattr(sys.modules["scrubvae"], "get"):
        return sys.modules["scrubvae"]
    import yaml

    for name in [
        "neuroposelib", "neuroposelib.read", "neuroposelib.visualization",
        "neuroposelib.DataStruct", "line_profiler", "h5py", "matplotlib",
        "matplotlib.pyplot", "matplotlib.lines", "seaborn", "colorcet", "wandb",
    ]:
        try:
            __import__(name)
        except Exception:
            _stub(name)
    npl = sys.modules["neuroposelib"]
    if isinstance(npl, _Anything):
        def _cfg(path):
            with open(path) as f:
                return yaml.safe_load(f)
        npl.read.config = _cfg
    lp = sys.modules["line_profiler"]
    if isinstance(lp, _Anything):
        lp.profile = lambda f: f
    if src not in sys.path:
        sys.path.insert(0, src)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import scrubvae  # noqa
        import scrubvae.get  # noqa  (not imported by the package __init__)
        import scrubvae.train.trainer  # noqa
    return sys.modules["scrubvae"]
