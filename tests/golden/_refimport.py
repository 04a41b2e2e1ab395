"""Import the UNMODIFIED reference (tdunnlab/scrubvae) from /root/reference.

Only usable in the build container (the GPU box has no /root/reference). Used by
make_golden.py and by the CPU-only cross-check tests that are skipped when the
reference tree is absent.  Six optional third-party imports of the reference
(neuroposelib, line_profiler, h5py, matplotlib, seaborn, colorcet, wandb) are
not installed here and carry no hot-path arithmetic, so they are stubbed.
"""
import sys
import types
import os

REF_ROOT = "/root/reference"
REF_SRC = os.path.join(REF_ROOT, "src")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "scrubvae"))


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
    if not available():
        raise RuntimeError("reference tree not present at " + REF_ROOT)
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
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import scrubvae  # noqa
        import scrubvae.get  # noqa  (not imported by the package __init__)
        import scrubvae.train.trainer  # noqa
    return sys.modules["scrubvae"]
