"""Import shim that lets the *unmodified* reference (IDAS-Durham/JUNE, pure Python)
run in the build container so that golden vectors can be generated from it.

TEST INFRASTRUCTURE ONLY.  Nothing in ``june_b200/`` imports this file.  It is used by
``tests/golden/make_golden.py`` (run by hand in the build container, where
``/root/reference`` exists) to produce the fixtures committed under ``tests/golden/``.
``/root/reference`` does not exist on the GPU box, so nothing that runs there imports
this module.

What it patches, and why (each item was hit when running the reference here):

* ``june/paths.py:106-117`` blocks on ``input()`` when there is no ``data/`` folder
  -> we chdir to a scratch directory that holds an empty ``data/``.
* ``june/__init__.py:7-16`` imports the whole package, which needs ``recordclass``,
  ``h5py``, ``tables``, ``matplotlib``... (not installed) -> a small pure-Python stand-in
  for ``recordclass.dataobject`` and ``MagicMock`` modules for the I/O / plotting deps.
* ``Interaction.print_transmission_statistics`` raises ``KeyError`` in single-process
  mode (``june/interaction/interaction.py:315`` vs ``:163-168``) -> no-op (diagnostics).
* ``Activities`` has 7 fields (``june/demography/person.py:20-27``) but is constructed with 6
  in ``epidemiology.py:195`` -> the stand-in pads missing fields with ``None``.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("JUNE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "june"))


class _DataObjectMeta(type):
    def __new__(mcls, name, bases, ns):
        ann = ns.get("__annotations__", {})
        fields = []
        for b in bases:
            fields.extend(getattr(b, "__fields__", ()))
        own = [k for k in ann if k not in fields]
        fields = tuple(fields) + tuple(own)
        defaults = {k: ns.pop(k) for k in list(own) if k in ns}
        ns["__slots__"] = tuple(own)
        ns["__fields__"] = fields
        ns["__defaults__map"] = defaults
        return super().__new__(mcls, name, bases, ns)


class dataobject(metaclass=_DataObjectMeta):
    """Minimal stand-in for ``recordclass.dataobject``: annotated fields become slots,
    positional/keyword init, missing fields padded with their default or ``None``."""

    def __init__(self, *args, **kwargs):
        fields = self.__fields__
        if len(args) > len(fields):
            raise TypeError("too many positional arguments")
        defaults = getattr(type(self), "__defaults__map", {})
        for i, f in enumerate(fields):
            if i < len(args):
                v = args[i]
            elif f in kwargs:
                v = kwargs[f]
            else:
                v = defaults.get(f)
            setattr(self, f, v)

    def __iter__(self):
        return iter(getattr(self, f) for f in self.__fields__)

    def __repr__(self):
        body = ", ".join(f"{f}={getattr(self, f)!r}" for f in self.__fields__)
        return f"{type(self).__name__}({body})"


_MOCKED = [
    "h5py", "tables", "matplotlib", "matplotlib.pyplot", "matplotlib.colors",
    "matplotlib.cm", "matplotlib.dates", "matplotlib.ticker", "matplotlib.patches",
    "matplotlib.animation", "matplotlib.gridspec", "matplotlib.lines",
    "matplotlib.collections", "matplotlib.colorbar", "matplotlib.figure",
    "matplotlib.backends", "matplotlib.backends.backend_agg", "matplotlib.patheffects",
    "mpl_toolkits", "mpl_toolkits.axes_grid1", "mpl_toolkits.mplot3d",
    "geopy", "geopy.distance", "geopandas", "shapely", "shapely.geometry", "shapely.ops",
    "rasterio", "pyproj", "contextily", "score_clustering", "rtree", "cycler", "PIL",
    "PIL.Image", "seaborn", "plotly", "plotly.graph_objects", "plotly.express",
    "plotly.subplots", "networkx", "folium", "imageio", "cv2", "tqdm.auto",
    "adjustText", "descartes", "mapclassify", "branca", "branca.colormap",
]

_state = {"ready": False}


def setup():
    """Install the shims and import the reference package.  Idempotent.  Returns the
    imported ``june`` module."""
    if _state["ready"]:
        import june  # type: ignore

        return june
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    scratch = tempfile.mkdtemp(prefix="june_ref_cwd_")
    os.makedirs(os.path.join(scratch, "data"), exist_ok=True)
    os.chdir(scratch)
    rc = types.ModuleType("recordclass")
    rc.dataobject = dataobject
    sys.modules["recordclass"] = rc
    for name in _MOCKED:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = MagicMock(name=name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # june.paths inspects sys.argv for --data/--configs
    argv = sys.argv
    sys.argv = [argv[0] if argv else "x"]
    try:
        import logging

        import june  # type: ignore
        from june.interaction import Interaction  # type: ignore

        Interaction.print_transmission_statistics = lambda self: None
        logging.disable(logging.CRITICAL)
    finally:
        sys.argv = argv
    _state["ready"] = True
    return june
