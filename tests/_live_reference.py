"""Load the UNMODIFIED reference package from /root/reference under the alias
``ref_torch_rechub`` (test/fixture-generation infrastructure only).

The reference's ``torch_rechub/__init__.py:5`` calls
``importlib.metadata.metadata("torch-rechub")`` so a throw-away dist-info shim is
put on ``sys.path`` first.  Because this repo's own package has the
same import name, the reference is imported under a private module name and its
``torch_rechub.*`` entries are moved out of ``sys.modules`` afterwards.

/root/reference exists only in the build container: every caller must guard with
``live_reference_available()``; nothing under ``-m gpu`` may use it.
"""
import importlib
import os
import sys
import tempfile

REFERENCE_ROOT = "/root/reference"
_ALIAS = "ref_torch_rechub"


def live_reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "torch_rechub"))


def load_reference():
    """Return the reference package object (cached in sys.modules under the alias)."""
    if _ALIAS in sys.modules:
        return sys.modules[_ALIAS]
    if not live_reference_available():
        raise RuntimeError("live reference not present at %s" % REFERENCE_ROOT)
    shim = tempfile.mkdtemp(prefix="rechub_ref_shim_")
    di = os.path.join(shim, "torch_rechub-0.8.0.dist-info")
    os.makedirs(di)
    with open(os.path.join(di, "METADATA"), "w") as f:
        f.write("Metadata-Version: 2.1\nName: torch-rechub\nVersion: 0.8.0\nLicense: MIT\n")
    # stash whatever currently answers to ``torch_rechub`` (this repo's package)
    stashed = {k: v for k, v in sys.modules.items() if k == "torch_rechub" or k.startswith("torch_rechub.")}
    for k in stashed:
        del sys.modules[k]
    sys.path.insert(0, shim)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        importlib.invalidate_caches()
        ref = importlib.import_module("torch_rechub")
        importlib.import_module("torch_rechub.models.ranking")
        importlib.import_module("torch_rechub.trainers")
        importlib.import_module("torch_rechub.basic.layers")
        importlib.import_module("torch_rechub.basic.features")
        importlib.import_module("torch_rechub.utils.data")
        importlib.import_module("torch_rechub.utils.match")
        importlib.import_module("torch_rechub.models.matching")
        importlib.import_module("torch_rechub.basic.loss_func")
        assert ref.__file__.startswith(REFERENCE_ROOT), ref.__file__
    finally:
        sys.path.remove(REFERENCE_ROOT)
        sys.path.remove(shim)
    # re-home the reference modules under the alias and restore ours
    for k in [k for k in sys.modules if k == "torch_rechub" or k.startswith("torch_rechub.")]:
        sys.modules[_ALIAS + k[len("torch_rechub"):]] = sys.modules.pop(k)
    sys.modules.update(stashed)
    importlib.invalidate_caches()
    return sys.modules[_ALIAS]


def ref_module(dotted):
    """``ref_module('basic.layers')`` -> the reference's torch_rechub.basic.layers."""
    load_reference()
    return sys.modules[_ALIAS + "." + dotted]
