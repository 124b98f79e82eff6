"""Loads the reference's torch-only modules from /root/reference under the alias ``_refmimo``
so they never collide with this repo's own ``mimo`` package.  TEST INFRASTRUCTURE ONLY.

Sources, in order: $MIMO_REFERENCE_ROOT, /root/reference (build container), oracle/_ref (staged copy of the five torch-only
files, git-ignored, shipped to the GPU box by gpurun). Callers must check ``reference_available()`` first.
"""
import importlib
import importlib.util
import os
import sys

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _has_reference(root: str) -> bool:
    return os.path.isfile(os.path.join(root, "mimo", "models", "mimo_components", "model.py"))


def _pick_root() -> str:
    """The live checkout in the build container, else the copy oracle/stage_reference.py staged into the git-ignored
    oracle/_ref/ (that copy travels to the GPU box with the gpurun snapshot)."""
    for root in (os.environ.get("MIMO_REFERENCE_ROOT"), "/root/reference", _STAGED):
        if root and _has_reference(root):
            return root
    return "/root/reference"


REF_ROOT = _pick_root()


def reference_available() -> bool:
    return _has_reference(REF_ROOT)


def load():
    """Returns a namespace with the reference's MimoUNet, components, losses, loss_buffer, utils."""
    if "_refmimo" not in sys.modules:
        pkg_dir = os.path.join(REF_ROOT, "mimo")
        spec = importlib.util.spec_from_file_location(
            "_refmimo", os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
        mod = importlib.util.module_from_spec(spec)
        sys.modules["_refmimo"] = mod
        spec.loader.exec_module(mod)

    class NS:
        model = importlib.import_module("_refmimo.models.mimo_components.model")
        components = importlib.import_module("_refmimo.models.mimo_components.components")
        loss_buffer = importlib.import_module("_refmimo.models.mimo_components.loss_buffer")
        losses = importlib.import_module("_refmimo.losses")
        utils = importlib.import_module("_refmimo.models.utils")

    return NS
