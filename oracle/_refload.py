"""Loads the reference's torch-only modules from /root/reference under the alias ``_refmimo``
so they never collide with this repo's own ``mimo`` package.  TEST INFRASTRUCTURE ONLY.

Only works in the build container (the GPU box has no /root/reference); callers must check
``reference_available()`` first.
"""
import importlib
import importlib.util
import os
import sys

REF_ROOT = os.environ.get("MIMO_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "mimo", "models", "mimo_components", "model.py"))


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
