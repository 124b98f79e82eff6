"""See ``mimo/__init__.py``: hot-path modules live here; anything else resolves from a reference checkout later on sys.path."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
