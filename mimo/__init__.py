"""Host-side mirror of the reference's ``mimo`` package for the B200 hot path (same import paths, names, signatures).

Only the modules on the hot path live here. Everything else the reference scripts import -- ``mimo.tasks``,
``mimo.datasets``, ``mimo.visualization``, ``mimo.regularization`` -- is host I/O / logging glue that stays the reference's own
code: ``extend_path`` lets those sub-packages resolve from a reference checkout that sits LATER on ``sys.path``
(``PYTHONPATH=<this repo>:<reference>``), while the modules of this repo shadow their reference namesakes.
"""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
