"""``paddle_edl`` -- the import name used by the reference's README and examples
(``from paddle_edl.distill.distill_reader import DistillReader``,
``python -m paddle_edl.collective.launch``).  It is an alias of :mod:`edl_b200`: every submodule of
``edl_b200`` is importable under the ``paddle_edl.`` prefix."""
from edl_b200._alias import install_alias as _install_alias

_install_alias(__name__)
