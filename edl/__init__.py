"""``edl`` -- the reference's source-package name (python/setup.py.in: ``name='edl'``; the unit
tests import ``edl.*``).  Alias of :mod:`edl_b200`."""
from edl_b200._alias import install_alias as _install_alias

_install_alias(__name__)
