"""Shadow of the reference's ``lib/models/models.py``.

``lib``, ``lib.models``, ``lib.tracker`` ... are PEP-420 namespace packages in the reference (no ``__init__.py``), so putting
this repository BEFORE the reference on ``PYTHONPATH`` makes ``import lib.models.models`` resolve here while
``lib.tracker``, ``lib.utils``, ``lib.dataset_loader`` still resolve to the reference.  ``scripts/test_usot.py`` then runs
unchanged (``models.__dict__['USOT']()`` at scripts/test_usot.py:138) on the sm_100a engine.  See INTEGRATION.md.
"""
from usot_b200.models import USOT, USOT_  # noqa: F401
