"""Shadow of the reference's ``lib/tracker/usot_tracker.py`` (PEP-420 namespace package, see lib/models/models.py in this
repository and INTEGRATION.md): ``from lib.tracker.usot_tracker import USOTTracker`` (scripts/test_usot.py:13,60) resolves to the
device-side tracker when this repository precedes the reference on ``PYTHONPATH``."""
from usot_b200.tracker import USOTConfig, USOTTracker, python2round  # noqa: F401
