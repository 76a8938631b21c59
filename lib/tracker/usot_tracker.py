"""Shadow of the reference's ``lib/tracker/usot_tracker.py`` (PEP-420 namespace package, see lib/models/models.py in this
repository and INTEGRATION.md): ``from lib.tracker.usot_tracker import USOTTracker`` (scripts/test_usot.py:13,60) resolves to the
device-side tracker when this repository precedes the reference on ``PYTHONPATH``.

``USOT_B200_HOST_TRACKER=1`` keeps the reference's OWN host-side tracker loop instead (same effect as deleting this directory):
the next ``lib/tracker/usot_tracker.py`` on the namespace path -- the unmodified reference file -- is executed in this module's
place, on top of the shadowed model (its ``.cpu()`` / ``torch.cat`` / ``.cuda()`` round trips of channels-last features included).
"""
import os as _os

if _os.environ.get("USOT_B200_HOST_TRACKER", "0") not in ("", "0"):
    import importlib.util as _ilu
    import lib.tracker as _pkg

    _here = _os.path.dirname(_os.path.abspath(__file__))
    _cands = [_os.path.join(p, "usot_tracker.py") for p in _pkg.__path__ if _os.path.abspath(p) != _here]
    _cands = [c for c in _cands if _os.path.exists(c)]
    if not _cands:
        raise ImportError("USOT_B200_HOST_TRACKER=1 needs the reference tree on PYTHONPATH after this repository")
    _spec = _ilu.spec_from_file_location("lib.tracker._reference_usot_tracker", _cands[0])
    _mod = _ilu.module_from_spec(_spec)
    import sys as _sys
    _sys.modules[_spec.name] = _mod
    _spec.loader.exec_module(_mod)
    REFERENCE_FILE = _cands[0]
    USOTTracker, USOTConfig = _mod.USOTTracker, _mod.USOTConfig
else:
    from usot_b200.tracker import USOTConfig, USOTTracker, python2round  # noqa: F401
