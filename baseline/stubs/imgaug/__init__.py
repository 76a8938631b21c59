"""Stand-in for `imgaug` (not installed here and incompatible with numpy 2): the reference tracker only uses
``iaa.Sequential([iaa.Fliplr(1)])`` on one HWC image with one bounding box (lib/tracker/usot_tracker.py:17-20,107-115); this
implements exactly that with imgaug 0.4.0 semantics (pinned against nothing else: the flip is an array reversal)."""
from . import augmenters, augmentables  # noqa: F401
