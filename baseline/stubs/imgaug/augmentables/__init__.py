from . import bbs  # noqa: F401
