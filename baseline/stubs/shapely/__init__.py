"""Stand-in for `shapely` (not installed in this image).  lib/utils/test_utils.py:2 imports Polygon and box at module load; only
the VOT overlap test (poly_iou) uses them.  Axis-aligned rectangles are implemented exactly; rotated polygons are not."""
