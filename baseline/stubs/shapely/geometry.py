"""Axis-aligned rectangles only (what poly_iou needs for (left, top, width, height) boxes)."""


class _Area:
    def __init__(self, area):
        self.area = area


class _Rect:
    def __init__(self, x1, y1, x2, y2):
        self.b = (x1, y1, x2, y2) if x2 > x1 and y2 > y1 else None

    @property
    def area(self):
        return 0.0 if self.b is None else (self.b[2] - self.b[0]) * (self.b[3] - self.b[1])

    def intersection(self, o):
        if self.b is None or o.b is None:
            return _Rect(0, 0, 0, 0)
        return _Rect(max(self.b[0], o.b[0]), max(self.b[1], o.b[1]), min(self.b[2], o.b[2]), min(self.b[3], o.b[3]))

    def union(self, o):
        return _Area(self.area + o.area - self.intersection(o).area)


def box(minx, miny, maxx, maxy):
    return _Rect(minx, miny, maxx, maxy)


def Polygon(pts):
    xs, ys = sorted(set(p[0] for p in pts)), sorted(set(p[1] for p in pts))
    if len(xs) > 2 or len(ys) > 2:
        raise NotImplementedError("stand-in shapely: only axis-aligned rectangles are supported")
    return _Rect(xs[0], ys[0], xs[-1], ys[-1])
