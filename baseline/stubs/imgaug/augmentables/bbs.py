class BoundingBox:
    def __init__(self, x1, y1, x2, y2):
        self.x1, self.y1, self.x2, self.y2 = x1, y1, x2, y2


class BoundingBoxesOnImage:
    def __init__(self, bounding_boxes, shape):
        self.bounding_boxes, self.shape = bounding_boxes, shape

    def __getitem__(self, i):
        return self.bounding_boxes[i]
