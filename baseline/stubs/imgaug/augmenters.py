from .augmentables.bbs import BoundingBox, BoundingBoxesOnImage


class Fliplr:
    def __init__(self, p):
        assert p == 1, "stand-in imgaug supports Fliplr(1) only"


class Sequential:
    def __init__(self, children):
        assert len(children) == 1 and isinstance(children[0], Fliplr), "stand-in imgaug supports Sequential([Fliplr(1)]) only"

    def __call__(self, image, bounding_boxes):
        w = bounding_boxes.shape[1]
        out = [BoundingBox(w - b.x2, b.y1, w - b.x1, b.y2) for b in bounding_boxes.bounding_boxes]
        return image[:, ::-1], BoundingBoxesOnImage(out, bounding_boxes.shape)
