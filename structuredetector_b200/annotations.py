"""Output types of the decoding path.

API-compatible with the reference's annotation classes (reference:
src/sdnet/utils/utils.py:12-60 ``Keypoint``, 63-148 ``Box``, 151-237 ``Object``,
240-308 ``ImageAnnotation``): same constructor arguments, attributes, method names
and JSON layout, so ``evaluate``/``detect``-style callers and the reference
``Evaluator`` can consume what our decoder returns.  Implementation is our own: a
small ``_Scalable`` mixin supplies the copy-returning ``resized``/``normalized``
variants, and all geometry goes through two helpers.
"""
from __future__ import annotations

import copy
import json
from pathlib import Path

import numpy as np

__all__ = ["Keypoint", "Box", "Object", "ImageAnnotation"]


def _ratio(src, dst):
    """(sx, sy) that maps coordinates in a ``src=(w, h)`` frame to a ``dst=(w, h)`` frame."""
    return dst[0] / src[0], dst[1] / src[1]


class _Scalable:
    """``resize``/``normalize`` mutate and return self; the ``-d`` forms work on a deep copy."""

    __slots__ = ()

    def resized(self, in_size, out_size):
        return copy.deepcopy(self).resize(in_size, out_size)

    def normalized(self, size=None):
        clone = copy.deepcopy(self)
        return clone.normalize(size) if size is not None else clone.normalize()


class Keypoint(_Scalable):
    # slots: the decoder builds ~300 of these per image; the C assembly loop (csrc/fastobj.c) stores straight
    # into the slot offsets.  Attribute names and constructor are the reference's (utils.py:12-17).
    __slots__ = ("kind", "x", "y", "score")

    def __init__(self, kind, x, y, score=None):
        self.kind, self.x, self.y, self.score = kind, x, y, score

    def resize(self, in_size, out_size):
        sx, sy = _ratio(in_size, out_size)
        self.x *= sx
        self.y *= sy
        return self

    def normalize(self, size):
        self.x /= size[0]
        self.y /= size[1]
        return self

    def distance(self, other):
        return np.hypot(self.x - other.x, self.y - other.y)

    def json_repr(self):
        return {"kind": self.kind, "location": {"x": self.x, "y": self.y}, "score": self.score}

    @staticmethod
    def from_json(json_dict):
        loc = json_dict["location"]
        return Keypoint(json_dict["kind"], loc["x"], loc["y"], json_dict.get("score"))

    def __repr__(self):
        return f"Keypoint(kind: {self.kind}, x: {self.x}, y: {self.y}, score: {self.score})"


class Box(_Scalable):
    _FIELDS = ("x_min", "y_min", "x_max", "y_max")

    def __init__(self, x_min, y_min, x_max, y_max):
        self.x_min, self.y_min, self.x_max, self.y_max = x_min, y_min, x_max, y_max

    x_mid = property(lambda self: (self.x_min + self.x_max) / 2)
    y_mid = property(lambda self: (self.y_min + self.y_max) / 2)
    width = property(lambda self: abs(self.x_max - self.x_min))
    height = property(lambda self: abs(self.y_max - self.y_min))

    def _scale(self, sx, sy):
        self.x_min *= sx
        self.x_max *= sx
        self.y_min *= sy
        self.y_max *= sy
        return self

    def resize(self, in_size, out_size):
        return self._scale(*_ratio(in_size, out_size))

    def normalize(self, size):
        self.x_min /= size[0]
        self.y_min /= size[1]
        self.x_max /= size[0]
        self.y_max /= size[1]
        return self

    def yolo_coords(self, size):
        w, h = size
        return self.x_mid / w, self.y_mid / h, self.width / w, self.height / h

    def standardize(self):
        self.x_min, self.x_max = min(self.x_min, self.x_max), max(self.x_min, self.x_max)
        self.y_min, self.y_max = min(self.y_min, self.y_max), max(self.y_min, self.y_max)
        return self

    def standardized(self):
        return copy.deepcopy(self).standardize()

    def json_repr(self):
        return {name: getattr(self, name) for name in self._FIELDS}

    @staticmethod
    def from_json(json_dict):
        if json_dict is None:
            return None
        return Box(*(json_dict[name] for name in Box._FIELDS))

    def __repr__(self):
        return f"Box(x_min: {self.x_min}, y_min: {self.y_min}, x_max: {self.x_max}, y_max: {self.y_max})"


class Object(_Scalable):
    """One detected structure: a named anchor keypoint plus the parts grouped onto it."""

    __slots__ = ("name", "anchor", "parts", "box")

    def __init__(self, name, anchor, parts=None, box=None):
        self.name, self.anchor, self.box = name, anchor, box
        self.parts = parts or []

    @property
    def x(self):
        return self.anchor.x

    @x.setter
    def x(self, value):
        self.anchor.x = value

    @property
    def y(self):
        return self.anchor.y

    @y.setter
    def y(self, value):
        self.anchor.y = value

    @property
    def nb_parts(self):
        return len(self.parts)

    def _members(self):
        yield self.anchor
        if self.box is not None:
            yield self.box
        yield from self.parts

    def resize(self, in_size, out_size):
        for member in self._members():
            member.resize(in_size, out_size)
        return self

    def normalize(self, size):
        for member in self._members():
            member.normalize(size)
        return self

    def distance(self, other):
        return self.anchor.distance(other.anchor)

    def json_repr(self):
        return {
            "label": self.name,
            "box": self.box.json_repr() if self.box else None,
            "parts": [kp.json_repr() for kp in (self.anchor, *self.parts)],
        }

    @staticmethod
    def from_json(json_dict, anchor_name):
        keypoints = [Keypoint.from_json(entry) for entry in json_dict["parts"]]
        anchors = [kp for kp in keypoints if kp.kind == anchor_name]
        assert len(anchors) <= 1, "More than one anchor found for object, achor must be unique."
        assert anchors, f"Anchor part with name '{anchor_name}' not found while decoding JSON file."
        others = [kp for kp in keypoints if kp.kind != anchor_name]
        return Object(json_dict["label"], anchors[0], others, Box.from_json(json_dict["box"]))

    def __repr__(self):
        return f"Object(name: {self.name}, anchor: {self.anchor}, parts: {self.parts}, box: {self.box})"


class ImageAnnotation(_Scalable):
    def __init__(self, image_path, objects=None, img_size=None):
        self.image_path = Path(image_path)
        self.objects = objects or []
        self.img_size = img_size

    image_name = property(lambda self: self.image_path.name)
    image_stem = property(lambda self: self.image_path.stem)
    nb_parts = property(lambda self: sum(obj.nb_parts for obj in self.objects))
    is_empty = property(lambda self: not self.objects)

    def __len__(self):
        return len(self.objects)

    def resize(self, in_size, out_size):
        for obj in self.objects:
            obj.resize(in_size, out_size)
        return self

    def normalize(self, size=None):
        size = size or self.img_size
        assert size, f"Annotation for '{self.image_path}' does not have a size."
        for obj in self.objects:
            obj.normalize(size)
        return self

    def json_repr(self):
        return {
            "image_path": str(self.image_path.expanduser().resolve()),
            "img_size": self.img_size,
            "objects": [obj.json_repr() for obj in self.objects],
        }

    def save_json(self, save_dir=None):
        target = Path(save_dir or "detections/")
        target.mkdir(parents=True, exist_ok=True)
        (target / self.image_path.with_suffix(".json").name).write_text(json.dumps(self.json_repr(), indent=2))

    @staticmethod
    def from_json(file: Path, anchor_name: str):
        data = json.loads(Path(file).read_text())
        objects = [Object.from_json(entry, anchor_name) for entry in data["objects"]]
        return ImageAnnotation(Path(data["image_path"]), objects, data.get("img_size", None))

    def __repr__(self):
        return f"ImageAnnotation(name: {self.image_name}, objects: {self.objects}, img_size: {self.img_size})"
