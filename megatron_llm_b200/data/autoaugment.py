"""AutoAugment ImageNet policy (Cubuk et al. 2018; parity: megatron/data/autoaugment.py -- unused by any shipped
model, kept for API completeness).  A policy is 25 sub-policies of two (operation, probability, magnitude index)
steps; one sub-policy is drawn per image."""
from __future__ import annotations

import random

import numpy as np

_POLICY = [  # (op1, p1, mag1, op2, p2, mag2)
    ("posterize", 0.4, 8, "rotate", 0.6, 9), ("solarize", 0.6, 5, "autocontrast", 0.6, 5),
    ("equalize", 0.8, 8, "equalize", 0.6, 3), ("posterize", 0.6, 7, "posterize", 0.6, 6),
    ("equalize", 0.4, 7, "solarize", 0.2, 4), ("equalize", 0.4, 4, "rotate", 0.8, 8),
    ("solarize", 0.6, 3, "equalize", 0.6, 7), ("posterize", 0.8, 5, "equalize", 1.0, 2),
    ("rotate", 0.2, 3, "solarize", 0.6, 8), ("equalize", 0.6, 8, "posterize", 0.4, 6),
    ("rotate", 0.8, 8, "color", 0.4, 0), ("rotate", 0.4, 9, "equalize", 0.6, 2),
    ("equalize", 0.0, 7, "equalize", 0.8, 8), ("invert", 0.6, 4, "equalize", 1.0, 8),
    ("color", 0.6, 4, "contrast", 1.0, 8), ("rotate", 0.8, 8, "color", 1.0, 2),
    ("color", 0.8, 8, "solarize", 0.8, 7), ("sharpness", 0.4, 7, "invert", 0.6, 8),
    ("shearX", 0.6, 5, "equalize", 1.0, 9), ("color", 0.4, 0, "equalize", 0.6, 3),
    ("equalize", 0.4, 7, "solarize", 0.2, 4), ("solarize", 0.6, 5, "autocontrast", 0.6, 5),
    ("invert", 0.6, 4, "equalize", 1.0, 8), ("color", 0.6, 4, "contrast", 1.0, 8),
    ("equalize", 0.8, 8, "equalize", 0.6, 3)]

_RANGES = {"shearX": np.linspace(0, 0.3, 10), "shearY": np.linspace(0, 0.3, 10),
           "translateX": np.linspace(0, 150 / 331, 10), "translateY": np.linspace(0, 150 / 331, 10),
           "rotate": np.linspace(0, 30, 10), "color": np.linspace(0.0, 0.9, 10),
           "posterize": np.round(np.linspace(8, 4, 10), 0).astype(int), "solarize": np.linspace(256, 0, 10),
           "contrast": np.linspace(0.0, 0.9, 10), "sharpness": np.linspace(0.0, 0.9, 10),
           "brightness": np.linspace(0.0, 0.9, 10), "autocontrast": [0] * 10, "equalize": [0] * 10, "invert": [0] * 10}


def _apply(op, img, magnitude, fillcolor):
    from PIL import Image, ImageEnhance, ImageOps
    sign = random.choice([-1, 1])
    if op in ("shearX", "shearY", "translateX", "translateY"):
        m = magnitude * sign
        coeffs = {"shearX": (1, m, 0, 0, 1, 0), "shearY": (1, 0, 0, m, 1, 0),
                  "translateX": (1, 0, m * img.size[0], 0, 1, 0), "translateY": (1, 0, 0, 0, 1, m * img.size[1])}[op]
        return img.transform(img.size, Image.AFFINE, coeffs, Image.BICUBIC, fillcolor=fillcolor)
    if op == "rotate":          # rotate on an RGBA canvas so the corners take the fill colour
        rot = img.convert("RGBA").rotate(magnitude * sign)
        return Image.composite(rot, Image.new("RGBA", rot.size, (128,) * 4), rot).convert(img.mode)
    if op in ("color", "contrast", "sharpness", "brightness"):
        enh = {"color": ImageEnhance.Color, "contrast": ImageEnhance.Contrast, "sharpness": ImageEnhance.Sharpness,
               "brightness": ImageEnhance.Brightness}[op]
        return enh(img).enhance(1 + magnitude * sign)
    if op == "posterize":
        return ImageOps.posterize(img, int(magnitude))
    if op == "solarize":
        return ImageOps.solarize(img, magnitude)
    return {"autocontrast": ImageOps.autocontrast, "equalize": ImageOps.equalize, "invert": ImageOps.invert}[op](img)


class SubPolicy:
    def __init__(self, operation1, probability1, magnitude_idx1, operation2, probability2, magnitude_idx2,
                 fillcolor=(128, 128, 128)):
        self.steps = [(operation1, probability1, _RANGES[operation1][magnitude_idx1]),
                      (operation2, probability2, _RANGES[operation2][magnitude_idx2])]
        self.fillcolor = fillcolor

    def __call__(self, img):
        for op, p, mag in self.steps:
            if random.random() < p:
                img = _apply(op, img, mag, self.fillcolor)
        return img


class ImageNetPolicy:
    """``transform = transforms.Compose([transforms.Resize(256), ImageNetPolicy(), transforms.ToTensor()])``"""

    def __init__(self, fillcolor=(128, 128, 128)):
        self.policies = [SubPolicy(o1, p1, m1, o2, p2, m2, fillcolor) for o1, p1, m1, o2, p2, m2 in _POLICY]

    def __call__(self, img):
        return random.choice(self.policies)(img)

    def __repr__(self):
        return "ImageNetPolicy"
