"""Folder-of-classes image dataset with class / per-class subsampling (parity: megatron/data/image_folder.py -- kept
for API completeness; like in the reference no vision model uses it).

``root/<class>/**/<image>``; ``classes_fraction`` keeps the first fraction of the (sorted) classes and
``data_per_class_fraction`` the first fraction of every class's files."""
from __future__ import annotations

import os
from typing import Any, Callable, Dict, List, Optional, Tuple

from torch.utils.data import Dataset

IMG_EXTENSIONS = (".jpg", ".jpeg", ".png", ".ppm", ".bmp", ".pgm", ".tif", ".tiff", ".webp")


def has_file_allowed_extension(filename: str, extensions: Tuple[str, ...]) -> bool:
    return filename.lower().endswith(extensions)


def is_image_file(filename: str) -> bool:
    return has_file_allowed_extension(filename, IMG_EXTENSIONS)


def make_dataset(directory: str, class_to_idx: Dict[str, int], data_per_class_fraction: float,
                 extensions: Optional[Tuple[str, ...]] = None,
                 is_valid_file: Optional[Callable[[str], bool]] = None) -> List[Tuple[str, int]]:
    if (extensions is None) == (is_valid_file is None):
        raise ValueError("Both extensions and is_valid_file cannot be None or not None at the same time")
    valid = is_valid_file or (lambda p: has_file_allowed_extension(p, extensions))
    directory = os.path.expanduser(directory)
    instances = []
    for cls in sorted(class_to_idx):
        cls_dir = os.path.join(directory, cls)
        if not os.path.isdir(cls_dir):
            continue
        files = [os.path.join(r, f) for r, _, fs in sorted(os.walk(cls_dir, followlinks=True)) for f in sorted(fs)]
        files = [f for f in files if valid(f)]
        keep = int(len(files) * data_per_class_fraction)
        instances.extend((f, class_to_idx[cls]) for f in files[:keep])
    return instances


def pil_loader(path: str):
    from PIL import Image
    with open(path, "rb") as f:
        return Image.open(f).convert("RGB")


def default_loader(path: str) -> Any:
    return pil_loader(path)


class DatasetFolder(Dataset):
    def __init__(self, root: str, loader: Callable[[str], Any], extensions: Optional[Tuple[str, ...]] = None,
                 transform: Optional[Callable] = None, target_transform: Optional[Callable] = None,
                 classes_fraction=1.0, data_per_class_fraction=1.0, is_valid_file: Optional[Callable] = None):
        self.root, self.transform, self.target_transform = root, transform, target_transform
        self.classes_fraction, self.data_per_class_fraction = classes_fraction, data_per_class_fraction
        self.classes, self.class_to_idx = self._find_classes(root)
        self.samples = self.make_dataset(root, self.class_to_idx, data_per_class_fraction, extensions, is_valid_file)
        if not self.samples:
            raise RuntimeError(f"Found 0 files in subfolders of: {root}"
                               + (f"\nSupported extensions are: {','.join(extensions)}" if extensions else ""))
        self.loader, self.extensions = loader, extensions
        self.total = len(self.samples)
        self.targets = [t for _, t in self.samples]

    make_dataset = staticmethod(make_dataset)

    def _find_classes(self, directory: str):
        classes = sorted(d.name for d in os.scandir(directory) if d.is_dir())
        classes = classes[:int(len(classes) * self.classes_fraction)]
        return classes, {c: i for i, c in enumerate(classes)}

    def __getitem__(self, index: int):
        for attempt in range(self.total):        # skip unreadable files instead of killing the epoch
            path, target = self.samples[(index + attempt) % self.total]
            try:
                sample = self.loader(path)
                break
            except Exception as e:
                print(f"could not read {path}: {e}")
        if self.transform is not None:
            sample = self.transform(sample)
        if self.target_transform is not None:
            target = self.target_transform(target)
        return sample, target

    def __len__(self) -> int:
        return self.total


class ImageFolder(DatasetFolder):
    def __init__(self, root: str, transform=None, target_transform=None, classes_fraction=1.0,
                 data_per_class_fraction=1.0, loader: Callable[[str], Any] = default_loader, is_valid_file=None):
        super().__init__(root, loader, IMG_EXTENSIONS if is_valid_file is None else None, transform=transform,
                         target_transform=target_transform, classes_fraction=classes_fraction,
                         data_per_class_fraction=data_per_class_fraction, is_valid_file=is_valid_file)
        self.imgs = self.samples
