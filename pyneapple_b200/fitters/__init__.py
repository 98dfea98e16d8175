"""B200 fitters with the reference's names and registry (fitters/__init__.py:9-43)."""

from __future__ import annotations

from .base import BaseFitter, PixelIndices
from .ideal import IDEALFitter
from .pixelwise import PixelWiseFitter
from .segmentationwise import SegmentationWiseFitter
from .segmented import SegmentedFitter

_REGISTRY: dict[str, type] = {
    "pixelwise": PixelWiseFitter,
    "segmentationwise": SegmentationWiseFitter,
    "ideal": IDEALFitter,
    "segmented": SegmentedFitter,
}


def get_fitter(name: str, **kwargs) -> BaseFitter:
    key = name.lower()
    if key not in _REGISTRY:
        raise ValueError(f"Unknown fitter: {name!r}. Available: {sorted(_REGISTRY)}")
    return _REGISTRY[key](**kwargs)


__all__ = ["BaseFitter", "IDEALFitter", "PixelIndices", "PixelWiseFitter", "SegmentationWiseFitter",
           "SegmentedFitter", "get_fitter"]
