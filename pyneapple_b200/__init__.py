"""pyneapple_b200 — B200-native batched voxel-fitting engine behind Pyneapple's solver/fitter API."""

__version__ = "0.1.0"
