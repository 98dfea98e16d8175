"""Input checks and dict -> array packing with the reference's error behaviour.

Mirrors ``utility/validation.py:84-112`` (shape checks), ``:153-173``
(parameter-name checks), ``:177-203`` / ``:248-297`` (p0 / bounds packing).
Unlike the reference, scalar p0 / bounds are NOT tiled to ``(n_params,
n_pixels)``: the device broadcasts a single vector.
"""

from __future__ import annotations

import logging

import numpy as np

log = logging.getLogger("pyneapple_b200")


def validate_xdata(xdata: np.ndarray) -> None:
    if xdata.ndim != 1:
        raise ValueError(f"xdata must be a 1D array, but got shape {xdata.shape}.")


def validate_data_shapes(xdata: np.ndarray, ydata: np.ndarray) -> None:
    if xdata.ndim != 1:
        raise ValueError(f"xdata must be a 1D array, but got shape {xdata.shape}.")
    if ydata.ndim == 1:
        if ydata.shape[0] != xdata.shape[0]:
            raise ValueError(
                f"ydata length {ydata.shape[0]} does not match xdata length {xdata.shape[0]}."
            )
    elif ydata.ndim >= 2:
        if ydata.shape[-1] != xdata.shape[0]:
            raise ValueError(
                f"ydata second dimension {ydata.shape[-1]} does not match xdata length {xdata.shape[0]}."
            )
    else:
        raise ValueError(f"ydata must be 1D or 2D array, but got shape {ydata.shape}.")


def validate_segmentation(segmentation: np.ndarray, image_shape: tuple) -> np.ndarray:
    if segmentation.ndim != len(image_shape) - 1:
        if segmentation.shape[-1] == 1:
            log.warning("Segmentation has a singleton channel dimension %s; squeezing.", segmentation.shape)
            segmentation = np.squeeze(segmentation, axis=-1)
        else:
            raise ValueError(
                f"Segmentation must have one less dimension than image shape {image_shape}, "
                f"but got shape {segmentation.shape}."
            )
    if segmentation.shape != image_shape[:-1]:
        raise ValueError(
            f"Segmentation shape {segmentation.shape} does not match expected image shape {image_shape[:-1]}."
        )
    return segmentation


def validate_parameter_names(parameters: dict, param_names: list[str]) -> None:
    missing = set(param_names) - set(parameters.keys())
    if missing:
        raise ValueError(
            f"Missing bounds for required parameters: {missing}. Required: {param_names}"
        )
    extra = set(parameters.keys()) - set(param_names)
    if extra:
        log.warning("Extra bounds will be ignored: %s", extra)


def validate_fixed_params(fixed_params: dict, all_param_names: list[str]) -> None:
    if not fixed_params:
        return
    unknown = set(fixed_params.keys()) - set(all_param_names)
    if unknown:
        raise ValueError(
            f"Unknown fixed parameter(s): {sorted(unknown)}. Valid names: {all_param_names}"
        )
    if len(fixed_params) >= len(all_param_names):
        raise ValueError(
            "Cannot fix all parameters — at least one must remain free. "
            f"Fixed: {sorted(fixed_params.keys())}, all: {all_param_names}"
        )


def validate_fixed_param_maps(fixed_param_maps: dict, spatial_shape: tuple, all_param_names) -> None:
    validate_fixed_params(fixed_param_maps, all_param_names)
    for name, arr in fixed_param_maps.items():
        if arr.shape != spatial_shape:
            raise ValueError(
                f"Fixed param map '{name}' has shape {arr.shape}, expected spatial shape {spatial_shape}."
            )


def p0_vector(p0: dict, param_names: list[str]) -> np.ndarray:
    """dict of scalars -> ``(n_params,)`` in ``param_names`` order."""
    validate_parameter_names(p0, param_names)
    if not isinstance(p0[param_names[0]], (int, float)):
        raise ValueError(
            "p0 must be provided as a dictionary with parameter names as keys and float values."
        )
    return np.array([p0[name] for name in param_names], dtype=float)


def bounds_vectors(bounds: dict, param_names: list[str]) -> tuple[np.ndarray, np.ndarray]:
    validate_parameter_names(bounds, param_names)
    first = bounds[param_names[0]]
    if not (isinstance(first[0], (int, float)) and isinstance(first[1], (int, float))):
        raise ValueError(
            "Bounds must be provided as a dictionary with parameter names as keys and "
            "(lower, upper) tuples as values, where lower and upper are floats."
        )
    lower = np.array([bounds[p][0] for p in param_names], dtype=float)
    upper = np.array([bounds[p][1] for p in param_names], dtype=float)
    return lower, upper
