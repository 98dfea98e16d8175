"""GPU: the resampler against cv2.resize outputs (bit-exact for the FP64 paths)."""

import numpy as np
import pytest

from _util import load

pytestmark = pytest.mark.gpu

from pyneapple_b200.resize import interpolate_array  # noqa: E402


def test_cubic_f64_bit_exact():
    g = load("resize_cv2")
    for n in range(int(g["n_cases"])):
        src, ref = g[f"src{n}"], g[f"cubic{n}"]
        got = interpolate_array(src, ref.shape[:2], "cubic")
        assert got.dtype == np.float64
        assert np.array_equal(got, ref), f"case {n}: max diff {np.abs(got - ref).max()}"


def test_linear_f64_bit_exact_vs_opencv_generic_path():
    g = load("resize_cv2")
    for n in range(int(g["n_cases"])):
        src, ref = g[f"src{n}"], g[f"linear{n}"]
        got = interpolate_array(src, ref.shape[:2], "linear")
        assert np.array_equal(got, ref), f"case {n}: max diff {np.abs(got - ref).max()}"


def test_label_mask_f32_dyadic():
    g = load("resize_cv2")
    seg = g["seg_src"]
    for n, t in enumerate((4, 8, 16, 32)):
        got = interpolate_array(seg, (t, t), "cubic")  # integer input -> float32, like ideal.py:310-311
        assert got.dtype == np.float32
        assert np.array_equal(got, g[f"seg{n}"])
        assert np.array_equal(got[..., 0] > 0.2, g[f"seg{n}"][..., 0] > 0.2)


def test_device_tensor_path_and_identity():
    import torch

    g = load("resize_cv2")
    src = g["src2"]
    host = interpolate_array(src, (64, 64), "cubic")
    dev = interpolate_array(torch.as_tensor(src).cuda(), (64, 64), "cubic").cpu().numpy()
    assert np.array_equal(host, dev)
    same = interpolate_array(src, src.shape[:2], "cubic")
    np.testing.assert_allclose(same, src, rtol=1e-4)  # tests/test_fitter_ideal.py:269-318 identity check
    with pytest.raises(ValueError):
        interpolate_array(src, (4, 4), "nearest")
