"""Device-side per-label mean signals against NumPy's
np.mean(image[segmentation == seg], axis=0) — fitters/segmentationwise.py:112-137."""

from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _numpy_means(image, seg):
    labels = np.unique(seg)
    return labels, np.array([np.mean(image[seg == s], axis=0) for s in labels])


@pytest.mark.parametrize("shape,n_b,n_labels", [((40, 40, 8), 16, 5), ((33, 17, 3), 40, 3), ((64, 64, 4), 16, 300),
                                                ((7, 5, 1), 1, 2), ((128, 128, 16), 16, 9)])
def test_means_match_numpy_and_are_deterministic(shape, n_b, n_labels):
    from pyneapple_b200 import engine

    rng = np.random.default_rng(n_labels)
    image = rng.uniform(0, 2000, shape + (n_b,))
    seg = rng.integers(0, n_labels, shape) * 3  # non-contiguous label values, background 0 included
    seg[: shape[0] // 2] = np.sort(seg[: shape[0] // 2], axis=0)  # long runs of equal labels
    labels, ref = _numpy_means(image, seg)
    got_labels, means, counts, inverse = engine.segment_means(image, seg)
    assert np.array_equal(got_labels, labels)
    assert np.array_equal(counts, [np.count_nonzero(seg == s) for s in labels])
    assert np.array_equal(labels[inverse].reshape(shape), seg)
    np.testing.assert_allclose(means, ref, rtol=1e-13, atol=0)
    again = engine.segment_means(image, seg)[1]
    assert np.array_equal(means, again)  # no floating-point atomics: bit-reproducible


def test_device_tensors_and_the_fitter():
    import torch

    from _util import load
    from pyneapple_b200 import engine

    g = load("fitter_pixelwise")
    image, seg = g["image"], g["seg"]
    labels, ref = _numpy_means(image, seg)
    lab_d, means_d, counts_d, _ = engine.segment_means(torch.as_tensor(image).cuda(), torch.as_tensor(seg).cuda())
    assert means_d.is_cuda and np.array_equal(lab_d, labels)
    np.testing.assert_allclose(means_d.cpu().numpy(), ref, rtol=1e-13)
    assert int(counts_d.sum()) == seg.size


def test_staged_transfers_round_trip():
    """pnb_upload / pnb_download (pageable host memory through page-locked bounce blocks)."""
    import torch

    from pyneapple_b200 import engine

    rng = np.random.default_rng(0)
    for shape, dtype in (((3, 1_000_003), np.float64), ((70 << 20,), np.uint8), ((513, 257, 9), np.float32),
                         ((12_345_678,), np.int32), ((10,), np.float64)):
        a = (rng.random(shape) * 200).astype(dtype)
        t = engine.to_device(a, "cuda:0")
        assert t.is_cuda and tuple(t.shape) == a.shape
        assert np.array_equal(t.cpu().numpy(), a)
        back = engine.to_host(t * 1)
        assert back.dtype == a.dtype and np.array_equal(back, a)
    nc = torch.arange(6_000_000, dtype=torch.float64, device="cuda").reshape(2000, 3000).T  # non-contiguous
    assert np.array_equal(engine.to_host(nc), nc.cpu().numpy())
