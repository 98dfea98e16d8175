"""GPU parity of the batched NNLS-spectrum post-processing against the outputs
of the reference's utility/spectrum.py (goldens) and against the oracle port on random spectra."""

from __future__ import annotations

import numpy as np
import pytest

from _util import load
from oracle import ref_port

pytestmark = pytest.mark.gpu

RTOL = 1e-12  # fractions / areas: same formulas, the summation order of <= 32 terms and pow() differ by ulps


def _check(r, ref, P, exact_d=True):
    assert np.array_equal(np.asarray(r["n_peaks"]), ref["n_peaks"])
    m = np.minimum(ref["n_peaks"], P)
    for v in range(len(m)):
        k = m[v]
        assert np.array_equal(r["peak_index"][v, :k], ref["peak_index"][v, :k]), v
        assert (r["peak_index"][v, k:] == -1).all()
        assert np.array_equal(r["d_values"][v, :k], ref["d_values"][v, :k]), v
        assert np.isnan(r["d_values"][v, k:]).all() and np.isnan(r["f_values"][v, k:]).all()
    full = ref["n_peaks"] <= P  # truncated voxels normalise over the stored peaks only
    np.testing.assert_allclose(r["f_values"][full], ref["f_values"][full][:, :P], rtol=RTOL, atol=0, equal_nan=True)
    if ref["d_cut"].shape[1]:
        np.testing.assert_allclose(r["d_cut"][full], ref["d_cut"][full], rtol=RTOL, atol=1e-15, equal_nan=True)
        np.testing.assert_allclose(r["f_cut"][full], ref["f_cut"][full], rtol=RTOL, atol=0, equal_nan=True)


@pytest.mark.parametrize("tag,height,reg", [("h0p1_reg", 0.1, True), ("h0p1_raw", 0.1, False), ("h5_reg", 5.0, True)])
def test_golden_parity(tag, height, reg):
    from pyneapple_b200 import spectrum

    g = load("spectrum_peaks")
    cut = [tuple(c) for c in g["cutoffs"]]
    r = spectrum.find_spectrum_peaks_batch(g["spectra"], g["bins"], height, reg, cutoffs=cut, max_peaks=32)
    ref = dict(n_peaks=g[f"{tag}_n_peaks"], peak_index=g[f"{tag}_idx"], d_values=g[f"{tag}_d"],
               f_values=g[f"{tag}_f"], d_cut=g[f"{tag}_d_cut"], f_cut=g[f"{tag}_f_cut"])
    _check(r, ref, 32)


def test_truncation_reports_the_true_count():
    from pyneapple_b200 import spectrum

    g = load("spectrum_peaks")
    r = spectrum.find_spectrum_peaks_batch(g["spectra"], g["bins"], 0.1, True, max_peaks=4)
    assert np.array_equal(r["n_peaks"], g["h0p1_reg_n_peaks"])
    assert r["n_peaks"].max() > 4 and r["d_values"].shape == (g["spectra"].shape[0], 4)
    few = g["h0p1_reg_n_peaks"] <= 4
    np.testing.assert_allclose(r["f_values"][few], g["h0p1_reg_f"][few][:, :4], rtol=RTOL, equal_nan=True)


def test_random_spectra_against_the_oracle_port_and_chunking():
    from pyneapple_b200 import spectrum

    rng = np.random.default_rng(3)
    n_vox, n = 6000, 250
    bins = np.geomspace(0.0008, 0.5, n)
    centers = rng.uniform(10, 240, (n_vox, 3)); widths = rng.uniform(1.5, 12, (n_vox, 3)); amps = rng.uniform(0, 50, (n_vox, 3))
    j = np.arange(n)[None, None, :]
    x = (amps[..., None] * np.exp(-0.5 * ((j - centers[..., None]) / widths[..., None]) ** 2)).sum(1)
    x[x < 1e-3] = 0.0
    x[::7] = np.round(x[::7])  # plateaus
    cut = [(0.0008, 0.004), (0.004, 0.06), (0.06, 0.5), (1.0, 2.0)]
    ref = ref_port.spectrum_peaks(x, bins, 0.5, True, cut, 8)
    r = spectrum.find_spectrum_peaks_batch(x, bins, 0.5, True, cutoffs=cut, max_peaks=8)
    _check(r, ref, 8)
    r2 = spectrum.find_spectrum_peaks_batch(x, bins, 0.5, True, cutoffs=cut, max_peaks=8, chunk_vox=1000)
    for k in r:
        assert np.array_equal(r[k], r2[k], equal_nan=True), k


def test_device_tensors_stay_on_the_device_and_chain_with_nnls():
    import torch

    from pyneapple_b200 import engine, models, spectrum, synth
    from pyneapple_b200.solvers.nnls import regularization_matrix

    cfg = synth.CONFIGS["C3"]
    b, y, _ = synth.sample_voxels(cfg, 512)
    model = models.NNLSModel(d_range=(0.0008, 0.5), n_bins=250)
    fit = engine.nnls_fit(model.get_basis(b), regularization_matrix(250, 2, 0.02), torch.as_tensor(y).cuda(), 250)
    coef = fit["coefficients"]
    cut = [(0.0008, 0.003), (0.003, 0.05), (0.05, 0.5)]
    r = spectrum.find_spectrum_peaks_batch(coef, model.bins, 0.1, True, cutoffs=cut)
    assert all(v.is_cuda for v in r.values())
    ref = ref_port.spectrum_peaks(coef.cpu().numpy(), model.bins, 0.1, True, cut, 8)
    _check({k: v.cpu().numpy() for k, v in r.items()}, ref, 8)
    assert (ref["n_peaks"] > 0).mean() > 0.9


def test_single_spectrum_mirrors_of_the_reference_functions():
    from scipy import signal as scipy_signal

    from pyneapple_b200 import spectrum

    g = load("spectrum_peaks")
    bins = g["bins"]
    for v in (0, 5, 40, 100, 130, 137, 139, 140):
        x = g["spectra"][v]
        d, f = spectrum.find_spectrum_peaks(x, bins, height=0.1, regularized=True)
        k = int(g["h0p1_reg_n_peaks"][v])
        assert d.shape == (k,) and f.shape == (k,)
        np.testing.assert_array_equal(d, g["h0p1_reg_d"][v, :k])
        np.testing.assert_allclose(f, g["h0p1_reg_f"][v, :k], rtol=RTOL)
        idx, props = scipy_signal.find_peaks(x, height=0.1)
        if k:
            fw = scipy_signal.peak_widths(x, idx, rel_height=0.5)[0]
            want = [float(h * w / (2 * np.sqrt(2 * np.log(2))) * np.sqrt(2 * np.pi)) for h, w in zip(props["peak_heights"], fw)]
            np.testing.assert_allclose(spectrum.calculate_peak_area(x, idx, props["peak_heights"]), want, rtol=1e-15)
            dn, fn = spectrum.apply_cutoffs(d, f, [tuple(c) for c in g["cutoffs"]])
            np.testing.assert_allclose(dn, g["h0p1_reg_d_cut"][v], rtol=RTOL, atol=1e-15, equal_nan=True)
            np.testing.assert_allclose(fn, g["h0p1_reg_f_cut"][v], rtol=RTOL, equal_nan=True)
    # geometric_mean_peak (utility/spectrum.py:106-136)
    pos, h = spectrum.geometric_mean_peak([1e-3, 4e-3], [1.0, 1.0])
    assert pos == pytest.approx(np.log10(2e-3), rel=1e-12) and h == pytest.approx(2.0)
    pos, h = spectrum.geometric_mean_peak([2e-3], [0.7])
    assert pos == pytest.approx(np.log10(2e-3), rel=1e-12) and h == pytest.approx(0.7)
    d, f = spectrum.find_spectrum_peaks(np.zeros(250), bins)
    assert d.size == 0 and f.size == 0
    dn, fn = spectrum.apply_cutoffs(d, f, [(0.001, 0.01)])
    assert np.isnan(dn).all() and np.isnan(fn).all()


def test_solver_fit_spectrum_peaks_equals_fit_then_postprocess():
    from pyneapple_b200 import models, spectrum, synth
    from pyneapple_b200.solvers import NNLSSolver

    cfg = synth.CONFIGS["C3"]
    b, y, _ = synth.sample_voxels(cfg, 3000)
    model = models.NNLSModel(d_range=(0.0008, 0.5), n_bins=250)
    cut = [(0.0008, 0.003), (0.003, 0.05), (0.05, 0.5)]
    solver = NNLSSolver(model=model, reg_order=2, mu=0.02, max_iter=250)
    fused = solver.fit_spectrum_peaks(b, y, height=0.1, cutoffs=cut, chunk_vox=1024)
    solver.fit(b, y)
    two = spectrum.find_spectrum_peaks_batch(solver.params_["coefficients"], model.bins, 0.1, True, cutoffs=cut)
    for k, v in two.items():
        assert np.array_equal(fused[k], v, equal_nan=True), k
    assert np.array_equal(fused["status"], solver.status_)
    np.testing.assert_array_equal(fused["residual"], solver.diagnostics_["residual"])


def test_fused_pipeline_with_large_chunks_orders_uploads_after_the_kernels():
    """Chunks of 8 MB take the staged (pageable) ``pnb_upload`` path, and with four of them the
    caching allocator hands chunk i the block of chunk i-2 while that chunk's NNLS kernel may still
    be reading it: the upload has to be ordered after the work queued on the stream."""
    from pyneapple_b200 import models, spectrum, synth
    from pyneapple_b200.solvers import NNLSSolver

    cfg = synth.CONFIGS["C3"]
    b, img, _ = synth.make_volume(cfg, 0, 5)
    y = np.ascontiguousarray(img.reshape(-1, 16)[: 4 * 65536 + 777])
    assert y[:65536].nbytes >= (4 << 20)
    model = models.NNLSModel(d_range=(0.0008, 0.5), n_bins=250)
    solver = NNLSSolver(model=model, reg_order=2, mu=0.02, max_iter=250)
    fused = solver.fit_spectrum_peaks(b, y, height=0.1, chunk_vox=65536)
    solver.fit(b, y)
    two = spectrum.find_spectrum_peaks_batch(solver.params_["coefficients"], model.bins, 0.1, True)
    for k, v in two.items():
        if v is not None:  # no cut-off ranges were asked for
            assert np.array_equal(fused[k], v, equal_nan=True), k
    assert np.array_equal(fused["status"], solver.status_)
    np.testing.assert_array_equal(fused["residual"], solver.diagnostics_["residual"])


def test_to_device_loop_does_not_overwrite_memory_a_queued_kernel_still_reads():
    import torch

    from pyneapple_b200 import engine, models, synth
    from pyneapple_b200.solvers.nnls import regularization_matrix

    cfg = synth.CONFIGS["C3"]
    b, img, _ = synth.make_volume(cfg, 0, 3)
    y = np.ascontiguousarray(img.reshape(-1, 16)[: 3 * 49152])
    model = models.NNLSModel(d_range=(0.0008, 0.5), n_bins=250)
    basis, R = model.get_basis(b), regularization_matrix(250, 2, 0.02)
    whole = engine.nnls_fit(basis, R, torch.as_tensor(y).cuda(), 250)["coefficients"].cpu().numpy()
    got = []
    for s0 in range(0, y.shape[0], 49152):
        yd = engine.to_device(y[s0:s0 + 49152], torch.device("cuda", 0))  # 6 MB: staged upload
        fit = engine.nnls_fit(basis, R, yd, 250)
        got.append(fit["coefficients"])
        del yd, fit  # the block goes back to the allocator while the kernel is still queued
    got = torch.cat(got).cpu().numpy()
    assert np.array_equal(got, whole)
