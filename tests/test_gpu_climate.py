"""GPU parity of the per-vertex climate kernels (climate.py:345-597) against golden
vectors computed by the reference and, at a larger size, against the CPU oracle.

Tolerance: the reference accumulates float64 terms into float32 arrays.  The device evaluates asin /
atan2 / cos with CUDA's libm (<= 2 ulp of float64) instead of glibc's, so a term can differ in its
last float64 bit and the float32 rounding of a sum can flip: results must agree to 4 float32 ulps of
the result's magnitude (2.4e-7 relative) after 360 accumulations, and the fraction of entries that
are not bit-identical is reported."""
import numpy as np
import pytest

from oracle import icosphere

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cl():
    import torch
    assert torch.cuda.is_available()
    from nixis_b200 import climate
    return climate


def _close(got, ref, what, ulps=4):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.dtype == np.float32 and got.shape == ref.shape, what
    tol = ulps * np.spacing(np.maximum(np.abs(ref), np.float32(1e-3)).astype(np.float32))
    bad = np.abs(got.astype(np.float64) - ref.astype(np.float64)) > tol
    assert not bad.any(), (what, int(bad.sum()), float(np.abs(got - ref).max()))
    return float((got != ref).mean())


@pytest.mark.parametrize("k", [8, 12])
def test_climate_vs_reference_golden(cl, golden, k):
    g = golden("climate")
    t = f"k{k}"
    R = float(g[f"{t}_R"][0])
    pts, _ = icosphere.icosa_sphere(k)
    P = np.ascontiguousarray(pts * R)
    h = g[f"{t}_height"]
    assert np.array_equal(g["seasonal_tilt"], np.array([cl.calculate_seasonal_tilt(23.44, d) for d in range(360)]))
    inexact = []
    for i, tilt in enumerate(g[f"{t}_tilts"]):
        inexact.append(_close(cl.assign_surface_temp(P, h, R, tilt), g[f"{t}_temp_{i}"], "temp"))
        inexact.append(_close(cl.calc_insolation_slice(R, tilt), g[f"{t}_slice_{i}"], "slice"))
        inexact.append(_close(cl.calc_daily_insolation(P, h, R, tilt), g[f"{t}_daily_{i}"], "daily"))
        for j, rot in enumerate((0, 37.5, -180.0)):
            inexact.append(_close(cl.calc_instant_insolation(P, h, R, rot, tilt), g[f"{t}_instant_{i}_{j}"], "instant"))
    inexact.append(_close(cl.brute_daily_insolation(P, h, R, g[f"{t}_tilts"][0]), g[f"{t}_brute_0"], "brute"))
    inexact.append(_close(cl.calc_yearly_insolation(P, h, R, 23.44), g[f"{t}_yearly"], "yearly", ulps=8))
    print(f"k={k}: fraction of entries not bit-identical to the reference, per call: max {max(inexact):.4f}")
    # in-place forms (climate.py:415, 551)
    arr = g[f"{t}_instant_0_0"].copy()
    cl.sample_insolation(arr, P, R, 37.5, g[f"{t}_tilts"][0])
    ref = g[f"{t}_instant_0_0"].astype(np.float64) + g[f"{t}_instant_0_1"].astype(np.float64)
    assert np.allclose(arr, ref, rtol=1e-6, atol=1e-6)
    out = np.zeros(len(P), dtype=np.float32)
    cl.interpolate_insolation(P, g[f"{t}_slice_1"], out, R)
    assert np.array_equal(out, g[f"{t}_daily_1"])      # fed the reference's own table: only asin differs


def test_climate_vs_oracle_k96_device_resident(cl):
    """92 162 vertices, Earth radius; positions stay on the device (float64 CUDA tensor in, CUDA out)."""
    import torch
    from oracle import oracle
    from nixis_b200 import runtime as rt
    k, R = 96, 6378100.0
    pts, _ = icosphere.icosa_sphere(k)
    P = np.ascontiguousarray(pts * R)
    dev = rt.mesh_points(k, f32=False, f64=True)[1] * R
    assert np.array_equal(dev.cpu().numpy(), P)
    tilt = cl.calculate_seasonal_tilt(23.44, 36)
    h = np.zeros(len(P))
    y = cl.calc_yearly_insolation(dev, None, R, 23.44)
    assert isinstance(y, torch.Tensor) and y.is_cuda
    _close(y.cpu().numpy(), oracle.calc_yearly_insolation(P, h, R, 23.44), "yearly", ulps=8)
    _close(cl.brute_daily_insolation(dev, None, R, tilt).cpu().numpy(), oracle.brute_daily_insolation(P, h, R, tilt), "brute")
    _close(cl.assign_surface_temp(dev, None, R, tilt).cpu().numpy(), oracle.assign_surface_temp(P, np.arange(len(P), dtype=np.float64), R, tilt), "temp")
    # physical sanity: the yearly mean is symmetric about the equator and largest there
    lat = np.degrees(np.arcsin(np.clip(P[:, 2] / R, -1, 1)))
    yy = y.cpu().numpy()
    assert yy[np.abs(lat) < 5].mean() > yy[np.abs(lat) > 70].mean()


def test_climate_timing_d1000(cl):
    """Config-scale run (d=1000, 10 000 002 vertices): finite results, one launch per driver."""
    import torch
    from nixis_b200 import runtime as rt, _lib
    k, R = 1000, 6378100.0
    dev = rt.mesh_points(k, f32=False, f64=True)[1] * R
    tilt = cl.calculate_seasonal_tilt(23.44, 36)
    res = {}
    for name, fn in (("assign_surface_temp", lambda: cl.assign_surface_temp(dev, None, R, tilt)),
                     ("calc_instant_insolation", lambda: cl.calc_instant_insolation(dev, None, R, 0, tilt)),
                     ("calc_daily_insolation", lambda: cl.calc_daily_insolation(dev, None, R, tilt)),
                     ("calc_yearly_insolation", lambda: cl.calc_yearly_insolation(dev, None, R, 23.44)),
                     ("brute_daily_insolation", lambda: cl.brute_daily_insolation(dev, None, R, tilt))):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        res[name] = (e0.elapsed_time(e1), _lib.launch_count - n0)
        assert torch.isfinite(out).all()
    print("d=1000 climate kernels (ms, launches):", {k_: (round(v[0], 3), v[1]) for k_, v in res.items()})
    assert res["brute_daily_insolation"][1] == 1 and res["calc_yearly_insolation"][1] == 2
