"""kx-ky domain filters: oracle and host logic against outputs of the reference's own functions
(tests/golden/reference_postprocessing.npz, made by oracle/make_golden_postprocessing.py), GPU path against both."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import postprocessing_oracle as orc                      # noqa: E402
from oracle.make_golden_postprocessing import CASES, make_slice     # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "reference_postprocessing.npz")
RTOL32 = 2e-6        # fp32 transform pair against the float64 reference (linear operator: no conditioning issue)


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a, dtype=np.complex128) - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _oracle_filter(c):
    kw = dict(c["kw"])
    if c["fn"] == "footprint":
        kw["direction"] = {"both": "both", "iline": "horizontal", "xline": "vertical"}[kw["direction"]]
        return orc.footprint_filter(c["shape"], **kw)
    f = kw["factors_upsampling"]
    return orc.antialias_filter(c["shape"], {"iline": "horizontal", "xline": "vertical"}[kw["direction"]], f["iline"], f["xline"], kw["sigma"])


@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference(c, gold):
    f = _oracle_filter(c)
    np.testing.assert_allclose(f, gold[c["name"] + "__filter"], rtol=0, atol=1e-13)
    y = orc.apply(make_slice(c), f)
    np.testing.assert_allclose(y, gold[c["name"] + "__y"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_host_filter_construction_matches_reference(c, gold):
    """the product's host-side filter planes (numpy/scipy, no GPU involved) equal the reference's."""
    from pseudo_3d_interpolation_b200 import cube_postprocessing_3D as pp
    if c["fn"] == "footprint":
        f = pp.footprint_filter(c["shape"], **c["kw"])
    else:
        f = pp.antialiasing_filter(c["shape"], **c["kw"])
    np.testing.assert_allclose(f, gold[c["name"] + "__filter"], rtol=0, atol=1e-13)


def test_host_helpers():
    from pseudo_3d_interpolation_b200 import cube_postprocessing_3D as pp
    k = pp.gaussian_kernel_2d(sigma=3)
    assert k.shape == (25, 25) and np.allclose(k, orc.kernel(3))
    assert pp.gaussian_kernel_2d(sigma=2, n=(4, 6), normalized=False).shape == (5, 7)
    assert pp.gaussian_kernel_2d(sigma=2, orientation="iline").shape == (5, 17)
    a = np.array([2.0, 4.0, 6.0])
    assert np.allclose(pp.rescale(a), [0, .5, 1]) and np.allclose(pp.rescale(a, 1e-3, 1)[0], 1e-3)
    assert pp.rescale(np.ones(3)) is not None and np.array_equal(pp.rescale(np.ones(3)), np.ones(3))
    with pytest.raises(ValueError):
        pp.antialiasing_filter((32, 32), "iline", {"il": 2, "xl": 1})


@pytest.mark.gpu
@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_gpu_matches_reference_golden(c, gold):
    from pseudo_3d_interpolation_b200 import cube_postprocessing_3D as pp
    d = make_slice(c)
    fn = pp.remove_acquisition_footprint if c["fn"] == "footprint" else pp.spatial_antialiasing
    y, f = fn(d, return_filter=True, verbose=0, **c["kw"])
    assert y.shape == d.shape and not np.iscomplexobj(y)
    np.testing.assert_allclose(f, gold[c["name"] + "__filter"], rtol=0, atol=1e-13)
    e = rel_l2(y, gold[c["name"] + "__y"])
    assert e <= RTOL32, e


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(256, 256), (200, 200), (1000, 64), (90, 2000), (1201, 847)])
def test_gpu_stack_against_oracle(shape):
    """register-resident, mixed and generic (Bluestein) plans; a stack of slices sharing one filter."""
    from pseudo_3d_interpolation_b200 import cube_postprocessing_3D as pp
    rng = np.random.default_rng(11)
    n = 3
    d = (rng.standard_normal((n,) + shape) + 1j * rng.standard_normal((n,) + shape)).astype(np.complex64)
    y, f = pp.remove_acquisition_footprint(d, sigma=5, direction="both", return_filter=True)
    ref = orc.apply(d.astype(np.complex128), orc.footprint_filter(shape, sigma=5))
    assert np.allclose(f, orc.footprint_filter(shape, sigma=5), atol=1e-13)
    assert y.dtype == np.float32 and y.shape == d.shape
    assert rel_l2(y, ref) <= RTOL32, rel_l2(y, ref)
    # identity filter returns the real part of the input
    from pseudo_3d_interpolation_b200 import get_plan
    z = get_plan(*shape).kxky_filter(d, np.ones(shape, np.float32))
    assert rel_l2(z, d.astype(np.complex128)) <= RTOL32
