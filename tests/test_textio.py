"""Host text I/O: byte-identical to np.savetxt, round-trips np.loadtxt.
CPU only -- these routines need no GPU."""
import numpy as np

from wisp_mdanalysis_implementation_b200 import _lib


def test_savetxt_matches_numpy_bytes(tmp_path):
    rng = np.random.default_rng(0)
    a = rng.normal(size=(137, 53)) * 10.0 ** rng.integers(-12, 12, size=(137, 53))
    a[0, 0] = 0.0
    a[1, 1] = -0.0
    a[2, 2] = np.inf
    a[3, 3] = -np.inf
    a[4, 4] = np.nan
    np.savetxt(tmp_path / "np.dat", a)
    _lib.savetxt(str(tmp_path / "c.dat"), a)
    assert (tmp_path / "np.dat").read_bytes() == (tmp_path / "c.dat").read_bytes()
    b = (rng.random((60, 60)) < 0.3).astype(np.int64)
    np.savetxt(tmp_path / "npi.dat", b, fmt='%i')
    _lib.savetxt(str(tmp_path / "ci.dat"), b, fmt='%i')
    assert (tmp_path / "npi.dat").read_bytes() == (tmp_path / "ci.dat").read_bytes()
    v = rng.normal(size=17)
    np.savetxt(tmp_path / "npv.dat", v, header='Shape: nNodes x 3 (%d x 3)' % 17)
    _lib.savetxt(str(tmp_path / "cv.dat"), v, header='Shape: nNodes x 3 (%d x 3)' % 17)
    assert (tmp_path / "npv.dat").read_bytes() == (tmp_path / "cv.dat").read_bytes()


def test_savetxt_extreme_values_match_numpy_bytes(tmp_path):
    """Every decade of the double range, denormals, the largest double, halfway decimal cases."""
    rng = np.random.default_rng(7)
    wide = rng.normal(size=6000) * 10.0 ** rng.integers(-300, 300, size=6000)
    special = np.array([0.0, -0.0, 5e-324, -5e-324, 4.35e-320, 2.2250738585072014e-308, 2.225073858507201e-308,
                        1.7976931348623157e308, -1.7976931348623157e308, 1.0, -1.0, 0.1, 1 / 3, 0.5, 2.5, 1e22,
                        9.999999999999999e22, 1e23, 123456789.123456789, 1.0000000000000002, 0.9999999999999999])
    a = np.concatenate([wide, special])
    a = np.concatenate([a, np.zeros((-len(a)) % 3)]).reshape(-1, 3)
    np.savetxt(tmp_path / "np.dat", a)
    _lib.savetxt(str(tmp_path / "c.dat"), a)
    assert (tmp_path / "np.dat").read_bytes() == (tmp_path / "c.dat").read_bytes()
    assert np.array_equal(_lib.loadtxt(str(tmp_path / "c.dat")), a)


def test_loadtxt_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    a = rng.normal(size=(301, 77))
    np.savetxt(tmp_path / "a.dat", a, header="a header line")
    got = _lib.loadtxt(str(tmp_path / "a.dat"))
    assert np.array_equal(got, a)            # %.18e round-trips float64 exactly
    assert np.array_equal(got, np.loadtxt(tmp_path / "a.dat"))
    v = rng.normal(size=40)
    np.savetxt(tmp_path / "v.dat", v)
    assert np.array_equal(_lib.loadtxt(str(tmp_path / "v.dat")), v)


def test_large_matrix_speed(tmp_path):
    a = np.random.default_rng(2).normal(size=(1500, 1500))
    import time
    t0 = time.perf_counter()
    _lib.savetxt(str(tmp_path / "big.dat"), a)
    t_c = time.perf_counter() - t0
    t0 = time.perf_counter()
    b = _lib.loadtxt(str(tmp_path / "big.dat"))
    t_l = time.perf_counter() - t0
    assert np.array_equal(a, b)
    assert t_c < 10 and t_l < 10


def test_loadtxt_rejects_ragged_rows_and_squeezes_single_row(tmp_path):
    """a short row must not borrow numbers from the next line, extra values are an error,
    and a one-row file comes back 1-D like np.loadtxt."""
    import pytest
    from wisp_mdanalysis_implementation_b200 import _lib
    f = tmp_path / "ragged.dat"
    f.write_text("1.0 2.0 3.0\n4.0 5.0\n6.0 7.0 8.0 9.0\n")
    with pytest.raises(_lib.WispError):
        _lib.loadtxt(str(f))
    f.write_text("1.0 2.0 3.0\n4.0 5.0 6.0 7.0\n")
    with pytest.raises(_lib.WispError):
        _lib.loadtxt(str(f))
    f.write_text("1.0 2.0 x\n")
    with pytest.raises(_lib.WispError):
        _lib.loadtxt(str(f))
    f.write_text("# header\n1.5e+00 -2.5e-01 inf nan +3.0\n")
    got = _lib.loadtxt(str(f))
    want = np.loadtxt(str(f))
    assert got.shape == want.shape == (5,)
    assert np.array_equal(got, want, equal_nan=True)
    f.write_text("1 2\n3 4\n\n")
    assert np.array_equal(_lib.loadtxt(str(f)), np.array([[1.0, 2.0], [3.0, 4.0]]))
