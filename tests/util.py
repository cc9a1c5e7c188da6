"""Shared helpers of the parity tests."""
import numpy as np

# north_star tolerance for the streamfunction: 1e-9 relative OR 1e-6 Sv absolute (fp64 summation order only)
RTOL_PSI = 1e-9
ATOL_PSI_SV = 1e-6


def assert_psi_close(got, ref, what=""):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    nan_g, nan_r = np.isnan(got), np.isnan(ref)
    assert np.array_equal(nan_g, nan_r), f"{what}: NaN pattern differs ({nan_g.sum()} vs {nan_r.sum()})"
    inf_r = np.isinf(ref)
    assert np.array_equal(got[inf_r], ref[inf_r]), f"{what}: Inf pattern differs"
    ok = ~(nan_r | inf_r)
    err = np.abs(got[ok] - ref[ok])
    tol = np.maximum(RTOL_PSI * np.abs(ref[ok]), ATOL_PSI_SV)
    bad = err > tol
    assert not bad.any(), f"{what}: {bad.sum()} values out of tolerance, max abs err {err.max():.3e} Sv"
    return float(err.max()) if err.size else 0.0


def case_inputs(oracle, mesh, synth, with_basins=True, zero_edges=True):
    ib = oracle.basin_masks(*synth.basin_mask_inputs(mesh, with_basins), zero_edges=zero_edges)
    e3m = oracle.mask_e3v(mesh.e3v_0, mesh.vmask.astype(np.float32))
    return ib, e3m
