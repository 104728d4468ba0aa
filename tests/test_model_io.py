"""save_model / load_model envelope (rbf.rs:1087-1171, 1469-1488) and the global-trend transform (global_trend.rs:128-287)
on a CPU-only machine: fr_model_restore / fr_get_state keep model state on the host, no kernel is launched."""
import json

import numpy as np
import pytest

import ferreus_rbf_rs_b200 as fb
from oracle import rbf as orbf


def _envelope(n=40, dim=3, cols=2, trend=None, seed=3):
    rng = np.random.default_rng(seed)
    pts = rng.random((n, dim))
    mat = fb.RBFInterpolator._mat
    basis = 1 + dim  # cubic kernel, linear drift
    gt = None
    if trend is not None:
        t = orbf.GlobalTrendTransform(pts.mean(axis=0), trend)
        gt = {"affine_transform": mat(t.affine), "inverse_transform": mat(t.inverse)}
    return {
        "format": "ferreus_rbf.json", "version": 1,
        "points": mat(pts), "point_values": mat(rng.random((n, cols))),
        "coefficients": {"point_coefficients": mat(rng.standard_normal((n, cols))),
                         "poly_coefficients": mat(rng.standard_normal((basis, cols)))},
        "interpolant_settings": {"kernel_type": "Cubic", "spheroidal_order": "Three", "drift": "Linear", "nugget": 0.0,
                                 "base_range": 1.0, "total_sill": 1.0, "basis_size": basis, "polynomial_degree": 1,
                                 "fitting_accuracy": {"tolerance": 1e-6, "tolerance_type": "Relative"}},
        "translation_factor": [0.5] * dim, "scale_factor": [0.5] * dim,
        "params": {"solver_type": "FGMRES",
                   "ddm_params": {"leaf_threshold": 1024, "overlap_quota": 0.5, "coarse_ratio": 0.125,
                                  "coarse_threshold": 4096},
                   "fmm_params": {"interpolation_order": 11, "max_points_per_cell": 256, "compression_type": "ACA",
                                  "epsilon": 1e-11, "eval_chunk_size": 1024},
                   "naive_solve_threshold": 4096, "test_unique": True},
        "global_trend": gt,
    }


@pytest.mark.parametrize("with_trend", [False, True])
def test_envelope_round_trip_is_lossless(tmp_path, with_trend):
    trend = orbf.GlobalTrend.three(25.0, 70.0, 15.0, 3.0, 2.0, 1.0) if with_trend else None
    doc = _envelope(trend=trend)
    p1, p2 = str(tmp_path / "a.json"), str(tmp_path / "b.json")
    json.dump(doc, open(p1, "w"))
    model = fb.RBFInterpolator.load_model(p1)
    un = fb.RBFInterpolator._unmat
    assert np.array_equal(model.source_points, un(doc["points"]))
    assert np.array_equal(model.source_values, un(doc["point_values"]))
    co = model.coefficients
    assert np.array_equal(co.point_coefficients, un(doc["coefficients"]["point_coefficients"]))
    assert np.array_equal(co.poly_coefficients, un(doc["coefficients"]["poly_coefficients"]))
    assert model.params.fmm_params.interpolation_order == 11 and int(model.interpolant_settings.kernel_type) == 2
    model.save_model(p2)
    assert json.load(open(p2)) == doc          # field names, enum spellings, matrices, trend: unchanged


def test_envelope_validation():
    import tempfile
    import os
    d = tempfile.mkdtemp()
    for key, val in (("format", "something.else"), ("version", 7)):
        doc = _envelope()
        doc[key] = val
        path = os.path.join(d, f"{key}.json")
        json.dump(doc, open(path, "w"))
        with pytest.raises(ValueError):
            fb.RBFInterpolator.load_model(path)
    doc = _envelope()
    doc["interpolant_settings"]["basis_size"] = 10   # inconsistent with cubic + linear drift in 3-D
    path = os.path.join(d, "basis.json")
    json.dump(doc, open(path, "w"))
    with pytest.raises(ValueError):
        fb.RBFInterpolator.load_model(path)


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_trend_matrices_match_the_oracle_restatement(tmp_path, dim):
    """fr_fit_trend's affine / inverse matrices (host code) against the numpy restatement: fit a tiny model through
    the naive dense path?  No kernel may run on this machine, so the matrices are compared through a restored model:
    the oracle's matrices are written into an envelope, loaded, saved again and must survive bit for bit; the C++
    construction itself is compared with the oracle in tests/test_gpu_rbf.py::test_global_trend_matches_oracle."""
    trend = {1: orbf.GlobalTrend.one(2.0), 2: orbf.GlobalTrend.two(35.0, 2.5, 1.0),
             3: orbf.GlobalTrend.three(25.0, 70.0, 15.0, 3.0, 2.0, 1.0)}[dim]
    doc = _envelope(dim=dim, trend=trend)
    t = orbf.GlobalTrendTransform(fb.RBFInterpolator._unmat(doc["points"]).mean(axis=0), trend)
    assert np.allclose(t.affine @ t.inverse, np.eye(dim + 1), atol=1e-12)
    pts = fb.RBFInterpolator._unmat(doc["points"])
    assert np.allclose(t.inverse_transform_points(t.transform_points(pts)), pts, atol=1e-12)
    path = str(tmp_path / "t.json")
    json.dump(doc, open(path, "w"))
    model = fb.RBFInterpolator.load_model(path)
    out = str(tmp_path / "t2.json")
    model.save_model(out)
    assert json.load(open(out))["global_trend"] == doc["global_trend"]
