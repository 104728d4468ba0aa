"""Generates tests/golden/albatite_SD_points.npz (points float64 (35801, 3), values float64 (35801,)) from the
reference's shipped dataset `datasets/albatite_SD_points.csv` (columns X, Y, Z, SignedDistance) — the input of
BASELINE.json config C4.  Run in the build container, where /root/reference exists; the GPU box only sees the fixture.

    python tools/make_albatite_fixture.py [/root/reference/datasets/albatite_SD_points.csv]
"""
import sys

import numpy as np

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/datasets/albatite_SD_points.csv"
data = np.loadtxt(src, delimiter=",", skiprows=1, dtype=np.float64)
assert data.shape[1] == 4
np.savez_compressed("tests/golden/albatite_SD_points.npz", points=np.ascontiguousarray(data[:, :3]),
                    values=np.ascontiguousarray(data[:, 3]))
print(data.shape, data[:, :3].min(0), data[:, :3].max(0))
