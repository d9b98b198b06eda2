"""Randomised S1 configurations with clouds built to sit on every rounding edge (shared by the GPU
parity test and the oracle-vs-live-reference test)."""
import numpy as np


def adversarial_case(seed):
    """-> (point_cloud (3, n) f32 or f64, voxel_size, area_extents, height_lo, height_hi, num_slices).
    Points are snapped to voxel multiples, to slice boundaries +- one ulp, to the open extents;
    there are exact duplicates and points outside the extents."""
    rng = np.random.default_rng(900 + seed)
    voxel = float(rng.choice([np.float32(0.1), np.float32(0.05), np.float32(0.2), 0.25, np.float32(0.16)]))
    S = int(rng.integers(1, 7))
    lo = float(np.float32(rng.choice([-0.2, 0.0, -0.5, 0.3])))
    hi = float(np.float32(lo + rng.choice([1.0, 2.5, 1.7])))
    ext = [[-float(rng.integers(8, 21)), float(rng.integers(8, 21))], [-5, 3], [0, float(rng.integers(10, 31))]]
    n = int(rng.integers(2000, 30000))
    x = rng.uniform(ext[0][0] - 1, ext[0][1] + 1, n)
    z = rng.uniform(ext[2][0] - 1, ext[2][1] + 1, n)
    h = rng.uniform(lo - 0.3, hi + 0.3, n)
    hpd = (hi - lo) / S
    k = n // 5
    x[:k] = np.round(x[:k] / voxel) * voxel                      # on voxel edges
    z[k:2 * k] = np.round(z[k:2 * k] / voxel) * voxel
    edges = lo + hpd * rng.integers(0, S + 1, k)
    h[2 * k:3 * k] = edges                                       # on slice boundaries
    y = 1.65 - h
    y[2 * k:2 * k + k // 3] = np.nextafter(y[2 * k:2 * k + k // 3], 10.0)
    y[2 * k + k // 3:2 * k + 2 * (k // 3)] = np.nextafter(y[2 * k + k // 3:2 * k + 2 * (k // 3)], -10.0)
    x[3 * k:3 * k + 20] = rng.choice([ext[0][0], ext[0][1]], 20)  # on the open extents
    z[3 * k + 20:3 * k + 40] = rng.choice([ext[2][0], ext[2][1]], 20)
    dup = rng.integers(0, n, k)                                   # exact duplicates (tie -> lower index)
    x[4 * k:5 * k], y[4 * k:5 * k], z[4 * k:5 * k] = x[dup], y[dup], z[dup]
    pc = np.stack([x, y, z])
    if seed % 2:
        pc = pc.astype(np.float32)
    return pc, voxel, ext, lo, hi, S
