"""BASELINE.json configs[4] (dense 128-beam LiDAR: 500k points, 0.05 m BEV 1400x1600): device time
of S1 (binning atomics + resolve) and S2 (integral image + 89 600-anchor filter) against their
algorithmic bytes (S1 16 N + 4 (S+1) H W = 61.76 MB, S2 17.0 MB), next to config A.
CUDA-graph replays over three rotating buffer sets, CUDA events.
usage: python tools/time_config_e.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dodt_b200 import ops, synth  # noqa: E402

PEAK = 6554.2
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
dev = torch.device("cuda", 0)
anchors = torch.from_numpy(synth.car_anchors()).to(dev)
nA = anchors.shape[0]
for name, n_pts, voxel in (("A (120k pts, 0.1 m, 700x800)", 120000, synth.VOXEL_SIZE),
                           ("E (500k pts, 0.05 m, 1400x1600)", 500000, synth.VOXEL_SIZE_DENSE)):
    nx, _, nz, min_x, _, min_z = ops.bev_grid(synth.AREA_EXTENTS, voxel)
    params = ops.make_bev_params(synth.GROUND_PLANE, synth.AREA_EXTENTS, voxel, synth.HEIGHT_LO,
                                 synth.HEIGHT_HI, synth.NUM_SLICES, True, 0.2, 2.0)
    sets = []
    for k in range(3):
        pc = torch.from_numpy(synth.point_cloud(5, k, n_points=n_pts)).to(dev)
        sets.append(dict(
            pts=pc.contiguous(), maps=torch.empty(synth.NUM_SLICES + 1, nz, nx, device=dev),
            occ=torch.empty(nx, nz, dtype=torch.uint8, device=dev),
            stats=torch.empty(24, dtype=torch.int32, device=dev),
            ws=torch.empty(max(ops.bev_workspace_bytes(n_pts, synth.NUM_SLICES, nx, nz), 256), dtype=torch.uint8, device=dev),
            ii=torch.empty(nx + 1, nz + 1, dtype=torch.int32, device=dev),
            ws_ii=torch.empty(max(ops.integral_workspace_bytes(nx, nz), 256), dtype=torch.uint8, device=dev),
            keep=torch.empty(nA, dtype=torch.uint8, device=dev)))

    def s1(x):
        ops.bev_slices(x["pts"], params, x["maps"], x["occ"], x["stats"], x["ws"])

    def s2(x):
        ops.integral_image_2d(x["occ"], x["ii"], x["ws_ii"])
        ops.anchor_filter_2d(anchors, x["ii"], nx, nz, min_x, min_z, voxel, 1, keep=x["keep"])

    for stage, fn, nbytes in (("S1", s1, 16 * n_pts + 4 * (synth.NUM_SLICES + 1) * nx * nz),
                              ("S2", s2, nx * nz + 4 * (nx + 1) * (nz + 1) + 65 * nA)):
        graphs = []
        for x in sets:
            fn(x)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn(x)
            graphs.append(g)
        for g in graphs:
            g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 60
        a.record()
        for i in range(reps):
            graphs[i % 3].replay()
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / reps
        print("config %-32s %s  %7.1f us  %6.2f MB algorithmic  %5.0f GB/s  %.3f of %.0f GB/s"
              % (name, stage, us, nbytes / 1e6, nbytes / us / 1e3, nbytes / us / 1e3 / PEAK, PEAK))
    kept = int(sets[0]["keep"].sum())
    print("config %-32s anchors kept %d of %d, occupied cells %d" % (name, kept, nA, int(sets[0]["occ"].sum())))
