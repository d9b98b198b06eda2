"""Configuration constants of the DODT car config shared by the frame runner, bench.py and tests."""
import numpy as np

# KITTI Car clusters of avod/configs/pyramid_cars_with_aug_dt_5_tracking.config (2 clusters)
CAR_ANCHOR_SIZES = [[3.514, 1.581, 1.511], [4.236, 1.653, 1.547]]
# P2 of avod/tests/datasets/Kitti/tracking/training/calib/0000.txt (public KITTI calibration)
KITTI_P2 = np.array([[721.5377, 0.0, 609.5593, 44.85728],
                     [0.0, 721.5377, 172.854, 0.2163791],
                     [0.0, 0.0, 1.0, 0.002745884]])
