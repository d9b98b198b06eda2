"""Generates tests/golden/f3_three_d_iou.npz from the LIVE reference (wavedata evaluation.three_d_iou
with this image's Pillow, imported through oracle/ref_shim.py) — TEST INFRASTRUCTURE.
    python -m oracle.make_f3_golden
Seeded car-sized boxes [ry, l, h, w, tx, ty, tz] and six neighbours each (shifted, rotated,
rescaled: overlaps from none to almost complete), plus degenerate pairs (identical, disjoint,
touching, vertically disjoint). The reference rasterises the bases at 0.01 m (evaluation.py:164-261);
the exact polygon intersection differs from it by at most ~0.01 in IoU on these cases: the tests
state 0.02."""
import numpy as np

from . import ref_shim


def cases(seed=0, n=250):
    rng = np.random.default_rng(seed)
    boxes, others = [], []
    for _ in range(n):
        l, w, h = rng.uniform(3.0, 5.0), rng.uniform(1.4, 2.0), rng.uniform(1.3, 1.9)
        box = np.array([rng.uniform(-np.pi, np.pi), l, h, w, rng.uniform(-20, 20), rng.uniform(1.4, 1.9),
                        rng.uniform(5, 60)])
        o = []
        for _k in range(6):
            d = rng.normal(0, [1.2, 0.1, 1.5])
            o.append([box[0] + rng.normal(0, 0.5), l * rng.uniform(0.8, 1.2), h * rng.uniform(0.8, 1.2),
                      w * rng.uniform(0.8, 1.2), box[4] + d[0], box[5] + d[1], box[6] + d[2]])
        boxes.append(box)
        others.append(o)
    b = np.array([0.3, 4.0, 1.5, 1.6, 1.0, 1.65, 20.0])
    boxes.append(b)
    others.append([b, b + [0, 0, 0, 0, 30, 0, 0], b + [0, 0, 0, 0, 0, 5.0, 0], b + [np.pi / 2, 0, 0, 0, 0, 0, 0],
                   b + [0, 0, 0, 0, 0.5, 0, 0.5], b * [1, 0.5, 1, 0.5, 1, 1, 1]])
    return np.array(boxes), np.array(others)


def main():
    assert ref_shim.install()
    from wavedata.tools.obj_detection.evaluation import three_d_iou
    boxes, others = cases()
    iou = np.array([three_d_iou(b, o) for b, o in zip(boxes, others)])
    np.savez_compressed("tests/golden/f3_three_d_iou.npz", boxes=boxes, others=others, iou=iou)
    print("wrote", iou.shape, "IoUs; non-zero:", int((iou > 0).sum()), "max", float(iou.max()))


if __name__ == "__main__":
    main()
