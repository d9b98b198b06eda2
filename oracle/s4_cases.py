"""The S4 parity cases shared by oracle/make_s4_golden.py and tests/test_s4_reference_kernel.py —
TEST INFRASTRUCTURE. Inputs are seeded; only the reference kernels' outputs are frozen.

Two latent races of the reference, for the record (neither bites at these cases):
  * PadData (pad.cu.cc:32-35): the threads past the last pixel of the last 16-thread block write 0.0
    to row `in_height` of the padded map — a position other blocks fill with data whenever
    in_width*in_height is not a multiple of 16 and the column is >= padding. Every case below has
    H*W % 16 == 0 or is flagged by pad_race_possible(); the generator re-runs each case and records
    whether the two runs were bit-identical.
  * CorrelateData (correlation_kernel.cu.cc:73,99-104): thread 0 reads sum[0..31] after the block
    barrier while the other 31 threads of the (single) warp may already zero sum[] for the next
    output channel. nvcc reconverges the warp after the `if (ch_off == 0)` block, so in practice
    the read is ordered; the re-run check above covers it.
"""
import numpy as np

DODT = dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=5)

SMALL_CASES = [
    ((1, 20, 24, 8), dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=5)),   # DODT
    ((2, 11, 13, 16), dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=5)),
    ((1, 12, 12, 8), dict(kernel_size=1, max_displacement=2, stride_1=1, stride_2=2, padding=2)),   # r = 1
    ((1, 14, 12, 8), dict(kernel_size=1, max_displacement=4, stride_1=1, stride_2=2, padding=6)),   # shift < 0
    ((1, 14, 15, 8), dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=4)),   # shift > 0
    ((1, 12, 10, 5), dict(kernel_size=1, max_displacement=4, stride_1=1, stride_2=1, padding=4)),
    ((1, 15, 13, 3), dict(kernel_size=3, max_displacement=4, stride_1=2, stride_2=2, padding=4)),
    ((2, 9, 10, 4), dict(kernel_size=3, max_displacement=3, stride_1=1, stride_2=1, padding=5)),
    ((1, 8, 8, 2), dict(kernel_size=1, max_displacement=20, stride_1=1, stride_2=2, padding=20)),   # defaults
    ((1, 40, 72, 32), dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=5)),
    ((2, 19, 150, 32), dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=5)),
    ((1, 33, 70, 24), dict(kernel_size=1, max_displacement=3, stride_1=1, stride_2=2, padding=3)),
    ((1, 17, 66, 16), dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=3)),
    ((1, 17, 66, 16), dict(kernel_size=1, max_displacement=4, stride_1=1, stride_2=2, padding=7)),
    ((1, 16, 48, 64), dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=5)),  # C > 32: 2 terms/thread
    ((1, 24, 40, 40), dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=5)),  # ragged channel loop
]


def pad_race_possible(shape):
    return (shape[1] * shape[2]) % 16 != 0


def out_shape(H, W, kernel_size, max_displacement, stride_1, stride_2, padding):
    """correlation_kernel.cc:39-57"""
    kr = (kernel_size - 1) // 2
    border = max_displacement + kr
    oh = int(np.ceil(np.float32(H + 2 * padding - 2 * border) / np.float32(stride_1)))
    ow = int(np.ceil(np.float32(W + 2 * padding - 2 * border) / np.float32(stride_1)))
    r = max_displacement // stride_2
    return oh, ow, (2 * r + 1) ** 2


def inputs(shape, kw, seed=0):
    rng = np.random.default_rng(seed + shape[1] * 31 + shape[3])
    a = rng.standard_normal(shape).astype(np.float32)
    b = rng.standard_normal(shape).astype(np.float32)
    oh, ow, oc = out_shape(shape[1], shape[2], kw["kernel_size"], kw["max_displacement"], kw["stride_1"],
                           kw["stride_2"], kw["padding"])
    g = rng.standard_normal((shape[0], oh, ow, oc)).astype(np.float32)
    return a, b, g


def full_inputs():
    """Config C: the bench's synthetic BEV feature pair [1,700,800,32] and a seeded gradient."""
    from oracle import synth_ref as synth
    f0, f1 = synth.feature_pair(3, 0)
    g = np.random.default_rng(11).standard_normal((1, 700, 800, 25)).astype(np.float32)
    return f0, f1, g


def full_sample_pixels(n=4096):
    """Seeded pixel subset of the 700x800 map, with the four corners and border rows forced in."""
    rng = np.random.default_rng(2024)
    px = rng.choice(700 * 800, size=n, replace=False)
    forced = [0, 799, 699 * 800, 699 * 800 + 799, 4 * 800 + 4, 5 * 800 + 5, 694 * 800 + 794, 350 * 800 + 1]
    px[:len(forced)] = forced
    return np.sort(np.unique(px))
