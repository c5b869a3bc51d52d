"""Tolerances of the parity suite (stated once, used everywhere).

Bit-exact classes (no tolerance): concatenate, interweave, difference (copies / one
subtraction rounded once), hard argmin/argmax on a given volume, and every reduction op on
dyadic inputs (k/8), whose fp32 sums are exact and therefore order independent.

Tolerance classes (random N(0,1) inputs), following SURVEY.md section 8c:
"""
import numpy as np

# inner product / groupwise, fp32 in/out:  |delta| <= 2e-5 * sqrt(C) * max|L| * max|R|
def corr_atol_fp32(c, lmax, rmax):
    return 2e-5 * np.sqrt(c) * lmax * rmax

# 16-bit inputs: compare with the fp32 oracle evaluated on the same rounded inputs;
# relative tolerance of the output cast plus the reference's own product rounding.
RTOL_16 = {"fp16": 2.0 ** -8, "bf16": 2.0 ** -6}

# soft-argmax disparity, fp32:  |delta| <= 1e-4 * D  (pixels)
def soft_argmax_atol(d):
    return 1e-4 * max(d, 1)

# v4 tail (trilinear x softmax x expectation), fp32: <= 1e-3 px
V4_TAIL_ATOL = 1e-3

# gradients: relative to the gradient's own scale
GRAD_RTOL = 2e-4
