"""Shared helpers of the GPU tests."""
import torch


def same_occupancy(a, b, tol=1e-5):
    """Do two dense grids hold voxels in the same cells? `a != 0` is not a stable test of that: a feature of a nearly dead
    channel can be EXACTLY 0.0 in one evaluation (max == mean == 0) and 1e-7 in another one whose BatchNorm sums were added in a
    different order (tools/split_stress.py found one such cell). Cells where exactly one side is zero must therefore hold a value
    below `tol` x max|b| on the other side; everything else is left to the value comparison next to this call."""
    a, b = a.detach(), b.detach().to(a.device)
    m = (a != 0) != (b != 0)
    if not bool(m.any()):
        return True
    scale = max(float(b.abs().max()), 1e-30)
    return float(a[m].abs().max()) <= tol * scale and float(b[m].abs().max()) <= tol * scale
