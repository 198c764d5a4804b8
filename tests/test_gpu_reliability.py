"""Reliability of the fused tcgen05 MLP kernels' mbarrier protocol (the round-1 watchdog trap)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu


def test_late_weight_chunks_change_nothing():
    """Producer fault injection (test twin of the kernels, -DNR_FAULT_INJECT): tile 1's weight chunks of every
    single-M-tile step arrive 20 us late -- the timing under which round 1's idle MMA issuer met its next w_full barrier
    a phase early, fell through the parity wait and desynchronised the ring.  All outputs must be bit-identical to the
    un-delayed launch (and the launch must return)."""
    import soak_mlp
    soak_mlp.inject(1 << 17)


def test_soak_back_to_back_launches():
    """600 back-to-back launches (sdf + normals with feature image, radiance pass, sdf only) at 1 M points, every output
    compared bit for bit with the first launch: a mis-fed MMA or a lost barrier phase shows as a mismatch or a trap."""
    import soak_mlp
    soak_mlp.soak(600, 1 << 20)


def test_soak_split_precision_kernel():
    """the same soak for the split-precision kernel (csrc/mlp_rev_split.cu: sdf + normals with the feature image, the radiance
    pass it feeds, sdf only); its late-chunk fault injection runs in test_late_weight_chunks_change_nothing"""
    import soak_mlp
    soak_mlp.soak(300, 1 << 20, soak_mlp.SPLIT_PIPE)
