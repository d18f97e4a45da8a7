"""Seeded sweep of odd convolution geometries on the tcgen05 implicit-GEMM kernels (property-style coverage of
sizes and tails, beyond the hand-picked cases of selftest.CONV_CASES): ragged M tails (N*P*Q not a multiple of 128), channel
counts that leave partial 64-wide K blocks and partial N tiles, non-square maps, stride 2, dilation, 5x5 -- each against
F.conv2d / conv_transpose2d / autograd in fp32.  Runs last (file name) so a miss here cannot hide the rest of the suite."""
import random

import pytest

pytestmark = pytest.mark.gpu


def _case(seed):
    rnd = random.Random(1000 + seed)
    R = rnd.choice([1, 1, 3, 3, 3, 5])
    stride = rnd.choice([1, 1, 2])
    dil = 2 if (R == 3 and rnd.random() < 0.25) else 1
    pad = rnd.choice([0, dil * (R // 2)])
    span = dil * (R - 1) + 1
    H = rnd.randint(max(span, 5), 33)
    W = rnd.randint(max(span, 5), 33)
    return dict(N=rnd.randint(1, 5), H=H, W=W, C=8 * rnd.randint(1, 41), K=8 * rnd.randint(1, 41), R=R, stride=stride, pad=pad,
                dil=dil)


@pytest.mark.parametrize("seed", range(12))
def test_conv_fprop_wgrad_random_geometry(seed):
    from distribuuuu_b200 import selftest
    kw = _case(seed)
    selftest.check_conv_fprop(**kw)
    selftest.check_conv_wgrad(**kw, tol=3e-2)


@pytest.mark.parametrize("seed", range(8))
def test_conv_dgrad_random_geometry(seed):
    from distribuuuu_b200 import selftest
    kw = _case(100 + seed)
    kw.pop("stride")
    kw["H"], kw["W"] = kw["H"] + 2, kw["W"] + 2      # stride-1 data gradient; keep the output map >= 1 after "valid" padding
    selftest.check_conv_dgrad(**kw)
