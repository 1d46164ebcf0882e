"""GPU op-level tests of the fused ResBlock1-pair kernel (summertts_b200/csrc/rb_fused.cuh) through the C-ABI test hook,
against the oracle's conv restatement (pinned to the compiled reference by tests/test_oracle.py):

    y = act(x + conv2(leaky_0.1(conv1_dil(leaky_0.1(x)))))          ResBlock1::forward, src/modules/ResBlock1.cpp:55-69

Ragged packed batches put utterance ends on both sides of the 256-(k-1)-row super-tile boundaries; the long case gives
every persistent CTA several tiles (buffer / barrier phase wrap-around).
Tolerances: mode 0 (split-fp16, fp32-accurate) max|a-b|/max|b| < 2e-5; mode 1 (single fp16 MMA) < 5e-3."""
import numpy as np
import pytest
from parity_util import rel_err

from oracle import vits_numpy as vn
from summertts_b200 import engine

pytestmark = pytest.mark.gpu


def _rec(rng, c, k, wscale):
    w = (rng.standard_normal((c, k, c)) * wscale).astype(np.float32)
    b = (rng.standard_normal(c) * 0.1).astype(np.float32)
    rec = np.concatenate([np.array([c, c, k, (k - 1) // 2, 1, 1], np.float32), w.ravel(), b]).astype(np.float32)
    return rec, dict(outCh=c, inCh=c, k=k, pad=(k - 1) // 2, dil=1, hasBias=1, w=w, b=b, stride=1)


def _oracle(x, cv1, cv2, d, seg, out_leaky):
    outs = []
    for i in range(len(seg) - 1):
        xs = x[seg[i]:seg[i + 1]]
        t = vn.conv1d(vn.leaky_relu(xs, 0.1), cv1, pad=d * (cv1["k"] - 1) // 2, dil=d)
        y = xs + vn.conv1d(vn.leaky_relu(t, 0.1), cv2)
        outs.append(vn.leaky_relu(y, 0.1) if out_leaky else y)
    return np.concatenate(outs, axis=0)


CASES = [(32, 3, 1), (32, 7, 3), (32, 11, 5), (64, 3, 1), (64, 7, 3), (64, 11, 5), (64, 5, 1), (32, 3, 3), (16, 3, 1), (16, 7, 3), (16, 11, 5)]


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("case", CASES)
def test_fused_pair_vs_oracle(native_lib, case, mode):
    c, k, d = case
    rng = np.random.default_rng(1000 * c + 10 * k + d)
    rec1, cv1 = _rec(rng, c, k, 0.6 / np.sqrt(c * k))
    rec2, cv2 = _rec(rng, c, k, 0.6 / np.sqrt(c * k))
    ov = 256 - (k - 1)
    lens = [1, 5, ov - 1, ov, ov + 1, 2 * ov, 2 * ov + 3, 37, 600]
    seg = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    x = (rng.standard_normal((seg[-1], c)) * 2).astype(np.float32)
    for out_leaky in (True, False):
        y, flags = engine.test_rbpair(rec1, rec2, x, dil1=d, mode=mode, seg_off=seg, out_leaky=out_leaky)
        want = _oracle(x, cv1, cv2, d, seg, out_leaky)
        assert flags == 0
        assert np.isfinite(y).all()
        e = rel_err(y, want)
        assert e < (2e-5 if mode == 0 else 5e-3), e
        # every utterance separately too (an error confined to a short utterance must not hide behind max|b|)
        for i in range(len(lens)):
            a, b = y[seg[i]:seg[i + 1]], want[seg[i]:seg[i + 1]]
            assert np.abs(a - b).max() < (2e-5 if mode == 0 else 5e-3) * max(np.abs(want).max(), 1.0), (i, lens[i])


@pytest.mark.parametrize("case", [(32, 11, 5), (64, 7, 3)])
def test_fused_pair_many_tiles_per_cta(native_lib, case):
    """~420 super-tiles on 148 persistent CTAs: 2-3 tiles per CTA (x-tile double buffer, ring / barrier parity wrap)."""
    c, k, d = case
    rng = np.random.default_rng(7)
    rec1, cv1 = _rec(rng, c, k, 0.6 / np.sqrt(c * k))
    rec2, cv2 = _rec(rng, c, k, 0.6 / np.sqrt(c * k))
    lens = [20000, 31, 30000, 25000, 9, 26000]
    seg = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    x = (rng.standard_normal((seg[-1], c)) * 2).astype(np.float32)
    y, flags = engine.test_rbpair(rec1, rec2, x, dil1=d, mode=0, seg_off=seg, out_leaky=True)
    want = _oracle(x, cv1, cv2, d, seg, True)
    assert flags == 0 and rel_err(y, want) < 2e-5
    y2, _ = engine.test_rbpair(rec1, rec2, x, dil1=d, mode=0, seg_off=seg, out_leaky=True)
    assert np.array_equal(y, y2)      # deterministic


def test_fused_pair_overflow_flag(native_lib):
    """|x| beyond the split-fp16 range (8 |x| > 65504) raises the overflow flag instead of silently saturating."""
    rng = np.random.default_rng(3)
    rec1, _ = _rec(rng, 32, 3, 0.1)
    rec2, _ = _rec(rng, 32, 3, 0.1)
    x = (rng.standard_normal((300, 32)) * 2).astype(np.float32)
    x[100, 5] = 3.0e4
    _, flags = engine.test_rbpair(rec1, rec2, x, dil1=1, mode=0)
    assert flags & 1
