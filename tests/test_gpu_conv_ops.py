"""GPU op-level known-answer tests of the two conv paths (fp32 FFMA tiles / tcgen05 split-fp16 tiles)
through the C-ABI test hook, against the oracle's conv restatement (itself pinned to the compiled
reference by tests/test_oracle.py) on seeded inputs, incl. ragged packed batches."""
import numpy as np
import pytest
from parity_util import rel_err

from oracle import vits_numpy as vn
from summertts_b200 import engine

pytestmark = pytest.mark.gpu


def _rec(rng, o, c, k, p, d, stride=None, wscale=0.2):
    w = (rng.standard_normal((o, k, c)) * wscale).astype(np.float32)
    b = rng.standard_normal(o).astype(np.float32)
    hdr = [o, c, k, p, d, 1] + ([stride] if stride else [])
    rec = np.concatenate([np.array(hdr, np.float32), w.ravel(), b]).astype(np.float32)
    cv = dict(outCh=o, inCh=c, k=k, pad=p, dil=d, hasBias=1, w=w, b=b, stride=stride or 1)
    return rec, cv


def _oracle_packed(x, cv, seg, fn):
    return np.concatenate([fn(x[seg[i]:seg[i + 1]], cv) for i in range(len(seg) - 1)], axis=0)


SHAPES = [  # (Cout, Cin, k, dil)
    (64, 64, 3, 1), (64, 64, 7, 3), (64, 64, 11, 5), (32, 32, 11, 5), (32, 32, 3, 1), (384, 192, 5, 1),
    (192, 96, 1, 1), (96, 192, 1, 1), (768, 192, 3, 1), (192, 768, 3, 1), (128, 192, 7, 1), (72, 32, 7, 1),
    (16, 16, 7, 3), (576, 192, 1, 1), (256, 192, 3, 1),
]


@pytest.mark.parametrize("use_tc", [0, 1])
@pytest.mark.parametrize("shape", SHAPES)
def test_conv1d_paths_vs_oracle(native_lib, shape, use_tc):
    o, c, k, d = shape
    rng = np.random.default_rng(hash(shape) % 2**31)
    rec, cv = _rec(rng, o, c, k, d * (k - 1) // 2, d)
    seg = np.array([0, 5, 140, 141, 400, 777], np.int32)  # ragged: 5, 135, 1, 259, 377 rows
    x = (rng.standard_normal((seg[-1], c)) * 3).astype(np.float32)
    y = engine.test_conv1d(rec, x, use_tc=use_tc, seg_off=seg)
    want = _oracle_packed(x, cv, seg, vn.conv1d)
    assert y.shape == want.shape
    assert rel_err(y, want) < (1e-5 if use_tc else 2e-6)


@pytest.mark.parametrize("use_tc", [0, 1])
def test_leaky_relu_gate_tanh_epilogues(native_lib, use_tc):
    rng = np.random.default_rng(77)
    seg = np.array([0, 300, 333], np.int32)
    x = (rng.standard_normal((333, 64)) * 2).astype(np.float32)
    rec, cv = _rec(rng, 128, 64, 5, 2, 1)
    y = engine.test_conv1d(rec, x, use_tc=use_tc, seg_off=seg, in_act=1, slope=0.1, epi=1)
    want = np.maximum(_oracle_packed(vn.leaky_relu(x, 0.1), cv, seg, vn.conv1d), 0)
    assert rel_err(y, want) < 1e-5
    # WN gate (WN.cpp:85-98): tanh(a[:, :H]) * sigmoid(a[:, H:])
    y = engine.test_conv1d(rec, x, use_tc=use_tc, seg_off=seg, epi=4)
    a = _oracle_packed(x, cv, seg, vn.conv1d)
    want = vn.tanh_ref(a[:, :64]) * vn.sigmoid(a[:, 64:])
    assert y.shape == want.shape and rel_err(y, want) < 2e-5
    y = engine.test_conv1d(rec, x, use_tc=use_tc, seg_off=seg, epi=6)
    assert rel_err(y, vn.tanh_ref(a)) < 2e-5


@pytest.mark.parametrize("use_tc", [0, 1])
@pytest.mark.parametrize("cfg", [(64, 128, 16, 4), (32, 64, 16, 8), (16, 32, 4, 2)])
def test_conv_transposed_phase_expansion(native_lib, cfg, use_tc):
    """ConvTranspose1d as a stride*Cout-wide dense conv == nn_conv1d_transposed::forward."""
    o, c, k, s = cfg
    rng = np.random.default_rng(5)
    rec, cv = _rec(rng, o, c, k, 0, 1, stride=1)
    cv.update(stride=s, pad=(k - s) // 2)
    seg = np.array([0, 70, 71, 200], np.int32)
    x = rng.standard_normal((200, c)).astype(np.float32)
    y = engine.test_conv1d(rec, x, use_tc=use_tc, transposed=True, stride=s, pad=(k - s) // 2, seg_off=seg)
    want = _oracle_packed(x, cv, seg, vn.conv1d_transposed)
    assert y.shape == want.shape and rel_err(y, want) < 1e-5


def test_large_magnitudes_saturate_not_nan(native_lib):
    """fp16 split conversion saturates (cvt.rn.satfinite) instead of producing inf/NaN."""
    rng = np.random.default_rng(1)
    rec, cv = _rec(rng, 32, 32, 3, 1, 1)
    x = (rng.standard_normal((256, 32)) * 2000).astype(np.float32)  # |x|*8 stays below 65504 mostly
    y = engine.test_conv1d(rec, x, use_tc=1)
    assert np.isfinite(y).all()
    assert rel_err(y, vn.conv1d(x, cv)) < 1e-4
