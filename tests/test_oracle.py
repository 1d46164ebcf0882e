"""CPU tests that PIN the oracle: the numpy restatement (oracle/vits_numpy.py) against
(a) the committed golden vectors generated from the compiled, unmodified reference and
(b) the compiled reference itself (oracle/_ref) when it is present."""
import glob
import json
import os

import numpy as np
import pytest
from parity_util import GOLDEN, find_model, lsb_diff, rel_err

from oracle import ref, vits_numpy as vn
from summertts_b200 import binfmt

OPS = np.load(os.path.join(GOLDEN, "ops.npz"))


def _conv_rec(rec, transposed=False):
    n = 7 if transposed else 6
    h = [int(v) for v in rec[:n]]
    o, c, k, p, d, hb = h[:6]
    w = rec[n:n + o * k * c].reshape(o, k, c)
    b = rec[n + o * k * c:] if hb else None
    return dict(outCh=o, inCh=c, k=k, pad=p, dil=d, hasBias=hb, w=w, b=b, stride=h[6] if transposed else 1)


@pytest.mark.parametrize("tag", ["dense", "dil", "k1"])
def test_conv1d_kat(tag):
    y = vn.conv1d(OPS["x"], _conv_rec(OPS["conv_%s_rec" % tag]))
    assert y.shape == OPS["conv_%s_y" % tag].shape
    assert rel_err(y, OPS["conv_%s_y" % tag]) < 1e-5


def test_depthwise_kat():
    cv = _conv_rec(OPS["dw_rec"])
    cv.update(sep=1, pad=3, dil=3)
    assert rel_err(vn.conv1d(OPS["x"], cv), OPS["dw_y"]) < 1e-5


def test_conv_transposed_kat():
    cv = _conv_rec(OPS["convT_rec"], transposed=True)
    cv.update(stride=4, pad=2)
    y = vn.conv1d_transposed(OPS["x"], cv)
    assert y.shape == OPS["convT_y"].shape and rel_err(y, OPS["convT_y"]) < 1e-5


def test_layernorm_istft_pqmf_eltwise_kat():
    rec = OPS["ln_rec"]
    n = int(rec[0])
    assert rel_err(vn.layer_norm(OPS["x"], dict(gamma=rec[1:1 + n], beta=rec[1 + n:])), OPS["ln_y"]) < 1e-5
    assert rel_err(vn.istft(OPS["istft_mag"], OPS["istft_ph"]), OPS["istft_y"][0]) < 1e-5
    Hs = vn.pqmf_synthesis_filters()
    cv = dict(outCh=1, inCh=4, k=63, pad=31, dil=1, hasBias=0, w=Hs.T.reshape(1, 63, 4), b=None)
    assert rel_err(vn.conv1d(vn._zero_stuff4(OPS["pqmf_x"]), cv)[:, 0], OPS["pqmf_y"][:, 0]) < 1e-5
    assert rel_err(vn.tanh_ref(OPS["elt_x"]), OPS["tanh_y"]) < 1e-6
    assert rel_err(vn.gelu(OPS["elt_x"]), OPS["gelu_y"]) < 1e-6


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "synth_*.npz"))))
def test_numpy_oracle_vs_golden_synthetic(path):
    g = np.load(path)
    blob = binfmt.synthetic_model(seed=int(g["seed"]), **json.loads(str(g["hp"])))
    from tests.golden.make_golden import blob_sum
    if blob_sum(blob) != str(g["sha"]):
        pytest.skip("numpy RNG stream differs from the one the fixture was made with")
    M = binfmt.parse_model(blob)
    for pre, forced in (("a", None), ("b", g["forced"])):
        o = vn.infer(M, g["ids"], sid=int(g["sid"]), forced_w=forced)
        assert np.array_equal(o["w_ceil"], g[pre + "_wceil"])
        assert o["F"] == int(g[pre + "_F"])
        assert rel_err(o["z"], g[pre + "_z"]) < 1e-4
        assert rel_err(o["o"], g[pre + "_o"]) < 1e-4
        assert lsb_diff(o["pcm"], g[pre + "_pcm"]) <= 1


def test_numpy_oracle_vs_golden_shipped_fast():
    """single_speaker_fast + test.txt ids: frame count exact, waveform <= 1e-3, PCM within 2 LSB."""
    blob = find_model("single_speaker_fast")
    if blob is None:
        pytest.skip("shipped model not available here")
    g = np.load(os.path.join(GOLDEN, "real_single_speaker_fast.npz"))
    o = vn.infer(binfmt.parse_model(blob), g["ids"], sid=0, length_scale=float(g["ls"]))
    assert o["F"] == int(g["F"]) == 396 and np.array_equal(o["w_ceil"], g["wceil"])
    assert rel_err(o["o"], g["o"]) < 1e-3
    assert lsb_diff(o["pcm"], g["pcm"]) <= 2


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
def test_compiled_reference_reproduces_golden():
    """The fixtures really are what the compiled reference produces (guards against stale goldens)."""
    ref.set_threads(1)
    g = np.load(os.path.join(GOLDEN, "synth_ms_fix.npz"))
    blob = binfmt.synthetic_model(seed=int(g["seed"]), **json.loads(str(g["hp"])))
    r = ref.RefModel(blob).infer(g["ids"], sid=int(g["sid"]))
    assert np.array_equal(r.pcm, g["a_pcm"]) and r.F == int(g["a_F"])
    x = OPS["x"]
    assert np.array_equal(ref.conv1d(OPS["conv_dil_rec"], x), OPS["conv_dil_y"])
