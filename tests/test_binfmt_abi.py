"""CPU tests: .bin reader/writer, the C++ parser behind the C ABI, and the ABI surface itself."""
import ctypes as C
import os
import re

import numpy as np
import pytest
from parity_util import ROOT, find_model

from summertts_b200 import binfmt, engine

HP_SMALL = dict(nLayers=1, preCh=32)


def test_writer_reader_roundtrip():
    for hp in (dict(decType=1, durPredType=1), dict(decType=0, durPredType=0, isMS=1, spkNum=3, gin=16,
                                                     upRates=(4, 2), upK=(8, 4)),
               dict(decType=2), dict(decType=3)):
        hp = dict(HP_SMALL, **hp)
        blob = binfmt.synthetic_model(seed=1, **hp)
        M = binfmt.parse_model(blob)
        assert M["nn_end"] == blob.size
        assert M["decType"] == hp.get("decType", 1) and M["isMS"] == hp.get("isMS", 0)
        assert len(M["flow"]["layers"]) == 4 and M["flow"]["layers"][0]["wn"]["in_layers"][0]["pad"] == 2
        assert len(M["dec"]["resblocks"]) == len(M["dec"]["ups"]) * 3
        # deterministic
        assert np.array_equal(blob, binfmt.synthetic_model(seed=1, **hp))


def test_cabi_parser_matches_python(native_lib):
    for hp in (dict(decType=1), dict(decType=0, durPredType=0, isMS=1, spkNum=3, gin=16, upRates=(4, 2), upK=(8, 4))):
        blob = binfmt.synthetic_model(seed=2, **dict(HP_SMALL, **hp))
        txt, end = engine.describe_model(blob)
        M = binfmt.parse_model(blob)
        assert end == M["nn_end"]
        assert "dec=%d" % M["decType"] in txt.splitlines()[0]
        n_conv_lines = sum(1 for l in txt.splitlines() if " conv " in l)
        assert n_conv_lines > 50
        # trailing bytes after the NN section (the frontend tail) are ignored by the parser
        tail = np.concatenate([blob, np.arange(100, dtype=np.float32)])
        assert engine.describe_model(tail)[1] == M["nn_end"]


def test_cabi_rejects_truncated_and_garbage(native_lib):
    blob = binfmt.synthetic_model(seed=3, **HP_SMALL)
    with pytest.raises(engine.SttsError) as ei:
        engine.describe_model(blob[: blob.size // 2])
    assert ei.value.code == -2  # STTS_E_FORMAT
    bad = blob.copy()
    bad[3] = 9  # unknown decoder type (SynthesizerTrn.cpp:126-132)
    with pytest.raises(engine.SttsError) as ei:
        engine.describe_model(bad)
    assert ei.value.code == -2


@pytest.mark.parametrize("name,end", [("single_speaker_fast", 14771313), ("single_speaker_mid", 17388305),
                                      ("single_speaker_english_fast", 14763441), ("multi_speakers", 16744854)])
def test_shipped_models_parse(native_lib, name, end):
    """NN-section end offsets measured from the reference constructors."""
    blob = find_model(name)
    if blob is None:
        pytest.skip("shipped model not available here")
    assert binfmt.parse_model(blob)["nn_end"] == end
    assert engine.describe_model(blob)[1] == end


def test_abi_exports_every_declared_symbol(native_lib):
    hdr = open(os.path.join(ROOT, "include", "stts_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(stts_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    for sym in declared:
        assert hasattr(native_lib, sym), "libstts_b200.so does not export " + sym
    assert sorted(engine.ABI_SYMBOLS) == declared
    assert b"sm_100a" in native_lib.stts_version()


def test_create_fails_loudly_without_gpu(native_lib):
    """No CPU fallback: without a usable sm_100 device stts_create must return STTS_E_CUDA."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    blob = binfmt.synthetic_model(seed=4, **HP_SMALL)
    h = C.c_void_p()
    rc = native_lib.stts_create(blob.ctypes.data, blob.nbytes, 0, C.byref(h))
    assert rc == -4 and not h.value
    assert b"no CUDA device" in native_lib.stts_last_error()
    with pytest.raises(engine.SttsError):
        engine.SynthesizerTrn(blob)


@pytest.mark.parametrize("shape", [(384, 5, 192), (64, 7, 64), (32, 11, 32), (72, 7, 32), (192, 1, 96), (96, 1, 192), (20, 3, 48)])
def test_tc_weight_packing(native_lib, shape):
    """Host side of the tensor-core path (conv_tc.cuh, tc_pack_weights_host): tiling choice, power-of-two scale and the
    split-fp16 UMMA core-matrix stage layout, against an independent numpy restatement.  Bit-exact."""
    from summertts_b200 import engine

    o, k, c = shape
    rng = np.random.default_rng(o * 131 + k * 17 + c)
    w = (rng.standard_normal(shape) * rng.uniform(0.01, 2.0)).astype(np.float32)
    meta, halves = engine.debug_pack_weights(w)
    assert meta["eligible"] == 1
    NC, nch, KC, kch, merged = meta["NC"], meta["nchunks"], meta["KC"], meta["kchunks"], meta["merged"]
    assert KC * kch == c and NC % 16 == 0 and NC * nch >= o and NC * (nch - 1) < o
    assert meta["colsplit"] == (1 if (o >= 128 and KC == 64) else 0)
    assert merged == (1 if (kch == 1 and not meta["colsplit"] and k * (KC // 16) <= 24) else 0)
    # largest |w| * 2^e lands in [512, 1024): fp16-safe and far from the subnormal range
    e = meta["wexp"]
    assert 512.0 <= np.abs(w).max() * 2.0 ** e < 1024.0
    v = (w * np.float32(2.0 ** e)).astype(np.float32)
    hi = v.astype(np.float16)
    lo = (v - hi.astype(np.float32)).astype(np.float16)
    # hi + lo reproduces the scaled weight to ~2^-22 relative
    assert np.abs(hi.astype(np.float64) + lo.astype(np.float64) - v).max() <= np.abs(v).max() * 2.0 ** -21
    stage = 2 * KC * NC
    assert halves.size == nch * kch * k * stage
    exp = np.zeros(halves.size, np.uint16)
    hb, lb = hi.view(np.uint16), lo.view(np.uint16)
    for nc in range(nch):
        for kc in range(kch):
            for tap in range(k):
                base = ((nc * kch + kc) * k + tap) * stage
                for cc in range(KC):
                    n = np.arange(NC)
                    oo = nc * NC + n
                    ok = oo < o
                    src_h = np.where(ok, hb[np.minimum(oo, o - 1), tap, kc * KC + cc], 0)
                    src_l = np.where(ok, lb[np.minimum(oo, o - 1), tap, kc * KC + cc], 0)
                    if merged:     # [c/8][hi rows 0..NC-1 | lo rows NC..2NC-1][c%8]
                        ih = base + ((cc // 8) * 2 * NC + n) * 8 + cc % 8
                        exp[ih] = src_h
                        exp[ih + NC * 8] = src_l
                    else:          # hi plane [c/8][n][c%8], then the lo plane
                        ih = base + ((cc // 8) * NC + n) * 8 + cc % 8
                        exp[ih] = src_h
                        exp[ih + KC * NC] = src_l
    assert np.array_equal(halves, exp)


def test_tc_weight_packing_rejects_odd_shapes(native_lib):
    from summertts_b200 import engine

    assert engine.debug_pack_weights(np.ones((8, 3, 32), np.float32))[0]["eligible"] == 0     # < 16 output channels
    assert engine.debug_pack_weights(np.ones((32, 3, 20), np.float32))[0]["eligible"] == 0    # inCh % 16 != 0


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under summertts_b200/ (Python, CUDA, C++ host) and no product entry point may
    import, link or execute it; bench.py may only in its CPU legs, __graft_entry__ only to build it and in smoke()."""
    import re

    pkg = os.path.join(ROOT, "summertts_b200")
    bad = []
    for d, _, files in os.walk(pkg):
        if os.path.basename(d) in ("_build", "bin", "__pycache__"):
            continue
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")) and f != "Makefile":
                continue
            txt = open(os.path.join(d, f), errors="replace").read()
            if re.search(r"\boracle\b", txt):
                bad.append(os.path.join(d, f))
    assert not bad, "product files mention the oracle: %s" % bad
    hdr = open(os.path.join(ROOT, "include", "stts_b200.h")).read()
    assert "oracle" not in hdr
    # bench.py: the oracle only inside the CPU legs (cpu_worker_steps / --impl reference)
    for line in open(os.path.join(ROOT, "bench.py")):
        if re.search(r"^\s*(from|import)\s+oracle", line):
            raise AssertionError("bench.py imports the oracle at module level: " + line)
