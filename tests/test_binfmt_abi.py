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
