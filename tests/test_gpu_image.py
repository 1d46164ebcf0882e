"""Pre-packed device image cache (stts_create_cached) on the GPU: an engine built from the image must be
bit-identical to one built from the blob; stale, foreign and truncated images are rejected and rewritten."""
import os
import time

import numpy as np
import pytest
from parity_util import TEST_TXT_IDS, find_model

from summertts_b200 import binfmt, engine

pytestmark = pytest.mark.gpu


def test_image_roundtrip_and_rejection(native_lib, tmp_path):
    blob = find_model("single_speaker_fast")
    if blob is None:
        blob = binfmt.synthetic_model(seed=11)
    img = str(tmp_path / "fast.stts_image")
    t0 = time.time(); A = engine.SynthesizerTrn(blob); t_plain = time.time() - t0
    want = A.infer_ids(TEST_TXT_IDS)
    A.close()
    t0 = time.time(); B = engine.SynthesizerTrn(blob, image_path=img); t_rec = time.time() - t0
    assert not B.from_image and os.path.getsize(img) > 1 << 20
    assert np.array_equal(B.infer_ids(TEST_TXT_IDS), want)
    B.close()
    t0 = time.time(); Cc = engine.SynthesizerTrn(blob, image_path=img); t_img = time.time() - t0
    assert Cc.from_image
    assert np.array_equal(Cc.infer_ids(TEST_TXT_IDS), want)
    Cc.close()
    print("create: plain %.3f s, recording %.3f s, from image %.3f s (%d MB)" % (t_plain, t_rec, t_img, os.path.getsize(img) >> 20))
    # a different model must not accept this image: it is rewritten
    other = binfmt.synthetic_model(seed=12)
    D = engine.SynthesizerTrn(other, image_path=img)
    assert not D.from_image
    D.close()
    E = engine.SynthesizerTrn(other, image_path=img)
    assert E.from_image
    E.close()
    # truncated image: rejected, engine still correct, image rewritten
    with open(img, "r+b") as f:
        f.truncate(os.path.getsize(img) // 2)
    F = engine.SynthesizerTrn(other, image_path=img)
    assert not F.from_image
    F.close()
    assert engine.SynthesizerTrn(other, image_path=img).from_image
