"""Chunked / streaming synthesis (stts_infer_stream) on the GPU: the flow and the decoder run over frame
chunks with a receptive-field halo; the concatenated chunks must be BIT-IDENTICAL to the one-shot stts_infer_ids result (which
tests/test_gpu_parity.py pins to the compiled reference), for every decoder variant and with speaker conditioning."""
import glob
import json
import os

import numpy as np
import pytest
from parity_util import GOLDEN, TEST_TXT_IDS, find_model, synth_ids

from summertts_b200 import binfmt, engine

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fast_blob():
    b = find_model("single_speaker_fast")
    return b if b is not None else binfmt.synthetic_model(seed=11)


@pytest.mark.parametrize("chunk", [64, 100, 257, 100000])
def test_stream_equals_one_shot_full_size(native_lib, fast_blob, chunk):
    ids = (TEST_TXT_IDS[:-1] * 2) + [1]
    E = engine.SynthesizerTrn(fast_blob)
    whole = E.infer_ids(ids)
    chunks, first_ms = E.infer_stream(ids, chunk_frames=chunk)
    frames = whole.size // 256
    assert len(chunks) == (frames + chunk - 1) // chunk
    assert all(c.size == min(chunk, frames - i * chunk) * 256 for i, c in enumerate(chunks))
    assert np.array_equal(np.concatenate(chunks), whole)
    assert first_ms > 0
    E.close()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "synth_*.npz"))))
def test_stream_equals_one_shot_all_decoders(native_lib, path):
    """HiFi-GAN (+ speaker conditioning, stochastic duration predictor), multi-stream iSTFT, iSTFT and MB-iSTFT + PQMF decoders."""
    g = np.load(path)
    blob = binfmt.synthetic_model(seed=int(g["seed"]), **json.loads(str(g["hp"])))
    E = engine.SynthesizerTrn(blob)
    ids = synth_ids(np.random.default_rng(2), 60)
    E.set_forced_durations(np.full(60, 4.0, np.float32))
    whole = E.infer_ids(ids, int(g["sid"]), 1.0)
    chunks, _ = E.infer_stream(ids, int(g["sid"]), 1.0, chunk_frames=48)
    assert np.array_equal(np.concatenate(chunks), whole)
    E.close()


def test_first_chunk_latency_sanity(native_lib, fast_blob):
    """Latency to first audio.  Measured on B200 (profiles/r2_stream_latency.txt): a 15 s utterance takes 3.0 ms one-shot and the
    first 128-frame chunk is on the host after 3.0 ms as well -- at these sizes BOTH are bound by ~170 kernel launches plus one host
    round trip for the frame counts, not by GPU work, so chunking cannot cut the latency further (CUDA graphs would); it bounds the
    working set and lets the host consume PCM while later chunks run.  The test only guards against a regression."""
    ids = (TEST_TXT_IDS[:-1] * 8) + [1]
    E = engine.SynthesizerTrn(fast_blob)
    E.infer_ids(ids)
    E.infer_ids(ids)
    whole_ms = E.last_timing()["total"]
    _, first_ms = E.infer_stream(ids, chunk_frames=128)
    _, first_ms = E.infer_stream(ids, chunk_frames=128)
    assert 0 < first_ms < 2.0 * whole_ms
    E.close()
