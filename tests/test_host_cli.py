"""tts_b200 (summertts_b200/host/tts_cli.cpp): the demo CLI + WAV container (reference
test/main.cpp:7-65,75-148).  CPU: container bytes, argument / error behaviour, loud failure without a GPU.
GPU: phoneme-id batches through the CLI give exactly the PCM the C ABI returns."""
import os
import struct
import subprocess
import wave

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "summertts_b200", "host")
CLI = os.path.join(ROOT, "summertts_b200", "bin", "tts_b200_ids")


@pytest.fixture(scope="module")
def cli(native_lib):
    r = subprocess.run(["make", "-C", HOST, "ids-cli"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert os.path.exists(CLI)
    return CLI


def ref_header(data_bytes, rate=16000):
    """The 44 bytes convertAudioToWavBuf writes (test/main.cpp:13-60), restated with struct."""
    return (b"RIFF" + struct.pack("<I", data_bytes + 36) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, rate, rate * 2, 2, 16)
            + b"data" + struct.pack("<I", data_bytes))


def test_wav_container_bytes(cli, tmp_path):
    out = tmp_path / "tone.wav"
    assert subprocess.run([cli, "--selftest-wav", str(out)]).returncode == 0
    raw = out.read_bytes()
    assert len(raw) == 44 + 2 * 16000
    assert raw[:44] == ref_header(2 * 16000)
    with wave.open(str(out)) as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 16000, 16000)
        pcm = np.frombuffer(w.readframes(16000), dtype="<i2")
    n = np.arange(16000)
    assert np.array_equal(pcm, np.rint(8000.0 * np.sin(2.0 * np.pi * 440.0 * n / 16000)).astype(np.int16))


def test_cli_argument_errors(cli, tmp_path):
    assert subprocess.run([cli], capture_output=True).returncode == 2
    assert subprocess.run([cli, "--bogus", "a", "b", "c"], capture_output=True).returncode == 2
    ids = tmp_path / "ids.txt"
    ids.write_text("1 2 3 4 5 6\n")
    r = subprocess.run([cli, "--ids", str(ids), str(tmp_path / "missing.bin"), str(tmp_path / "o.wav")], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot read model" in r.stderr
    junk = tmp_path / "junk.bin"
    junk.write_bytes(np.arange(64, dtype=np.float32).tobytes())
    r = subprocess.run([cli, "--ids", str(ids), str(junk), str(tmp_path / "o.wav")], capture_output=True, text=True)
    assert r.returncode == 1 and "model rejected" in r.stderr


def _synthetic_model(tmp_path):
    from summertts_b200 import binfmt

    blob = binfmt.synthetic_model(7)
    p = tmp_path / "synthetic.bin"
    np.asarray(blob, dtype=np.float32).tofile(p)
    return p, blob


def test_cli_needs_ids_or_frontend_and_fails_loudly_without_gpu(cli, tmp_path):
    model, _ = _synthetic_model(tmp_path)
    ids = tmp_path / "ids.txt"
    ids.write_text("0 9 30 3 11 40 3 0 1\n0 8 29 3 12 41 3 13 50 3 0 0 1\n")
    r = subprocess.run([cli, str(ids), str(model), str(tmp_path / "o.wav")], capture_output=True, text=True)
    assert r.returncode == 1 and "no text frontend" in r.stderr
    bad = tmp_path / "bad.txt"
    bad.write_text("1 2 x\n")
    r = subprocess.run([cli, "--ids", str(bad), str(model), str(tmp_path / "o.wav")], capture_output=True, text=True)
    assert r.returncode == 1 and "bad id line" in r.stderr
    import torch

    if not torch.cuda.is_available():      # product rule: no CPU fallback
        r = subprocess.run([cli, "--ids", str(ids), str(model), str(tmp_path / "o.wav")], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stderr
        assert not (tmp_path / "o.wav").exists()


@pytest.mark.gpu
def test_cli_ids_batch_matches_c_abi(cli, tmp_path):
    """Three ragged phoneme-id lines -> <out>_%04d.wav; payloads bit-identical to stts_infer_ids, headers canonical."""
    from summertts_b200 import engine

    model, blob = _synthetic_model(tmp_path)
    rng = np.random.default_rng(3)
    utts = [[0] + [int(v) for v in rng.integers(1, 40, size=n)] + [0, 1] for n in (9, 23, 14)]
    ids = tmp_path / "ids.txt"
    ids.write_text("".join(" ".join(map(str, u)) + "\n" for u in utts))
    out = tmp_path / "o.wav"
    r = subprocess.run([cli, "--ids", str(ids), str(model), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    tts = engine.SynthesizerTrn(np.asarray(blob, dtype=np.float32))
    for i, u in enumerate(utts):
        raw = (tmp_path / ("o.wav_%04d.wav" % i)).read_bytes()
        pcm = tts.infer_ids(np.asarray(u, np.int32), 0, 1.0)
        assert raw[:44] == ref_header(2 * pcm.size)
        assert np.array_equal(np.frombuffer(raw[44:], dtype="<i2"), pcm)
    # --concat: one file, utterances in input order
    r = subprocess.run([cli, "--ids", "--concat", str(ids), str(model), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = out.read_bytes()
    allpcm = np.concatenate([tts.infer_ids(np.asarray(u, np.int32), 0, 1.0) for u in utts])
    assert raw[:44] == ref_header(2 * allpcm.size) and np.array_equal(np.frombuffer(raw[44:], dtype="<i2"), allpcm)
    tts.close()


REF_CLI = os.path.join(HOST, "_build", "tts_b200")
REF_ROOT = "/root/reference"


@pytest.mark.skipif(not (os.path.exists(REF_CLI) and os.path.isdir(os.path.join(REF_ROOT, "models"))),
                    reason="frontend-enabled tts_b200 needs the reference tree (make -C summertts_b200/host REF=...)")
def test_cli_text_frontend_reproduces_reference_ids(tmp_path):
    """Text -> ids through the CLI (reference frontend behind frontend.hpp) == the id sequence the reference's own
    SynthesizerTrn::infer produces for test.txt, English: 700 ids for test_eng.txt; several frontend
    worker threads give the same ids as one."""
    from parity_util import TEST_TXT_IDS

    def dump(args):
        r = subprocess.run([REF_CLI, "--dump-ids"] + args, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        return [[int(v) for v in l.split()] for l in r.stdout.splitlines() if l.strip()]

    chs = os.path.join(REF_ROOT, "models", "single_speaker_fast.bin")
    got = dump([os.path.join(REF_ROOT, "test.txt"), chs, str(tmp_path / "o.wav")])
    assert got == [list(TEST_TXT_IDS)]
    eng = dump([os.path.join(REF_ROOT, "test_eng.txt"), os.path.join(REF_ROOT, "models", "single_speaker_english_fast.bin"),
                str(tmp_path / "o.wav")])
    assert len(eng) == 1 and len(eng[0]) == 700
    corpus = tmp_path / "corpus.txt"
    corpus.write_text("今天天气不错,我们一起去公园散步吧。\n会议将于2024年5月20日上午10点举行。\n银行的行长今天在银行门口行走。\n" * 4)
    one = dump(["--per-line", "--frontend-threads", "1", str(corpus), chs, str(tmp_path / "o.wav")])
    # the heuristic would pick 1 worker for 12 lines; exercise the parallel path with an explicit count on a longer file
    corpus.write_text(corpus.read_text() * 70)      # 840 lines -> 3 workers allowed
    many = dump(["--per-line", "--frontend-threads", "3", str(corpus), chs, str(tmp_path / "o.wav")])
    assert len(one) == 12 and len(many) == 840 and many[:12] == one and many[12:24] == one
