"""Golden output of the reference's OWN demo program for the drop-in test (tests/test_dropin.py).

Runs oracle/_ref/tts_test_ref (the unmodified reference: test/main.cpp + SynthesizerTrn.cpp + Eigen NN + host frontend, built by
`make -C oracle tts_test_ref`) on /root/reference/test.txt with single_speaker_fast.bin, OMP_NUM_THREADS=1, and stores the WAV it
writes (44-byte header + PCM, test/main.cpp:7-65) together with the input text.  Build container only (needs /root/reference).

    python tests/golden/make_dropin_golden.py
"""
import hashlib
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "tts_test_ref")
    text = open(os.path.join(REF, "test.txt"), "rb").read()
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "out.wav")
        env = dict(os.environ, OMP_NUM_THREADS="1")
        subprocess.run([exe, os.path.join(REF, "test.txt"), os.path.join(REF, "models", "single_speaker_fast.bin"), out], check=True,
                       env=env, stdout=subprocess.DEVNULL)
        wav = open(out, "rb").read()
    hdr = np.frombuffer(wav[:44], dtype=np.uint8)
    pcm = np.frombuffer(wav[44:], dtype=np.int16)
    np.savez_compressed(os.path.join(HERE, "dropin_single_speaker_fast.npz"), text=np.frombuffer(text, dtype=np.uint8), header=hdr, pcm=pcm,
                        md5=hashlib.md5(wav).hexdigest())
    print("wav bytes", len(wav), "samples", pcm.size, "md5", hashlib.md5(wav).hexdigest())


if __name__ == "__main__":
    main()
