"""oracle/cpu_worker_steps.py — TEST/BENCH INFRASTRUCTURE: times the reference's own CPU
implementation (oracle/_ref = the unmodified reference objects compiled by oracle/Makefile) on a list
of utterances, for bench.py's cpu_baseline / --impl reference legs.  One process = one host thread;
bench.py starts one per core.  Prints one JSON line with per-rep start/end wall-clock stamps."""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    spec = json.loads(sys.argv[1])
    from oracle import ref

    ref.set_threads(int(spec.get("threads", 1)))
    M = ref.RefModel(np.fromfile(spec["model"], dtype=np.float32))
    utts, forced = spec["utts"], spec.get("forced")
    M.infer(utts[0][:16], dumps=False)  # untimed warm-up (allocator / OpenMP pool)
    while time.time() < float(spec.get("start_at", 0)):  # crude start barrier so the workers overlap
        time.sleep(0.001)
    t0s, t1s, samples = [], [], []
    for _rep in range(int(spec.get("reps", 1))):
        t0 = time.time()
        n = 0
        for i, ids in enumerate(utts):
            r = M.infer(ids, sid=int(spec.get("sid", 0)), length_scale=float(spec.get("ls", 1.0)),
                        forced_w=None if forced is None else forced[i], dumps=False)
            n += r.S
        t0s.append(t0)
        t1s.append(time.time())
        samples.append(n)
    print(json.dumps({"t0": t0s, "t1": t1s, "samples": samples}))


if __name__ == "__main__":
    main()
