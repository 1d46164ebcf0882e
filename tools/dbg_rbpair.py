"""Staged diagnosis of the fused ResBlock1-pair kernel on the GPU box (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from oracle import vits_numpy as vn
from summertts_b200 import engine, build
build.build_native()

def rec(w, b, k):
    c = w.shape[0]
    return np.concatenate([np.array([c, c, k, (k - 1) // 2, 1, 1], np.float32), w.ravel(), b]).astype(np.float32), \
        dict(outCh=c, inCh=c, k=k, pad=(k - 1) // 2, dil=1, hasBias=1, w=w, b=b, stride=1)

def oracle(x, cv1, cv2, d, seg, out_leaky):
    outs = []
    for i in range(len(seg) - 1):
        xs = x[seg[i]:seg[i + 1]]
        t = vn.conv1d(vn.leaky_relu(xs, 0.1), cv1, pad=d * (cv1["k"] - 1) // 2, dil=d)
        y = xs + vn.conv1d(vn.leaky_relu(t, 0.1), cv2)
        outs.append(vn.leaky_relu(y, 0.1) if out_leaky else y)
    return np.concatenate(outs, axis=0)

def run(c, k, d, mode, lens, stage, out_leaky=False):
    rng = np.random.default_rng(5)
    s = 0.6 / np.sqrt(c * k)
    w1 = (rng.standard_normal((c, k, c)) * s).astype(np.float32); b1 = (rng.standard_normal(c) * 0.1).astype(np.float32)
    w2 = (rng.standard_normal((c, k, c)) * s).astype(np.float32); b2 = (rng.standard_normal(c) * 0.1).astype(np.float32)
    if stage == "res":   w1[:] = 0; w2[:] = 0; b1[:] = 0
    if stage == "c2const": w1[:] = 0
    if stage == "c2id":  # conv1 = identity (center tap), conv2 random
        w1[:] = 0
        for o in range(c): w1[o, k // 2, o] = 1.0
        b1[:] = 0
    if stage == "c1only":  # conv2 = identity
        w2[:] = 0
        for o in range(c): w2[o, k // 2, o] = 1.0
        b2[:] = 0
    r1, cv1 = rec(w1, b1, k); r2, cv2 = rec(w2, b2, k)
    seg = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    x = (rng.standard_normal((seg[-1], c)) * 2).astype(np.float32)
    y, fl = engine.test_rbpair(r1, r2, x, dil1=d, mode=mode, seg_off=seg, out_leaky=out_leaky)
    want = oracle(x, cv1, cv2, d, seg, out_leaky)
    err = np.abs(y - want)
    r, col = np.unravel_index(err.argmax(), err.shape)
    u = int(np.searchsorted(seg, r, side="right") - 1)
    print("C=%d k=%d d=%d mode=%d stage=%-8s flags=%d  max err %.3e (max|want| %.2f) at row %d (utt %d, local %d) col %d ; y=%.4f want=%.4f x=%.4f" % (
        c, k, d, mode, stage, fl, err.max(), np.abs(want).max(), r, u, r - seg[u], col, y[r, col], want[r, col], x[r, col]))
    # error profile over local rows of the first long utterance and over columns
    big = int(np.argmax(lens)); a0, a1 = seg[big], seg[big + 1]
    e = err[a0:a1]
    rows = e.max(axis=1)
    bad = np.where(rows > 1e-3 * max(np.abs(want).max(), 1))[0]
    print("   utt %d (len %d): bad rows %d; first %s last %s ; per-col max err: %s" % (big, a1 - a0, bad.size, bad[:6], bad[-6:],
          np.array2string(e.max(axis=0)[:16], precision=3)))
    if bad.size and bad.size < e.shape[0]:
        good = np.setdiff1d(np.arange(e.shape[0]), bad)
        print("   good rows sample:", good[:8], "...", good[-8:])
    return err.max()

def run2(c, k, d, mode, lens, out_leaky):
    rng = np.random.default_rng(5)
    s = 0.6 / np.sqrt(c * k)
    w1 = (rng.standard_normal((c, k, c)) * s).astype(np.float32); b1 = (rng.standard_normal(c) * 0.1).astype(np.float32)
    w2 = (rng.standard_normal((c, k, c)) * s).astype(np.float32); b2 = (rng.standard_normal(c) * 0.1).astype(np.float32)
    r1, cv1 = rec(w1, b1, k); r2, cv2 = rec(w2, b2, k)
    seg = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    x = (rng.standard_normal((seg[-1], c)) * 2).astype(np.float32)
    y, fl = engine.test_rbpair(r1, r2, x, dil1=d, mode=mode, seg_off=seg, out_leaky=out_leaky)
    want = oracle(x, cv1, cv2, d, seg, out_leaky)
    err = np.abs(y - want)
    print("C=%d k=%d d=%d mode=%d leaky=%d lens=%s: max err %.3e" % (c, k, d, mode, out_leaky, lens, err.max()))
    for u in range(len(lens)):
        e = err[seg[u]:seg[u + 1]].max(axis=1)
        bad = np.where(e > 1e-4)[0]
        if bad.size:
            print("   utt %d len %d: %d bad rows, first %s last %s, max %.3e ; sample y/want row %d: %s / %s" % (
                u, lens[u], bad.size, bad[:5], bad[-5:], e.max(), bad[0], y[seg[u] + bad[0], :4], want[seg[u] + bad[0], :4]))


if __name__ == "__main__":
    ov = 254
    run2(32, 3, 1, 0, [1, 5, ov - 1, ov, ov + 1, 2 * ov, 2 * ov + 3, 37, 600], True)
    run2(32, 3, 1, 0, [1, 5, ov - 1, ov, ov + 1, 2 * ov, 2 * ov + 3, 37, 600], False)
    run2(32, 3, 1, 0, [600, 40], True)
    run2(32, 3, 1, 0, [5, 600], False)
    run2(32, 3, 1, 0, [254, 600], False)
    run2(32, 3, 1, 0, [255, 600], False)
    run2(32, 3, 1, 0, [600, 255, 7], False)
    sys.exit(0)
if __name__ == "__main__":
    lens = [600, 40]
    for (c, k, d) in [(32, 3, 1)]:
        for mode in (1, 0):
            for stage in ("res", "c2const", "c2id", "c1only", "full"):
                run(c, k, d, mode, lens, stage)
    for (c, k, d) in [(32, 7, 3), (64, 3, 1), (64, 11, 5)]:
        for mode in (1, 0):
            run(c, k, d, mode, lens, "full")
