#!/usr/bin/env python3
"""Discrete-event model of pc_fused.cuh's mbarrier protocol (issuers x4, epilogue warps x16, weight producer, tile producer), CPU only.

mbarrier parity waits cannot tell phase n from phase n + 2; a role that skips a phase of a barrier deadlocks or — worse — passes
early.  Every wait here carries the phase it MEANS; the model asserts that the wait passes on exactly that phase, and that every
role terminates, under randomised latencies.  Mirrors pc_issuer / pc_kernel; keep in sync when the protocol changes.
"""
import heapq, random, sys


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.tx, self.phase, self.waiters = name, count, count, 0, 0, []

    def _check(self, sim):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count
            ws, self.waiters = self.waiters, []
            for w in ws:
                sim.try_wake(self, *w)

    def arrive(self, sim, tx=0):
        self.tx += tx
        self.pending -= 1
        assert self.pending >= 0, "too many arrivals on " + self.name
        self._check(sim)

    def complete_tx(self, sim, n):
        self.tx -= n
        self._check(sim)


class Sim:
    def __init__(self, seed):
        self.t, self.q, self.n, self.rng, self.errors, self.live = 0.0, [], 0, random.Random(seed), [], {}

    def at(self, dt, fn):
        self.n += 1
        heapq.heappush(self.q, (self.t + dt, self.n, fn))

    def spawn(self, name, gen):
        self.live[name] = "start"
        self._step(name, gen)

    def try_wake(self, bar, name, gen, parity, want):
        if (bar.phase & 1) != parity:
            if bar.phase != want + 1:
                self.errors.append("%s passed %s at phase %d, meant completion of phase %d" % (name, bar.name, bar.phase, want))
            self.at(self.rng.uniform(1, 30), lambda: self._step(name, gen))
        else:
            if bar.phase > want + 1:
                self.errors.append("%s waits on %s for phase %d but the barrier is already in phase %d (will alias)" % (name, bar.name, want, bar.phase))
            bar.waiters.append((name, gen, parity, want))
            self.live[name] = "wait %s phase %d" % (bar.name, want)

    def _step(self, name, gen):
        try:
            op = next(gen)
        except StopIteration:
            self.live.pop(name, None)
            return
        if op[0] == "wait":
            self.try_wake(op[1], name, gen, op[2], op[3])
        elif op[0] == "delay":
            self.at(op[1], lambda: self._step(name, gen))
        elif op[0] == "poll":       # cooperative polling loop iteration
            self.at(self.rng.uniform(5, 40), lambda: self._step(name, gen))

    def run(self, limit=5e7):
        while self.q and self.t < limit:
            self.t, _, fn = heapq.heappop(self.q)
            fn()
        return not self.live


def simulate(seed, nst, nb, npairs, nchunks, nitems, mode, rs, SPU=2, acc_first=False):
    sim = Sim(seed)
    R = sim.rng
    a_full, a_empty = Bar("a_full", 1), Bar("a_empty", 4)
    acc_full = [Bar("acc_full%d" % i, 1) for i in range(4)]
    acc_empty = [Bar("acc_empty%d" % i, 8) for i in range(4)]
    r_full = [Bar("r_full%d" % i, 1) for i in range(2)]
    s_free = [Bar("s_free%d" % i, 1) for i in range(2)]
    b_full = [Bar("b_full%d" % i, 1) for i in range(nb)]
    b_empty = [Bar("b_empty%d" % i, 2) for i in range(nb)]
    NU = 2 if mode else (nst + SPU - 1) // SPU
    w0, w1 = 1, 1 + nitems            # start mid-tile on purpose

    def test(bar, parity):           # mbarrier.test_wait
        return (bar.phase & 1) != parity

    def issuer(slot, b):
        af_k = 0; e_k = -1; ring0 = 0; prev_tg = -1; mma_done = [0.0]
        fills = {}                                           # ring slot -> fills seen
        rs0, rp0 = [slot], [0]                               # the kernel's incremental ring position

        def commit(bar):                                     # tcgen05.commit: arrives when this thread's prior MMAs are done
            sim.at(max(0.0, mma_done[0] - sim.t) + R.uniform(20, 200), lambda: bar.arrive(sim))
        for w in range(w0, w1):
            pair, tg = w % npairs, w // npairs
            active = pair * 2 + slot < nchunks
            if tg != prev_tg:
                yield ("wait", a_full, af_k & 1, af_k); af_k += 1
            prev_tg = tg
            for st in range(nst):
                own = ((st & 1) == b) if mode else (((st // SPU) & 1) == b)
                first = (st == b) if mode else (st % SPU == 0)
                last = (st + 2 >= nst) if mode else (st % SPU == SPU - 1 or st == nst - 1)
                if own and active and first:
                    yield ("wait", acc_empty[slot * 2 + b], (e_k & 1), e_k); e_k += 1
                q = ring0 + 2 * st + slot
                r = q % nb
                k = fills.get(r, 0); fills[r] = k + 1
                assert (q // nb) == k, "ring walk mismatch"
                assert r == rs0[0] and (k & 1) == rp0[0], "incremental ring walk mismatch"
                rs0[0] += 2
                if rs0[0] >= nb:
                    rs0[0] -= nb; rp0[0] ^= 1
                yield ("wait", b_full[r], k & 1, k)
                if own and active:
                    yield ("delay", R.uniform(100, 600))
                    mma_done[0] = max(mma_done[0], sim.t) + R.uniform(200, 900)
                commit(b_empty[r])
                if own and active and last:
                    commit(acc_full[slot * 2 + b])
            ring0 += 2 * nst
            if w + 1 == w1 or (w + 1) // npairs != tg:
                commit(a_empty)

    def epilogue(slot, wi):
        fk = [0, 0]; rk = 0
        for w in range(w0, w1):
            pair = w % npairs
            chunk = pair * 2 + slot
            if chunk >= nchunks:
                continue
            for un in range(NU):
                b = un & 1
                yield ("wait", acc_full[slot * 2 + b], fk[b] & 1, fk[b]); fk[b] += 1
                yield ("delay", R.uniform(50, 400))
                acc_empty[slot * 2 + b].arrive(sim)
            if rs:
                accin = (chunk % 2 == 0) or not acc_first      # some chunks accumulate, some store fresh
                yield ("wait", r_full[slot], rk & 1, rk); rk += 1
                yield ("delay", R.uniform(200, 1500))
                if wi == 0:
                    yield ("delay", R.uniform(100, 800))        # bulk store + wait_read
                    s_free[slot].arrive(sim)
            else:
                yield ("delay", R.uniform(200, 2500))

    def weight_producer():                      # warp 20: tight blocking loop over the ring
        fill = 0
        for w in range(w0, w1):
            for st in range(nst):
                for q in range(2):
                    ws = fill % nb; k = fill // nb     # fill k of ring slot ws needs release k - 1
                    if k >= 1:
                        yield ("wait", b_empty[ws], (k - 1) & 1, k - 1)
                    bar = b_full[ws]
                    bar.arrive(sim, tx=1)
                    sim.at(R.uniform(300, 3000), (lambda bb: (lambda: bb.complete_tx(sim, 1)))(bar))
                    fill += 1
                    yield ("delay", R.uniform(20, 120))

    def tile_producer():                        # warp 21: activation tile + residual tiles, in item order, blocking
        ae_k = -1; sf_k = [-1, -1]; prev_tg = -1
        for w in range(w0, w1):
            pair, tg = w % npairs, w // npairs
            if tg != prev_tg:
                if ae_k >= 0:
                    yield ("wait", a_empty, ae_k & 1, ae_k)
                ae_k += 1
                a_full.arrive(sim, tx=1)
                sim.at(R.uniform(500, 4000), lambda: a_full.complete_tx(sim, 1))
            prev_tg = tg
            if rs:
                for rslot in range(2):
                    chunk = pair * 2 + rslot
                    if chunk >= nchunks:
                        continue
                    need = (chunk % 2 == 0) or not acc_first
                    if sf_k[rslot] >= 0:
                        yield ("wait", s_free[rslot], sf_k[rslot] & 1, sf_k[rslot])
                    sf_k[rslot] += 1
                    bar = r_full[rslot]
                    if need:
                        bar.arrive(sim, tx=1)
                        sim.at(R.uniform(300, 2500), (lambda bb: (lambda: bb.complete_tx(sim, 1)))(bar))
                    else:
                        bar.arrive(sim)
            yield ("delay", R.uniform(20, 200))

    for s in range(2):
        for b in range(2):
            sim.spawn("issuer(s%d,b%d)" % (s, b), issuer(s, b))
        for wi in range(8):
            sim.spawn("epi(s%d,w%d)" % (s, wi), epilogue(s, wi))
    sim.spawn("weights", weight_producer())
    sim.spawn("tiles", tile_producer())
    ok = sim.run()
    return ok, sim


def main():
    cases = [dict(nst=3, nb=4, npairs=3, nchunks=6, mode=0, rs=True), dict(nst=3, nb=4, npairs=2, nchunks=3, mode=0, rs=True),
             dict(nst=3, nb=4, npairs=3, nchunks=6, mode=1, rs=True), dict(nst=15, nb=6, npairs=3, nchunks=6, mode=0, rs=False),
             dict(nst=15, nb=6, npairs=3, nchunks=6, mode=1, rs=False), dict(nst=15, nb=4, npairs=3, nchunks=6, mode=0, rs=False),
             dict(nst=15, nb=2, npairs=3, nchunks=6, mode=0, rs=False), dict(nst=5, nb=8, npairs=1, nchunks=1, mode=0, rs=True),
             dict(nst=2, nb=8, npairs=2, nchunks=4, mode=1, rs=True, acc_first=True)]
    bad = 0
    for c in cases:
        for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
            ok, sim = simulate(seed, nitems=7, **c)
            if not ok or sim.errors:
                bad += 1
                print("FAIL", c, "seed", seed, "deadlock" if not ok else "", sim.errors[:3], dict(list(sim.live.items())[:6]))
                break
        else:
            print("ok  ", c)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
