"""Randomised model check of the split flash-attention kernel's synchronisation protocol (csrc/flash_attn.cu,
flash_attn_kernel): producer warp, MMA issuer, the in-order asynchronous tensor-core queue and 2 x 4 softmax warps run as
independent actors under random interleavings; every buffer (K/V ring stages, three S buffers, P[2], O_blk[2]) carries the
tile number it holds, every read asserts that it sees the tile it expects and every write asserts that all readers of the
previous contents are done.  mbarriers follow the hardware rule: wait(parity) succeeds iff the phase with that parity has
completed, i.e. iff the barrier's current phase parity differs from it (which is what makes two-phase aliasing possible).

    python tools/fa_protocol_check.py --runs 2000

A clean run says the barrier protocol is sound AS WRITTEN HERE; it says nothing about proxy fences or hardware semantics."""
import argparse
import random

NT, STAGES = 12, 4


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier expects"
        if self.pending == 0:
            self.pending = self.count
            self.phase += 1

    def ready(self, parity):
        return (self.phase & 1) != parity


class Sim:
    def __init__(self, rng):
        self.rng = rng
        self.kv_full = [Bar(1) for _ in range(STAGES)]
        self.kv_empty = [Bar(1) for _ in range(STAGES)]
        self.s_full = [Bar(1) for _ in range(3)]
        self.s_empty = [Bar(4) for _ in range(3)]
        self.p_full = [Bar(4) for _ in range(2)]
        self.o_full = [Bar(1) for _ in range(2)]
        self.o_empty = [Bar(4) for _ in range(2)]
        self.stage_tile = [None] * STAGES        # tile whose K/V the stage holds
        self.stage_reads = [0] * STAGES          # outstanding tensor-core reads (QK + PV) of the stage's tile: 2 when loaded
        self.S = [None] * 3                      # tile whose scores the buffer holds
        self.S_readers = [0] * 3                 # softmax warps that still have to finish reading it
        self.P = [None] * 2
        self.P_writers = [0, 0]                  # warps that have written their rows of the current P
        self.P_read_pending = [False, False]     # a P V MMA that will read it has been issued and not completed
        self.O = [None] * 2
        self.O_readers = [0, 0]
        self.queue = []                          # issued, not yet completed tensor-core ops, in order
        self.done_tiles = {0: [], 1: []}

    # ---- actors are generators yielding ("wait", bar, parity) or ("step",)
    def producer(self):
        stage, phase = 0, 0
        for j in range(NT):
            yield ("wait", self.kv_empty[stage], phase ^ 1)
            assert self.stage_reads[stage] == 0, f"TMA overwrites stage {stage} (tile {self.stage_tile[stage]}) while the tensor core still reads it"
            self.stage_tile[stage] = ("loading", j)
            yield ("step",)
            self.stage_tile[stage] = j
            self.stage_reads[stage] = 2
            self.kv_full[stage].arrive()
            stage += 1
            if stage == STAGES:
                stage, phase = 0, phase ^ 1

    def mma(self):
        def qk(j):
            st, sb = j % STAGES, j % 3
            yield ("wait", self.kv_full[st], (j // STAGES) & 1)
            yield ("wait", self.s_empty[sb], ((j // 3) & 1) ^ 1)
            self.queue.append(("qk", j))
            self.queue.append(("commit", self.s_full[sb]))
            yield ("step",)

        def pv(i):
            st, pb = i % STAGES, i & 1
            par = (i >> 1) & 1
            yield ("wait", self.p_full[pb], par)
            yield ("wait", self.o_empty[pb], par ^ 1)
            self.P_read_pending[pb] = True
            self.queue.append(("pv", i))
            self.queue.append(("commit", self.o_full[pb]))
            self.queue.append(("commit", self.kv_empty[st]))
            yield ("step",)
        yield from qk(0)
        yield from qk(1)
        for j in range(NT):
            if j + 2 < NT:
                yield from qk(j + 2)
            yield from pv(j)

    def tensor_core(self):
        while True:
            if not self.queue:
                yield ("idle",)
                continue
            op = self.queue.pop(0)
            if op[0] == "qk":
                j = op[1]
                st, sb = j % STAGES, j % 3
                assert self.stage_tile[st] == j, f"QK({j}) reads stage {st} holding {self.stage_tile[st]}"
                assert self.S_readers[sb] == 0, f"QK({j}) overwrites S[{sb}] (tile {self.S[sb]}) with {self.S_readers[sb]} softmax warps still reading"
                self.S[sb] = j
                self.S_readers[sb] = 4
                self.stage_reads[st] -= 1
            elif op[0] == "pv":
                i = op[1]
                st, pb = i % STAGES, i & 1
                assert self.stage_tile[st] == i, f"PV({i}) reads V of stage {st} holding {self.stage_tile[st]}"
                assert self.P[pb] == i and self.P_writers[pb] == 4, f"PV({i}) reads P[{pb}] = tile {self.P[pb]} written by {self.P_writers[pb]} warps"
                assert self.O_readers[pb] == 0, f"PV({i}) overwrites O[{pb}] (tile {self.O[pb]}) with {self.O_readers[pb]} warps still reading"
                self.O[pb] = i
                self.O_readers[pb] = 4
                self.P_read_pending[pb] = False
                self.stage_reads[st] -= 1
            else:
                op[1].arrive()
            yield ("step",)

    def softmax_warp(self, grp, w):
        k = 0
        for j in range(grp, NT, 2):
            sb = j % 3
            yield ("wait", self.s_full[sb], (j // 3) & 1)
            assert self.S[sb] == j, f"group {grp} warp {w}: pass 1 of tile {j} reads S[{sb}] holding tile {self.S[sb]}"
            yield ("step",)
            if k >= 1:
                yield from self.accumulate(grp, w, k - 1, j - 2)
            # pass 2: read S again, write P[grp]
            assert self.S[sb] == j, f"group {grp} warp {w}: pass 2 of tile {j} reads S[{sb}] holding tile {self.S[sb]}"
            assert not self.P_read_pending[grp], f"group {grp} warp {w}: writes P[{grp}] for tile {j} while PV({self.P[grp]}) is still pending"
            if self.P[grp] != j:
                self.P[grp] = j
                self.P_writers[grp] = 0
            self.P_writers[grp] += 1
            yield ("step",)
            self.S_readers[sb] -= 1
            self.s_empty[sb].arrive()
            self.p_full[grp].arrive()
            k += 1
        yield from self.accumulate(grp, w, k - 1, grp + 2 * (k - 1))
        if grp == 1:   # the merge: this warp's rows of (o, m, l) go into the P buffers
            assert not any(self.P_read_pending) and not any(op[0] == "pv" for op in self.queue), \
                f"group 1 warp {w} writes the merge buffer while a P V MMA may still read P"
        self.done_tiles[grp].append(w)

    def accumulate(self, grp, w, k, tile):
        yield ("wait", self.o_full[grp], k & 1)
        assert self.O[grp] == tile, f"group {grp} warp {w}: accumulate of tile {tile} reads O[{grp}] holding tile {self.O[grp]}"
        yield ("step",)
        self.O_readers[grp] -= 1
        self.o_empty[grp].arrive()

    def run(self):
        actors = {"producer": self.producer(), "mma": self.mma(), "tc": self.tensor_core()}
        for g in range(2):
            for w in range(4):
                actors[f"g{g}w{w}"] = self.softmax_warp(g, w)
        blocked = {}
        steps = 0
        self.speed = {n: self.rng.choice([0.05, 0.2, 0.5, 1.0]) for n in actors}
        while True:
            live = [n for n in actors if n != "tc"]
            if not live:
                return steps
            runnable = []
            for n in actors:
                b = blocked.get(n)
                if b is None or b[0].ready(b[1]):
                    runnable.append(n)
            progressed = False
            # every actor has its own speed in this run (a slow tensor core, one straggling softmax warp, a fast producer ...)
            chosen = [n for n in runnable if self.rng.random() < self.speed[n]] or [self.rng.choice(runnable)]
            self.rng.shuffle(chosen)
            for n in chosen:
                blocked.pop(n, None)
                try:
                    ev = next(actors[n])
                except StopIteration:
                    del actors[n]
                    progressed = True
                    continue
                if ev[0] == "wait":
                    if not ev[1].ready(ev[2]):
                        blocked[n] = (ev[1], ev[2])
                    progressed = True
                elif ev[0] == "step":
                    progressed = True
            steps += 1
            if not progressed:
                stuck = all(n == "tc" or (n in blocked and not blocked[n][0].ready(blocked[n][1])) for n in actors)
                if stuck and not self.queue:
                    raise RuntimeError("deadlock: " + ", ".join(f"{n} waits parity {b[1]} (phase {b[0].phase})" for n, b in blocked.items()))
            if steps > 200000:
                raise RuntimeError("no termination")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--runs", type=int, default=2000)
    a = ap.parse_args()
    for seed in range(a.runs):
        try:
            Sim(random.Random(seed)).run()
        except (AssertionError, RuntimeError) as e:
            print(f"seed {seed}: {type(e).__name__}: {e}")
            raise SystemExit(1)
    print(f"{a.runs} random interleavings: no hazard, no deadlock")
