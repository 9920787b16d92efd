"""Randomised discrete-event model of the mbarrier protocol of attention_pipe_kernel (csrc/encoder.cu).

The kernel was written without GPU access, so its synchronisation is checked here on the CPU: the producer/issuer thread
and the four softmax/epilogue warps are coroutines that execute the SAME sequence of waits / arrives / asynchronous
completions as the CUDA code (same barrier counts, same parities), a scheduler interleaves them at random, asynchronous
agents (TMA completion, tensor-core commit) fire after random delays in issue order, and a model of the two smem buffers /
TMEM regions asserts that nothing is overwritten while still in use.  A deadlock or a hazard fails the test.

mbarrier model: `wait(parity)` succeeds once the phase with that parity has completed (a fresh barrier has "completed"
parity 1), `arrive` decrements the pending count of the current phase, the phase flips when it reaches zero; expect_tx
adds bytes that complete_tx must deliver before the phase can flip.
"""
import random


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.tx, self.cur = count, count, 0, 0

    def _maybe_flip(self):
        if self.pending == 0 and self.tx == 0:
            self.cur ^= 1
            self.pending = self.count

    def arrive(self):
        assert self.pending > 0, "more arrivals than the barrier's count in one phase"
        self.pending -= 1
        self._maybe_flip()

    def arrive_expect_tx(self, nbytes):
        self.tx += nbytes
        self.arrive()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        self._maybe_flip()

    def done(self, parity):
        return self.cur != parity


def simulate(n_items, seed, mutation=None):
    rnd = random.Random(seed)
    full, s_ready, p_ready, o_ready, free_ = ([MBar(1) for _ in range(2)], [MBar(1) for _ in range(2)], [MBar(4) for _ in range(2)],
                                              [MBar(1) for _ in range(2)], [MBar(4) for _ in range(2)])
    # resource model: which item's data each buffer currently holds, and who may still be using it
    smem = [{"item": None, "state": "free"} for _ in range(2)]      # free -> loading -> qkv -> p (P over Q|K) -> free
    tmem = [{"item": None, "state": "free"} for _ in range(2)]      # free -> s -> o -> free
    readers_left = [0, 0]                                           # softmax warps that still have to read this buffer's TMEM
    async_q = []                                                    # (fire_time, seqno, callback): TMA / tensor-core completions
    mma_chain = []                                                  # tensor-core work retires in issue order
    clock = [0]
    seqno = [0]
    done_items = []

    def later(cb, lo=1, hi=12):
        seqno[0] += 1
        async_q.append([clock[0] + rnd.randint(lo, hi), seqno[0], cb])

    def issue_mma(effect, commit_bar):
        """MMAs retire in order; the commit arrives after everything issued before it has retired"""
        t = max([clock[0]] + [m[0] for m in mma_chain]) + rnd.randint(1, 6)
        entry = [t, effect, commit_bar]
        mma_chain.append(entry)

    def issuer():
        def load(it):
            b = it & 1
            want = ((it >> 1) & 1) ^ 1
            if mutation == "free_parity":
                want ^= 1                                # (mutant) waits for the wrong phase
            while mutation != "no_free_wait" and not free_[b].done(want):
                yield
            assert smem[b]["state"] == "free" and tmem[b]["state"] == "free", f"load({it}) into a buffer still in use: {smem[b]}, {tmem[b]}"
            smem[b].update(item=it, state="loading")
            full[b].arrive_expect_tx(48)

            def landed(b=b, it=it):
                assert smem[b] == {"item": it, "state": "loading"}
                smem[b]["state"] = "qkv"
                full[b].complete_tx(48)
            later(landed)
        if n_items > 0:
            yield from load(0)
        for it in range(n_items):
            b, ph = it & 1, (it >> 1) & 1
            while not full[b].done(ph):
                yield
            assert smem[b] == {"item": it, "state": "qkv"} and tmem[b]["state"] == "free"
            tmem[b].update(item=it, state="s_pending")

            def qk_done(b=b, it=it):
                assert tmem[b] == {"item": it, "state": "s_pending"}
                tmem[b]["state"] = "s"
                readers_left[b] = 4
            issue_mma(qk_done, s_ready[b])
            if it + 1 < n_items:
                yield from load(it + 1)
            while mutation != "no_p_wait" and not p_ready[b].done(ph):
                yield
            assert smem[b] == {"item": it, "state": "p"}, f"PV({it}) issued before P is complete: {smem[b]}"
            assert tmem[b]["state"] == "s" and readers_left[b] == 0, "PV overwrites scores that are still being read"
            tmem[b]["state"] = "o_pending"

            def pv_done(b=b, it=it):
                tmem[b]["state"] = "o"
                readers_left[b] = 4
            issue_mma(pv_done, o_ready[b])

    def softmax_warp(w):
        for it in range(n_items):
            b, ph = it & 1, (it >> 1) & 1
            while not s_ready[b].done(ph):
                yield
            assert tmem[b] == {"item": it, "state": "s"}, f"warp {w} reads scores of item {it} but TMEM holds {tmem[b]}"
            for _ in range(rnd.randint(1, 4)):
                yield                                    # two passes over the scores, P written to smem (over Q|K)
            assert smem[b]["item"] == it and smem[b]["state"] in ("qkv", "p_partial")
            smem[b]["state"] = "p_partial"
            readers_left[b] -= 1
            if readers_left[b] == 0:
                smem[b]["state"] = "p"
            p_ready[b].arrive()
            while not o_ready[b].done(ph):
                yield
            assert tmem[b] == {"item": it, "state": "o"}, f"warp {w} reads the output of item {it} but TMEM holds {tmem[b]}"
            for _ in range(rnd.randint(1, 3)):
                yield
            readers_left[b] -= 1
            if readers_left[b] == 0:                     # last reader: buffer (smem + TMEM) is dead
                smem[b].update(item=None, state="free")
                tmem[b].update(item=None, state="free")
                done_items.append(it)
            free_[b].arrive()

    actors = [issuer()] + [softmax_warp(w) for w in range(4)]
    alive = list(actors)
    idle_rounds = 0
    while alive:
        clock[0] += 1
        # asynchronous completions that are due
        progressed = False
        for ev in sorted([e for e in async_q if e[0] <= clock[0]], key=lambda e: e[1]):
            async_q.remove(ev)
            ev[2]()
            progressed = True
        while mma_chain and mma_chain[0][0] <= clock[0]:
            _, effect, bar = mma_chain.pop(0)
            effect()
            bar.arrive()
            progressed = True
        a = rnd.choice(alive)
        try:
            before = (tuple(b.cur for bars in (full, s_ready, p_ready, o_ready, free_) for b in bars), len(done_items))
            next(a)
            after = (tuple(b.cur for bars in (full, s_ready, p_ready, o_ready, free_) for b in bars), len(done_items))
            progressed |= before != after
        except StopIteration:
            alive.remove(a)
            progressed = True
        idle_rounds = 0 if (progressed or async_q or mma_chain) else idle_rounds + 1
        assert idle_rounds < 2000, f"deadlock: items done {done_items}, smem {smem}, tmem {tmem}"
    assert done_items == list(range(n_items)), done_items
    assert all(s["state"] == "free" for s in smem) and all(t["state"] == "free" for t in tmem)
    return True


def test_attention_pipe_protocol_model():
    for n_items in (0, 1, 2, 3, 4, 7, 20):
        for seed in range(150):
            assert simulate(n_items, seed)


def test_the_model_detects_broken_protocols():
    """the model is only worth something if it fails on wrong protocols: three mutants of the kernel's waits"""
    import pytest
    for mutation in ("no_free_wait", "free_parity", "no_p_wait"):
        failures = 0
        for seed in range(40):
            try:
                simulate(7, seed, mutation)
            except AssertionError:
                failures += 1
        assert failures > 0, f"mutant {mutation} was not detected"


if __name__ == "__main__":
    test_attention_pipe_protocol_model()
    test_the_model_detects_broken_protocols()
    print("attention_pipe protocol model: ok")
