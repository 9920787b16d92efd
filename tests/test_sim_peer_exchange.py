"""Randomised model of the peer-memory exchange protocol (csrc/peer.cu + parallel.PeerExchange), checked on the CPU.

Each rank is a coroutine that runs its steps in stream order -- store my embeddings into slot `rank` of buffer `par` on
every peer, publish the step number, wait for everybody's flag, read all slots (local search), store block g of the
candidates into slot `rank` on rank g, publish, wait, read (merge) -- while a scheduler interleaves the ranks with random,
strongly skewed speeds.  The model asserts what the double-buffering argument in parallel.py claims: a slot is never
overwritten before its reader is done with the previous contents, a reader never sees data of another step, nobody
deadlocks.  (With this strict every-step dependency chain even a single buffer passes the model; the second buffer is
margin.)  Mutants that must be caught: flags that do not carry the step number, and a merge that does not wait.
"""
import random


def simulate(G, steps, seed, nbuf=2, flags_carry_step=True, wait_for_candidates=True):
    rnd = random.Random(seed)
    speed = [rnd.choice([1, 1, 2, 5, 20]) for _ in range(G)]            # relative slowness of every rank
    # buffers[rank][kind][par][slot] = step whose data is stored there (None = never written / being written)
    buf = [{k: [[None] * G for _ in range(nbuf)] for k in ("q", "c")} for _ in range(G)]
    reading = [{k: [None] * nbuf for k in ("q", "c")} for _ in range(G)]   # step a rank is currently reading from (kind, par)
    flag = [{k: [0] * G for k in ("q", "c")} for _ in range(G)]           # flag[rank][kind][src] = last published step
    finished = [0] * G

    def store(dst, kind, par, slot, step):
        r = reading[dst][kind][par]
        assert r is None, f"rank {slot} overwrites {kind}[{par}] of rank {dst} (step {step}) while rank {dst} reads step {r}"
        old = buf[dst][kind][par][slot]
        assert old is None or finished_reading[dst][kind] >= old, (
            f"rank {slot} overwrites {kind}[{par}][{slot}] of rank {dst} holding step {old}, which rank {dst} has not consumed yet")
        buf[dst][kind][par][slot] = step

    finished_reading = [{"q": 0, "c": 0} for _ in range(G)]

    def rank_program(r):
        for step in range(1, steps + 1):
            par = step % nbuf
            published = step if flags_carry_step else 1
            # (1) embeddings -> every peer, then the flags
            for p in rnd.sample(range(G), G):
                store(p, "q", par, r, step)
                yield
            for p in range(G):
                flag[p]["q"][r] = published if flags_carry_step else flag[p]["q"][r] + 0 or 1
            yield
            # (2) wait for all sources
            while not all(flag[r]["q"][g] >= published for g in range(G)):
                yield
            # (3) local search reads all slots
            reading[r]["q"][par] = step
            for _ in range(rnd.randint(1, 4)):
                assert all(buf[r]["q"][par][g] == step for g in range(G)), (
                    f"rank {r} searches step {step} but its query buffer holds {buf[r]['q'][par]}")
                yield
            reading[r]["q"][par] = None
            finished_reading[r]["q"] = step
            # (4) candidates: block g -> rank g
            for p in rnd.sample(range(G), G):
                store(p, "c", par, r, step)
                yield
            for p in range(G):
                flag[p]["c"][r] = published if flags_carry_step else 1
            yield
            while wait_for_candidates and not all(flag[r]["c"][g] >= published for g in range(G)):
                yield
            reading[r]["c"][par] = step
            for _ in range(rnd.randint(1, 3)):
                assert all(buf[r]["c"][par][g] == step for g in range(G)), (
                    f"rank {r} merges step {step} but its candidate buffer holds {buf[r]['c'][par]}")
                yield
            reading[r]["c"][par] = None
            finished_reading[r]["c"] = step
            finished[r] = step

    progs = {r: rank_program(r) for r in range(G)}
    idle = 0
    while progs:
        r = rnd.choice(list(progs))
        if rnd.randint(1, speed[r]) != 1:           # slow ranks are scheduled less often
            idle += 1
            assert idle < 2_000_000, f"deadlock: finished steps {finished}"
            continue
        try:
            before = list(finished)
            next(progs[r])
            idle = 0 if before != finished else idle + 1
        except StopIteration:
            del progs[r]
            idle = 0
        assert idle < 2_000_000, f"deadlock: finished steps {finished}"
    assert finished == [steps] * G
    return True


def test_peer_exchange_protocol_model():
    for G in (2, 3, 8):
        for seed in range(60):
            assert simulate(G, steps=9, seed=seed)


def test_the_model_detects_broken_protocols():
    caught_single, caught_noseq = 0, 0
    for seed in range(40):
        try:
            simulate(3, steps=9, seed=seed, wait_for_candidates=False)   # merge without waiting: reads stale candidates
        except AssertionError:
            caught_single += 1
        try:
            simulate(3, steps=9, seed=seed, flags_carry_step=False)  # boolean flags: a stale flag satisfies a later wait
        except AssertionError:
            caught_noseq += 1
    assert caught_single > 0 and caught_noseq > 0, (caught_single, caught_noseq)


if __name__ == "__main__":
    test_peer_exchange_protocol_model()
    test_the_model_detects_broken_protocols()
    print("peer exchange protocol model: ok")
