"""Offline check of the fused transformer-block kernel's table-driven weight ring (csrc/tblock.cu, TBLOCK_WIDE_FF).

The kernel's weight boxes travel through 16 KB slots: ring slots 4..8 in every phase and, in the FF chunk loop only, the
four AH boxes 0..3 as well.  The slot and use number of load i of a tile are a closed form (``place`` below restates it);
three producer warps issue the loads i % 3 in order, each after a PARITY wait on the release barrier of the slot's previous
use; one consumer (the MMA warp) takes the boxes strictly in sequence order and releases each slot on the barrier
[use & 1][slot]; the first use of an AH slot in a tile also waits for ``stage_free`` (the out-proj epilogue has read the u
tile out of the AH boxes).  A parity wait cannot tell "my phase completed" from "two phases ago completed", so the protocol
is only correct if no producer can run that far ahead.  This script

  * checks the static invariant: when a warp reaches load L, what its own earlier waits prove to be consumed covers the
    third-previous use of L's slot (two barriers per slot => aliasing needs a distance of four uses), and
  * runs a randomised discrete-event simulation of the real wait conditions (mbarrier phase counters, parity tests) over
    several tiles per CTA for the three load sequences (head / tail 0 / tail 1) and reports any overwrite of an unconsumed
    box, any consumption of the wrong box and any deadlock.

``python profiles/ring_protocol_sim.py`` prints one line per mode; tests/test_host_cpu.py runs a short version.
"""
import random

N_SLOTS = 9
AH_PER_TILE = 26  # AH-slot loads of one FF chunk loop: 6 full patterns of (4 AH + 5 ring) + 2


def loads_per_tile(mode):
    return {"head": 48, "tail0": 128, "tail1": 80}[mode]


def place(mode, t, i):
    """-> (slot, use) of load i of tile t: restates the closed form of the kernel's producer warps.  Ring slots 4..8 and AH
    slots 0..3 are two cyclic sub-rings with running counters; out-proj / FF1 chunks 0, 1 (i < 24), the QKV phase
    (i >= 80) and head mode use the ring only, the 56 loads of the FF chunk loop follow the pattern 4 x AH, 5 x ring."""
    if mode == "head":
        cnt = t * 48 + i
        return 4 + cnt % 5, cnt // 5
    r_tile = 102 if mode == "tail0" else 54
    if i < 24:
        rc = i
    elif i < 80:
        f = i - 24
        q9, m9 = divmod(f, 9)
        if m9 < 4:
            cnt = t * AH_PER_TILE + q9 * 4 + m9
            return cnt % 4, cnt // 4
        rc = 24 + q9 * 5 + (m9 - 4)
    else:
        rc = 54 + (i - 80)
    cnt = t * r_tile + rc
    return 4 + cnt % 5, cnt // 5


def needs_stage_free(mode, i):
    """The first four AH loads of a tile wait for the out-proj epilogue to be done with the u tile."""
    return mode != "head" and 24 <= i < 80 and (i - 24) < 4


def static_invariant(mode, tiles=4, producers=3):
    """Number of loads at which a producer warp could alias a parity wait (0 = safe): with two release barriers per slot
    (even / odd uses) a wait is ambiguous only if the THIRD-previous use of the slot may still be unreleased, so what the
    warp's own earlier waits prove to be consumed (the consumer works strictly in order) must cover that load."""
    P = loads_per_tile(mode)
    last, prev = {}, {}
    for t in range(tiles):
        for i in range(P):
            L = t * P + i
            sl = place(mode, t, i)[0]
            prev[L] = last.get(sl)
            last[sl] = L
    know = {w: -1 for w in range(producers)}
    bad = 0
    for t in range(tiles):
        for i in range(P):
            L, w = t * P + i, i % producers
            p = L
            for _ in range(3):
                p = prev.get(p) if p is not None else None
            kn = know[w]
            if needs_stage_free(mode, i):  # every out-proj box (loads < 16) of this tile is consumed
                kn = max(kn, t * P + 15)
            if p is not None and kn < p:
                bad += 1
            if prev[L] is not None:
                kn = max(kn, prev[L])
            know[w] = kn
    return bad


def simulate(mode, tiles=3, seed=0, producers=3, max_steps=2_000_000):
    rnd = random.Random(seed)
    P = loads_per_tile(mode)
    full_ph = [0] * N_SLOTS
    emp_ph = [[0] * N_SLOTS, [0] * N_SLOTS]
    stage_ph = 0
    content = [None] * N_SLOTS
    cons = (0, 0)
    prod = [[0, w] for w in range(producers)]
    inflight = []

    def done(count, parity):  # mbarrier.try_wait.parity: true iff the phase of that parity has completed
        return (count & 1) != parity

    for _ in range(max_steps):
        if cons[0] >= tiles:
            return "ok"
        acts = []
        for w in range(producers):
            t, i = prod[w]
            while i >= P:
                t, i = t + 1, i - P
            prod[w] = [t, i]
            if t >= tiles:
                continue
            s, use = place(mode, t, i)
            if needs_stage_free(mode, i) and not done(stage_ph, t & 1):
                continue
            if use > 0 and not done(emp_ph[(use - 1) & 1][s], ((use - 1) >> 1) & 1):
                continue
            acts.append(("issue", w, t, i, s))
        acts += [("land",) + L for L in inflight]
        t, i = cons
        s, use = place(mode, t, i)
        if done(full_ph[s], use & 1):
            acts.append(("consume", t, i, s, use))
        if not acts:
            return f"deadlock at {cons}"
        a = rnd.choice(acts)
        if a[0] == "issue":
            _, w, t, i, s = a
            if content[s] is not None:
                return f"overwrite of slot {s} holding {content[s]} by load {(t, i)}"
            content[s] = ("pending", t, i)
            inflight.append((t, i, s))
            prod[w] = [t, i + producers]
        elif a[0] == "land":
            _, t, i, s = a
            inflight.remove((t, i, s))
            content[s] = (t, i)
            full_ph[s] += 1
        else:
            _, t, i, s, use = a
            if content[s] != (t, i):
                return f"slot {s} holds {content[s]}, expected {(t, i)}"
            content[s] = None
            emp_ph[use & 1][s] += 1
            if mode != "head" and i == 15:  # the out-proj has retired: its epilogue follows and hands the AH boxes over
                stage_ph += 1
            i += 1
            cons = (t + 1, 0) if i == P else (t, i)
    return "no progress"


if __name__ == "__main__":
    for mode in ("head", "tail0", "tail1"):
        res = {simulate(mode, 4, seed) for seed in range(40)}
        print(f"{mode}: loads per tile {loads_per_tile(mode)}, static alias risks {static_invariant(mode)}, "
              f"simulation {sorted(res)}")
