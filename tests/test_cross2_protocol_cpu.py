"""Protocol model of attn_cross2_kernel (consistentid_b200/csrc/attn_cross2.cuh): the producer, the MMA issuer and the three softmax
warpgroups restated as coroutines over mbarrier objects with the hardware's phase-parity semantics, run under randomised schedules and
randomised completion delays of the asynchronous engines (TMA loads, tcgen05 commits complete later, in order).

It checks what the GPU tests cannot enumerate: for any unit range / tiles-per-(sample, head) / ring depth the protocol never deadlocks,
every S = Q K^T reads the Q slot and the K/V slot that hold ITS unit / group, every P.V still finds its V in place, no TMEM buffer is
overwritten before its reader is done - and it documents why the host only selects the kernel when a (sample, head) run has >= 2 query
tiles (with one tile per run the S stream, three units ahead, needs a third K/V group in the two-slot ring: the model deadlocks)."""
import random

import pytest

NB = 3          # TMEM buffers = softmax warpgroups


class Bar:
    def __init__(self):
        self.done = 0               # completed phases

    def passes(self, parity):       # mbarrier.try_wait.parity: true once the phase with this parity has completed
        return (self.done & 1) != parity


def simulate(g_beg, g_end, tiles, NQ, seed, alias=False, max_steps=200000):
    rng = random.Random(seed)
    q_full, q_free = [Bar() for _ in range(NQ)], [Bar() for _ in range(NQ)]
    kv_full, kv_free = [Bar(), Bar()], [Bar(), Bar()]
    s_full, p_ready, o_full, o_free = ([Bar() for _ in range(NB)] for _ in range(4))
    q_slot, kv_slot = [None] * NQ, [None, None]            # what the shared-memory slots hold (unit / K-V group)
    buf = [dict(S=None, P=None, O=None, O_drained=True) for _ in range(NB)]
    tma, mma = [], []                                       # in-order completion queues of the asynchronous engines
    group0 = g_beg // tiles

    def grp(g):
        return g // tiles - group0 if (g_beg % tiles == 0) else (g // tiles - group0)

    def producer():
        n_kv, uq = -1, 0
        for g in range(g_beg, g_end):
            t = g % tiles
            if g == g_beg or t == 0:
                n_kv += 1
                s = n_kv & 1
                yield ("wait", kv_free[s], ((n_kv >> 1) & 1) ^ 1)
                kv_slot[s] = ("loading", g // tiles)
                tma.append(("kv", s, g // tiles))
            slot = uq % NQ
            yield ("wait", q_free[slot], ((uq // NQ) & 1) ^ 1)
            q_slot[slot] = ("loading", g)
            tma.append(("q", slot, g))
            uq += 1

    def mma_warp():
        n = g_end - g_beg

        class Cur:
            def __init__(s):
                s.g, s.t, s.nkv = g_beg, g_beg % tiles, 0

            def advance(s):
                s.g += 1
                s.t += 1
                if s.t == tiles:
                    s.t = 0
                if s.t == 0:
                    s.nkv += 1

        def issue_S(cu, u):
            if cu.g == g_beg or cu.t == 0:
                yield ("wait", kv_full[cu.nkv & 1], (cu.nkv >> 1) & 1)
            kvs, slot, j = cu.nkv & 1, u % NQ, u % NB
            yield ("wait", q_full[slot], (u // NQ) & 1)
            if alias:
                yield ("wait", o_free[j], ((u // NB) & 1) ^ 1)
            assert q_slot[slot] == ("ready", cu.g), (q_slot[slot], cu.g)
            assert kv_slot[kvs] == ("ready", cu.g // tiles), (kv_slot[kvs], cu.g)
            b = buf[j]
            assert b["S"] is None and b["P"] is None, f"S({cu.g}) over live scores / probabilities of buffer {j}: {b}"
            if alias:
                assert b["O_drained"], f"S({cu.g}) over an undrained accumulator"
            b["S"] = cu.g
            mma.append(("s_full", j))
            mma.append(("q_free", slot))

        def issue_PV(cu, x):
            j, par = x % NB, (x // NB) & 1
            yield ("wait", p_ready[j], par)
            if not alias:
                yield ("wait", o_free[j], par ^ 1)
            b = buf[j]
            assert b["P"] == cu.g, (b, cu.g)
            assert kv_slot[cu.nkv & 1] == ("ready", cu.g // tiles), "V_cat of the unit's group was overwritten before its P.V"
            assert b["O_drained"], f"P.V({cu.g}) over an undrained accumulator"
            b["P"], b["O"], b["O_drained"] = None, cu.g, False       # in-order tensor pipe: P is consumed before any later S of this buffer
            mma.append(("o_full", j))
            if cu.g + 1 == g_end or cu.t + 1 == tiles:
                mma.append(("kv_free", cu.nkv & 1))

        cs, cp, us = Cur(), Cur(), 0
        while us < n and us < NB:
            yield from issue_S(cs, us)
            cs.advance()
            us += 1
        for x in range(n):
            yield from issue_PV(cp, x)
            cp.advance()
            if us < n:
                yield from issue_S(cs, us)
                cs.advance()
                us += 1

    def softmax_wg(w):
        k = 0
        for g in range(g_beg + w, g_end, NB):
            par = k & 1
            yield ("wait", s_full[w], par)
            b = buf[w]
            assert b["S"] == g, (b, g)
            b["S"], b["P"] = None, g                       # scores in registers, probabilities written over them
            p_ready[w].done += 1
            yield ("wait", o_full[w], par)
            assert b["O"] == g, (b, g)
            b["O"], b["O_drained"] = None, True
            o_free[w].done += 1
            k += 1

    agents = {"producer": producer(), "mma": mma_warp(), **{f"wg{w}": softmax_wg(w) for w in range(NB)}}
    blocked = {}
    for _ in range(max_steps):
        runnable = [a for a in agents if a not in blocked or blocked[a][0].passes(blocked[a][1])]
        choices = runnable + (["tma"] if tma else []) + (["mma_done"] if mma else [])
        if not choices:
            if not agents:
                return True
            return False                                    # deadlock
        pick = rng.choice(choices)
        if pick == "tma":
            kind, slot, what = tma.pop(0)
            if kind == "kv":
                kv_slot[slot] = ("ready", what); kv_full[slot].done += 1
            else:
                q_slot[slot] = ("ready", what); q_full[slot].done += 1
            continue
        if pick == "mma_done":
            kind, idx = mma.pop(0)
            {"s_full": s_full, "q_free": q_free, "o_full": o_full, "kv_free": kv_free}[kind][idx].done += 1
            continue
        blocked.pop(pick, None)
        try:
            ev = next(agents[pick])
            blocked[pick] = (ev[1], ev[2])
        except StopIteration:
            del agents[pick]
    raise AssertionError("simulation did not terminate")


@pytest.mark.parametrize("tiles", [2, 3, 5, 8, 32])
@pytest.mark.parametrize("alias", [False, True])
def test_protocol_completes_and_keeps_operands_in_place(tiles, alias):
    rng = random.Random(tiles * 7 + alias)
    for trial in range(40):
        g_beg = rng.randrange(0, 4 * tiles)
        n = rng.choice([1, 2, 3, 4, 7, 9, 14, 27, 28])
        NQ = 4 if alias else 8                              # Cross2Cfg::NQ: 8 slots for one 64-wide head-dim chunk, 4 for two (d = 80, the aliased layout)
        assert simulate(g_beg, g_beg + n, tiles, NQ, seed=trial, alias=alias), (g_beg, n, tiles, alias)


def test_one_tile_per_head_needs_a_third_kv_slot():
    """Why cid_attn_cross only selects the kernel for N > 128: with one query tile per (sample, head) every unit is its own K/V group, the S
    stream runs three groups ahead of the P.V stream and the two-slot K/V ring deadlocks (the producer waits for a P.V the issuer has not
    reached because it is itself waiting for that K/V)."""
    assert not simulate(0, 9, tiles=1, NQ=8, seed=0)
