"""Host model of the operand ring of the fused SAGE layer (buckgnn_b200/csrc/gemm_tc.cuh: producer, gather warps, MMA
issuer over a 4-slot ring with parity-waited mbarriers) under random interleavings.

The rule it pins: a parity wait can only tell the barrier's current phase from the previous one, so every agent that
waits on a slot's `empty` barrier must TAKE PART in every use of the slot (wait, then arrive on `full`).
  * "skip": the first fused-layer build let the gather warps wait only before the slots they fill (the aggregate K
    blocks); between two of those a slot is used twice by the producer alone, the gather warps -- when fast -- read
    the parity of the first skipped phase as theirs and overwrote a tile the tensor core had not consumed (GPU
    symptom: correct first launch, corrupt or trapped later launches);
  * "watch": the first fix made them wait on every use but arrive only on their own (the producer arriving twice on
    the others).  Safe while they poll often enough, but nothing waits for them there: a gather warp held up for four
    slot times finds the barrier two phases ahead and waits forever.  The model deadlocks it in most schedules;
  * "participate" (what the kernel does now): their arrival is required on every use; safe under any schedule."""
import random

import pytest

K_STAGES, AGG_KB, X_KB = 4, 8, 8
PER_TILE = AGG_KB + X_KB


class Barrier:
    """mbarrier with an arrival count; try_wait(parity) is true iff the phase with that parity has completed, i.e. the
    barrier is currently in a phase of the OTHER parity (what `mbarrier.try_wait.parity` reports)."""

    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier expects in one phase"
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase + 1

    def try_wait(self, parity):
        return (self.phase & 1) != parity


def simulate(n_tiles, mode, seed, gather_bias=0.5):
    """Runs producer / gather / MMA as state machines stepped in random order; returns the list of protocol
    violations (a slot written while its previous contents were still unread, or consumed with the wrong contents)."""
    rng = random.Random(seed)
    full = [Barrier(2) for _ in range(K_STAGES)]       # producer + gather (or the producer standing in) per use
    empty = [Barrier(1) for _ in range(K_STAGES)]
    slot_a = [None] * K_STAGES                          # which use's A tile the slot holds
    slot_b = [None] * K_STAGES
    last_consumed = [s_ - K_STAGES for s_ in range(K_STAGES)]     # the use of each slot the MMA read last
    bad = []
    total = n_tiles * PER_TILE
    pos = {"producer": 0, "gather": 0, "mma": 0}
    assert mode in ("skip", "watch", "participate")
    # the uses the gather warps act on, in order
    gather_uses = [u for u in range(total) if (u % PER_TILE) < AGG_KB or mode != "skip"]

    def write(slot, which, use):
        if last_consumed[slot] != use - K_STAGES:       # the slot's previous use has not been read yet (or was skipped)
            bad.append(f"use {use}: {which} of slot {slot} written while the MMA has only consumed use {last_consumed[slot]} of it")
        if which == "a":
            slot_a[slot] = use
        else:
            slot_b[slot] = use

    def step(agent):
        if agent == "producer":
            u = pos[agent]
            if u >= total:
                return False
            s, parity = u % K_STAGES, (u // K_STAGES) & 1
            if not empty[s].try_wait(parity ^ 1):
                return False
            is_agg = (u % PER_TILE) < AGG_KB
            write(s, "b", u)
            if not is_agg:
                write(s, "a", u)                        # TMA loads the root rows
                if mode != "participate":
                    full[s].arrive()                    # ... and the producer stands in for the gather warps
            full[s].arrive()
            pos[agent] += 1
            return True
        if agent == "gather":
            i = pos[agent]
            if i >= len(gather_uses):
                return False
            u = gather_uses[i]
            s, parity = u % K_STAGES, (u // K_STAGES) & 1
            if not empty[s].try_wait(parity ^ 1):
                return False
            if (u % PER_TILE) < AGG_KB:
                write(s, "a", u)
                full[s].arrive()
            elif mode == "participate":
                full[s].arrive()
            pos[agent] += 1
            return True
        u = pos[agent]                                  # MMA issuer
        if u >= total:
            return False
        s, parity = u % K_STAGES, (u // K_STAGES) & 1
        if not full[s].try_wait(parity):
            return False
        if slot_a[s] != u or slot_b[s] != u:
            bad.append(f"use {u}: consumed slot {s} holding A of {slot_a[s]}, B of {slot_b[s]}")
        last_consumed[s] = u
        empty[s].arrive()                               # tcgen05.commit
        pos[agent] += 1
        return True

    idle = 0
    while pos["mma"] < total and idle < 10000 and len(bad) < 5:
        r = rng.random()
        agent = "gather" if r < gather_bias else ("producer" if r < gather_bias + (1 - gather_bias) / 2 else "mma")
        try:
            progressed = step(agent)
        except AssertionError as e:                     # a barrier saw an arrival that belongs to another phase
            bad.append(str(e))
            break
        idle = 0 if progressed else idle + 1
    if pos["mma"] < total and not bad:
        bad.append(f"deadlock at use {pos['mma']}")
    return bad


@pytest.mark.parametrize("seed", range(20))
def test_ring_is_safe_when_every_waiter_takes_part_in_every_use(seed):
    assert simulate(6, "participate", seed, gather_bias=0.05 + 0.045 * seed) == []      # slow ... eager gather warps


def test_skipping_the_producer_only_phases_aliases_the_parity():
    """the round-2 bug, reproduced: eager gather warps that wait only on their own uses corrupt the ring"""
    outcomes = [simulate(6, "skip", seed, gather_bias=0.8) for seed in range(20)]
    assert any(o and "deadlock" not in o[0] for o in outcomes)


def test_watching_without_arriving_can_fall_two_phases_behind():
    outcomes = [simulate(6, "watch", seed, gather_bias=0.1) for seed in range(20)]
    assert any(o and "deadlock" in o[0] for o in outcomes)
