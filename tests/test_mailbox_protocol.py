"""CPU model check of the peer-memory mailbox protocol of the interface exchange (sem_b200/csrc/sem_comm.cu).

The device code cannot run here, but its safety argument is pure protocol logic and is checked by simulation: every rank
owns a mailbox with, per side, TWO line slots (epoch parity) and one `arrived` epoch flag written by the neighbour; an
exchange is  push (store the line into the neighbour's slot[e & 1], then release arrived = e)  followed, possibly much
later and after other work, by  wait (acquire arrived >= e) + read slot[e & 1].  The claim: because a rank finishes exchange e
(needs the neighbour's push e) before it starts exchange e + 1, a neighbour can be at most ONE exchange ahead, so two
slots suffice and a slot is never overwritten before it has been read.  The simulation interleaves the ranks' micro-steps
at random (including the worst case of a rank racing ahead as far as the protocol lets it) and checks every payload."""
import random

import pytest


class Rank:
    def __init__(self, r, world):
        self.r, self.world = r, world
        self.nb = [n for n in (r - 1, r + 1) if 0 <= n < world]
        # mailbox: per sending neighbour two slots and the arrived flag
        self.slot = {n: [None, None] for n in self.nb}
        self.arrived = {n: 0 for n in self.nb}
        self.sent = 0          # device-resident epoch counters of the real code
        self.consumed = 0
        self.pc = 0            # micro program counter inside the current exchange
        self.max_lead = 0


def run(world, exchanges, seed, bias=None):
    rng = random.Random(seed)
    ranks = [Rank(r, world) for r in range(world)]
    done = [False] * world
    steps = 0
    while not all(done):
        steps += 1
        assert steps < 10_000_000, "deadlock"
        # bias: let one rank run whenever it can (worst-case skew), else uniform choice
        order = list(range(world))
        rng.shuffle(order)
        if bias is not None and rng.random() < 0.9:
            order.remove(bias)
            order.insert(0, bias)
        progressed = False
        for r in order:
            k = ranks[r]
            if done[r]:
                continue
            e = k.sent + 1 if k.pc == 0 else k.sent       # epoch of the exchange in flight
            if k.pc == 0:                                  # push kernel: one micro-step per neighbour store, then the flags
                for n in k.nb:
                    peer = ranks[n]
                    # the slot about to be overwritten must have been consumed by the peer (or never used)
                    old = peer.slot[r][e & 1]
                    assert old is None or old[1] <= peer.consumed, f"rank {r} overwrites unread slot of rank {n}: {old}, consumed {peer.consumed}"
                    peer.slot[r][e & 1] = (r, e)
                for n in k.nb:
                    ranks[n].arrived[r] = e                # release after the data
                k.sent = e
                k.pc = 1
                progressed = True
                break
            # wait + add kernel: blocks until every neighbour's flag has reached this epoch
            if all(k.arrived[n] >= e for n in k.nb):
                for n in k.nb:
                    k.max_lead = max(k.max_lead, k.arrived[n] - e)
                    assert k.slot[n][e & 1] == (n, e), f"rank {r} reads {k.slot[n][e & 1]} instead of ({n}, {e})"
                k.consumed = e
                k.pc = 0
                if e == exchanges:
                    done[r] = True
                progressed = True
                break
        assert progressed, "no rank can make progress: deadlock"
    return max(k.max_lead for k in ranks)


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_two_slots_suffice_under_any_interleaving(world):
    lead = 0
    for seed in range(40):
        lead = max(lead, run(world, 25, seed))
        for bias in range(world):
            lead = max(lead, run(world, 25, 1000 * seed + bias, bias=bias))
    # a neighbour is never more than one exchange ahead of the rank that reads its slot
    assert lead <= 1


def test_single_slot_would_not_be_enough():
    """Why there are two slots: the schedule  r0 push 1, r1 push 1, r0 read 1, r0 push 2  is legal (rank 0 has everything it
    needs for exchange 2) while rank 1 has not read epoch 1 yet -- with a single slot per side rank 0 would overwrite it."""
    a, b = Rank(0, 2), Rank(1, 2)
    b.slot[0][1] = (0, 1); b.arrived[0] = 1; a.sent = 1            # rank 0 pushes epoch 1 (parity 1)
    a.slot[1][1] = (1, 1); a.arrived[1] = 1; b.sent = 1            # rank 1 pushes epoch 1, does not read yet
    assert a.slot[1][1] == (1, 1); a.consumed = 1                  # rank 0 reads epoch 1
    unread = b.slot[0][1]
    assert unread == (0, 1) and unread[1] > b.consumed             # rank 1 still needs this line ...
    b.slot[0][0] = (0, 2); b.arrived[0] = 2; a.sent = 2            # ... and rank 0 pushes epoch 2: parity 0, another slot
    assert b.slot[0][1] == (0, 1)                                  # untouched; with one slot it would now hold (0, 2)


# ---------------------------------------------------------------------------------------------------------------------
# The one-launch partitioned apply (sem_march3_kernel<.., XCH = true>, sem_capi.cu::fused_apply): the exchange happens INSIDE
# the operator kernel, per strip.  Model: a launch is a list of one-warp CTAs dispatched in order into `slots` resident
# places -- first the edge CTAs (left edge strips, then right edge strips), then interior CTAs that only do work.  An edge CTA
# of strip s on side d: march (a few work steps), PUSH its segment into the neighbour's slot[d'][s][e & 1] and release the
# neighbour's arrived[d'][s] = e, then WAIT for its own arrived[d][s] >= e, read slot[d][s][e & 1], set epoch[d][s] = e and
# retire.  A waiting CTA keeps its slot.  Launch e + 1 of a rank starts when every CTA of its launch e has retired (stream
# order).  Claims checked under random interleavings: no deadlock when the 2 x strips edge CTAs fit the resident slots; every
# read returns the neighbour's segment of the SAME epoch; a slot is never overwritten before it was read (two parities suffice).
class FusedRank:
    def __init__(self, r, world, strips):
        self.r = r
        self.sides = [d for d, n in ((0, r - 1), (1, r + 1)) if 0 <= n < world]     # 0: left line, 1: right line
        self.nb = {0: r - 1, 1: r + 1}
        self.slot = {d: [[None, None] for _ in range(strips)] for d in self.sides}
        self.arrived = {d: [0] * strips for d in self.sides}
        self.epoch = {d: [0] * strips for d in self.sides}
        self.launch = 0
        self.pending, self.resident = [], []


def run_fused(world, strips, slots, interior, launches, seed, bias=None, other_path=None):
    rng = random.Random(seed)
    ranks = [FusedRank(r, world, strips) for r in range(world)]

    def start_launch(k):
        k.launch += 1
        k.pending = [["edge", d, s, rng.randint(1, 3), "march"] for d in k.sides for s in range(strips)]
        k.pending += [["interior", None, None, rng.randint(2, 6), "march"] for _ in range(interior)]
        k.resident = []

    for k in ranks:
        start_launch(k)
    steps, lead = 0, 0
    while any(k.launch <= launches for k in ranks):
        steps += 1
        assert steps < 5_000_000, "deadlock"
        order = [r for r in range(world) if ranks[r].launch <= launches]
        rng.shuffle(order)
        if bias is not None and bias in order and rng.random() < 0.9:
            order.remove(bias)
            order.insert(0, bias)
        progressed = False
        for r in order:
            k = ranks[r]
            while k.pending and len(k.resident) < slots:          # the hardware scheduler: in launch order
                k.resident.append(k.pending.pop(0))
            if not k.resident:                                     # launch complete
                if k.launch == launches:
                    k.launch += 1                                  # done
                else:
                    start_launch(k)
                progressed = True
                break
            cands = list(range(len(k.resident)))
            rng.shuffle(cands)
            for i in cands:
                cta = k.resident[i]
                kind, d, s, work, pc = cta
                e = k.launch
                if pc == "march":
                    cta[3] -= 1
                    if cta[3] <= 0:
                        cta[4] = "push" if kind == "edge" else "done"
                elif pc == "push" and r == other_path:
                    cta[4] = "wait"                                # this rank is on the other protocol: writes other flags
                elif pc == "push":
                    peer, dp = ranks[k.nb[d]], 1 - d
                    old = peer.slot[dp][s][e & 1]
                    assert old is None or old[1] <= peer.epoch[dp][s], f"rank {r} overwrites an unread segment: {old}"
                    peer.slot[dp][s][e & 1] = (r, e, s)
                    peer.arrived[dp][s] = e
                    cta[4] = "wait"
                elif pc == "wait":
                    if k.arrived[d][s] < e:
                        continue                                   # spinning: holds its slot, try another CTA
                    lead = max(lead, k.arrived[d][s] - e)
                    assert k.slot[d][s][e & 1] == (k.nb[d], e, s), (r, d, s, e, k.slot[d][s])
                    k.epoch[d][s] = e
                    cta[4] = "done"
                if cta[4] == "done":
                    k.resident.pop(i)
                progressed = True
                break
            if progressed:
                break
        assert progressed, "no CTA of any rank can make progress: deadlock"
    return lead


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_in_kernel_exchange_protocol(world):
    lead = 0
    for seed in range(12):
        # edge CTAs co-resident (what fused_apply_possible requires: 4 x strips <= resident slots)
        lead = max(lead, run_fused(world, strips=5, slots=20, interior=30, launches=12, seed=seed))
        for bias in (0, world - 1, world // 2):
            lead = max(lead, run_fused(world, strips=5, slots=20, interior=30, launches=12, seed=100 * seed + bias, bias=bias))
        # the tightest co-resident case: exactly the edge CTAs of a middle rank fit
        lead = max(lead, run_fused(world, strips=4, slots=8, interior=10, launches=8, seed=7000 + seed, bias=seed % world))
    assert lead <= 1


def test_in_kernel_exchange_needs_the_same_path_on_both_sides():
    """The world-3 bug of round 2: the choice between the one-launch exchange and the exchange kernels used per-rank numbers, so
    a rank with a narrow slab took the other path (other mailbox region, other flags) and both sides spun on flags the
    neighbour never writes.  In the model a rank that does not write the one-launch flags starves its neighbours' waiting
    edge CTAs: no CTA can make progress.  (The real kernel ends in its 120 s trap; the choice now only uses numbers all ranks
    agreed on at attach time, sem_capi.cu::fused_apply_possible.)"""
    with pytest.raises(AssertionError, match="deadlock"):
        run_fused(3, strips=3, slots=12, interior=4, launches=2, seed=1, other_path=1)
