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
