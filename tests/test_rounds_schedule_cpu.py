"""Model check of the rounds form's schedule (meshoptimizer_b200/csrc/mob200_decoder.cuh).

A unit's producer groups up to four consecutive blocks of its queue into a round; the unit's decoders start a round
when every member has been staged, publish a member's block aggregate after unpacking it, and finish the member when
its carry (look-back over the stream's previous blocks, decoded by other units) has been resolved.  The model below
replays those dependencies as a fix-point over events and answers one question: does every block finish?

  * staging all members of a round BEFORE resolving any of their carries (what the kernel does) always finishes;
  * resolving each member's carry right after its own copy (the first version) closes a cycle between units as soon
    as a carry can depend on a block that another unit decodes in a round of the same age -- e.g. one long stream.

CPU only; no CUDA, no oracle.
"""
import itertools

import pytest

ROUND = 4     # kRoundBlocks
SLOTS = 8     # Lay<true>::kSlots
BATCH = 16    # kProducerBatch


def finishes(n_streams: int, blocks_per_stream: int, units: int, stage_first: bool, run_major: bool = False) -> bool:
    total = n_streams * blocks_per_stream
    units = min(units, total)
    if not run_major:
        # level-major decode order: ticket t = block (t // n_streams) of stream (t % n_streams); unit u takes u, u + units, ...
        queue = {u: list(range(u, total, units)) for u in range(units)}
        pred = {t: (t - n_streams if t >= n_streams else None) for t in range(total)}
        block_of = {t: (t % n_streams, t // n_streams) for t in range(total)}
    else:
        # block mode + rounds (section 5.6): runs of ROUND consecutive blocks of a stream, level-major over the runs; the
        # order is dealt to the units in chunks of ROUND tickets (chunk c -> unit c % units)
        order = [(s, b) for b0 in range(0, blocks_per_stream, ROUND) for s in range(n_streams) for b in range(b0, min(b0 + ROUND, blocks_per_stream))]
        ticket = {sb: t for t, sb in enumerate(order)}
        block_of = dict(enumerate(order))
        pred = {t: (ticket[(s, b - 1)] if b else None) for t, (s, b) in block_of.items()}
        chunks = [list(range(c, min(c + ROUND, total))) for c in range(0, total, ROUND)]
        units = min(units, len(chunks))
        queue = {u: [t for c in chunks[u::units] for t in c] for u in range(units)}
    # rounds: consecutive members of a queue, never across a producer batch
    rounds = {}   # ticket -> (unit, round index, position in round, members of the round)
    unit_rounds = {}
    for u, q in queue.items():
        rs = []
        for b0 in range(0, len(q), BATCH):
            batch = q[b0 : b0 + BATCH]
            rs += [batch[k : k + ROUND] for k in range(0, len(batch), ROUND)]
        unit_rounds[u] = rs
        for ri, members in enumerate(rs):
            for pos, t in enumerate(members):
                rounds[t] = (u, ri, pos, members)
    index_in_queue = {t: i for q in queue.values() for i, t in enumerate(q)}

    staged, unpacked, carried, done = set(), set(), set(), set()

    def step() -> bool:
        progress = False
        for t in range(total):
            u, ri, pos, members = rounds[t]
            prev_round = unit_rounds[u][ri - 1] if ri > 0 else []
            i = index_in_queue[t]
            if t not in staged:
                ok = all(m in staged for m in members[:pos])                       # the producer works in order
                ok = ok and all(m in carried for m in prev_round)                  # ... and finished the previous round
                ok = ok and (i < SLOTS or queue[u][i - SLOTS] in done)             # the slot is free again
                if not stage_first:
                    ok = ok and all(m in carried for m in members[:pos])           # first version: copy g waits for carry g-1
                if ok:
                    staged.add(t); progress = True
            if t not in unpacked:
                ok = all(m in staged for m in members)                             # the decoders wait for every member
                ok = ok and all(m in done for m in prev_round)                     # ... after finishing the previous round
                if ok:
                    unpacked.add(t); progress = True
            chained = run_major and len(members) > 1 and all(
                block_of[m][0] == block_of[members[0]][0] and block_of[m][1] == block_of[members[0]][1] + k for k, m in enumerate(members))
            if t not in carried and chained and pos > 0:
                # chained round: the decoder warps hand the first member's carry on through the members' aggregates
                if members[0] in carried and all(m in unpacked for m in members[:pos]):
                    carried.add(t); progress = True
            elif t not in carried:
                ok = t in staged and all(m in carried for m in members[:pos])
                if stage_first:
                    ok = ok and all(m in staged for m in members)
                ok = ok and (pred[t] is None or pred[t] in unpacked)               # look-back needs the predecessor's aggregate
                if ok:
                    carried.add(t); progress = True
            if t not in done:
                if all(m in unpacked for m in members) and all(m in carried for m in members):
                    done.add(t); progress = True
        return progress

    while step():
        pass
    return len(done) == total


CASES = [(1, 64, 5), (1, 200, 7), (2, 40, 5), (3, 33, 4), (7, 12, 5), (9, 9, 6), (16, 8, 5), (40, 5, 6), (5, 50, 12)]


@pytest.mark.parametrize("n_streams,blocks,units", CASES)
def test_staging_before_carries_always_finishes(n_streams, blocks, units):
    assert finishes(n_streams, blocks, units, stage_first=True)


def test_sweep_small_shapes():
    for n_streams, blocks, units in itertools.product(range(1, 9), (1, 2, 5, 9, 17, 30), (1, 2, 3, 5, 8)):
        assert finishes(n_streams, blocks, units, stage_first=True), (n_streams, blocks, units)


def test_carry_after_each_copy_can_deadlock():
    """the schedule of the first version: one long stream, more than two rounds per unit"""
    assert not finishes(1, 64, 5, stage_first=False)
    assert not finishes(7, 12, 5, stage_first=False)


def test_run_major_chained_rounds_finish():
    """block mode + rounds: run-major order, carries chained inside a round by the decoder warps"""
    for n_streams, blocks, units in CASES + [(1, 7, 3), (3, 6, 2), (5, 3, 4), (2, 17, 16)]:
        assert finishes(n_streams, blocks, units, stage_first=True, run_major=True), (n_streams, blocks, units)
    for n_streams, blocks, units in itertools.product(range(1, 7), (1, 2, 5, 9, 17), (1, 2, 3, 5, 8)):
        assert finishes(n_streams, blocks, units, stage_first=True, run_major=True), (n_streams, blocks, units)
