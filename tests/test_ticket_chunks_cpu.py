"""The decode order is dealt to the decode units in chunks of 1 << ticket_shift tickets (meshoptimizer_b200/csrc/
mob200_decoder.cuh: unit_block_count / unit_ticket; shift 2 = the run-major order of block mode + rounds).  This is
the arithmetic of those two functions: every ticket is taken by exactly one unit, in ascending order, a chunk never
straddles a metadata batch of 16 blocks, and the host's run-major order keeps every block behind its predecessor."""
import random


def unit_block_count(total, units, sh, unit):
    chunks = (total + (1 << sh) - 1) >> sh
    if unit >= chunks:
        return 0
    mine = (chunks - unit + units - 1) // units
    last = unit + (mine - 1) * units
    return (mine << sh) - (((chunks << sh) - total) if last == chunks - 1 else 0)


def unit_ticket(units, sh, unit, i):
    return ((((i >> sh) * units) + unit) << sh) + (i & ((1 << sh) - 1))


def test_every_ticket_once():
    rng = random.Random(5)
    cases = [(0, 1, 2), (1, 1, 2), (3, 740, 2), (4, 740, 2), (5, 2, 2), (2963, 740, 2), (2960 * 4, 740, 2), (2960 * 4 + 1, 740, 2)]
    cases += [(rng.randrange(0, 5000), rng.randrange(1, 800), rng.choice([0, 1, 2, 4])) for _ in range(300)]
    for total, units, sh in cases:
        seen = []
        for u in range(units):
            c = unit_block_count(total, units, sh, u)
            t = [unit_ticket(units, sh, u, i) for i in range(c)]
            assert t == sorted(t)
            for i, x in enumerate(t):  # members of a chunk sit in one batch of 16 consecutive positions of the unit
                assert (i >> sh) == ((i - (x & ((1 << sh) - 1))) >> sh) and (i // 16) == ((i - (x & ((1 << sh) - 1))) // 16)
            seen += t
        assert sorted(seen) == list(range(total)), (total, units, sh)


def run_major_order(nblocks, run=4):
    """host order (mob200_api.cu plan_create_body): streams sorted by block count, descending"""
    order = []
    live = len(nblocks)
    for b0 in range(0, nblocks[0] if nblocks else 0, run):
        while live and nblocks[live - 1] <= b0:
            live -= 1
        for s in range(live):
            order += [(s, b) for b in range(b0, min(b0 + run, nblocks[s]))]
    return order


def test_run_major_order_is_a_valid_decode_order():
    rng = random.Random(9)
    for _ in range(100):
        nb = sorted((rng.randrange(0, 40) for _ in range(rng.randrange(1, 30))), reverse=True)
        order = run_major_order(nb)
        assert sorted(order) == sorted((s, b) for s, n in enumerate(nb) for b in range(n))
        pos = {sb: i for i, sb in enumerate(order)}
        for (s, b), i in pos.items():
            if b:
                assert pos[(s, b - 1)] < i
