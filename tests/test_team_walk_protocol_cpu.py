"""The hand-over protocol of the team walk (mob200_walk_team.cuh) as a dependency model: helper warps build the windows of a
stream into a ring of four slots, the chain warp waits for the windows it needs and releases the ones it has left behind.
Random walks (incl. runs of literal channels that jump over several windows) must never dead-lock, a slot must never be
overwritten while the chain may still read it, and the chain must never read a window that is not built."""
import random

W, H = 4, 2  # kTeamWindows, kTeamHelpers


def simulate(n_windows, moves, release_first=True):
    full_done = [0] * W   # completed phases of full[slot]
    empty_done = [0] * W  # completed phases of empty[slot]
    slot_holds = [None] * W
    helper_next = list(range(H))  # next window of each helper
    have_until = released = 0
    pos = 0  # index into moves: (first, last) window needs of successive team_need calls
    state = "release"
    steps = 0
    while pos < len(moves):
        progressed = False
        # helpers
        for h in range(H):
            k = helper_next[h]
            if k >= n_windows:
                continue
            if k >= W and empty_done[k % W] < k // W:  # waits for the release of window k - W
                continue
            assert slot_holds[k % W] is None or slot_holds[k % W] < released, "overwrote a live window"
            slot_holds[k % W] = k
            full_done[k % W] += 1
            helper_next[h] += H
            progressed = True
        # chain: one team_need call, in the order of the device code
        first, last = moves[pos]
        first = min(first, n_windows)
        last = min(last, n_windows - 1)
        if release_first:
            while released < first and released < have_until:
                empty_done[released % W] += 1
                released += 1
                progressed = True
        if have_until <= last and have_until < n_windows:
            if full_done[have_until % W] > have_until // W:
                assert slot_holds[have_until % W] == have_until
                have_until += 1
                progressed = True
        else:
            if not release_first:
                while released < first and released < have_until:
                    empty_done[released % W] += 1
                    released += 1
            # the chain now reads windows first .. last
            for w in range(first, last + 1):
                assert slot_holds[w % W] == w, "read a window that is not in the ring"
            pos += 1
            progressed = True
        steps += 1
        if not progressed:
            return False
    return True


def random_moves(rng, n_windows):
    moves, w = [], 0
    while w < n_windows:
        span = rng.choice([0, 0, 1])
        moves.append((w, min(n_windows - 1, w + span)))
        w += rng.choice([0, 0, 1, 1, 2, 5, 9])  # channels inside a window, the next window, jumps over literal runs
    return moves


def test_protocol_never_deadlocks():
    rng = random.Random(7)
    for trial in range(300):
        n = rng.randint(1, 60)
        assert simulate(n, random_moves(rng, n)), trial


def test_waiting_before_releasing_can_deadlock():
    # (the order that was tried first: a jump over more windows than the ring holds stops both sides)
    assert not simulate(40, [(0, 1), (9, 10)], release_first=False)
