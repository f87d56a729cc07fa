"""CPU model of the decoupled look-back the one-launch radius-graph builder uses to turn per-CTA edge counts into CSR
offsets (cmd_gen_b200/csrc/graph.cu radius_rows_fused_kernel, phase B): every CTA publishes its aggregate, then walks
its predecessors a window at a time (thread t reads CTA j - t; the kernel's window is its 1024 threads), adds
aggregates up to and including the nearest predecessor whose INCLUSIVE prefix is already known, and publishes its own
inclusive prefix.  Whatever order the CTAs get to that point in, the result must be the exclusive prefix sum."""
import random

AGG, INCL = 1, 2


def lookback(status, c, W=32):
    """What warp 0 of CTA c computes; status[j] = (flag, value) or None (not yet published: the real kernel spins)."""
    prefix, j = 0, c - 1
    while True:
        window = []
        for lane in range(W):
            idx = j - lane
            if idx >= 0:
                assert status[idx] is not None, "would spin forever: a predecessor never published"
                window.append(status[idx])
            else:
                window.append((INCL, 0))                       # before CTA 0: inclusive prefix 0
        incl_lanes = [l for l, (f, _) in enumerate(window) if f == INCL]
        first = incl_lanes[0] if incl_lanes else W - 1
        prefix += sum(v for l, (_, v) in enumerate(window) if l <= first)
        if incl_lanes:
            return prefix
        j -= W


def run(totals, order, rng, W=32):
    n = len(totals)
    status = [None] * n
    # every CTA has published its aggregate before anyone it blocks can finish; CTA 0 publishes inclusive at once
    for c in range(n):
        status[c] = (INCL, totals[0]) if c == 0 else (AGG, totals[c])
    got = [None] * n
    got[0] = 0
    for c in order:
        if c == 0:
            continue
        got[c] = lookback(status, c, W)
        if rng.random() < 0.8:                                 # some CTAs are slow to publish their inclusive prefix
            status[c] = (INCL, got[c] + totals[c])
    return got


def test_lookback_gives_exclusive_prefix_in_any_order():
    rng = random.Random(3)
    for n in (1, 2, 31, 32, 33, 64, 65, 158, 500, 1024, 1025, 2500):
        totals = [rng.randrange(0, 1000) for _ in range(n)]
        want, acc = [], 0
        for t in totals:
            want.append(acc)
            acc += t
        for trial in range(6):
            order = list(range(n))
            if trial:
                rng.shuffle(order)
            assert run(totals, order, rng, W=32 if trial % 2 else 1024) == want, (n, trial)
