"""CPU model of the in-kernel coordinate finish of the tcgen05 edge kernel (cmd_gen_b200/csrc/tc_edge.cu, COORD
instantiation): per 16-edge unit the epilogue sums trans = coord_diff * scalar over each CSR row run,
writes a row that lies inside the unit at once, and for a row that spans units leaves one partial per unit in
cpart[2 * unit + slot] and bumps cticket[row]; the unit that arrives LAST adds the pieces in unit order.  The model runs
the units in random arrival orders with integer messages (so every order must give the same exact result) and checks
the three invariants the kernel relies on: every row is written exactly once, a (unit, slot) cell is written by at
most one row, and every ticket is back at zero afterwards (the buffers are reused by the next launch without a reset)."""
import numpy as np
import pytest

UNIT = 16


def finish_units(rowptr, n_rows, vals, order, x0):
    E = int(rowptr[n_rows])
    n_units = (E + UNIT - 1) // UNIT
    erow = np.repeat(np.arange(n_rows), np.diff(rowptr[:n_rows + 1]))
    cpart = np.full((2 * n_units + 2,), np.iinfo(np.int64).min, dtype=np.int64)     # poison: a read of an unwritten cell shows
    written_by = {}
    ticket = np.zeros(n_rows, dtype=np.int64)
    x_next = np.full(n_rows, np.iinfo(np.int64).min, dtype=np.int64)
    writes = np.zeros(n_rows, dtype=np.int64)
    for un in order:
        u0 = un * UNIT
        lanes = [e for e in range(u0, min(u0 + UNIT, E))]
        # in-order sums over the lanes of the same row (what the 16-step shuffle loop computes on the run's last lane)
        for li, e in enumerate(lanes):
            r = erow[e]
            last_of_run = li == len(lanes) - 1 or erow[lanes[li + 1]] != r
            if not last_of_run:
                continue
            s = sum(int(vals[k]) for k in lanes[:li + 1] if erow[k] == r)
            rs, re = int(rowptr[r]), int(rowptr[r + 1])
            if rs >= u0 and re <= u0 + UNIT:                                       # the whole row lies in this unit
                x_next[r] = x0[r] + s
                writes[r] += 1
                continue
            fu, lu = rs // UNIT, (re - 1) // UNIT
            cell = 2 * un + (1 if un == fu else 0)
            assert cell not in written_by, f"cell {cell} written by rows {written_by[cell]} and {r}"
            written_by[cell] = r
            cpart[cell] = s
            old = ticket[r]
            ticket[r] += 1
            if old == lu - fu:                                                     # every other unit of the row has arrived
                tot = 0
                for t in range(fu, lu + 1):
                    piece = cpart[2 * t + (1 if t == fu else 0)]
                    assert piece != np.iinfo(np.int64).min, "piece read before it was written"
                    tot += int(piece)
                x_next[r] = x0[r] + tot
                writes[r] += 1
                ticket[r] = 0
    return x_next, writes, ticket


@pytest.mark.parametrize("seed", range(12))
def test_unit_tickets_finish_every_row_once_in_any_arrival_order(seed):
    rng = np.random.default_rng(seed)
    n_rows = int(rng.integers(1, 60))
    # degrees from 1 (self loop only) to 70 (a row over five units); a few rows aligned to unit boundaries on purpose
    deg = rng.choice([1, 2, 5, 11, 16, 17, 32, 33, 70], size=n_rows, p=[.1, .1, .2, .3, .1, .05, .05, .05, .05])
    if seed % 3 == 0:
        deg[0] = UNIT                                                              # row 0 fills unit 0 exactly
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    E = int(rowptr[-1])
    vals = rng.integers(-1000, 1000, size=E)
    x0 = rng.integers(-50, 50, size=n_rows)
    expect = x0 + np.array([vals[rowptr[r]:rowptr[r + 1]].sum() for r in range(n_rows)])
    n_units = (E + UNIT - 1) // UNIT
    for trial in range(6):
        order = rng.permutation(n_units) if trial else np.arange(n_units)
        x_next, writes, ticket = finish_units(rowptr, n_rows, vals, order, x0)
        assert np.array_equal(writes, np.ones(n_rows, dtype=np.int64))
        assert np.array_equal(x_next, expect)
        assert not ticket.any()


def test_only_the_first_rows_move():
    """Pocket conditioning: the coordinate kernel covers the E_p = rowptr[N_p] edges of the phar rows only — the row after
    the last moving one never contributes a unit, and the last unit is ragged."""
    rng = np.random.default_rng(99)
    deg = rng.integers(1, 30, size=40)
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    n_moving = 17
    vals = rng.integers(-9, 9, size=int(rowptr[-1]))
    x0 = np.zeros(40, dtype=np.int64)
    x_next, writes, ticket = finish_units(rowptr, n_moving, vals, rng.permutation((int(rowptr[n_moving]) + UNIT - 1) // UNIT), x0)
    assert np.array_equal(x_next[:n_moving], [vals[rowptr[r]:rowptr[r + 1]].sum() for r in range(n_moving)])
    assert np.array_equal(writes, np.ones(n_moving, dtype=np.int64)) and not ticket.any()
