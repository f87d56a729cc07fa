"""CPU model of the tcgen05 edge kernel's atomic-free segmented sum (cmd_gen_b200/csrc/common.cuh "segmented sum",
graph.cu edge_dst / row_agg_src, tc_edge.cu epilogue, tc_node.cu stage_agg): the bookkeeping the graph builder writes
(edst per edge, agg_src per node), replayed by a model of the kernel's walk over its units, must let a consumer
rebuild every row's sum — for both work splits (contiguous lanes / per-unit), any degree sequence, any edge count.
Integer "messages" make the check exact.  The GPU tests check the CUDA code against the oracle; this file pins the
SCHEME, so that a change of one of its three parties without the others fails without a GPU."""
import numpy as np
import pytest

UNIT = 16
AGG_EMPTY = -2 ** 31


def lane_first_unit(l, U, L):            # common.cuh
    return l * U // L


def lane_of_unit(u, U, L):               # common.cuh
    return ((u + 1) * L - 1) // U


def seg_lane_of(u, U, L):                # graph.cu (L == 0: a lane is a unit)
    return lane_of_unit(u, U, L) if L else u


def seg_lane_first(l, U, L):
    return lane_first_unit(l, U, L) if L else l


def bookkeeping(rowptr, N, L):
    """graph.cu fill pass: edst[E], agg_src[N]."""
    E = int(rowptr[-1])
    U = (E + UNIT - 1) // UNIT
    edst = np.full(E, -1, dtype=np.int64)
    agg_src = np.zeros(N, dtype=np.int64)
    for row in range(N):
        rs, re = int(rowptr[row]), int(rowptr[row + 1])
        if re <= rs:
            agg_src[row] = AGG_EMPTY
            continue
        lf, ll = seg_lane_of(rs // UNIT, U, L), seg_lane_of((re - 1) // UNIT, U, L)
        if lf == ll:
            agg_src[row] = row
            edst[re - 1] = row
            continue
        first_start = seg_lane_first(lf, U, L) * UNIT
        agg_src[row] = -(1 + ((((lf << 10) | (ll - lf)) << 1) | (0 if rs <= first_start else 1)))
        for pos in range(rs, re):
            l = seg_lane_of(pos // UNIT, U, L)
            lane_start, lane_end = seg_lane_first(l, U, L) * UNIT, seg_lane_first(l + 1, U, L) * UNIT
            if pos == re - 1 or pos == lane_end - 1:
                edst[pos] = N + 2 * l + (0 if rs <= lane_start else 1)
    return edst, agg_src, U


def kernel_walk(values, edst, E, U, L, n_ctas):
    """tc_edge.cu epilogue: every group walks its units in order with a carried running sum and stores where edst says."""
    buf = {}
    if L:                                                   # lanes: group g of CTA c owns lane 4 c + g
        walks = [range(lane_first_unit(l, U, L), lane_first_unit(l + 1, U, L)) for l in range(L)]
    else:                                                   # units: tile t -> CTA t mod n_ctas, group g owns unit 4 t + g
        walks = [[u] for u in range(U)]                     # the sum never carries across units there (unit ends flush)
    for units in walks:
        s = 0
        for u in units:
            for e in range(u * UNIT, min((u + 1) * UNIT, E)):
                s += int(values[e])
                if edst[e] >= 0:
                    assert edst[e] not in buf, "two stores to one destination row"
                    buf[int(edst[e])] = s
                    s = 0
        assert s == 0, "a running sum was never stored"
    return buf


def consumer(buf, agg_src, N, U, L):
    """tc_node.cu stage_agg / common.cuh agg_load4."""
    out = np.zeros(N, dtype=np.int64)
    for row in range(N):
        code = int(agg_src[row])
        if code == AGG_EMPTY:
            continue
        if code >= 0:
            out[row] = buf[code]
            continue
        k = -(code + 1)
        lf, extra, slot = k >> 11, (k >> 1) & 1023, k & 1
        for i in range(extra + 1):
            idx = N + 2 * (lf + i) + (slot if i == 0 else 0)
            if idx not in buf:
                # only a lane that owns no unit (fewer units than lanes) may be silent: its partial rows are cleared
                # by the graph builder before every denoiser call (graph.cu launch_build_edges), i.e. read as zero
                assert L and lane_first_unit(lf + i, U, L) == lane_first_unit(lf + i + 1, U, L), "missing partial row of a lane that owns units"
                continue
            out[row] += buf[idx]
    return out


@pytest.mark.parametrize("L,n_ctas", [(0, 148), (592, 148), (8, 2), (12, 3)])
@pytest.mark.parametrize("seed,N,max_deg", [(0, 200, 12), (1, 50, 90), (2, 400, 3), (3, 7, 700), (4, 1, 1), (5, 300, 40)])
def test_bookkeeping_lets_the_consumer_rebuild_every_row(L, n_ctas, seed, N, max_deg):
    rng = np.random.default_rng(seed)
    deg = rng.integers(0, max_deg + 1, size=N)
    deg[rng.integers(0, N)] = max_deg                       # at least one long row
    rowptr = np.concatenate([[0], np.cumsum(deg)])
    E = int(rowptr[-1])
    values = rng.integers(-1000, 1000, size=E)
    edst, agg_src, U = bookkeeping(rowptr, N, L)
    if L and U:                                             # lanes cover every unit exactly once, in order
        bounds = [lane_first_unit(l, U, L) for l in range(L + 1)]
        assert bounds[0] == 0 and bounds[-1] == U and all(b1 >= b0 for b0, b1 in zip(bounds, bounds[1:]))
        assert all(lane_first_unit(lane_of_unit(u, U, L), U, L) <= u < lane_first_unit(lane_of_unit(u, U, L) + 1, U, L) for u in range(U))
    buf = kernel_walk(values, edst, E, U, L, n_ctas)
    got = consumer(buf, agg_src, N, U, L)
    want = np.array([values[rowptr[r]:rowptr[r + 1]].sum() for r in range(N)], dtype=np.int64)
    assert np.array_equal(got, want)
    whole = int((agg_src >= 0).sum())
    if L:                                                   # lanes: all but at most L - 1 rows are stored whole
        assert N - whole - int((agg_src == AGG_EMPTY).sum()) <= max(L - 1, 0)
