"""Host-side planning of the tiled kernel, checked on the CPU (pure arithmetic in libtsgemm_b200.so, no device needed):
the unit decomposition with tail balancing must cover every (row tile, column) exactly once, and the progress groups of
the multi-GPU mode 2 must account for every arrival the kernel will make."""
import ctypes as C

import pytest

import __graft_entry__ as ge

SHAPES = [(4096, 4096), (64, 512), (8192, 14336), (16384, 16384), (129, 257), (33, 29), (1000, 100), (4096, 4096 * 8), (256, 4096), (32, 40)]


@pytest.fixture(scope="module")
def L():
    t = ge.load()
    lib = t.lib()
    lib.tsg_plan_units.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int * 5)]
    lib.tsg_plan_unit_at.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int * 3)]
    lib.tsg_plan_progress.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int * 9), C.POINTER(C.c_uint * 8)]
    return lib


def plan(L, M, N, sms):
    out = (C.c_int * 5)()
    assert L.tsg_plan_units(M, N, sms, C.byref(out)) == 0
    return list(out)


@pytest.mark.parametrize("sms", [148, 132, 1, 7])
@pytest.mark.parametrize("shape", SHAPES, ids=[f"M{m}N{n}" for m, n in SHAPES])
def test_units_cover_every_tile_exactly_once(L, shape, sms):
    M, N = shape
    mtiles, ntiles, units_full, sub, total = plan(L, M, N, sms)
    assert mtiles == -(-M // 128) and ntiles == -(-N // 256)
    assert units_full % sms == 0 and units_full <= mtiles * ntiles and sub in (1, 2, 4, 8)
    assert total == units_full + (mtiles * ntiles - units_full) * sub
    covered = {}
    o = (C.c_int * 3)()
    for u in range(total):
        assert L.tsg_plan_unit_at(M, N, sms, u, C.byref(o)) == 0
        mt, n0, cw = list(o)
        assert 0 <= mt < mtiles and cw in (16, 8, 4, 2) and n0 % 32 == 0 and n0 % (16 * cw) == 0
        assert (cw == 16) == (u < units_full or sub == 1)
        for c in range(n0, n0 + 16 * cw, 32):  # 32-column granules
            assert (mt, c) not in covered, f"unit {u} overlaps unit {covered.get((mt, c))}"
            covered[(mt, c)] = u
    assert len(covered) == mtiles * ntiles * 8  # every 32-column granule of every 256-column tile of every row tile
    # the tail is never longer than one round of full units would have been
    tail_units = total - units_full
    assert -(-tail_units // sms) / sub <= -(-(mtiles * ntiles - units_full) // sms) + 1e-9


@pytest.mark.parametrize("shape", SHAPES, ids=[f"M{m}N{n}" for m, n in SHAPES])
def test_progress_groups_account_for_every_arrival(L, shape):
    M, N = shape
    sms = 148
    mtiles, ntiles, units_full, sub, total = plan(L, M, N, sms)
    ng, gb, tg = C.c_int(), (C.c_int * 9)(), (C.c_uint * 8)()
    assert L.tsg_plan_progress(M, N, sms, C.byref(ng), C.byref(gb), C.byref(tg)) == 0
    g = ng.value
    bounds = list(gb)[: g + 1]
    assert 1 <= g <= 8 and bounds[0] == 0 and bounds[-1] == mtiles and all(b1 > b0 for b0, b1 in zip(bounds, bounds[1:]))
    assert sum(list(tg)[:g]) == total * 16  # one arrival per compute warp per unit
    # recount independently from the unit list
    per_group = [0] * g
    o = (C.c_int * 3)()
    for u in range(total):
        L.tsg_plan_unit_at(M, N, sms, u, C.byref(o))
        grp = max(i for i in range(g) if o[0] >= bounds[i])
        per_group[grp] += 16
    assert per_group == list(tg)[:g]
    # group sizes shrink towards the end once there are enough row tiles
    if mtiles >= 32:
        sizes = [b1 - b0 for b0, b1 in zip(bounds, bounds[1:])]
        assert sizes[0] >= sizes[-1] and sizes[-1] <= max(1, mtiles // 16)


def test_plan_rejects_bad_arguments(L):
    out = (C.c_int * 5)()
    assert L.tsg_plan_units(0, 10, 148, C.byref(out)) != 0
    o = (C.c_int * 3)()
    assert L.tsg_plan_unit_at(128, 256, 148, 8, C.byref(o)) != 0  # one tile, fewer tiles than SMs: cut into 8 units (0..7)


@pytest.mark.parametrize("sms", [148, 132, 7])
@pytest.mark.parametrize("units", [1, 32, 74, 75, 148, 149, 222, 223, 512, 1024, 3000])
@pytest.mark.parametrize("bpw", [1, 2, 16])
def test_bcsr_ring_plan_never_lengthens_the_last_round(L, units, sms, bpw):
    """gemm_bcsr_ring.cu: the left-over units of the last round are dealt as two column slices only when that shortens it."""
    L.tsg_dbg_bcsr_ring_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.tsg_dbg_bcsr_ring_plan.restype = None
    fu, sub = C.c_int(), C.c_int()
    L.tsg_dbg_bcsr_ring_plan(units, sms, bpw, C.byref(fu), C.byref(sub))
    fu, sub = fu.value, sub.value
    assert sub in (1, 2) and 0 <= fu <= units
    if sub == 1:
        assert fu == units
    else:
        assert bpw >= 2 and fu % sms == 0 and units - fu < sms and (units - fu) * 2 <= sms  # the slices fit ONE round
    rounds_plain = -(-units // sms)
    rounds_split = fu // sms + (-(-((units - fu) * sub) // sms)) / sub if sub == 2 else rounds_plain
    assert rounds_split <= rounds_plain
    if bpw >= 2 and 0 < units % sms <= sms // 2:
        assert sub == 2 and rounds_split == rounds_plain - 0.5


@pytest.mark.parametrize("M", [1, 31, 32, 100, 256, 257, 384, 511, 512, 600, 1024, 2048, 4096, 5000, 8192, 16384, 65536, 100000, 200001])
def test_host_slab_schedule_covers_all_rows(L, M):
    """staging.cu: the pipelined host-pointer GEMM cuts M into a ramp of slabs (128 rows at both ends, up to 1024+ in the middle)."""
    out = (C.c_int * 72)()
    L.tsg_dbg_host_slabs.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int * 72)]
    n = L.tsg_dbg_host_slabs(M, 0, C.byref(out))
    v = list(out)[:n]
    assert 1 <= n <= 72 and sum(v) == M and all(x > 0 for x in v)
    if M > 256:
        assert all(x >= 128 for x in v), v                       # no sliver slabs
        assert v[0] <= 255 and v[-1] <= 255                      # short head and tail
        assert sum(1 for x in v if x % 128) <= 1                 # only the ragged rest is not a whole number of row tiles
    else:
        assert v == [M]
    # the forced uniform schedule (TSG_HOST_SLAB_ROWS)
    n = L.tsg_dbg_host_slabs(M, 256, C.byref(out))
    u = list(out)[:n]
    assert sum(u) == M and all(x > 0 for x in u) and n <= 32
