"""Pins the join part of the oracle: og_join_sels against the reference's own GroupSels tests (pkg/vm/message/group_sels_test.go:47-118, transcribed
below), og_join_find / og_join_probe against a nested-loop join written straight from the join definitions.  No GPU."""
import numpy as np

import oracle_lib as O


def _sels(groups, group_count):
    g = np.asarray(groups, dtype=np.uint64)
    offsets = np.zeros(group_count + 2, np.int32); vals = np.zeros(max(len(g), 1), np.int32)
    m = O.go().og_join_sels(O.p(g), len(g), group_count, O.p(offsets), O.p(vals))
    return offsets, vals[:m]


def _get(offsets, vals, k):      # GroupSels.Get, joinMapMsg.go:127-132
    if k + 1 >= len(offsets):
        return []
    return list(vals[offsets[k]:offsets[k + 1]])


def test_group_sels_reference_cases():
    # TestGroupSels_Normal0Based (:47-61): Insert(0,10) Insert(1,11) Insert(0,12) Insert(1,13); Finalize(2, 4)
    groups = np.zeros(14, np.uint64); groups[[10, 12]] = 1; groups[[11, 13]] = 2
    off, vals = _sels(groups, 2)
    assert _get(off, vals, 0) == [10, 12] and _get(off, vals, 1) == [11, 13] and _get(off, vals, 2) == []
    # TestGroupSels_Dedup1Based (:63-77): Insert(1,0) Insert(2,1) Insert(1,2); Finalize(2, 3): keys 1..groupCount -> the + 2 in the offsets length
    off, vals = _sels([2, 3, 2], 2)       # ids are key + 1 here
    assert _get(off, vals, 1) == [0, 2] and _get(off, vals, 2) == [1] and _get(off, vals, 0) == []
    # TestGroupSels_NullsSkipped (:101-118): rows 0, 1, 3 -> groups 0, 1, 2; row 2 NULL
    off, vals = _sels([1, 2, 0, 3], 3)
    assert _get(off, vals, 0) == [0] and _get(off, vals, 1) == [1] and _get(off, vals, 2) == [3]
    assert list(off) == [0, 1, 2, 3, 3]


def test_find_and_probe_against_a_nested_loop_join():
    rng = np.random.default_rng(0)
    for trial in range(30):
        nb, npr = int(rng.integers(0, 40)), int(rng.integers(0, 40))
        build = rng.integers(0, 8, nb).astype(np.uint64); probe = rng.integers(0, 10, npr).astype(np.uint64)
        bnull = rng.random(nb) < 0.2; pnull = rng.random(npr) < 0.2
        # IntHashMap insert: first-seen ids, NULL keys get no group
        table, ids = [], np.zeros(nb, np.uint64)
        for i in range(nb):
            if bnull[i]:
                continue
            if build[i] not in table:
                table.append(build[i])
            ids[i] = table.index(build[i]) + 1
        table = np.array(table, np.uint64)
        off, sels = _sels(ids, len(table))
        vals = np.zeros(npr, np.uint64)
        pn = np.zeros((npr + 63) // 64 + 1, np.uint64)
        for i in range(npr):
            if pnull[i]:
                pn[i >> 6] |= np.uint64(1) << np.uint64(i & 63)
        O.go().og_join_find(O.p(table), len(table), O.p(probe), O.p(pn), npr, O.p(vals))
        for jt in range(4):
            want = []
            for i in range(npr):
                m = [] if pnull[i] else [j for j in range(nb) if not bnull[j] and build[j] == probe[i]]
                assert (vals[i] == 0) == (len(m) == 0)
                if jt == 0:
                    want += [(i, j) for j in m]
                elif jt == 1:
                    want += [(i, j) for j in m] if m else [(i, -1)]
                elif jt == 2:
                    want += [(i, -1)] if m else []
                else:
                    want += [] if m else [(i, -1)]
            op, ob = np.zeros(2000, np.int64), np.zeros(2000, np.int64)
            r = O.go().og_join_probe(O.p(vals), npr, O.p(off), O.p(sels), jt, O.p(op), O.p(ob), 2000)
            assert [(int(a), int(b)) for a, b in zip(op[:r], ob[:r])] == want, (trial, jt)
