"""Host-side sharding and partial-result merge for multi-GPU runs (one process per GPU).

Mirrors the reference's data-parallel split: a scan is cut into disjoint block ranges, one pipeline each
(pkg/sql/compile/scope.go:521-568 buildScanParallelRun), partial aggregate states travel to a merge scope
(pkg/sql/colexec/group/mergeGroup.go:132-247 MergeGroup, colexec/mergetop) -- here the exchange is one NCCL
all_gather of a fixed-size record per rank and every rank merges the records in RANK ORDER (deterministic).
Pure host logic: packs / unpacks / merges small records; no column data is touched here.
"""
import numpy as np

BLOCK_ROWS = 8192  # objectio.BlockMaxRows, pkg/objectio/const.go:26

Q6_REC_BYTES = 16
Q1_REC_BYTES = 8 * 8 * 8   # 8 groups x 8 float64 words


def block_range(rank, world, n_rows):
    """contiguous range of 8192-row blocks for `rank` (strong-scaling split of one table)"""
    nblocks = (n_rows + BLOCK_ROWS - 1) // BLOCK_ROWS
    per = (nblocks + world - 1) // world
    r0 = min(n_rows, rank * per * BLOCK_ROWS)
    r1 = min(n_rows, (rank + 1) * per * BLOCK_ROWS)
    return r0, r1


# ---------------------------------------------------------------------------------------------- Q6: (sum, count)
def pack_q6(sum_, count):
    return np.asarray([sum_], dtype=np.float64).tobytes() + np.asarray([count], dtype=np.int64).tobytes()


def merge_q6(buf, world):
    """buf: world * 16 bytes.  SUM partials are added in rank order; an empty partial (count 0) is NULL and skipped
    (sumAvgExec.BatchMerge, sumavg2.go:222-236)."""
    b = np.frombuffer(bytes(buf), dtype=np.uint8).reshape(world, Q6_REC_BYTES)
    total, count, isnull = 0.0, 0, True
    for r in range(world):
        s = b[r, :8].copy().view(np.float64)[0]
        c = int(b[r, 8:].copy().view(np.int64)[0])
        if c == 0:
            continue
        total = s if isnull else total + s
        isnull = False
        count += c
    return float(total), count, isnull


# ---------------------------------------------------------------------------------------------- Q1: grouped partials
_Q1_FIELDS = ("sum_qty", "sum_base_price", "sum_disc_price", "sum_charge", "sum_disc")


def pack_q1(groups, row_offset=0):
    out = np.zeros(64, dtype=np.float64)
    for i, g in enumerate(groups[:8]):
        out[i * 8:(i + 1) * 8] = [g["returnflag"] + 256 * g["linestatus"] + 65536, g["sum_qty"], g["sum_base_price"], g["sum_disc_price"],
                                  g["sum_charge"], g["sum_disc"], g["count_order"], g["first_row"] + row_offset]
    return out.tobytes()


def merge_q1(buf, world):
    """re-hash partial group keys and BatchMerge in rank order; groups come back in global first-seen (row) order"""
    b = np.frombuffer(bytes(buf), dtype=np.float64).reshape(world, 8, 8)
    acc = {}
    for r in range(world):
        for i in range(8):
            rec = b[r, i]
            if rec[0] < 65536 or rec[6] == 0:
                continue
            key = int(rec[0]) - 65536
            a = acc.get(key)
            if a is None:
                acc[key] = {"returnflag": key & 0xFF, "linestatus": key >> 8, "count_order": int(rec[6]), "first_row": int(rec[7]),
                            **{f: float(rec[1 + j]) for j, f in enumerate(_Q1_FIELDS)}}
            else:
                for j, f in enumerate(_Q1_FIELDS):
                    a[f] = a[f] + float(rec[1 + j])
                a["count_order"] += int(rec[6])
                a["first_row"] = min(a["first_row"], int(rec[7]))
    out = sorted(acc.values(), key=lambda g: g["first_row"])
    for g in out:
        c = float(g["count_order"])
        g["avg_qty"], g["avg_price"], g["avg_disc"] = g["sum_qty"] / c, g["sum_base_price"] / c, g["sum_disc"] / c
    return out


# ---------------------------------------------------------------------------------------------- top-k
def pack_topk(keys, dists):
    return np.ascontiguousarray(keys, dtype=np.int64).tobytes() + np.ascontiguousarray(dists, dtype=np.float64).tobytes()


def unpack_topk(buf, world, nq, k):
    per = nq * k * 16
    b = np.frombuffer(bytes(buf), dtype=np.uint8).reshape(world, per)
    keys = np.stack([b[r, :nq * k * 8].copy().view(np.int64) for r in range(world)])
    dists = np.stack([b[r, nq * k * 8:].copy().view(np.float64) for r in range(world)])
    return keys, dists


def merge_topk_host(keys, dists, nq, k):
    """reference k-way merge (used by the CPU test to check the device merge kernel's contract): ascending (dist, key),
    -1 keys are padding; fewer than k results pad (-1, 0) at the FRONT like brute_force.go:319-331."""
    world = keys.shape[0]
    ok = np.full((nq, k), -1, dtype=np.int64)
    od = np.zeros((nq, k), dtype=np.float64)
    kk = keys.reshape(world, nq, k)
    dd = dists.reshape(world, nq, k)
    for q in range(nq):
        cand = [(dd[r, q, j], kk[r, q, j]) for r in range(world) for j in range(k) if kk[r, q, j] >= 0]
        cand.sort()
        cand = cand[:k]
        pad = k - len(cand)
        for j, (d, key) in enumerate(cand):
            ok[q, pad + j] = key
            od[q, pad + j] = d
    return ok.reshape(-1), od.reshape(-1)
