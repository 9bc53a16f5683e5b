"""N > 1 host logic on CPU: two gloo ranks shard a table into block ranges, exchange partial records with all_gather and
merge them (matrixone_b200/shard.py).  The per-rank partials are produced by the CPU oracle here -- the GPU kernels are
covered by the -m gpu tests; this test pins the sharding / packing / rank-ordered merge that bench.py runs under NCCL."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_ROWS = 100_000
NQ, K = 6, 5


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as O
    from matrixone_b200 import datagen, shard
    r0, r1 = shard.block_range(rank, world, N_ROWS)
    cols = {k: np.ascontiguousarray(v) for k, v in datagen.lineitem(10, r0, r1 - r0).items()}
    P = datagen.q6_params()
    s, ns, _ = O.q6(cols, r1 - r0, P)
    g1 = O.q1(cols, r1 - r0, datagen.Q1_CUTOFF)
    rng = np.random.default_rng(1)
    ds = rng.standard_normal((400, 8)).astype(np.float32); qs = rng.standard_normal((NQ, 8)).astype(np.float32)
    lo, hi = rank * 200, (rank + 1) * 200
    keys, dists = O.bruteforce(ds[lo:hi], qs, K)
    keys = np.where(keys >= 0, keys + lo, keys)

    def gather(payload):
        send = torch.frombuffer(bytearray(payload), dtype=torch.uint8)
        buf = torch.zeros(len(payload) * world, dtype=torch.uint8)
        dist.all_gather_into_tensor(buf, send)
        return buf.numpy().tobytes()

    q6 = shard.merge_q6(gather(shard.pack_q6(s, ns)), world)
    q1 = shard.merge_q1(gather(shard.pack_q1(g1, r0)), world)
    tk = shard.merge_topk_host(*shard.unpack_topk(gather(shard.pack_topk(keys, dists)), world, NQ, K), NQ, K)
    if rank == 0:
        torch.save({"q6": q6, "q1": q1, "topk": tk}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_exchange_merge(tmp_path):
    import oracle_lib as O
    from matrixone_b200 import datagen, shard
    world = 2
    assert shard.block_range(0, 2, N_ROWS) == (0, 57344) and shard.block_range(1, 2, N_ROWS) == (57344, N_ROWS)   # whole 8192-row blocks
    assert shard.block_range(3, 8, 600_037_902)[0] % 8192 == 0
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(world, 29613, out), nprocs=world, join=True)
    res = torch.load(out, weights_only=False)
    cols = datagen.lineitem(10, 0, N_ROWS)
    P = datagen.q6_params()
    s, ns, _ = O.q6(cols, N_ROWS, P)
    assert res["q6"][1] == ns and abs(res["q6"][0] - s) <= 1e-12 * abs(s) and res["q6"][2] is False
    want = O.q1(cols, N_ROWS, datagen.Q1_CUTOFF)
    assert [(g["returnflag"], g["linestatus"], g["count_order"], g["first_row"]) for g in res["q1"]] == \
           [(g["returnflag"], g["linestatus"], g["count_order"], g["first_row"]) for g in want]
    for g, w in zip(res["q1"], want):
        for k in ("sum_qty", "sum_base_price", "sum_disc_price", "sum_charge", "avg_qty", "avg_price", "avg_disc"):
            assert abs(g[k] - w[k]) <= 1e-12 * abs(w[k])
    rng = np.random.default_rng(1)
    ds = rng.standard_normal((400, 8)).astype(np.float32); qs = rng.standard_normal((NQ, 8)).astype(np.float32)
    gk, gd = O.bruteforce(ds, qs, K)
    assert (res["topk"][1] == gd).all() and (res["topk"][0] == gk).all()


def test_merge_records_edge_cases():
    from matrixone_b200 import shard
    buf = shard.pack_q6(0.0, 0) + shard.pack_q6(5.5, 3) + shard.pack_q6(0.0, 0)
    assert shard.merge_q6(buf, 3) == (5.5, 3, False)
    assert shard.merge_q6(shard.pack_q6(0.0, 0) * 2, 2) == (0.0, 0, True)          # every partial NULL => SUM is NULL
    k = np.asarray([[-1, -1, 7], [3, 5, 9]]); d = np.asarray([[0.0, 0.0, 0.5], [0.1, 0.2, 0.9]])
    mk, md = shard.merge_topk_host(k, d, 1, 3)
    assert list(mk) == [3, 5, 7] and list(md) == [0.1, 0.2, 0.5]
    mk, md = shard.merge_topk_host(np.asarray([[-1, -1, 7], [-1, -1, -1]]), np.asarray([[0, 0, 0.5], [0, 0, 0.0]]), 1, 3)
    assert list(mk) == [-1, -1, 7] and list(md) == [0.0, 0.0, 0.5]                 # front padding survives the merge
