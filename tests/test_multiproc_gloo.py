"""N > 1 host logic on CPU: world_size-2 gloo run of the query partitioning + gather.  The lookup
itself is the CPU oracle here (no GPU in this container); on GPUs the same plumbing carries
VariantModel.find_variants_raw (see bench.py)."""
import os
import socket

import pytest
import torch.multiprocessing as mp

import workloads


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, queries, out_q):
    import torch.distributed as dist
    from analiticcl_b200 import parallel
    from oracle import orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = orc.OracleModel(alphabet_tsv=orc.TEST_ALPHABET_TSV)
    for w in ["rites", "tiers", "tires", "tries", "tyres", "rides", "brides", "dire", "huis", "huls", "think", "right"]:
        o.add_to_vocabulary(w)
    o.build()
    p = orc.make_params(max_anagram_distance=2, max_edit_distance=2, score_threshold=0.0, cutoff_threshold=0.0)
    res = parallel.lookup_partitioned(queries, lambda qs: o.find_variants_batch(qs, p, threads=1))
    ms, units = parallel.reduce_step_time(10.0 + rank, len(queries) // world)
    if rank == 0:
        out_q.put((res, ms, units))
    dist.barrier()
    dist.destroy_process_group()


def test_partition_is_a_partition():
    from analiticcl_b200.parallel import partition
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            parts = partition(n, w)
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def test_world_size_2_gloo():
    from oracle import orc
    queries = ["rite", "huys", "tink", "rihgt", "tyre", "bride", "dier", "x", "rides", "tiers", "thnik"]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, queries, q)) for r in range(2)]
    for p in procs:
        p.start()
    res, ms, units = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = orc.OracleModel(alphabet_tsv=orc.TEST_ALPHABET_TSV)
    for w in ["rites", "tiers", "tires", "tries", "tyres", "rides", "brides", "dire", "huis", "huls", "think", "right"]:
        o.add_to_vocabulary(w)
    o.build()
    exp = o.find_variants_batch(queries, orc.make_params(max_anagram_distance=2, max_edit_distance=2, score_threshold=0.0,
                                                         cutoff_threshold=0.0))
    assert res == exp                      # gathered in input order, identical to the single-process run
    assert ms == 11.0 and units == 10.0    # max over ranks, sum over ranks
