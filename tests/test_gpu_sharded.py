"""Lexicon-sharded mode: shard-local scoring + exchange + merge must equal the unsharded result
(and therefore the oracle), bit for bit.  The 1-GPU test emulates the all-gather by concatenation;
the 2-GPU test runs the real NCCL exchange (skipped when fewer than 2 GPUs are visible)."""
import os
import socket

import pytest

import workloads
from oracle import orc

pytestmark = pytest.mark.gpu


def _models(A, sharded, n_shards, lexicon, confusables=()):
    ms = []
    for s in range(n_shards):
        m = sharded.ShardedVariantModel(workloads.ALPHABET, A.Weights())
        m.read_lexicon(lexicon)
        for pat, w in confusables:
            m.add_to_confusables(pat, w)
        m.build(device=0, shard=s, n_shards=n_shards)
        ms.append(m)
    return ms


@pytest.mark.parametrize("n_shards,kw,conf", [(2, dict(), False), (3, dict(freq_weight=0.25), False),
                                               (2, dict(freq_weight=0.25, max_matches=5), True)],
                         ids=["2 shards", "3 shards freq", "2 shards confusables"])
def test_sharded_equals_oracle_single_gpu(n_shards, kw, conf):
    import analiticcl_b200 as A
    from analiticcl_b200 import sharded
    from test_gpu_parity import assert_same, to_orc_params
    lexicon = workloads.nld_freq_lexicon()
    confs = workloads.CFG2_CONFUSABLES if conf else ()
    ms = _models(A, sharded, n_shards, lexicon, confs)
    o = orc.OracleModel(alphabet_file=workloads.ALPHABET)
    o.read_lexicon(lexicon)
    for pat, w in confs:
        o.add_to_confusables(pat, w)
    o.build()
    qs = workloads.ocr_noise(workloads.read_words("nld"), 1200, 77) + ["", "a", "zzzzzzzz"]
    sp = A.SearchParameters(**kw)
    batches, exports = zip(*[m.score(qs, sp, 0) for m in ms])
    got = sharded.merge_exports_locally(ms, list(batches), list(exports), len(qs))
    exp = o.find_variants_batch(qs, to_orc_params(sp))
    assert_same(got, exp, qs, f"sharded x{n_shards}")
    assert sum(e.n_records for e in exports) > 0 and all(e.n_records > 0 for e in exports)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, queries, q):
    import torch
    import torch.distributed as dist
    import analiticcl_b200 as A
    from analiticcl_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    m = sharded.ShardedVariantModel(workloads.ALPHABET, A.Weights())
    m.read_lexicon(workloads.lexicon_path("eng"))
    m.build(device=rank, shard=rank, n_shards=world)
    res = m.find_variants_raw(queries, A.SearchParameters(), device=rank)  # the library's own NCCL exchange
    res_torch = m.find_variants_raw_torch(queries, A.SearchParameters(), device=rank)  # cross-check: torch collectives
    again = m.find_variants_raw(queries[:700], A.SearchParameters(freq_weight=0.2, max_matches=5), device=rank)
    if rank == 0:
        q.put((res, res_torch, again))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_nccl_two_gpus(eng_oracle):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from test_gpu_parity import assert_same
    qs = workloads.misspellings(workloads.read_words("eng"), 1500, 5150, min_len=2, max_len=16)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, qs, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, got_torch, again = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp = eng_oracle.find_variants_batch(qs, orc.make_params())
    assert_same(got, exp, qs, "nccl sharded (library exchange)")
    assert_same(got_torch, exp, qs, "nccl sharded (torch collectives)")
    assert_same(again, eng_oracle.find_variants_batch(qs[:700], orc.make_params(freq_weight=0.2, max_matches=5)), qs[:700],
                "nccl sharded, second batch on the same communicator")
