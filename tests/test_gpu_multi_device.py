"""One model, several GPUs of one process (anl_model_build_multi): every lookup call is spread over the replicas,
one dispatcher thread per device, and comes back in query order exactly as from one device (and the oracle)."""
import pytest

import workloads
from oracle import orc

pytestmark = pytest.mark.gpu


def _devices():
    import torch
    return list(range(torch.cuda.device_count()))


@pytest.mark.parametrize("chunk", ["1024", "65536"])
def test_replicas_equal_oracle(eng_oracle, monkeypatch, chunk):
    import analiticcl_b200 as A
    from test_gpu_parity import assert_same, to_orc_params
    devs = _devices()
    monkeypatch.setenv("ANL_CHUNK", chunk)
    monkeypatch.setenv("ANL_HIT_CAP", "256")  # some chunks re-run overflowed queries
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.read_lexicon(workloads.lexicon_path("eng"))
    m.build(devices=devs)  # all visible GPUs (on a 1-GPU box: the same code path with one replica)
    assert m.device_count() == len(devs)
    words = workloads.read_words("eng")
    qs = workloads.misspellings(words, 9000, 31, min_len=2, max_len=20) + ["", "q" * 300]
    qs += workloads.misspellings(words, 4000, 32, min_len=3, max_len=6)
    sp = A.SearchParameters()
    got = m.find_variants_raw(qs, sp)
    exp = eng_oracle.find_variants_batch(qs, to_orc_params(sp), threads=0)
    assert_same(got, exp, qs, f"{len(devs)} replicas chunk {chunk}")
    # find_all_matches rides on the same batch call: the same matches as from a single-device model
    text = " ".join(qs[:3000])
    a = m.find_all_matches(text, A.SearchParameters(max_ngram=1))
    single = A.VariantModel(workloads.ALPHABET, A.Weights())
    single.read_lexicon(workloads.lexicon_path("eng"))
    single.build(device=0)
    assert a == single.find_all_matches(text, A.SearchParameters(max_ngram=1)) and len(a) >= 3000


def test_replicas_with_confusables(monkeypatch):
    """The confusable stage and the host finish of flagged queries under the multi-device dispatch."""
    import analiticcl_b200 as A
    from test_gpu_parity import assert_same, to_orc_params
    devs = _devices()
    monkeypatch.setenv("ANL_CHUNK", "2048")
    lexicon = workloads.nld_freq_lexicon()
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.read_lexicon(lexicon)
    o = orc.OracleModel(alphabet_file=workloads.ALPHABET)
    o.read_lexicon(lexicon)
    for pat, w in workloads.CFG2_CONFUSABLES:
        m.add_to_confusables(pat, w)
        o.add_to_confusables(pat, w)
    m.build(devices=devs)
    o.build()
    qs = workloads.ocr_noise(workloads.read_words("nld"), 12000, 909)
    sp = A.SearchParameters(freq_weight=0.25)
    assert_same(m.find_variants_raw(qs, sp), o.find_variants_batch(qs, to_orc_params(sp)), qs, "replicas + confusables")


def test_bad_device_lists():
    import analiticcl_b200 as A
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.add_to_vocabulary("frog", 1, A.VocabParams())
    with pytest.raises(ValueError):
        m.build(devices=[])
    with pytest.raises(ValueError):
        m.build(devices=[0, 0])
    with pytest.raises(RuntimeError):
        m.build(devices=[0, 99])


def _partitioned_worker(rank, world, port, out):
    import os
    import torch
    import torch.distributed as dist
    import analiticcl_b200 as A
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)  # (plumbing only: the results are gathered as objects)
    dev = rank % torch.cuda.device_count()
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.read_lexicon(workloads.lexicon_path("eng"))
    m.build(device=dev)
    qs = workloads.misspellings(workloads.read_words("eng"), 3001, 77)
    res = m.find_variants_partitioned(qs, A.SearchParameters())
    if rank == 0:
        out.put(res)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_query_partitioned_processes_equal_oracle(eng_oracle):
    """SURVEY 8e mode 1 as the Python mirror exposes it: two processes (one replica each, on GPU rank % n_gpus), every
    rank looks up its contiguous slice, rank 0 gets the whole list in input order."""
    import torch.multiprocessing as mp
    from test_gpu_parity import assert_same, to_orc_params
    import analiticcl_b200 as A
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_partitioned_worker, args=(r, 2, 29533, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    qs = workloads.misspellings(workloads.read_words("eng"), 3001, 77)
    sp = A.SearchParameters()
    assert_same(res, eng_oracle.find_variants_batch(qs, to_orc_params(sp), threads=0), qs, "2 query-partitioned processes")
