#!/usr/bin/env python3
"""bench.py -- throughput of the variant-lookup hot path on B200 (queries/s, DP GCUPS, roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|cfg4]

Workload (BASELINE.json configs[1], the one `metric` is quoted on): query mode on
nld.aspell.lexicon + simple.alphabet.tsv with a synthetic Zipf frequency column, 1M synthetic
OCR-noise queries, max anagram / edit distance 3, late confusables, freq_weight 0.25.
A "step" is one pass of the hot path over the whole 1M-query batch.

  value : queries/s with the encoded batch already resident in HBM (probe + score kernels on the
          device, results left in HBM); CUDA events on the launching stream, max over ranks.
  e2e   : the same metric through the C-ABI call anl_find_variants_batch with HOST buffers:
          host normalisation, H2D, both kernels, D2H, host post-pass (confusables, cut-off).
  N > 1 : one process per GPU (torchrun), index replicated, every rank gets its own 1M-query batch
          (different seed) -> weak scaling, no data-path collective.
  --impl reference : the CPU oracle (C++ restatement of the reference algorithm, OpenMP over queries,
          all host threads) on a bounded sample of the same workload.  The Rust reference itself
          cannot be built in this image (no cargo/rustc).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import workloads  # noqa: E402

METRIC = "queries/sec and DP GCUPS at 1/2/4/8 B200 vs reference CPU (cores stated)"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, workload, n):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed `ncu --set full` capture
    (profiles/traffic.json, written by tools/summarize_ncu.py), if it was taken on this workload and launch size."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(p))
    except (OSError, ValueError):
        return None, None
    if "probe" in d and "workload" in d.get("probe", {}):  # round-1 layout: one workload at the top level
        d = {d["probe"]["workload"]: d}
    e = d.get(workload, {}).get(kernel)
    if not e or int(e.get("queries", -1)) != int(n):
        return None, None
    return float(e["dram_bytes"]), e.get("source")


def ncu_entry(kernel, workload):
    """The whole profiles/traffic.json entry of one kernel on one workload (issue utilisation, lanes per instruction,
    L2 sectors...), or {}."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except (OSError, ValueError):
        return {}
    if "probe" in d and "workload" in d.get("probe", {}):
        d = {d["probe"]["workload"]: d}
    return d.get(workload, {}).get(kernel, {})


def random_read_peak(footprint_bytes, width):
    """Measured random-read rate of this GPU (G reads/s) for independent `width`-byte loads spread over a footprint of
    that many bytes: tools/micro/randread.cu, table committed as profiles/r02d_randread.txt (log-interpolated)."""
    rows = []
    try:
        for ln in open(os.path.join(ROOT, "profiles", "r02d_randread.txt")):
            f = ln.split()
            if len(f) == 5 and f[0].isdigit():
                rows.append((float(f[0]) * 2 ** 20, {8: float(f[1]), 16: float(f[2]), 32: float(f[3])}[width]))
    except OSError:
        return None
    if not rows:
        return None
    if footprint_bytes <= rows[0][0]:
        return rows[0][1]
    for (b0, r0), (b1, r1) in zip(rows, rows[1:]):
        if footprint_bytes <= b1:
            t = (np.log(footprint_bytes) - np.log(b0)) / (np.log(b1) - np.log(b0))
            return float(np.exp(np.log(r0) + t * (np.log(r1) - np.log(r0))))
    return rows[-1][1]


def survey_probe_keys(spec, queries, k, sample=400):
    """SURVEY.md 8(d)'s P: distinct neighbourhood keys per query under canonical generation WITHOUT the
    symmetric-delete level, P(q,k) = sum over distinct non-empty sub-multisets D (d <= k deletions) of
    sum_{j <= k-d} C(a_D + j - 1, j), a_D = alphabet_size - (distinct classes deleted).  Mean over a sample."""
    from math import comb
    from itertools import combinations
    from oracle import orc
    o = orc.OracleModel(alphabet_file=workloads.ALPHABET)
    asize = 32
    tot = 0
    qs = queries[:sample]
    for q in qs:
        f = sorted(o.normalize(q))
        n = len(f)
        seen = set()
        for d in range(0, min(k, n - 1) + 1):
            for pos in combinations(range(n), d):
                rest = tuple(f[i] for i in range(n) if i not in pos)
                if not rest or (d, rest) in seen:
                    continue
                seen.add((d, rest))
                a = asize - len({f[i] for i in pos})
                tot += sum(comb(a + j - 1, j) for j in range(0, k - d + 1))
    return tot / max(1, len(qs))


def workload_spec(name, rank=0):
    """-> dict(lexicon builder, queries, search kwargs, confusables, label)."""
    if name == "cfg2":
        return dict(label="cfg2: nld.aspell.lexicon + Zipf freq, 1M OCR-noise queries, k=3, late confusables, freq_weight=0.25",
                    lexicon=workloads.nld_freq_lexicon(), queries=lambda n: workloads.cfg2_queries(n, 2003 + rank),
                    n=1_000_000, params=dict(max_anagram_distance=3, max_edit_distance=3, freq_weight=0.25),
                    confusables=workloads.CFG2_CONFUSABLES)
    if name == "cfg1":
        return dict(label="cfg1: eng.aspell.lexicon, 10k synthetic misspellings, k=2",
                    lexicon=workloads.lexicon_path("eng"), queries=lambda n: workloads.cfg1_queries(n, 1001 + rank),
                    n=10_000, params=dict(max_anagram_distance=2, max_edit_distance=2), confusables=[])
    if name == "eng3":
        # the north-star target sentence: English aspell, query mode, max edit distance 3
        return dict(label="eng3: eng.aspell.lexicon, 1M synthetic misspellings (cfg1 generator), k=3",
                    lexicon=workloads.lexicon_path("eng"), queries=lambda n: workloads.cfg1_queries(n, 1003 + rank),
                    n=1_000_000, params=dict(max_anagram_distance=3, max_edit_distance=3), confusables=[])
    if name == "cfg4":
        return dict(label="cfg4: eng.aspell.lexicon, 1M misspellings (len>=8, 2-4 edits), k=4",
                    lexicon=workloads.lexicon_path("eng"), queries=lambda n: workloads.cfg4_queries(n, 4001 + rank),
                    n=1_000_000, params=dict(max_anagram_distance=4, max_edit_distance=4), confusables=[])
    if name.startswith("cfg5"):
        # cfg5 or cfg5:<entries> -- synthetic corpus lexicon (default 10 M entries), HBM-resident index
        entries = int(name.split(":")[1]) if ":" in name else 10_000_000
        return dict(label=f"cfg5: {entries}-entry synthetic corpus lexicon (concatenated eng entries, Zipf freq), "
                          "1M misspellings per step, k=3",
                    lexicon=workloads.cfg5_lexicon(entries), n=1_000_000,
                    queries=lambda n: workloads.cfg5_queries(n, 5002 + rank, entries),
                    params=dict(max_anagram_distance=3, max_edit_distance=3), confusables=[])
    raise SystemExit(f"unknown workload {name}")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def make_oracle(spec):
    from oracle import orc
    o = orc.OracleModel(alphabet_file=workloads.ALPHABET)
    o.read_lexicon(spec["lexicon"])
    for pat, w in spec["confusables"]:
        o.add_to_confusables(pat, w)
    o.build()
    return o, orc.make_params(**spec["params"])


def time_oracle(o, oparams, queries, threads):
    blob, offs = workloads.pack(queries)
    offs_c = offs.ctypes.data_as(C.POINTER(C.c_uint64))
    t0 = time.perf_counter()
    total, st = o.find_variants_batch_raw(blob, offs_c, len(queries), oparams, threads)
    dt = time.perf_counter() - t0
    return dt, total, st


def run_reference(args):
    """--impl reference: the CPU restatement of the reference algorithm, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import orc
    spec = workload_spec(args.workload)
    o, oparams = make_oracle(spec)
    threads = host_cores()  # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core
    sample_n = min(spec["n"], args.ref_sample)
    queries = spec["queries"](spec["n"])[:sample_n]
    for _ in range(args.warmup):
        time_oracle(o, oparams, queries[: max(64, sample_n // 8)], threads)
    times, cells = [], 0
    for _ in range(args.steps):
        dt, _, st = time_oracle(o, oparams, queries, threads)
        times.append(dt)
        cells = st.dl_cells
    tot = sum(times)
    qps = sample_n * args.steps / tot
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 multi-limb integer / u8 DP / f64 score", "data": "synthetic",
        "config": {"workload": spec["label"], "batch_queries": sample_n,
                   "note": "CPU port of the reference algorithm (oracle/oracle.cpp); the Rust reference cannot be built here"},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": f"first {sample_n} queries of the workload per step, OpenMP over queries",
                         "gcups": cells * 1e-9 / (tot / args.steps)},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import analiticcl_b200 as A
    from analiticcl_b200 import _capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # the host side of the e2e path is multi-threaded: split the box's cores between the ranks
    os.environ.setdefault("ANL_HOST_THREADS", str(max(1, host_cores() // world)))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(0)
    dev = torch.cuda.current_device()
    L = _capi.lib()

    spec = workload_spec(args.workload, rank)
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.read_lexicon(spec["lexicon"])
    for pat, w in spec["confusables"]:
        m.add_to_confusables(pat, w)
    # `--gpus N` without torchrun: ONE process, one model with a replica on each of the N GPUs (anl_model_build_multi);
    # the e2e call is then spread over all of them by the library.  `value` stays the device-resident pass on GPU 0.
    single_process_gpus = args.gpus if (world == 1 and args.gpus > 1) else 1
    t0 = time.perf_counter()
    if single_process_gpus > 1:
        m.build(devices=list(range(single_process_gpus)))
    else:
        m.build(device=dev)
    build_s = time.perf_counter() - t0
    n = args.queries or spec["n"]
    queries = spec["queries"](n)
    sp = A.SearchParameters(**spec["params"])
    blob, offs = _capi.pack(queries)
    offs_p = _capi.u64ptr(offs)

    def check(st):
        if st != 0:
            raise RuntimeError(L.anl_last_error().decode())

    # ---- value: device-resident batch ----------------------------------------------------------------
    batch = C.c_void_p()
    check(L.anl_device_batch_create(m._h, blob, offs_p, n, C.byref(sp.data), C.byref(batch)))
    stream = torch.cuda.Stream()
    sh = C.c_void_p(stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        check(L.anl_device_batch_run(m._h, batch, sh))
    torch.cuda.synchronize()
    pm, sm_, xm = C.c_float(), C.c_float(), C.c_float()
    check(L.anl_device_batch_timings(m._h, batch, C.byref(pm), C.byref(sm_), C.byref(xm)))  # resets the per-run event window
    sampler = ClockSampler(dev)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    launches0 = L.anl_kernel_launches()
    ev0.record(stream)
    for _ in range(args.steps):
        check(L.anl_device_batch_run(m._h, batch, sh))
    ev1.record(stream)
    torch.cuda.synchronize()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    stage = (C.c_float * 7)()
    check(L.anl_device_batch_stage_timings(m._h, batch, stage))
    bloom_ms, exact_ms, prefilter_ms, rank_ms, conf_ms, finish_ms, export_ms = (float(v) for v in stage)
    probe_ms, score_ms, rescore_ms = bloom_ms + exact_ms, prefilter_ms + rank_ms, conf_ms + finish_ms
    ctr = _capi.Counters()
    check(L.anl_device_batch_counters(m._h, batch, C.byref(ctr)))
    launches = L.anl_kernel_launches() - launches0  # counted by the library: every kernel of the timed passes

    # ---- e2e: host buffers through the public C-ABI call ------------------------------------------------
    e2e_steps = args.e2e_steps  # 0 = skip (profiling runs)
    n_results = 0
    e2e_s = float("nan")
    launches_e2e0 = 0
    e2e_n = n
    if single_process_gpus > 1:  # every GPU gets a batch of its own size: weak scaling inside one call
        e2e_queries = []
        for g in range(single_process_gpus):
            e2e_queries += workload_spec(args.workload, g)["queries"](n)
        blob, offs = _capi.pack(e2e_queries)
        offs_p = _capi.u64ptr(offs)
        e2e_n = len(e2e_queries)
    n_e2e_call = e2e_n
    if e2e_steps > 0:
        n = n_e2e_call
        rs = C.c_void_p()
        check(L.anl_find_variants_batch(m._h, blob, offs_p, n, C.byref(sp.data), C.byref(rs)))  # warm-up
        n_results = L.anl_result_set_offsets(rs)[n]
        L.anl_result_set_free(rs)
        barrier()
        torch.cuda.synchronize()
        launches_e2e0 = L.anl_kernel_launches()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            rs = C.c_void_p()
            check(L.anl_find_variants_batch(m._h, blob, offs_p, n, C.byref(sp.data), C.byref(rs)))
            L.anl_result_set_free(rs)
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        n = args.queries or spec["n"]
    clocks = sampler.stop()
    if e2e_steps > 0:
        launches += L.anl_kernel_launches() - launches_e2e0  # the batch call works in chunks of 65536 queries
    ist = m.index_stats()
    max_q_bytes = int(np.max(np.diff(offs.astype(np.int64)))) if n else 0
    stride = (min(max_q_bytes, 254) + 2 + 15) & ~15
    # the query text + u32 offsets go up (the rows are encoded on the device); the final arrays come back: u64 offsets,
    # u32 flags, 32-byte variant records (+ 48 bytes of summary per 65536-query chunk)
    h2d = len(blob) + 4 * (n_e2e_call + 1)
    d2h = n_e2e_call * (8 + 4) + 32 * int(n_results) + 48 * ((n_e2e_call + 65535) // 65536)

    # ---- reduce over ranks: max time, summed work --------------------------------------------------------
    from analiticcl_b200 import parallel
    (dev_ms_max, e2e_s_max, probe_ms_max, score_ms_max), (total_q, total_cells) = parallel.reduce_max_sum(
        [dev_ms, e2e_s, probe_ms, score_ms], [float(n), float(ctr.dl_cells)], device="cuda")

    if rank == 0:
        peak, peak_src = load_peaks()
        step_ms = dev_ms_max / args.steps
        value = total_q / (step_ms / 1000.0)
        # algorithmic bytes of one launch of each kernel on this rank (DESIGN.md "rooflines")
        probe_bytes = (n * stride + 16 * (ctr.probes - ctr.deletion_keys) + 8 * ctr.probes + 16 * ctr.probe_steps +
                       5 * ctr.postings + 24 * ctr.postings + 8 * ctr.anagram_hits + 4 * ctr.instance_pairs + 8 * n)
        score_bytes = (n * stride + 4 * ctr.instance_pairs + ist["norm_stride"] * ctr.instance_pairs + 8 * ctr.survivors +
                       24 * ctr.results + 8 * n)
        # the probe kernel is the memory-system-bound one and is reported against the measured HBM
        # peak; the score kernel is integer-issue bound and is reported as DP GCUPS (`dp` object below).
        achieved = probe_bytes / (probe_ms / 1000.0) / 1e9
        wl_key = args.workload
        traffic, traffic_src = ncu_traffic("probe", wl_key, n)
        stage_list = (("bloom_kernel", bloom_ms), ("exact_kernel", exact_ms), ("pair_list_kernels", prefilter_ms),
                      ("dp_kernel+rank_kernel", rank_ms), ("confusable_kernels", conf_ms), ("finish_kernel", finish_ms))
        time_dominant = max(stage_list, key=lambda kv: kv[1])[0]
        index_bytes = ist["table_bytes"] + ist["bloom_bytes"] + 5 * ist["postings"] + 32 * ist["anagrams"] + ist["instance_bytes"]
        hbm_resident = index_bytes > 126e6  # beyond the 126 MB L2
        # one line per kernel (group), each against the limit that binds IT:
        #  * Bloom stage: one independent 8-byte read per neighbourhood node over the Bloom filter's footprint -> measured
        #    random-read rate of this GPU at that footprint (tools/micro/randread.cu); also instruction-issue bound (ncu)
        #  * exact stage: one 16-byte slot read per probe step over the table's footprint -> random-read rate, same source
        #  * DP: u8 cells, integer-issue bound -> GCUPS against SURVEY 8(d)'s nominal 148 SM x 128 lanes x f / 12 ops per cell
        sm_ghz = (clocks.get("sm_mhz") or 1965.0) / 1e3
        dp_peak = 148 * 128 * sm_ghz / 12.0  # GCUPS
        kr = []
        rr8 = random_read_peak(ist["bloom_bytes"], 8)
        if rr8 and bloom_ms > 0:
            e = ncu_entry("bloom_kernel", wl_key)
            kr.append({"kernel": "bloom_kernel", "ms": bloom_ms, "bound": "random 8-byte reads (Bloom words) / instruction issue",
                       "achieved": ctr.probes / (bloom_ms / 1e3) / 1e9, "peak": rr8, "unit": "G reads/s",
                       "frac": ctr.probes / (bloom_ms / 1e3) / 1e9 / rr8, "footprint_bytes": ist["bloom_bytes"],
                       "peak_source": "profiles/r02d_randread.txt (measured on this pool's B200, independent loads)",
                       "issue_active_pct": e.get("issue_active_pct"), "lanes_per_instruction": e.get("lanes_per_instruction")})
        rr16 = random_read_peak(ist["table_bytes"], 16)
        rr32 = random_read_peak(32 * ist["anagrams"], 32)
        if rr16 and rr32 and exact_ms > 0:
            # two dependent random reads dominate a staged node: the 16-byte table slot(s) by fingerprint, then the 32-byte
            # anagram record of every posting that is verified (posting-array reads of multi-posting slots and the staged-node
            # queue itself are not counted).  Roof = the time those reads take at the measured random-read rates of their footprints.
            e = ncu_entry("exact_kernel", wl_key)
            t_min_ms = (ctr.probe_steps / rr16 + ctr.postings / rr32) / 1e6
            kr.append({"kernel": "exact_kernel", "ms": exact_ms, "bound": "random 16-byte reads (table slots) + random 32-byte reads (anagram records)",
                       "achieved": (ctr.probe_steps + ctr.postings) / (exact_ms / 1e3) / 1e9,
                       "peak": (ctr.probe_steps + ctr.postings) / (t_min_ms / 1e3) / 1e9, "unit": "G reads/s",
                       "frac": t_min_ms / exact_ms, "footprint_bytes": [ist["table_bytes"], 32 * ist["anagrams"]],
                       "peak_source": "profiles/r02d_randread.txt (16-byte reads at the table's footprint, 32-byte reads at the anagram records')",
                       "issue_active_pct": e.get("issue_active_pct"), "lanes_per_instruction": e.get("lanes_per_instruction")})
        if score_ms > 0:
            e = ncu_entry("dp_kernel", wl_key)
            gc = total_cells * 1e-9 / (score_ms_max / 1000.0)
            kr.append({"kernel": "pair_list_kernels + dp_kernel + rank_kernel (score stage)", "ms": score_ms, "bound": "integer issue (u8 DP cells)",
                       "achieved": gc, "peak": dp_peak, "unit": "GCUPS", "frac": gc / dp_peak,
                       "peak_source": "nominal: 148 SMs x 128 lanes x %.3f GHz / 12 integer operations per cell (SURVEY 8d)" % sm_ghz,
                       "cells": "reference cells (len_q x len_c of every pair that passes the length check); the DP itself runs "
                                "only on the pairs the bit-parallel prefilter cannot reject (counters.dp_cells)",
                       "issue_active_pct": e.get("issue_active_pct"), "lanes_per_instruction": e.get("lanes_per_instruction")})
        dominant_frac = None
        for r_ in kr:
            if r_["kernel"].startswith(time_dominant.split("+")[0].split("_kernels")[0]):
                dominant_frac = r_["frac"]
        try:
            survey_p = survey_probe_keys(spec, queries, int(spec["params"]["max_anagram_distance"]))
        except Exception:
            survey_p = None
        if hbm_resident:
            note = ("algorithmic bytes = exact per-launch counters (B_probe, the design's own byte model: DESIGN.md section 5); the "
                    "index of this lexicon (%.1f GB) is HBM-resident: the probes are random 8/16/32-byte reads, so the binding limit is "
                    "the GPU's random-read rate at that footprint (kernel_rooflines), not streaming bandwidth" % (index_bytes / 1e9))
        else:
            note = ("algorithmic bytes = exact per-launch counters (B_probe, the design's own byte model: DESIGN.md section 5); the "
                    "index of this lexicon (%.0f MB: Bloom words, table, postings) is L2-resident, so DRAM traffic is far below the "
                    "algorithmic bytes and the binding limit is instruction issue / L2 random-read rate, not HBM" % (index_bytes / 1e6))
        cpu = None
        if world == 1:
            # bounded CPU baseline on rank 0's host cores (N=1 only): the oracle port, all threads
            o, oparams = make_oracle(spec)
            threads = host_cores()
            sample_n = min(n, args.cpu_sample)
            dt, _, st = time_oracle(o, oparams, queries[:sample_n], threads)
            cpu = {"value": sample_n / dt, "unit": "queries/s", "cores": threads, "kind": "port",
                   "sample": f"first {sample_n} queries of the same workload, one pass, OpenMP over queries "
                             "(C++ restatement of the reference algorithm; the Rust binary cannot be built here)",
                   "gcups": st.dl_cells * 1e-9 / dt, "seconds": dt}
        line = {
            "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64 multi-limb integer / u8 DP / f64 score", "data": "synthetic",
            "config": {"workload": spec["label"], "batch_queries_per_gpu": n, "parallelism": f"query-partitioned replicas x{world}",
                       "l2": "per-step working set (query rows + hit lists + staged nodes + results ~ GBs) exceeds the 126 MB L2; "
                             + ("the index is HBM-resident" if hbm_resident else "the index is L2-resident by design"),
                       "index": ist, "index_bytes": index_bytes, "build_seconds": build_s,
                       "value_scope": "probe + score/rank (+ confusable + finish) kernels, encoded batch resident in HBM, "
                                      "results left in HBM"},
            "dp_gcups": total_cells * 1e-9 / (score_ms_max / 1000.0),
            "dp_gcups_of_step": total_cells * 1e-9 / (step_ms / 1000.0),
            "kernels": {"probe_ms": probe_ms, "score_ms": score_ms, "rescore_ms": rescore_ms,
                        "stages_ms": {"bloom_kernel": bloom_ms, "exact_kernel": exact_ms, "prefilter_kernel": prefilter_ms,
                                      "score_kernel": rank_ms, "confusable_kernel": conf_ms, "finish_kernel": finish_ms,
                                      "export_kernels": export_ms},
                        "stage_kernels": {"prefilter_kernel": "pairfilter_kernel + pairscan_kernel + pairscatter_kernel (shape-sorted pair list)",
                                          "score_kernel": "dp_kernel + rank_kernel", "confusable_kernel": "triage_kernel + confusable_kernel + confusable_wide_kernel"},
                        "probe_share": probe_ms / (probe_ms + score_ms + rescore_ms + export_ms),
                        "probe_algorithmic_bytes": probe_bytes, "score_algorithmic_bytes": score_bytes,
                        "probe_gbs": probe_bytes / (probe_ms / 1e3) / 1e9, "score_gbs": score_bytes / (score_ms / 1e3) / 1e9,
                        "probes_per_s": ctr.probes / (probe_ms / 1e3)},
            "counters": {f: getattr(ctr, f) for f, _ in ctr._fields_},
            "roofline": {"bound": "hbm", "kernel": "bloom_kernel + exact_kernel (candidate generation)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "time_dominant_kernel": time_dominant, "time_dominant_frac": dominant_frac,
                         "byte_model": {"B_probe_per_query": probe_bytes / n, "nodes_per_query": ctr.probes / n,
                                        "survey_8d_P_per_query": survey_p,
                                        "survey_8d_bytes_per_query": (survey_p * 32 + (ctr.anagram_hits * 8 + 4 * ctr.instance_pairs) / n)
                                        if survey_p else None,
                                        "why": "symmetric-delete depth 1 removes the last insertion level, so the kernels test "
                                               "`nodes_per_query` keys where SURVEY 8(d)'s canonical generation would test P"},
                         "note": note},
            "kernel_rooflines": kr,
            "dp": {"kernel": "pair list + dp_kernel + rank_kernel", "gcups": total_cells * 1e-9 / (score_ms_max / 1000.0), "unit": "GCUPS",
                   "peak": dp_peak, "frac": total_cells * 1e-9 / (score_ms_max / 1000.0) / dp_peak,
                   "cells_per_launch": ctr.dl_cells, "ms": score_ms,
                   "bound": "integer issue (u8 DP cells in shared memory, no tensor cores); see profiles/ for issue utilisation"},
            "cpu_baseline": cpu,
            "e2e": ({"value": (n_e2e_call if single_process_gpus > 1 else total_q) / e2e_s_max, "unit": "queries/s",
                     "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps, "results_per_step": int(n_results),
                     "gpus_in_one_process": single_process_gpus, "queries_per_call": int(n_e2e_call)} if e2e_steps > 0 else None),
            "gpu_launches": launches,
            "clocks": clocks,
        }
        print(json.dumps(line))
    L.anl_device_batch_free(m._h, batch)
    if world > 1:
        dist.destroy_process_group()


def run_search(args):
    """--workload cfg3: search mode.  find_all_matches over synthetic running text with n-gram spans
    (max_ngram = 3): host segmentation + pipelined GPU batches per window.  A "query" is one n-gram segment
    lookup.  The sequence consolidation (anl_match_set_consolidate) is a host post-pass
    and, as SURVEY 8d asks, kept out of the GPU number: it is timed separately after the timed region and
    reported as `consolidate_ms` / `tokens_per_s_consolidated`."""
    import torch
    import analiticcl_b200 as A
    from analiticcl_b200 import _capi
    torch.cuda.set_device(0)
    L = _capi.lib()
    n_tokens = args.queries or 2_000_000
    text = workloads.cfg3_text(n_tokens, 3001)
    raw = text.encode("utf-8")
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.read_lexicon(workloads.lexicon_path("eng"))
    m.build(device=0)
    sp = A.SearchParameters(max_ngram=3, max_anagram_distance=3, max_edit_distance=3)

    def once():
        ms = C.c_void_p()
        st = L.anl_find_all_matches(m._h, raw, len(raw), C.byref(sp.data), C.byref(ms))
        if st != 0:
            raise RuntimeError(L.anl_last_error().decode())
        n = L.anl_match_set_len(ms)
        return ms, n

    for _ in range(max(1, args.warmup // 3)):
        ms, n = once()
        L.anl_match_set_free(ms)
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = L.anl_kernel_launches()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ms, n = once()
        if _ < args.steps - 1:
            L.anl_match_set_free(ms)
    dt = (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop()
    a, b = C.c_uint64(), C.c_uint64()
    L.anl_match_set_lookup_counts(ms, C.byref(a), C.byref(b))
    lookups, distinct = a.value, b.value
    frac = lookups / max(1, n)
    # host post-pass, outside the timed region: most likely sequence per hard-delimited batch
    tc = time.perf_counter()
    best = C.c_void_p()
    if L.anl_match_set_consolidate(ms, raw, len(raw), C.byref(sp.data), C.byref(best)) != 0:
        raise RuntimeError(L.anl_last_error().decode())
    consolidate_s = time.perf_counter() - tc
    n_best = L.anl_match_set_len(best)
    L.anl_match_set_free(best)
    L.anl_match_set_free(ms)
    line = {
        "metric": METRIC, "value": lookups / dt, "unit": "queries/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 multi-limb integer / u8 DP / f64 score", "data": "synthetic",
        "config": {"workload": f"cfg3: find_all_matches over {n_tokens} tokens of synthetic running text, max_ngram=3, k=3 (eng)",
                   "segments": n, "segment_lookups": lookups, "looked_up_fraction": frac,
                   "distinct_strings_sent_to_gpu": distinct, "text_bytes": len(raw),
                   "value_scope": "end to end through anl_find_all_matches (host segmentation, pipelined GPU batches per "
                                  "window, result assembly); query = one n-gram segment lookup"},
        "tokens_per_s": n_tokens / dt,
        "consolidate_ms": consolidate_s * 1e3, "consolidated_matches": n_best,
        "tokens_per_s_consolidated": n_tokens / (dt + consolidate_s),
        "e2e": {"value": lookups / dt, "unit": "queries/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None},
        "gpu_launches": L.anl_kernel_launches() - launches0, "clocks": clocks,
    }
    print(json.dumps(line))


def run_sharded(args):
    """--sharded: lexicon-sharded mode (SURVEY 8e mode 2).  Every rank holds 1/N of the anagram keys, scores the WHOLE
    batch against its shard; the library exchanges the survivor lists over NCCL (one 8-byte all-gather of sizes + ONE
    grouped collective with exact sizes, csrc/shard_comm.cu) and merges on every rank.  A step = anl_shard_batch_step:
    score kernels + exchange + merge kernel + export kernels, batch resident in HBM; CUDA events per part."""
    import torch
    import torch.distributed as dist
    import analiticcl_b200 as A
    from analiticcl_b200 import _capi, sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("ANL_HOST_THREADS", str(max(1, host_cores() // world)))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = _capi.lib()
    spec = workload_spec(args.workload, 0)  # every rank sees the same queries
    m = sharded.ShardedVariantModel(workloads.ALPHABET, A.Weights())
    m.read_lexicon(spec["lexicon"])
    for pat, w in spec["confusables"]:
        m.add_to_confusables(pat, w)
    t0 = time.perf_counter()
    m.build(device=local_rank, shard=rank, n_shards=world)
    build_s = time.perf_counter() - t0
    m.init_comm()  # the library's own NCCL communicator (torch.distributed only carries the id)
    n = args.queries or spec["n"]
    queries = spec["queries"](n)
    sp = A.SearchParameters(**spec["params"])
    dev = torch.device("cuda", local_rank)

    def check(st):
        if st != 0:
            raise RuntimeError(L.anl_last_error().decode())

    blob, offs = _capi.pack(queries)
    batch = C.c_void_p()
    check(L.anl_device_batch_create(m._h, blob, _capi.u64ptr(offs), n, C.byref(sp.data), C.byref(batch)))
    stats = _capi.ShardStepStats()

    def step():
        check(L.anl_shard_batch_step(m._h, batch, C.byref(stats), None))

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    dist.barrier()
    torch.cuda.synchronize()
    launches0 = L.anl_kernel_launches()
    acc = [0.0, 0.0, 0.0]
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
        acc[0] += stats.score_ms
        acc[1] += stats.exchange_ms
        acc[2] += stats.merge_ms
    torch.cuda.synchronize()
    dist.barrier()
    dt = (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop()
    from analiticcl_b200 import parallel
    (dt_max, score_ms, exchange_ms, merge_ms), _ = parallel.reduce_max_sum([dt] + [a / args.steps for a in acc], [n], device=dev)
    if rank == 0:
        line = {
            "metric": METRIC, "value": n / dt_max, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt_max * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u64 multi-limb integer / u8 DP / f64 score", "data": "synthetic",
            "config": {"workload": spec["label"], "batch_queries": n,
                       "parallelism": f"lexicon-sharded x{world} (hash(key) mod N) + NCCL exchange inside the library + merge",
                       "index": m.index_stats(), "build_seconds": build_s,
                       "value_scope": "per step: score kernels on the shard, 8-byte all-gather of sizes, one grouped NCCL collective "
                                      "with every shard's survivors (exact sizes), merge kernel, export kernels; batch resident in HBM; "
                                      "host wall clock over the steps (max over ranks), parts by CUDA events"},
            "kernels": {"score_ms": score_ms, "exchange_ms": exchange_ms, "merge_and_export_ms": merge_ms,
                        "exchange_and_merge_ms": exchange_ms + merge_ms,
                        "nvlink_bytes_received_per_rank": int(stats.bytes_received),
                        "nvlink_gbs_received": stats.bytes_received / (exchange_ms / 1e3) / 1e9 if exchange_ms > 0 else None,
                        "survivor_records_this_rank": int(stats.records_local), "survivor_records_all_ranks": int(stats.records_total)},
            "gpu_launches": L.anl_kernel_launches() - launches0, "clocks": clocks,
        }
        print(json.dumps(line))
    L.anl_device_batch_free(m._h, batch)
    L.anl_shard_comm_free(m._h)
    dist.destroy_process_group()


def keep_stdout_clean():
    """NCCL writes its version / debug lines to stdout by default; the driver reads ONE JSON line from stdout."""
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")


def main():
    keep_stdout_clean()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--queries", type=int, default=0, help="override the batch size (default: the config's)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-sample", type=int, default=20000)
    ap.add_argument("--ref-sample", type=int, default=4000)
    ap.add_argument("--sharded", action="store_true", help="lexicon-sharded mode (needs torchrun with N > 1)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.sharded:
        run_sharded(args)
    elif args.workload == "cfg3":
        run_search(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
