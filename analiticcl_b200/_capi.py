"""ctypes declarations of include/analiticcl_b200.h (the same stub a maintainer of the reference's
Python binding would write)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("ANL_LIB_PATH") or os.path.join(_HERE, "libanaliticcl_b200.so")  # ANL_LIB_PATH: developer override (A/B of builds)

OK, ERR_INVALID, ERR_IO, ERR_NOT_BUILT, ERR_CUDA, ERR_UNSUPPORTED, ERR_EMPTY_INPUT = range(7)
THRESHOLD_RATIO, THRESHOLD_RATIO_WITH_LIMIT, THRESHOLD_ABSOLUTE = 0, 1, 2
STOP_EXHAUSTIVE, STOP_AT_EXACT_MATCH = 0, 1
VOCAB_NONE, VOCAB_INDEXED, VOCAB_LM, VOCAB_TRANSPARENT = 0, 1, 2, 4
FREQ_SUM, FREQ_MAX, FREQ_MIN, FREQ_REPLACE = 0, 1, 2, 3
NO_VIA = 0xFFFFFFFFFFFFFFFF


class Weights(C.Structure):
    _fields_ = [("ld", C.c_double), ("lcs", C.c_double), ("prefix", C.c_double), ("suffix", C.c_double),
                ("case_", C.c_double)]


class Threshold(C.Structure):
    _fields_ = [("kind", C.c_int32), ("ratio", C.c_float), ("value", C.c_uint32)]


class SearchParams(C.Structure):
    _fields_ = [
        ("max_anagram_distance", Threshold), ("max_edit_distance", Threshold), ("max_matches", C.c_uint64),
        ("score_threshold", C.c_double), ("cutoff_threshold", C.c_double), ("stop_criterion", C.c_int32),
        ("max_ngram", C.c_uint32), ("lm_order", C.c_uint32), ("max_seq", C.c_uint64), ("single_thread", C.c_int32),
        ("context_weight", C.c_float), ("variantmodel_weight", C.c_float), ("lm_weight", C.c_float),
        ("contextrules_weight", C.c_float), ("freq_weight", C.c_float), ("consolidate_matches", C.c_int32),
        ("unicodeoffsets", C.c_int32),
    ]


class VocabParams(C.Structure):
    _fields_ = [("text_column", C.c_uint32), ("freq_column", C.c_int32), ("freq_handling", C.c_int32),
                ("vocab_type", C.c_uint32), ("index", C.c_uint32)]


class Variant(C.Structure):
    _fields_ = [("vocab_id", C.c_uint64), ("dist_score", C.c_double), ("freq_score", C.c_double), ("via", C.c_uint64)]


class VocabInfo(C.Structure):
    _fields_ = [("text", C.c_void_p), ("text_len", C.c_uint32), ("frequency", C.c_uint32), ("lexindex", C.c_uint32),
                ("vocabtype", C.c_uint32), ("tokencount", C.c_uint32), ("norm_len", C.c_uint32)]


class Match(C.Structure):
    _fields_ = [("begin", C.c_uint64), ("end", C.c_uint64), ("n", C.c_uint32), ("selected", C.c_int32),
                ("n_variants", C.c_uint64), ("variants", C.POINTER(Variant))]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("queries", "deletion_keys", "probes", "filter_pass", "probe_steps", "postings", "anagram_hits",
                 "instance_pairs", "dl_pairs", "dl_cells", "survivors", "results", "reruns", "dp_pairs",
                 "dp_cells")]


class ShardStepStats(C.Structure):
    _fields_ = [("score_ms", C.c_float), ("exchange_ms", C.c_float), ("merge_ms", C.c_float), ("bytes_received", C.c_uint64),
                ("records_local", C.c_uint64), ("records_total", C.c_uint64)]


class IndexStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("table_slots", "table_bytes", "slot_bytes", "table_keys", "bloom_bytes", "postings", "anagrams",
                 "instances", "instance_bytes", "norm_stride", "mset_entries", "mset_bytes")] + \
               [(n, C.c_uint32) for n in ("max_key_bits", "max_charcount", "active_classes", "sd")]


# every symbol include/analiticcl_b200.h declares: name -> (restype, argtypes)
_vp, _cp, _u64, _i64, _i32, _u32, _sz = C.c_void_p, C.c_char_p, C.c_uint64, C.c_int64, C.c_int32, C.c_uint32, C.c_size_t
_P = C.POINTER
SIGNATURES = {
    "anl_weights_default": (None, [_P(Weights)]),
    "anl_search_params_default": (None, [_P(SearchParams)]),
    "anl_vocab_params_default": (None, [_P(VocabParams)]),
    "anl_last_error": (_cp, []),
    "anl_version": (_cp, []),
    "anl_model_new": (_i32, [_cp, _P(Weights), _i32, _P(_vp)]),
    "anl_model_new_from_tsv": (_i32, [_cp, _sz, _P(Weights), _i32, _P(_vp)]),
    "anl_model_free": (None, [_vp]),
    "anl_model_read_vocabulary": (_i32, [_vp, _cp, _P(VocabParams)]),
    "anl_model_add_to_vocabulary": (_i32, [_vp, _cp, _sz, _i32, _u32, _P(VocabParams), _P(_u64)]),
    "anl_model_add_variant": (_i32, [_vp, _u64, _cp, _sz, C.c_double, _i32, _u32, _P(VocabParams), _P(_i32)]),
    "anl_model_read_variants": (_i32, [_vp, _cp, _P(VocabParams), _i32]),
    "anl_model_read_confusablelist": (_i32, [_vp, _cp]),
    "anl_model_add_to_confusables": (_i32, [_vp, _cp, C.c_double]),
    "anl_model_set_confusables_before_pruning": (None, [_vp]),
    "anl_model_build": (_i32, [_vp, _i32]),
    "anl_model_build_multi": (_i32, [_vp, _P(C.c_int32), _u32]),
    "anl_model_build_on": (_i32, [_vp, _i32, _i32]),
    "anl_debug_index_digest": (None, [_vp, _P(_u64), _sz]),
    "anl_model_device_count": (_u32, [_vp]),
    "anl_model_has": (_i32, [_vp, _cp, _sz]),
    "anl_model_vocab_id": (_i64, [_vp, _cp, _sz]),
    "anl_model_vocab_size": (_u64, [_vp]),
    "anl_model_get_vocab": (_i32, [_vp, _u64, _P(VocabInfo)]),
    "anl_model_lexicon_count": (_u32, [_vp]),
    "anl_model_lexicon_name": (_cp, [_vp, _u32]),
    "anl_model_alphabet_size": (_u32, [_vp]),
    "anl_model_index_size": (_u64, [_vp]),
    "anl_model_instance_count": (_u64, [_vp]),
    "anl_model_anagram_count_of_length": (_u64, [_vp, _u32]),
    "anl_model_max_key_bits": (_u32, [_vp]),
    "anl_normalize": (_i64, [_vp, _cp, _sz, _P(C.c_uint8), _sz]),
    "anl_anahash": (_i64, [_vp, _cp, _sz, _P(_u64), _sz]),
    "anl_shortest_edit_script": (_i64, [_cp, _sz, _cp, _sz, C.c_char_p, _sz]),
    "anl_shortest_edit_script_fixed": (_i64, [_cp, _sz, _cp, _sz, C.c_char_p, _sz]),
    "anl_confusable_found_in": (_i32, [_cp, _cp, _sz, _cp, _sz]),
    "anl_find_variants_batch": (_i32, [_vp, _cp, _P(_u64), _u64, _P(SearchParams), _P(_vp)]),
    "anl_result_set_len": (_u64, [_vp]),
    "anl_result_set_get": (_P(Variant), [_vp, _u64, _P(_u64)]),
    "anl_result_set_offsets": (_P(_u64), [_vp]),
    "anl_result_set_variants": (_P(Variant), [_vp]),
    "anl_result_set_flags": (_u32, [_vp, _u64]),
    "anl_result_set_free": (None, [_vp]),
    "anl_find_all_matches": (_i32, [_vp, _cp, _sz, _P(SearchParams), _P(_vp)]),
    "anl_match_set_len": (_u64, [_vp]),
    "anl_match_set_get": (_i32, [_vp, _u64, _P(Match)]),
    "anl_kernel_launches": (_u64, []),
    "anl_match_set_free": (None, [_vp]),
    "anl_match_set_lookup_counts": (None, [_vp, _P(C.c_uint64), _P(C.c_uint64)]),
    "anl_model_save_index": (_i32, [_vp, _cp]),
    "anl_model_load_index": (_i32, [_vp, _cp, _i32]),
    "anl_model_shard": (None, [_vp, _P(_u32), _P(_u32)]),
    "anl_match_set_consolidate": (_i32, [_vp, _cp, _sz, _P(SearchParams), _P(_vp)]),
    "anl_debug_match_set_build": (_i32, [_cp, _sz, C.c_uint32, C.c_int32, _P(C.c_uint8), _P(C.c_uint64), _P(Variant), _u64, _P(_vp)]),
    "anl_model_learn_variants": (_i32, [_vp, _cp, _P(_u64), _u64, _P(SearchParams), _i32, _i32, _P(_u64)]),
    "anl_debug_learn_apply": (_i32, [_vp, _cp, _P(_u64), _u64, _P(_u64), _P(C.c_double), _P(_u64)]),
    "anl_debug_vocab_links": (_i64, [_vp, _u64, _i32, _P(_u64), _P(C.c_double), _sz]),
    "anl_model_consolidate": (_i32, [_vp, _vp, _cp, _sz, _P(SearchParams), _P(_vp)]),
    "anl_match_set_tags": (_u64, [_vp, _u64, _P(_P(C.c_uint16)), _P(_P(C.c_uint8))]),
    "anl_model_have_lm": (_i32, [_vp]),
    "anl_model_ngram_count": (_u64, [_vp]),
    "anl_model_read_contextrules": (_i32, [_vp, _cp]),
    "anl_model_add_contextrule": (_i32, [_vp, _cp, C.c_float, _P(_cp), C.c_uint32, _P(_cp), C.c_uint32]),
    "anl_model_contextrule_count": (C.c_uint32, [_vp]),
    "anl_model_tag_count": (C.c_uint32, [_vp]),
    "anl_model_tag_name": (_cp, [_vp, C.c_uint32]),
    "anl_debug_lm_score_tokens": (None, [_vp, _P(C.c_int64), _u64, _P(C.c_float), _P(C.c_double)]),
    "anl_debug_find_boundaries": (_i64, [_cp, _sz, _P(C.c_uint64), _P(C.c_uint64), _P(C.c_int32), _sz]),
    "anl_debug_segment_text": (_i64, [_cp, _sz, C.c_uint32, _P(C.c_uint64), _P(C.c_uint64), _P(C.c_uint32), _P(C.c_uint32), _sz]),
    "anl_debug_segment_text_device": (_i64, [C.c_int32, _cp, _sz, C.c_uint32, _P(C.c_uint64), _P(C.c_uint64), _P(C.c_uint32), _P(C.c_uint32), _sz,
                                             _P(C.c_uint64), _P(C.c_uint64), _P(C.c_int32), _sz, _P(C.c_uint64)]),
    "anl_device_batch_create": (_i32, [_vp, _cp, _P(_u64), _u64, _P(SearchParams), _P(_vp)]),
    "anl_device_batch_run": (_i32, [_vp, _vp, _vp]),
    "anl_device_batch_timings": (_i32, [_vp, _vp, _P(C.c_float), _P(C.c_float), _P(C.c_float)]),
    "anl_device_batch_stage_timings": (_i32, [_vp, _vp, _P(C.c_float)]),
    "anl_device_batch_fetch": (_i32, [_vp, _vp, _P(_vp)]),
    "anl_device_batch_free": (None, [_vp, _vp]),
    "anl_device_batch_counters": (_i32, [_vp, _vp, _P(Counters)]),
    "anl_model_index_stats": (_i32, [_vp, _P(IndexStats)]),
    "anl_model_build_sharded": (_i32, [_vp, _i32, _u32, _u32]),
    "anl_shard_export_size": (_i32, [_vp, _vp, _P(_u64), _P(_u32)]),
    "anl_shard_export": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "anl_shard_merge": (_i32, [_vp, _vp, _u32, _vp, _vp, _vp, _vp, _u64, _u32, _P(_vp)]),
    "anl_shard_comm_id": (_i32, [_P(C.c_uint8)]),
    "anl_shard_comm_init": (_i32, [_vp, _P(C.c_uint8), _i32, _i32]),
    "anl_shard_comm_free": (None, [_vp]),
    "anl_shard_batch_step": (_i32, [_vp, _vp, _P(ShardStepStats), _P(_vp)]),
    "anl_shard_find_variants_batch": (_i32, [_vp, _cp, _P(_u64), _u64, _P(SearchParams), _P(_vp)]),
}

_lib = None


def lib():
    """Loads libanaliticcl_b200.so.  Fails loudly when it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(f"{SO_PATH} not found: build it with `python -m analiticcl_b200.build` "
                              "(the variant-lookup path has no CPU fallback)")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def pack(inputs):
    """list[str] -> (blob bytes, numpy uint64 offsets[n+1]) as the C ABI wants them."""
    enc = [s.encode("utf-8") for s in inputs]
    offs = np.zeros(len(enc) + 1, dtype=np.uint64)
    if enc:
        np.cumsum([len(e) for e in enc], out=offs[1:])
    return b"".join(enc), offs


def u64ptr(arr):
    return arr.ctypes.data_as(C.POINTER(C.c_uint64))


def variant_score(dist, freq, freq_weight):
    """VariantResult::score, src/types.rs:335-341 (freq_weight is an f32 widened to f64)."""
    fw = C.c_float(freq_weight).value
    if fw == 0.0:
        return dist
    return (dist + fw * freq) / (1.0 + fw)
