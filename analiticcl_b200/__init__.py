"""analiticcl_b200 -- B200-native variant lookup behind analiticcl's `VariantModel` API.

This module mirrors the surface of the reference's Python binding for the variant-lookup path
(bindings/python/src/lib.rs, analiticcl.pyi): `VariantModel`, `Weights`, `SearchParameters`,
`VocabParams`, with the same method names, keyword arguments, result dictionaries and error types,
so `bindings/python/tests/tests.py` and `bindings/python/examples/example.py` read the same against
either implementation.  It is a thin ctypes layer over the C ABI (include/analiticcl_b200.h); all
work happens in libanaliticcl_b200.so (hand-written sm_100a kernels + host glue).

There is no CPU fallback: importing works anywhere (the library only needs the CUDA driver when
`build()` uploads the index), but `VariantModel.build()` raises if no B200-class GPU is present.
"""
import ctypes as C
import os
import sys

from . import _capi
from ._capi import lib as _lib

__all__ = ["VariantModel", "Weights", "SearchParameters", "VocabParams"]


class Weights:
    """Weights(ld=0.5, lcs=0.125, prefix=0.125, suffix=0.125, case=0.125) -- bindings/python/src/lib.rs:10-113"""

    _FIELDS = ("ld", "lcs", "prefix", "suffix", "case")

    def __init__(self, **kwargs):
        w = _capi.Weights()
        _lib().anl_weights_default(C.byref(w))
        self.ld, self.lcs, self.prefix, self.suffix, self.case = w.ld, w.lcs, w.prefix, w.suffix, w.case_
        for key, value in kwargs.items():
            if key in self._FIELDS:
                if value is not None:
                    setattr(self, key, float(value))
            else:
                print(f"Ignored unknown kwargs option {key}", file=sys.stderr)

    def to_dict(self):
        return {k: getattr(self, k) for k in self._FIELDS}

    def _c(self):
        return _capi.Weights(self.ld, self.lcs, self.prefix, self.suffix, self.case)


def _distance_threshold(value):
    """extract_distance_threshold, bindings/python/src/lib.rs:116-134 (+ FromStr, src/types.rs:85-108)."""
    err = ("Must be an integer expressing an absolute value, or float in range 0-1 expressing a ratio. "
           "Or a two-tuple expression a ratio with an absolute limit (float, int)")
    if isinstance(value, (tuple, list)) and len(value) == 2:
        return _capi.Threshold(_capi.THRESHOLD_RATIO_WITH_LIMIT, float(value[0]), int(value[1]))
    if isinstance(value, bool):
        raise ValueError(err)
    if isinstance(value, int):
        if not 0 <= value <= 255:
            raise ValueError(err)
        return _capi.Threshold(_capi.THRESHOLD_ABSOLUTE, 0.0, value)
    if isinstance(value, float):
        return _capi.Threshold(_capi.THRESHOLD_RATIO, value, 0)
    if isinstance(value, str):
        s = value
        try:
            if ";" in s:
                a, b = s.split(";")
                return _capi.Threshold(_capi.THRESHOLD_RATIO_WITH_LIMIT, float(a), int(b))
            try:
                v = int(s)
                if 0 <= v <= 255:
                    return _capi.Threshold(_capi.THRESHOLD_ABSOLUTE, 0.0, v)
            except ValueError:
                pass
            f = float(s)
            if 0.0 <= f <= 1.0:
                return _capi.Threshold(_capi.THRESHOLD_RATIO, f, 0)
        except ValueError:
            pass
        raise ValueError(f"Unable to convert from string ({value}). " + err)
    raise ValueError(err)


def _threshold_value(t):
    if t.kind == _capi.THRESHOLD_ABSOLUTE:
        return int(t.value)
    if t.kind == _capi.THRESHOLD_RATIO:
        return float(t.ratio)
    return (float(t.ratio), int(t.value))


class SearchParameters:
    """SearchParameters(**kwargs) -- bindings/python/src/lib.rs:136-446.

    Defaults are the library defaults (max_anagram_distance=3, max_edit_distance=3, max_matches=20,
    score_threshold=0.25, cutoff_threshold=2.0, max_ngram=3, freq_weight=0.0, ...), unknown keyword
    arguments are ignored with a warning on stderr, exactly like the reference binding.
    """

    _SIMPLE = {
        "max_matches": ("max_matches", int), "score_threshold": ("score_threshold", float),
        "cutoff_threshold": ("cutoff_threshold", float), "max_ngram": ("max_ngram", int), "max_seq": ("max_seq", int),
        "single_thread": ("single_thread", bool), "unicodeoffsets": ("unicodeoffsets", bool),
        "freq_weight": ("freq_weight", float), "lm_weight": ("lm_weight", float),
        "contextrules_weight": ("contextrules_weight", float), "variantmodel_weight": ("variantmodel_weight", float),
        "context_weight": ("context_weight", float), "consolidate_matches": ("consolidate_matches", bool),
    }

    def __init__(self, **kwargs):
        self.data = _capi.SearchParams()
        _lib().anl_search_params_default(C.byref(self.data))
        for key, value in kwargs.items():
            if key in ("max_anagram_distance", "max_edit_distance"):
                try:
                    setattr(self.data, key, _distance_threshold(value))
                except ValueError as e:  # the reference prints and keeps the default
                    print(e, file=sys.stderr)
            elif key == "stop_at_exact_match":
                if value is not None:
                    self.data.stop_criterion = _capi.STOP_AT_EXACT_MATCH if value else _capi.STOP_EXHAUSTIVE
            elif key in self._SIMPLE:
                field, conv = self._SIMPLE[key]
                if value is None:
                    print(f"No value specified for {key} parameter", file=sys.stderr)
                else:
                    setattr(self.data, field, conv(value))
            else:
                print(f"Ignored unknown kwargs option {key}", file=sys.stderr)

    # attributes with the setters the reference binding has (#[setter], bindings/python/src/lib.rs:262-446): the two
    # thresholds, max_matches, max_ngram, max_seq, single_thread, the five weights, consolidate_matches, unicodeoffsets,
    # stop_at_exact_match; score_threshold and cutoff_threshold are read-only there too
    def _set_threshold(self, field, value):
        setattr(self.data, field, _distance_threshold(value))  # (ValueError on a bad threshold, like the reference)

    max_anagram_distance = property(lambda self: _threshold_value(self.data.max_anagram_distance),
                                    lambda self, v: self._set_threshold("max_anagram_distance", v))
    max_edit_distance = property(lambda self: _threshold_value(self.data.max_edit_distance),
                                 lambda self, v: self._set_threshold("max_edit_distance", v))
    max_matches = property(lambda self: int(self.data.max_matches), lambda self, v: setattr(self.data, "max_matches", int(v)))
    score_threshold = property(lambda self: float(self.data.score_threshold))
    cutoff_threshold = property(lambda self: float(self.data.cutoff_threshold))
    max_ngram = property(lambda self: int(self.data.max_ngram), lambda self, v: setattr(self.data, "max_ngram", int(v)))
    max_seq = property(lambda self: int(self.data.max_seq), lambda self, v: setattr(self.data, "max_seq", int(v)))
    single_thread = property(lambda self: bool(self.data.single_thread), lambda self, v: setattr(self.data, "single_thread", int(bool(v))))
    unicodeoffsets = property(lambda self: bool(self.data.unicodeoffsets), lambda self, v: setattr(self.data, "unicodeoffsets", int(bool(v))))
    freq_weight = property(lambda self: float(self.data.freq_weight), lambda self, v: setattr(self.data, "freq_weight", float(v)))
    lm_weight = property(lambda self: float(self.data.lm_weight), lambda self, v: setattr(self.data, "lm_weight", float(v)))
    contextrules_weight = property(lambda self: float(self.data.contextrules_weight),
                                   lambda self, v: setattr(self.data, "contextrules_weight", float(v)))
    variantmodel_weight = property(lambda self: float(self.data.variantmodel_weight),
                                   lambda self, v: setattr(self.data, "variantmodel_weight", float(v)))
    context_weight = property(lambda self: float(self.data.context_weight), lambda self, v: setattr(self.data, "context_weight", float(v)))
    consolidate_matches = property(lambda self: bool(self.data.consolidate_matches),
                                   lambda self, v: setattr(self.data, "consolidate_matches", int(bool(v))))
    stop_at_exact_match = property(
        lambda self: self.data.stop_criterion == _capi.STOP_AT_EXACT_MATCH,
        lambda self, v: setattr(self.data, "stop_criterion", _capi.STOP_AT_EXACT_MATCH if v else _capi.STOP_EXHAUSTIVE))

    def to_dict(self):
        keys = ["max_anagram_distance", "max_edit_distance", "max_matches", "score_threshold", "cutoff_threshold",
                "max_ngram", "max_seq", "single_thread", "freq_weight", "lm_weight", "contextrules_weight",
                "variantmodel_weight", "consolidate_matches", "unicodeoffsets"]
        return {k: getattr(self, k) for k in keys}


class VocabParams:
    """VocabParams(**kwargs) -- bindings/python/src/lib.rs:448-546."""

    def __init__(self, **kwargs):
        self.data = _capi.VocabParams()
        _lib().anl_vocab_params_default(C.byref(self.data))
        for key, value in kwargs.items():
            if key == "text_column":
                if value is not None:
                    self.data.text_column = int(value)
            elif key == "freq_column":
                if value is not None:
                    self.data.freq_column = int(value)
            elif key == "index":
                if value is not None:
                    self.data.index = int(value)
            elif key == "freqhandling":
                m = {"sum": _capi.FREQ_SUM, "max": _capi.FREQ_MAX, "min": _capi.FREQ_MIN, "replace": _capi.FREQ_REPLACE}
                if value in m:
                    self.data.freq_handling = m[value]
                else:
                    print(f"WARNING: Ignored unknown value for VocabParams.freqhandling ({value})", file=sys.stderr)
            elif key == "vocabtype":
                m = {"NONE": _capi.VOCAB_NONE, "INDEXED": _capi.VOCAB_INDEXED, "LM": _capi.VOCAB_LM,
                     "TRANSPARENT": _capi.VOCAB_TRANSPARENT | _capi.VOCAB_INDEXED}
                if value in m:
                    self.data.vocab_type = m[value]
                else:
                    print(f"WARNING: Ignored unknown value for VocabParams.vocabtype ({value})", file=sys.stderr)
            else:
                print(f"WARNING: Ignored unknown VocabParams kwargs option {key}", file=sys.stderr)

    text_column = property(lambda self: int(self.data.text_column),
                           lambda self, v: setattr(self.data, "text_column", int(v)))
    freq_column = property(lambda self: None if self.data.freq_column < 0 else int(self.data.freq_column),
                           lambda self, v: setattr(self.data, "freq_column", -1 if v is None else int(v)))
    index = property(lambda self: int(self.data.index), lambda self, v: setattr(self.data, "index", int(v)))


def _check(status):
    if status != _capi.OK:
        msg = _lib().anl_last_error().decode("utf-8", "replace")
        if status in (_capi.ERR_INVALID, _capi.ERR_EMPTY_INPUT):
            raise ValueError(msg)
        raise RuntimeError(msg)


class VariantModel:
    """VariantModel(alphabet_file, weights, debug=0) -- bindings/python/src/lib.rs:548-812.

    read_lexicon, read_vocabulary, read_lm, add_to_vocabulary, read_variants, add_variant, read_confusablelist,
    set_confusables_before_pruning, read_contextrules, add_contextrule, build, __contains__, find_variants,
    find_variants_par, find_all_matches (with the whole sequence stage: variant model, language model, context rules),
    plus learn_variants, which the reference's library has and its binding lacks.
    """

    def __init__(self, alphabet_file, weights=None, debug=0, alphabet_tsv=None):
        self._h = C.c_void_p()
        w = (weights or Weights())._c()
        if alphabet_tsv is not None:  # VariantModel::new_with_alphabet (src/lib.rs:132)
            raw = alphabet_tsv.encode("utf-8") if isinstance(alphabet_tsv, str) else bytes(alphabet_tsv)
            _check(_lib().anl_model_new_from_tsv(raw, len(raw), C.byref(w), int(debug), C.byref(self._h)))
        else:
            _check(_lib().anl_model_new(os.fsencode(alphabet_file), C.byref(w), int(debug), C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib().anl_model_free(h)
            except Exception:
                pass
            self._h = None

    # -- loading ---------------------------------------------------------------------------------
    def read_lexicon(self, filename):
        self.read_vocabulary(filename, VocabParams())

    def read_vocabulary(self, filename, params=None):
        p = (params or VocabParams()).data
        _check(_lib().anl_model_read_vocabulary(self._h, os.fsencode(filename), C.byref(p)))

    def add_to_vocabulary(self, text, frequency=None, params=None):
        p = (params or VocabParams()).data
        raw = text.encode("utf-8")
        vid = C.c_uint64()
        _check(_lib().anl_model_add_to_vocabulary(self._h, raw, len(raw), frequency is not None, int(frequency or 0),
                                                  C.byref(p), C.byref(vid)))
        return vid.value

    def add_variant(self, ref_id, variant, score, frequency=None, params=None):
        """add_variant (src/lib.rs:460): link `variant` (added to the vocabulary with `params`) to the entry `ref_id`."""
        p = (params or VocabParams()).data
        raw = variant.encode("utf-8")
        added = C.c_int32()
        _check(_lib().anl_model_add_variant(self._h, int(ref_id), raw, len(raw), float(score), frequency is not None,
                                            int(frequency or 0), C.byref(p), C.byref(added)))
        return bool(added.value)

    def read_variants(self, filename, transparent=False, params=None):
        """read_variants (bindings/python/src/lib.rs:669-684): weighted variant list; transparent=True for an error list."""
        p = (params or VocabParams()).data
        _check(_lib().anl_model_read_variants(self._h, os.fsencode(filename), C.byref(p), int(bool(transparent))))

    def read_confusablelist(self, filename):
        _check(_lib().anl_model_read_confusablelist(self._h, os.fsencode(filename)))

    def add_to_confusables(self, editscript, weight=1.0):
        _check(_lib().anl_model_add_to_confusables(self._h, editscript.encode("utf-8"), float(weight)))

    def set_confusables_before_pruning(self):
        _lib().anl_model_set_confusables_before_pruning(self._h)

    def learn_variants(self, inputs, params, strict=True, auto_build=True):
        """VariantModel::learn_variants (src/lib.rs:1062-1139; the reference's Python binding does not expose it): look
        the inputs up and store the found variants in the model.  Returns the number of variant links added."""
        blob, offs = _capi.pack(list(inputs))
        count = C.c_uint64()
        _check(_lib().anl_model_learn_variants(self._h, blob, _capi.u64ptr(offs), len(offs) - 1, C.byref(params.data),
                                               int(bool(strict)), int(bool(auto_build)), C.byref(count)))
        return count.value

    def _vocab_size(self):
        return _lib().anl_model_vocab_size(self._h)

    def vocab_links(self, vid):
        """([(reference id, score), ...] this entry is a variant of, [variant ids this entry is the reference for])."""
        ids, sc = (C.c_uint64 * 256)(), (C.c_double * 256)()
        n = _lib().anl_debug_vocab_links(self._h, vid, 0, ids, sc, 256)
        of = [(ids[i], sc[i]) for i in range(n)]
        n = _lib().anl_debug_vocab_links(self._h, vid, 1, ids, sc, 256)
        return of, [ids[i] for i in range(n)]

    def read_lm(self, filename):
        """bindings/python/src/lib.rs:659-667: read_vocabulary with VocabType::LM -- n-grams `w1 w2 ..<TAB>count` for
        the language-model term of find_all_matches (collected by build())."""
        self.read_vocabulary(filename, VocabParams(vocabtype="LM"))

    def read_contextrules(self, filename):
        """bindings/python/src/lib.rs:691-696 (src/lib.rs:570-656)."""
        _check(_lib().anl_model_read_contextrules(self._h, os.fsencode(filename)))

    def add_contextrule(self, pattern, score, tag=(), tagoffset=()):
        """bindings/python/src/lib.rs:630-643 (src/lib.rs:658-765)."""
        tg = [t.encode("utf-8") for t in tag]
        to = [t.encode("utf-8") for t in tagoffset]
        a = (C.c_char_p * max(1, len(tg)))(*tg)
        b = (C.c_char_p * max(1, len(to)))(*to)
        _check(_lib().anl_model_add_contextrule(self._h, pattern.encode("utf-8"), float(score), a, len(tg), b, len(to)))

    def have_lm(self):
        return bool(_lib().anl_model_have_lm(self._h))

    def tags(self):
        return [_lib().anl_model_tag_name(self._h, i).decode("utf-8") for i in range(_lib().anl_model_tag_count(self._h))]

    def build(self, device=-1, devices=None, gpu_build=None):
        """Build the anagram index and upload it to the GPU (`device` = CUDA ordinal, -1 = current).  With
        `devices=[0, 1, ...]` every listed GPU of this process gets a replica and each lookup call is spread over
        all of them (anl_model_build_multi).  `gpu_build=True/False` forces the index construction onto the device /
        the host cores (default: the device for lexicons of a million entries and more)."""
        if devices is not None:
            arr = (C.c_int32 * len(devices))(*[int(d) for d in devices])
            _check(_lib().anl_model_build_multi(self._h, arr, len(devices)))
        elif gpu_build is not None:
            _check(_lib().anl_model_build_on(self._h, int(device), int(bool(gpu_build))))
        else:
            _check(_lib().anl_model_build(self._h, int(device)))

    def device_count(self):
        return _lib().anl_model_device_count(self._h)

    def save_index(self, filename):
        """Write the built index to `filename` (not in the reference: its build takes seconds; see the C header)."""
        _check(_lib().anl_model_save_index(self._h, os.fsencode(filename)))

    def load_index(self, filename, device=-1):
        """Instead of build(): read an index written by save_index for the same alphabet and vocabulary (same order)
        and upload it to the GPU.  Raises RuntimeError on a foreign, mismatching or corrupt file."""
        _check(_lib().anl_model_load_index(self._h, os.fsencode(filename), int(device)))

    def __contains__(self, text):
        raw = text.encode("utf-8")
        return bool(_lib().anl_model_has(self._h, raw, len(raw)))

    # -- helpers -----------------------------------------------------------------------------------
    def _lexicon_names(self):
        n = _lib().anl_model_lexicon_count(self._h)
        return [_lib().anl_model_lexicon_name(self._h, i).decode("utf-8") for i in range(n)]

    def _vocab(self, vid):
        info = _capi.VocabInfo()
        _check(_lib().anl_model_get_vocab(self._h, vid, C.byref(info)))
        return info

    def _variant_dict(self, v, freq_weight, lexnames):
        """variantresult_to_dict, bindings/python/src/lib.rs:554-588."""
        info = self._vocab(v.vocab_id)
        d = {
            "text": C.string_at(info.text, info.text_len).decode("utf-8"),
            "score": _capi.variant_score(v.dist_score, v.freq_score, freq_weight),
            "dist_score": v.dist_score,
            "freq_score": v.freq_score,
        }
        if v.via != _capi.NO_VIA:
            vi = self._vocab(v.via)
            d["via"] = C.string_at(vi.text, vi.text_len).decode("utf-8")
        d["lexicons"] = [name for i, name in enumerate(lexnames) if info.lexindex & (1 << i)]
        return d

    def _run(self, inputs, params):
        blob, offs = _capi.pack(inputs)
        rs = C.c_void_p()
        _check(_lib().anl_find_variants_batch(self._h, blob, _capi.u64ptr(offs), len(inputs), C.byref(params.data),
                                              C.byref(rs)))
        return rs

    # -- lookup --------------------------------------------------------------------------------------
    def find_variants(self, input, params):
        """bindings/python/src/lib.rs:704-717: list of variant dicts for one input string."""
        return self.find_variants_par([input], params)[0]["variants"]

    def find_variants_par(self, input, params):
        """bindings/python/src/lib.rs:720-749: [{"input": str, "variants": [dict, ...]}, ...]."""
        inputs = list(input)
        rs = self._run(inputs, params)
        try:
            lexnames = self._lexicon_names()
            fw = params.data.freq_weight
            out = []
            cnt = C.c_uint64()
            for i, text in enumerate(inputs):
                ptr = _lib().anl_result_set_get(rs, i, C.byref(cnt))
                out.append({"input": text,
                            "variants": [self._variant_dict(ptr[j], fw, lexnames) for j in range(cnt.value)]})
            return out
        finally:
            _lib().anl_result_set_free(rs)

    def find_variants_raw(self, inputs, params, with_via=False):
        """Batch lookup returning plain tuples: [[(vocab_id, dist_score, freq_score[, via or None]), ...], ...]."""
        rs = self._run(list(inputs), params)
        try:
            n = len(inputs)
            offs = _lib().anl_result_set_offsets(rs)
            var = _lib().anl_result_set_variants(rs)
            if with_via:
                return [[(var[j].vocab_id, var[j].dist_score, var[j].freq_score, None if var[j].via == _capi.NO_VIA else var[j].via)
                         for j in range(offs[i], offs[i + 1])] for i in range(n)]
            return [[(var[j].vocab_id, var[j].dist_score, var[j].freq_score) for j in range(offs[i], offs[i + 1])]
                    for i in range(n)]
        finally:
            _lib().anl_result_set_free(rs)

    def find_variants_partitioned(self, inputs, params, group=None, dst=0):
        """Query-partitioned lookup over the ranks of a torch.distributed job (one process per GPU, every rank holding a
        replica of this model; SURVEY 8e mode 1): each rank looks up its contiguous slice, rank `dst` gets the whole
        list in input order (others None).  No data-path collective -- only the gather of the results."""
        from . import parallel
        return parallel.lookup_partitioned(list(inputs), lambda qs: self.find_variants_raw(qs, params), group=group, dst=dst)

    def find_all_matches(self, text, params):
        """bindings/python/src/lib.rs:752-805: [{"input", "offset": {"begin","end"}, "variants": [...]}, ...]
        with the selected variant first (+ "tag" / "seqnr" where context rules tagged the match).  With max_ngram > 1, a
        language model or context rules the matches are the most likely sequence per hard-delimited batch
        (most_likely_sequence, src/lib.rs:1912-1924, 2088-2495), like the reference.  `consolidate_matches=False` (which
        the reference's library carries but never reads) returns the producer's view instead: every looked-up segment
        of every order."""
        raw = text.encode("utf-8")
        ms = C.c_void_p()
        _check(_lib().anl_find_all_matches(self._h, raw, len(raw), C.byref(params.data), C.byref(ms)))
        try:
            sequence = params.data.max_ngram > 1 or self.have_lm() or _lib().anl_model_contextrule_count(self._h) > 0  # :1912
            if sequence and params.data.consolidate_matches:
                best = C.c_void_p()
                _check(_lib().anl_model_consolidate(self._h, ms, raw, len(raw), C.byref(params.data), C.byref(best)))
                _lib().anl_match_set_free(ms)
                ms = best
            return self._match_list(ms, text, raw, params)
        finally:
            _lib().anl_match_set_free(ms)

    def _match_list(self, ms, text, raw, params):
        lexnames = self._lexicon_names()
        fw = params.data.freq_weight
        cpmode = bool(params.data.unicodeoffsets)
        out = []
        m = _capi.Match()
        tagnames = self.tags()
        tg, sq = C.POINTER(C.c_uint16)(), C.POINTER(C.c_uint8)()
        for i in range(_lib().anl_match_set_len(ms)):
            _check(_lib().anl_match_set_get(ms, i, C.byref(m)))
            if not m.variants and params.data.max_ngram > 1 and m.n > 1:
                continue  # higher-order segment skipped as redundant: no lookup happened
            seg = text[m.begin:m.end] if cpmode else raw[m.begin:m.end].decode("utf-8")
            variants = [self._variant_dict(m.variants[j], fw, lexnames) for j in range(m.n_variants)]
            if m.selected > 0:
                variants.insert(0, variants.pop(m.selected))
            d = {"input": seg, "offset": {"begin": int(m.begin), "end": int(m.end)}}
            nt = _lib().anl_match_set_tags(ms, i, C.byref(tg), C.byref(sq)) if tagnames else 0
            if nt:  # bindings/python/src/lib.rs:768-783
                d["tag"] = [tagnames[tg[k]] for k in range(nt)]
                d["seqnr"] = [int(sq[k]) for k in range(nt)]
            d["variants"] = variants
            out.append(d)
        return out

    # -- introspection used by tests / benchmarks -------------------------------------------------------
    def vocab_text(self, vid):
        info = self._vocab(vid)
        return C.string_at(info.text, info.text_len).decode("utf-8")

    def index_size(self):
        return _lib().anl_model_index_size(self._h)

    def instance_count(self):
        return _lib().anl_model_instance_count(self._h)

    def anagram_count_of_length(self, cc):
        return _lib().anl_model_anagram_count_of_length(self._h, cc)

    def max_key_bits(self):
        return _lib().anl_model_max_key_bits(self._h)

    def anahash(self, text):
        raw = text.encode("utf-8")
        limbs = (C.c_uint64 * 64)()
        n = _lib().anl_anahash(self._h, raw, len(raw), limbs, 64)
        return sum(int(limbs[i]) << (64 * i) for i in range(n))

    def normalize(self, text):
        raw = text.encode("utf-8")
        buf = (C.c_uint8 * 1024)()
        n = _lib().anl_normalize(self._h, raw, len(raw), buf, 1024)
        return list(buf[:n])

    def index_stats(self):
        st = _capi.IndexStats()
        _check(_lib().anl_model_index_stats(self._h, C.byref(st)))
        return {f: getattr(st, f) for f, _ in st._fields_}
