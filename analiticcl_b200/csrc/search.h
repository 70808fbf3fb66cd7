// search.h -- host segmentation for find_all_matches (src/search.rs:190-336).
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "hostpool.h"

namespace anl {

enum { BOUNDARY_NONE = 0, BOUNDARY_WEAK = 1, BOUNDARY_NORMAL = 2, BOUNDARY_HARD = 3 };  // src/search.rs:176-185

struct Boundary {
  size_t begin, end;  // byte offsets
  int strength;
};
struct SegmentSpan {
  size_t begin, end;  // byte offsets into the text
  uint32_t n;         // n-gram order
};
// All segments of a text: batch by batch (a batch = the segments between two hard boundaries,
// src/lib.rs:1840-1903), orders 1..max_ngram ascending inside a batch -- so a batch's unigrams come first.
struct SegmentedText {
  PodBuffer<SegmentSpan> segs;      // (recycled blocks: no page faults after the first call)
  PodBuffer<uint64_t> batch_first;  // index of each batch's first segment; n_batches + 1 entries
};

// A batch = the text between two hard boundaries (src/lib.rs:1822): it starts at byte `begin`, its boundaries are
// bounds[begin_index .. end_index] and bounds[end_index] is the hard boundary that closes it.
struct BatchDesc {
  size_t begin, begin_index, end_index;
};
void list_batches(const std::vector<Boundary>& bounds, std::vector<BatchDesc>* out);

const std::vector<Boundary>& find_boundaries(const std::string& text);  // valid until the calling thread's next call
void find_match_ngrams(const std::string& text, const Boundary* bounds, size_t nbounds, uint32_t order, size_t begin, size_t end,
                       std::vector<SegmentSpan>* out);
void segment_text(const std::string& text, uint32_t max_ngram, SegmentedText* out, PodBuffer<Boundary>* bounds_out = nullptr,
                  PodBuffer<BatchDesc>* batches_out = nullptr);
// The same segmentation computed on `device` (gpu_segment.cu): boundary detection, strengths, batches and n-gram spans
// as kernels over the text in HBM.  false + *err on a CUDA error or a text of 2 GiB and more.
// bounds_out / batches_out (optional): the boundaries (find_boundaries) and batch descriptors (list_batches) as well.
bool segment_text_device(int device, const std::string& text, uint32_t max_ngram, SegmentedText* out, std::string* err,
                         PodBuffer<Boundary>* bounds_out = nullptr, PodBuffer<BatchDesc>* batches_out = nullptr);
// Which producer a text of `len` bytes gets: the device for running text (>= 64 KiB, ANL_SEGMENT_DEVICE_MIN), the host
// loop for short strings where a kernel launch would be the whole cost; ANL_SEGMENT=host|device forces one.
bool segment_on_device(size_t len);

// A text's segmentation as the match set keeps it for anl_match_set_consolidate (24 B per segment and per boundary):
// the producer's spans plus the boundaries and batch descriptors they were cut from.
struct Segmentation {
  SegmentedText st;
  PodBuffer<Boundary> bounds;  // (recycled blocks like the spans: a 100 MB array is not page-faulted in on every call)
  PodBuffer<BatchDesc> batches;
  size_t text_len = 0;
  uint32_t max_ngram = 0;
};
// device < 0 or a short text: the host producer; else the device's.  false + *err on a CUDA error.
bool segment_any(int device, const std::string& text, uint32_t max_ngram, Segmentation* out, std::string* err);

// One step of the most likely sequence of a batch: segment `seg` (index into the batch's segments) rendered as its
// variant `variant`, or copied from the input (out of vocabulary) when variant < 0.
struct SequenceStep {
  uint32_t seg;
  int32_t variant;
};
// Per segment of the batch: how many variants its lookup returned (0 also for segments that were not looked up)
// and the f64 score (VariantResult::score, src/types.rs:335-341) of variant j at score[first[seg] + j].
struct BatchVariants {
  const uint32_t* count;
  const uint64_t* first;
  const double* score;
  const uint64_t* vocab_id = nullptr;  // per variant, parallel to score (needed by the language model / context rules)
};
// most_likely_sequence (src/lib.rs:2088-2495) for a model without language model and context rules: the
// lowest-cost path through the batch's segment lattice.  Returns false when the lattice has no arcs (the
// reference then returns the matches unchanged, :2261-2267).
bool most_likely_sequence(const Boundary* bounds, size_t nbounds, size_t end_offset, const SegmentSpan* segs, size_t nsegs,
                          const BatchVariants& variants, std::vector<SequenceStep>* out);
// The whole of most_likely_sequence (sequence.cpp): the max_seq shortest paths, each scored by the model's language
// model (lm_score, src/lib.rs:2570-2674) and context rules (test_context_rules, :2501-2566), best weighted sum kept
// (:2381-2425).  `text` is the whole input (boundary tokens are read from it).  out_tags (optional): per step the tags
// and sequence numbers its context rules assign (Match.tag / Match.seqnr).  Falls back to the single cheapest path when
// the model has neither (or hm == nullptr).
class HostModel;
struct SequenceWeights {  // src/types.rs:139-156
  uint64_t max_seq = 250;
  float lm_weight = 1.0f, variantmodel_weight = 3.0f, contextrules_weight = 1.0f;
};
struct StepTags {
  std::vector<uint16_t> tag;
  std::vector<uint8_t> seqnr;
};
bool most_likely_sequence_full(const HostModel* hm, const std::string& text, const Boundary* bounds, size_t nbounds, size_t end_offset,
                               const SegmentSpan* segs, size_t nsegs, const BatchVariants& variants, const SequenceWeights& w,
                               std::vector<SequenceStep>* out, std::vector<StepTags>* out_tags);
std::vector<uint64_t> byte_to_codepoint_map(const std::string& text);

}  // namespace anl
