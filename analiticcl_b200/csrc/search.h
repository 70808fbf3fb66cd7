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

const std::vector<Boundary>& find_boundaries(const std::string& text);  // valid until the calling thread's next call
void find_match_ngrams(const std::string& text, const Boundary* bounds, size_t nbounds, uint32_t order, size_t begin, size_t end,
                       std::vector<SegmentSpan>* out);
void segment_text(const std::string& text, uint32_t max_ngram, SegmentedText* out);
std::vector<uint64_t> byte_to_codepoint_map(const std::string& text);

}  // namespace anl
