// search.h -- host segmentation for find_all_matches (src/search.rs:190-336).
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

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
struct SpanBatch {  // the segments between two hard boundaries, orders 1..max_ngram (src/lib.rs:1840-1903)
  std::vector<SegmentSpan> segments;
};

std::vector<Boundary> find_boundaries(const std::string& text);
std::vector<SegmentSpan> find_match_ngrams(const std::string& text, const Boundary* bounds, size_t nbounds, uint32_t order,
                                           size_t begin, size_t end);
std::vector<SpanBatch> segment_text(const std::string& text, uint32_t max_ngram);
std::vector<uint64_t> byte_to_codepoint_map(const std::string& text);

}  // namespace anl
