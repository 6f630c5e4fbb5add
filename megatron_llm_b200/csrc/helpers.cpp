// Dataset index builders (CPU, pybind11 + numpy).
//
// Same contracts as megatron/data/helpers.cpp of the reference (build_sample_idx :83-169, build_blending_indices
// :20-80, build_mapping :187-450, build_blocks_mapping :453-693) so cached index files are interchangeable, but
// organised differently: one generic "sentence span walker" shared by the BERT and ICT mappings that emits rows
// into a std::vector in a single pass (the reference runs every builder twice: count, then fill), 64-bit row
// storage chosen at run time, and numpy arrays that own their buffer through a capsule.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <iostream>
#include <limits>
#include <random>
#include <stdexcept>
#include <vector>

namespace py = pybind11;

namespace {

constexpr int32_t kLongSentenceLen = 512;

template <typename T>
py::array vector_to_array(std::vector<T>* v, int64_t rows, int64_t cols) {
  py::capsule owner(v, [](void* p) { delete reinterpret_cast<std::vector<T>*>(p); });
  return py::array_t<T>({rows, cols}, {static_cast<int64_t>(cols * sizeof(T)), static_cast<int64_t>(sizeof(T))},
                        v->data(), owner);
}

// Fisher-Yates over fixed-width rows, seeded exactly like the reference (mt19937_64(seed + 1), i from the end)
template <typename T>
void shuffle_rows(std::vector<T>& data, int64_t rows, int width, uint64_t seed) {
  std::mt19937_64 gen(seed);
  for (int64_t i = rows - 1; i > 0; --i) {
    const int64_t j = static_cast<int64_t>(gen() % static_cast<uint64_t>(i + 1));
    for (int c = 0; c < width; ++c) std::swap(data[i * width + c], data[j * width + c]);
  }
}

}  // namespace

// dataset_index[i] / dataset_sample_index[i]: which dataset sample i comes from and its index inside that dataset,
// chosen greedily so the running mix tracks `weights` (largest deficit first).
void build_blending_indices(py::array_t<uint8_t>& dataset_index, py::array_t<int64_t>& dataset_sample_index,
                            const py::array_t<double>& weights, const int32_t num_datasets, const int64_t size,
                            const bool verbose) {
  if (verbose) std::cout << "> building indices for blendable datasets ..." << std::endl;
  auto di = dataset_index.mutable_unchecked<1>();
  auto dsi = dataset_sample_index.mutable_unchecked<1>();
  auto w = weights.unchecked<1>();
  std::vector<int64_t> taken(num_datasets, 0);
  for (int64_t i = 0; i < size; ++i) {
    const double denom = std::max(static_cast<double>(i), 1.0);
    int64_t best = 0;
    double best_deficit = w[0] * denom - static_cast<double>(taken[0]);
    for (int64_t d = 1; d < num_datasets; ++d) {
      const double deficit = w[d] * denom - static_cast<double>(taken[d]);
      if (deficit > best_deficit) { best_deficit = deficit; best = d; }
    }
    di[i] = static_cast<uint8_t>(best);
    dsi[i] = taken[best]++;
  }
  if (verbose) {
    std::cout << " > sample ratios:" << std::endl;
    for (int64_t d = 0; d < num_datasets; ++d)
      std::cout << "   dataset " << d << ", input: " << w[d]
                << ", achieved: " << static_cast<double>(taken[d]) / static_cast<double>(size) << std::endl;
  }
}

// GPT sample index: documents (in doc_idx order) are laid end to end; sample k covers seq_length+1 tokens starting
// at (doc_idx position, offset) = row k and ending at row k+1 (rows overlap by one token).
py::array build_sample_idx(const py::array_t<int32_t>& sizes_, const py::array_t<int32_t>& doc_idx_,
                           const int32_t seq_length, const int32_t num_epochs, const int64_t tokens_per_epoch) {
  if (seq_length <= 1 || num_epochs <= 0 || tokens_per_epoch <= 1) throw std::invalid_argument("build_sample_idx");
  auto sizes = sizes_.unchecked<1>();
  auto doc_idx = doc_idx_.unchecked<1>();
  const int64_t num_samples = (static_cast<int64_t>(num_epochs) * tokens_per_epoch - 1) / seq_length;
  auto* out = new std::vector<int32_t>(2 * (num_samples + 1));
  int64_t cursor = 0;   // position in doc_idx
  int32_t offset = 0;   // token offset inside that document
  (*out)[0] = 0;
  (*out)[1] = 0;
  for (int64_t s = 1; s <= num_samples; ++s) {
    int32_t need = seq_length + 1;
    while (need > 0) {
      const int32_t avail = sizes[doc_idx[cursor]] - offset;
      if (avail >= need) {
        offset += need - 1;  // the last token is shared with the next sample
        need = 0;
      } else {
        need -= avail;
        ++cursor;
        offset = 0;
      }
    }
    (*out)[2 * s] = static_cast<int32_t>(cursor);
    (*out)[2 * s + 1] = offset;
  }
  return vector_to_array<int32_t>(out, num_samples + 1, 2);
}

namespace {

inline int32_t draw_target_len(int32_t short_seq_ratio, int32_t max_length, std::mt19937& gen) {
  if (short_seq_ratio == 0) return max_length;
  const auto r = gen();
  if (r % short_seq_ratio == 0) return 2 + static_cast<int32_t>(r % (max_length - 1));
  return max_length;
}

// Walk every document `num_epochs` times and cut it into spans of consecutive sentences.  `Policy` decides the
// target length of the next span, when a span may be closed, and what row to emit.
template <typename T, int WIDTH, typename Policy>
py::array walk_spans(const py::array_t<int64_t>& docs_, const py::array_t<int32_t>& sizes_, int32_t num_epochs,
                     uint64_t max_num_samples, int32_t min_num_sent, int32_t min_remaining_to_close, uint64_t seed,
                     bool verbose, Policy& policy) {
  auto docs = docs_.unchecked<1>();
  auto sizes = sizes_.unchecked<1>();
  const int64_t num_docs = docs_.shape(0) - 1;
  auto* rows = new std::vector<T>();
  uint64_t count = 0, empty_docs = 0, one_sent_docs = 0, long_sent_docs = 0;
  for (int32_t epoch = 0; epoch < num_epochs && count < max_num_samples; ++epoch) {
    policy.begin_epoch();
    for (int64_t doc = 0; doc < num_docs; ++doc) {
      const int64_t first = docs[doc], last = docs[doc + 1];
      int64_t remaining = last - first;
      if (epoch == 0) {
        empty_docs += (remaining == 0);
        one_sent_docs += (remaining == 1);
      }
      bool has_long = false;
      if (remaining >= policy.long_check_min) {
        for (int64_t s = first; s < last; ++s)
          if (sizes[s] > kLongSentenceLen) { has_long = true; break; }
        if (has_long && epoch == 0) ++long_sent_docs;
      }
      if (remaining < min_num_sent || has_long) continue;
      int64_t span_start = first;
      int32_t span_tokens = 0, span_sents = 0;
      int32_t target = policy.next_target(doc);
      for (int64_t s = first; s < last; ++s) {
        span_tokens += sizes[s];
        ++span_sents;
        --remaining;
        const bool full = span_tokens >= target && remaining >= min_remaining_to_close && span_sents >= min_num_sent;
        if (full || remaining == 0) {
          policy.emit(*rows, span_start, s + 1, doc, target);
          ++count;
          span_start = s + 1;
          span_tokens = 0;
          span_sents = 0;
          target = policy.next_target(doc);
        }
      }
    }
  }
  if (verbose) {
    std::cout << "   number of empty documents: " << empty_docs << "\n   number of documents with one sentence: "
              << one_sent_docs << "\n   number of documents with long sentences: " << long_sent_docs
              << "\n   will create mapping for " << count << " samples" << std::endl;
  }
  shuffle_rows<T>(*rows, static_cast<int64_t>(count), WIDTH, seed + 1);
  return vector_to_array<T>(rows, static_cast<int64_t>(count), WIDTH);
}

template <typename T>
struct BertPolicy {  // rows: (first sentence, one past last sentence, target sequence length)
  int32_t short_seq_ratio, max_seq_length;
  std::mt19937 gen;
  int64_t long_check_min = 2;
  BertPolicy(int32_t ratio, int32_t max_len, uint32_t seed) : short_seq_ratio(ratio), max_seq_length(max_len), gen(seed) {}
  void begin_epoch() {}
  int32_t next_target(int64_t) { return draw_target_len(short_seq_ratio, max_seq_length, gen); }
  void emit(std::vector<T>& rows, int64_t a, int64_t b, int64_t, int32_t target) {
    rows.push_back(static_cast<T>(a));
    rows.push_back(static_cast<T>(b));
    rows.push_back(static_cast<T>(target));
  }
};

template <typename T>
struct BlockPolicy {  // rows: (first sentence, one past last sentence, document, block id within the epoch)
  const py::detail::unchecked_reference<int32_t, 1> titles;
  int32_t max_seq_length;
  int32_t block_id = 0;
  int64_t long_check_min;
  BlockPolicy(const py::array_t<int32_t>& t, int32_t max_len, int32_t min_sent)
      : titles(t.unchecked<1>()), max_seq_length(max_len), long_check_min(min_sent) {}
  void begin_epoch() { block_id = 0; }
  int32_t next_target(int64_t doc) { return max_seq_length - titles[doc]; }
  void emit(std::vector<T>& rows, int64_t a, int64_t b, int64_t doc, int32_t) {
    rows.push_back(static_cast<T>(a));
    rows.push_back(static_cast<T>(b));
    rows.push_back(static_cast<T>(doc));
    rows.push_back(static_cast<T>(block_id++));
  }
};

}  // namespace

// BERT-style sentence-span samples.  NOTE (reference parity): the reference re-seeds its generator for a counting
// pass and a filling pass, so one pass with the same seed yields the same sequence of target lengths.
py::array build_mapping(const py::array_t<int64_t>& docs, const py::array_t<int32_t>& sizes, const int num_epochs,
                        const uint64_t max_num_samples, const int max_seq_length, const double short_seq_prob,
                        const int seed, const bool verbose, const int32_t min_num_sent) {
  if (num_epochs <= 0 || max_seq_length <= 1 || short_seq_prob < 0.0 || short_seq_prob > 1.0 || seed <= 0)
    throw std::invalid_argument("build_mapping: bad arguments");
  const int32_t ratio = short_seq_prob > 0 ? static_cast<int32_t>(std::round(1.0 / short_seq_prob)) : 0;
  if (sizes.size() > static_cast<py::ssize_t>(std::numeric_limits<uint32_t>::max())) {
    BertPolicy<uint64_t> p(ratio, max_seq_length, seed);
    return walk_spans<uint64_t, 3>(docs, sizes, num_epochs, max_num_samples, min_num_sent, 2, seed, verbose, p);
  }
  BertPolicy<uint32_t> p(ratio, max_seq_length, seed);
  return walk_spans<uint32_t, 3>(docs, sizes, num_epochs, max_num_samples, min_num_sent, 2, seed, verbose, p);
}

// ICT/REALM blocks: spans of sentences whose length is bounded by max_seq_length minus the document's title length.
py::array build_blocks_mapping(const py::array_t<int64_t>& docs, const py::array_t<int32_t>& sizes,
                               const py::array_t<int32_t>& titles_sizes, const int num_epochs,
                               const uint64_t max_num_samples, const int max_seq_length, const int seed,
                               const bool verbose, const bool use_one_sent_blocks) {
  if (num_epochs <= 0 || max_seq_length <= 1 || seed <= 0) throw std::invalid_argument("build_blocks_mapping");
  const int32_t min_sent = use_one_sent_blocks ? 1 : 2;
  if (sizes.size() > static_cast<py::ssize_t>(std::numeric_limits<uint32_t>::max())) {
    BlockPolicy<uint64_t> p(titles_sizes, max_seq_length, min_sent);
    return walk_spans<uint64_t, 4>(docs, sizes, num_epochs, max_num_samples, min_sent, min_sent, seed, verbose, p);
  }
  BlockPolicy<uint32_t> p(titles_sizes, max_seq_length, min_sent);
  return walk_spans<uint32_t, 4>(docs, sizes, num_epochs, max_num_samples, min_sent, min_sent, seed, verbose, p);
}

PYBIND11_MODULE(_helpers_b200, m) {
  m.def("build_mapping", &build_mapping);
  m.def("build_blocks_mapping", &build_blocks_mapping);
  m.def("build_sample_idx", &build_sample_idx);
  m.def("build_blending_indices", &build_blending_indices);
}
