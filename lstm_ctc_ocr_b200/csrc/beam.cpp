// CTC prefix beam search, host side.
//
// Replaces tf.nn.ctc_beam_search_decoder(logits, seq_len, merge_repeated=True) at lib/networks/network.py:656 and
// lib/lstm/test.py:30 (beam_width 100, top_paths 1, blank = C-1) + sparse_tensor_to_dense(default 0) at network.py:657.
// The reference's op is a CPU-only TensorFlow kernel that runs at validation / evaluation time only; so is this one: it
// takes HOST logits (the caller copies the [T,N,C] logits back once) and spreads the utterances over host threads.  The hot
// path's decoder is the greedy kernel in ctc.cu; this entry point exists for exact reproduction of the reference's decode on
// outputs that are not peaked.
//
// Algorithm [upstream-memory: tensorflow/core/util/ctc/ctc_beam_search.h, CTCBeamSearchDecoder::Step / TopPaths]: a prefix
// tree whose entries carry log P(prefix, ending in blank) and log P(prefix, ending in its last label) for the previous and the
// current frame; per frame (1) every entry of the beam is re-scored in place, (2) entries are expanded in descending order of
// their previous total against the running bottom of a beam_width-bounded list, a full list evicting its bottom.  The result is
// the label sequence of the best entry, with consecutive equal labels collapsed when merge_repeated is set (TF applies that to
// the decoded sequence, so "a, blank, a" also comes out as one "a").
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <deque>
#include <limits>
#include <thread>
#include <vector>

#include "../../include/crnn_ctc.h"

int crnn_fail(int status, const char* fmt, ...);

namespace {

constexpr double kLogZero = -std::numeric_limits<double>::infinity();

inline double log_add(double a, double b) {
  if (a == kLogZero) return b;
  if (b == kLogZero) return a;
  const double m = a > b ? a : b;
  return m + std::log1p(std::exp(-std::fabs(a - b)));
}

struct Prob {
  double total = kLogZero, blank = kLogZero, label = kLogZero;
  void reset() { total = blank = label = kLogZero; }
};

struct Entry {
  Entry* parent = nullptr;
  int label = -1;
  int first_child = -1;       // index into the arena of the first of (C-1) consecutive children, -1 = not populated
  Prob oldp, newp;
  bool active() const { return newp.total != kLogZero; }
};

// Beam-width bounded list of the current leaves; the "bottom" is the entry with the smallest newp.total.
struct Leaves {
  std::vector<Entry*> v;
  Entry* bottom() const {
    Entry* b = v[0];
    for (Entry* e : v) if (e->newp.total < b->newp.total) b = e;
    return b;
  }
  void remove(Entry* e) { v.erase(std::find(v.begin(), v.end(), e)); }
};

void decode_one(const float* logits, int stride_t, int len, int C, int beam_width, int merge_repeated, int strip, int* out,
                int max_out, int* out_len, float* log_prob) {
  const int blank = C - 1, nlab = C - 1;
  std::deque<Entry> arena;                       // stable addresses
  arena.emplace_back();
  Entry* root = &arena[0];
  root->newp.total = 0.0; root->newp.blank = 0.0; root->newp.label = kLogZero;
  Leaves leaves;
  leaves.v.push_back(root);
  std::vector<Entry*> branches;
  std::vector<double> lp(C);
  for (int t = 0; t < len; ++t) {
    const float* row = logits + (size_t)t * stride_t;
    double mx = row[0];
    for (int c = 1; c < C; ++c) mx = std::max<double>(mx, row[c]);
    double se = 0.0;
    for (int c = 0; c < C; ++c) se += std::exp((double)row[c] - mx);
    const double norm = mx + std::log(se);
    for (int c = 0; c < C; ++c) lp[c] = (double)row[c] - norm;

    branches = leaves.v;
    std::stable_sort(branches.begin(), branches.end(), [](const Entry* a, const Entry* b) { return a->newp.total > b->newp.total; });
    leaves.v.clear();
    for (Entry* b : branches) b->oldp = b->newp;
    for (Entry* b : branches) {
      if (b->parent != nullptr) {
        if (b->parent->active()) {
          const double prev = (b->label == b->parent->label) ? b->parent->oldp.blank : b->parent->oldp.total;
          b->newp.label = log_add(b->newp.label, prev);
        }
        b->newp.label += lp[b->label];
      }
      b->newp.blank = b->oldp.total + lp[blank];
      b->newp.total = log_add(b->newp.blank, b->newp.label);
      leaves.v.push_back(b);
    }
    auto is_candidate = [&](double total) {
      return total > kLogZero && ((int)leaves.v.size() < beam_width || total > leaves.bottom()->newp.total);
    };
    for (Entry* b : branches) {
      if (!is_candidate(b->oldp.total)) continue;
      if (b->first_child < 0) {
        b->first_child = (int)arena.size();
        for (int c = 0; c < nlab; ++c) {
          arena.emplace_back();
          arena.back().parent = b;
          arena.back().label = c;
        }
      }
      for (int c = 0; c < nlab; ++c) {
        Entry* ch = &arena[b->first_child + c];
        if (ch->active()) continue;
        const double prev = (c == b->label) ? b->oldp.blank : b->oldp.total;
        ch->newp.blank = kLogZero;
        ch->newp.label = lp[c] + prev;
        ch->newp.total = ch->newp.label;
        if (is_candidate(ch->newp.total)) {
          if ((int)leaves.v.size() == beam_width) {
            Entry* bt = leaves.bottom();
            bt->newp.reset();
            leaves.remove(bt);
          }
          leaves.v.push_back(ch);
        } else {
          ch->oldp.reset();
          ch->newp.reset();
        }
      }
    }
  }
  Entry* best = leaves.v[0];
  for (Entry* e : leaves.v) if (e->newp.total > best->newp.total) best = e;
  std::vector<int> seq;
  for (Entry* e = best; e->parent != nullptr; e = e->parent) seq.push_back(e->label);
  int n = 0, prev = -1;
  for (auto it = seq.rbegin(); it != seq.rend(); ++it) {
    const int l = *it;
    const bool keep = !(merge_repeated && l == prev);
    prev = l;
    if (keep && l != strip && n < max_out) out[n++] = l;
  }
  for (int i = n; i < max_out; ++i) out[i] = 0;
  *out_len = n;
  if (log_prob) *log_prob = (float)(-best->newp.total);
}

}  // namespace

extern "C" int crnn_ctc_beam_search(const float* logits_host, const int* input_len_host, int T, int N, int C, int beam_width,
                                    int merge_repeated, int strip, int* out_host, int* out_len_host, float* neg_log_prob_host,
                                    int num_threads) {
  if (!logits_host || !input_len_host || !out_host || !out_len_host) return crnn_fail(CRNN_INVALID_VALUE, "beam_search: null pointer");
  if (T <= 0 || N <= 0 || C < 2 || beam_width < 1) return crnn_fail(CRNN_INVALID_VALUE, "beam_search: bad shape");
  for (int n = 0; n < N; ++n)
    if (input_len_host[n] < 0 || input_len_host[n] > T) return crnn_fail(CRNN_INVALID_VALUE, "beam_search: input_len[%d] outside [0, T]", n);
  int nt = num_threads > 0 ? num_threads : (int)std::thread::hardware_concurrency();
  nt = std::max(1, std::min(nt, N));
  auto work = [&](int tid) {
    for (int n = tid; n < N; n += nt)
      decode_one(logits_host + (size_t)n * C, N * C, input_len_host[n], C, beam_width, merge_repeated, strip,
                 out_host + (size_t)n * T, T, out_len_host + n, neg_log_prob_host ? neg_log_prob_host + n : nullptr);
  };
  if (nt == 1) { work(0); return CRNN_OK; }
  std::vector<std::thread> th;
  for (int i = 0; i < nt; ++i) th.emplace_back(work, i);
  for (auto& t : th) t.join();
  return CRNN_OK;
}
