// Host-side construction of the per-user item sets from the id-mapped interaction list
// (reference semantics: data/rating.py:48-55 -- trainSet_u[user][item] = rating, a
// dict of dicts, so a repeated (user, item) line keeps the POSITION of its first occurrence and the
// VALUE of its last one -- and model/ranking/BPR.py:22-25, which keeps the items rated >= 1).
// A stable counting sort by user followed by a small sort inside each user replaces three full-length
// numpy sorts; user ranges are processed by several threads.
#include <algorithm>
#include <thread>
#include <vector>

#include "common.h"

namespace {

struct Entry {
  int32_t item;
  int64_t pos;       // index in the input list
};

int worker_count(int64_t n) {
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 1;
  if (hw > 16) hw = 16;
  const int64_t by_size = n / (1 << 18) + 1;   // not worth a thread below ~256 K entries each
  return (int)std::min<int64_t>(hw, by_size);
}

template <typename F>
void parallel_ranges(int64_t n_items, int workers, F fn) {
  if (workers <= 1) {
    fn(0, n_items, 0);
    return;
  }
  std::vector<std::thread> th;
  for (int w = 0; w < workers; ++w) {
    const int64_t lo = n_items * w / workers, hi = n_items * (w + 1) / workers;
    th.emplace_back([=] { fn(lo, hi, w); });
  }
  for (auto& t : th) t.join();
}

}  // namespace

extern "C" int qrec_build_rated_csr(int64_t n, const int64_t* u, const int64_t* i, const double* rating,
                                    int32_t num_users, int32_t num_items, double positive_threshold,
                                    int64_t* sorted_rowptr, int32_t* sorted_cols, int64_t* pos_rowptr,
                                    int32_t* pos_cols, int32_t* possorted_cols) {
  QREC_REQUIRE(n >= 0 && num_users >= 0 && num_items >= 0, "qrec_build_rated_csr: negative size");
  QREC_REQUIRE(sorted_rowptr && pos_rowptr, "qrec_build_rated_csr: null row pointer output");
  QREC_REQUIRE(n == 0 || (u && i && sorted_cols && pos_cols && possorted_cols), "qrec_build_rated_csr: null pointer");
  const int64_t U = num_users;
  // 1. stable counting sort of the input positions by user
  std::vector<int64_t> start((size_t)U + 1, 0);
  for (int64_t k = 0; k < n; ++k) {
    QREC_REQUIRE(u[k] >= 0 && u[k] < U && i[k] >= 0 && i[k] < num_items, "qrec_build_rated_csr: id out of range at entry %lld",
                 (long long)k);
    ++start[(size_t)u[k] + 1];
  }
  for (int64_t a = 0; a < U; ++a) start[(size_t)a + 1] += start[(size_t)a];
  std::vector<Entry> ent((size_t)n);
  {
    std::vector<int64_t> cur(start.begin(), start.end() - 1);
    for (int64_t k = 0; k < n; ++k) ent[(size_t)cur[(size_t)u[k]]++] = Entry{(int32_t)i[k], k};
  }
  // 2. per user: order by (item, position), collapse repeats (first position, last value), count
  std::vector<int64_t> n_rated((size_t)U, 0), n_pos((size_t)U, 0);
  std::vector<uint8_t> keep((size_t)n, 0);     // per collapsed entry (stored at the front of the user's segment)
  const int workers = worker_count(n);
  parallel_ranges(U, workers, [&](int64_t lo, int64_t hi, int) {
    for (int64_t a = lo; a < hi; ++a) {
      Entry* b = ent.data() + start[(size_t)a];
      Entry* e = ent.data() + start[(size_t)a + 1];
      std::sort(b, e, [](const Entry& x, const Entry& y) { return x.item != y.item ? x.item < y.item : x.pos < y.pos; });
      int64_t out = 0, pos_cnt = 0;
      for (Entry* p = b; p < e;) {
        Entry* q = p;
        while (q + 1 < e && (q + 1)->item == p->item) ++q;       // p = first occurrence, q = last
        const double val = rating ? rating[q->pos] : 1.0;
        b[out] = *p;
        const bool pos = val >= positive_threshold;
        keep[(size_t)(start[(size_t)a] + out)] = pos ? 1 : 0;
        pos_cnt += pos ? 1 : 0;
        ++out;
        p = q + 1;
      }
      n_rated[(size_t)a] = out;
      n_pos[(size_t)a] = pos_cnt;
    }
  });
  // 3. row pointers, then the three column arrays
  sorted_rowptr[0] = 0;
  pos_rowptr[0] = 0;
  for (int64_t a = 0; a < U; ++a) {
    sorted_rowptr[a + 1] = sorted_rowptr[a] + n_rated[(size_t)a];
    pos_rowptr[a + 1] = pos_rowptr[a] + n_pos[(size_t)a];
  }
  parallel_ranges(U, workers, [&](int64_t lo, int64_t hi, int) {
    std::vector<Entry> tmp;
    for (int64_t a = lo; a < hi; ++a) {
      const Entry* b = ent.data() + start[(size_t)a];
      const int64_t cnt = n_rated[(size_t)a];
      int32_t* sc = sorted_cols + sorted_rowptr[a];
      int32_t* ps = possorted_cols + pos_rowptr[a];
      tmp.clear();
      for (int64_t k = 0; k < cnt; ++k) {
        sc[k] = b[k].item;
        if (keep[(size_t)(start[(size_t)a] + k)]) {
          *ps++ = b[k].item;                         // positives, ascending ids
          tmp.push_back(b[k]);
        }
      }
      std::sort(tmp.begin(), tmp.end(), [](const Entry& x, const Entry& y) { return x.pos < y.pos; });
      int32_t* pc = pos_cols + pos_rowptr[a];
      for (size_t k = 0; k < tmp.size(); ++k) pc[k] = tmp[k].item;    // positives, insertion order
    }
  });
  return QREC_OK;
}
