// tracker.cpp -- host side of the post-detection filter stage: the centroid tracker of
// TrackFilter (watsor/filter/track.py:29-149) and the write-back of DetectionSieve (watsor/filter/sieve.py:21-52),
// so that one call turns the 100 rows + GPU verdicts of a frame into the rows the reference would publish.
//
// Pure C++, no CUDA: the work is sequential, stateful and tiny (<= 100 rows per frame).  Everything the
// reference leaves to its runtime is restated explicitly:
//   * dict / defaultdict iteration = insertion order (labels), list order (objects of a label);
//   * scipy cdist + np.amin/argmin/argsort: compared as exact integer squared distances (sqrt is monotonic and
//     the sums are < 2^53, so order and ties are those of the float64 distances); argmin = first minimum;
//     argsort = numpy's default `quicksort` kind for float64, restated in numpy_argsort(): an index introsort --
//     median-of-3 partitions down to 16 elements, insertion sort below, heapsort past the depth limit
//     (numpy/core/src/npysort/quicksort.cpp; the reference pins numpy 1.23, docker/Dockerfile.base:33).  Up to
//     16 rows it is a plain insertion sort, i.e. ties keep the lower index first; above that the tie order is
//     whatever the partitions produce, and it is reproduced here step for step (checked against numpy's scalar
//     code path in tests/test_tracker.py; numpy >= 1.25 on AVX2/AVX-512 machines sorts with SIMD networks and
//     orders ties differently -- that is not the reference's environment);
//   * `for col in unused_cols` and the zone union iterate CPython *sets* of small ints: slot order of the open-
//     addressing table (Objects/setobject.c: 8 slots minimum, fill*5 >= mask*3 -> resize to > 4*used,
//     9 linear probes, perturb shift 5), restated in PySmallIntSet and pinned against the interpreter in
//     tests/test_tracker.py.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <deque>
#include <new>
#include <vector>

#include "../../include/watsor_b200.h"

namespace {

// CPython set holding small non-negative ints (hash(i) == i): insertion, then iteration in slot order.
class PySmallIntSet {
 public:
  PySmallIntSet() : table_(8, -1), fill_(0) {}
  void add(int key) {
    const size_t mask = table_.size() - 1;
    size_t i = (size_t)key & mask, perturb = (size_t)key;
    while (true) {
      if (table_[i] < 0) break;
      if (table_[i] == key) return;
      bool found = false;
      if (i + kLinearProbes <= mask) {
        for (size_t j = 1; j <= kLinearProbes; ++j) {
          if (table_[i + j] < 0) {
            i += j;
            found = true;
            break;
          }
          if (table_[i + j] == key) return;
        }
      }
      if (found) break;
      perturb >>= kPerturbShift;
      i = (i * 5 + 1 + perturb) & mask;
    }
    table_[i] = key;
    ++fill_;
    if (fill_ * 5 < mask * 3) return;
    resize(fill_ * 4);  // used == fill (nothing is ever removed), used <= 50000
  }
  template <class F>
  void for_each(F f) const {
    for (int k : table_)
      if (k >= 0) f(k);
  }

 private:
  static constexpr size_t kLinearProbes = 9, kPerturbShift = 5;
  void resize(size_t minused) {
    size_t newsize = 8;
    while (newsize <= minused) newsize <<= 1;
    std::vector<int> old;
    old.swap(table_);
    table_.assign(newsize, -1);
    const size_t mask = newsize - 1;
    for (int key : old) {
      if (key < 0) continue;
      size_t i = (size_t)key & mask, perturb = (size_t)key;
      while (true) {  // set_insert_clean
        if (table_[i] < 0) break;
        bool found = false;
        if (i + kLinearProbes <= mask) {
          for (size_t j = 1; j <= kLinearProbes; ++j)
            if (table_[i + j] < 0) {
              i += j;
              found = true;
              break;
            }
        }
        if (found) break;
        perturb >>= kPerturbShift;
        i = (i * 5 + 1 + perturb) & mask;
      }
      table_[i] = key;
    }
  }
  std::vector<int> table_;
  size_t fill_;
};

// np.argsort(v) for float64 keys without NaNs, `kind='quicksort'` (the default): numpy's aquicksort_ / aheapsort_.
// `v` holds exact integer squared distances; their order and ties are those of the float64 distances.
void numpy_heapsort(const long long* v, int* tosort, int n) {
  int* a = tosort - 1;  // 1-based
  for (int l = n >> 1; l > 0; --l) {
    const int tmp = a[l];
    int i = l, j = l << 1;
    while (j <= n) {
      if (j < n && v[a[j]] < v[a[j + 1]]) ++j;
      if (v[tmp] < v[a[j]]) {
        a[i] = a[j];
        i = j;
        j += j;
      } else {
        break;
      }
    }
    a[i] = tmp;
  }
  while (n > 1) {
    const int tmp = a[n];
    a[n] = a[1];
    --n;
    int i = 1, j = 2;
    while (j <= n) {
      if (j < n && v[a[j]] < v[a[j + 1]]) ++j;
      if (v[tmp] < v[a[j]]) {
        a[i] = a[j];
        i = j;
        j += j;
      } else {
        break;
      }
    }
    a[i] = tmp;
  }
}

void numpy_argsort(const long long* v, int num, std::vector<int>* order) {
  constexpr int kSmall = 15;  // partitions of more than 16 elements are split, the rest is insertion-sorted
  order->resize((size_t)num);
  int* t = order->data();
  for (int i = 0; i < num; ++i) t[i] = i;
  if (num < 2) return;
  int* pl = t;
  int* pr = t + num - 1;
  int* stack[128];
  int depth[64];
  int** sptr = stack;
  int* psdepth = depth;
  int msb = 0;
  for (unsigned m = (unsigned)num >> 1; m != 0; m >>= 1) ++msb;
  int cdepth = msb * 2;
  while (true) {
    if (cdepth < 0) {
      numpy_heapsort(v, pl, (int)(pr - pl) + 1);
    } else {
      while (pr - pl > kSmall) {
        int* pm = pl + ((pr - pl) >> 1);
        if (v[*pm] < v[*pl]) std::swap(*pm, *pl);
        if (v[*pr] < v[*pm]) std::swap(*pr, *pm);
        if (v[*pm] < v[*pl]) std::swap(*pm, *pl);
        const long long vp = v[*pm];
        int* pi = pl;
        int* pj = pr - 1;
        std::swap(*pm, *pj);
        for (;;) {
          do { ++pi; } while (v[*pi] < vp);
          do { --pj; } while (vp < v[*pj]);
          if (pi >= pj) break;
          std::swap(*pi, *pj);
        }
        int* pk = pr - 1;
        std::swap(*pi, *pk);
        if (pi - pl < pr - pi) {  // the larger partition waits on the stack
          *sptr++ = pi + 1;
          *sptr++ = pr;
          pr = pi - 1;
        } else {
          *sptr++ = pl;
          *sptr++ = pi - 1;
          pl = pi + 1;
        }
        *psdepth++ = --cdepth;
      }
      for (int* pi = pl + 1; pi <= pr; ++pi) {  // insertion sort
        const int vi = *pi;
        const long long vp = v[vi];
        int* pj = pi;
        int* pk = pi - 1;
        while (pj > pl && vp < v[*pk]) *pj-- = *pk--;
        *pj = vi;
      }
    }
    if (sptr == stack) break;
    pr = *(--sptr);
    pl = *(--sptr);
    cdepth = *(--psdepth);
  }
}

// Iteration order of `set(range(n)).difference(used)` (track.py:90,98).  CPython copies the left set and removes
// when it is more than four times larger than the right one (the copy keeps the ascending slot order of
// set(range(n))); otherwise it inserts the survivors one by one into a fresh set.
void unused_in_set_order(int n, const std::vector<char>& used, int n_used, std::vector<int>* out) {
  out->clear();
  if ((n >> 2) > n_used) {
    for (int c = 0; c < n; ++c)
      if (!used[c]) out->push_back(c);
    return;
  }
  PySmallIntSet s;
  for (int c = 0; c < n; ++c)
    if (!used[c]) s.add(c);
  s.for_each([&](int c) { out->push_back(c); });
}

struct History {
  std::deque<wb_detection> rows;  // deque(maxlen=history): oldest first
};

struct LabelObjects {
  int label;
  std::vector<History> objects;
};

inline void centroid(const wb_detection& d, long long* cx, long long* cy) {
  // track.py:120-123: int((x_min + x_max) / 2.0) truncates toward zero
  *cx = (long long)(((double)d.bounding_box.x_min + (double)d.bounding_box.x_max) / 2.0);
  *cy = (long long)(((double)d.bounding_box.y_min + (double)d.bounding_box.y_max) / 2.0);
}

// track.py:125-149
wb_detection combine(const History& h) {
  wb_detection out;
  std::memset(&out, 0, sizeof(out));
  const wb_detection& first = h.rows.front();
  out.label = first.label;
  out.confidence = first.confidence;
  out.bounding_box = first.bounding_box;
  for (size_t i = 1; i < h.rows.size(); ++i) {
    const wb_detection& d = h.rows[i];
    out.confidence = std::max(out.confidence, d.confidence);
    out.bounding_box.x_min = std::min(out.bounding_box.x_min, d.bounding_box.x_min);
    out.bounding_box.y_min = std::min(out.bounding_box.y_min, d.bounding_box.y_min);
    out.bounding_box.x_max = std::max(out.bounding_box.x_max, d.bounding_box.x_max);
    out.bounding_box.y_max = std::max(out.bounding_box.y_max, d.bounding_box.y_max);
  }
  PySmallIntSet zones;
  for (const wb_detection& d : h.rows)
    for (int z = 0; z < WB_MAX_ZONES; ++z)
      if (d.zones[z] > 0) zones.add(d.zones[z]);
  int k = 0;
  zones.for_each([&](int z) {
    if (k < WB_MAX_ZONES) out.zones[k++] = z;
  });
  return out;
}

}  // namespace

struct wb_tracker {
  int sensitivity, history;
  std::vector<LabelObjects> by_label;  // dict in insertion order
};

extern "C" {

int wb_tracker_create(int sensitivity, int history, wb_tracker** out) {
  if (out == nullptr || history < 1) return 1;
  wb_tracker* t = new (std::nothrow) wb_tracker();
  if (t == nullptr) return 1;
  t->sensitivity = sensitivity;
  t->history = history;
  *out = t;
  return 0;
}

int wb_tracker_destroy(wb_tracker* t) {
  delete t;
  return 0;
}

int wb_tracker_update(wb_tracker* t, const wb_detection* rows, int n_rows, const uint32_t* verdicts, wb_detection* out,
                      int out_cap, int* n_out, int* suspicious_activity) {
  if (t == nullptr || (rows == nullptr && n_rows > 0) || n_rows < 0 || n_out == nullptr) return 1;
  try {
    // stage 1 result: the rows that passed `label > 0 and all(predicates)` (track.py:26)
    struct Group {
      int label;
      std::vector<const wb_detection*> dets;
    };
    std::vector<Group> groups;  // defaultdict(list) in first-appearance order
    for (int i = 0; i < n_rows; ++i) {
      const bool pass = verdicts != nullptr ? (verdicts[i] & WB_V_PASS) != 0 : rows[i].label > 0;
      if (!pass) continue;
      auto it = std::find_if(groups.begin(), groups.end(), [&](const Group& g) { return g.label == rows[i].label; });
      if (it == groups.end()) {
        groups.push_back(Group{rows[i].label, {}});
        it = groups.end() - 1;
      }
      it->dets.push_back(&rows[i]);
    }
    if (suspicious_activity != nullptr) *suspicious_activity = groups.empty() ? 0 : 1;

    // labels that are no longer detected (track.py:41-46)
    t->by_label.erase(std::remove_if(t->by_label.begin(), t->by_label.end(),
                                     [&](const LabelObjects& lo) {
                                       return std::none_of(groups.begin(), groups.end(),
                                                           [&](const Group& g) { return g.label == lo.label; });
                                     }),
                      t->by_label.end());

    std::vector<long long> in_c, ex_c, row_min;
    std::vector<int> row_arg, order, fresh;
    std::vector<char> used_r, used_c;
    for (const Group& g : groups) {
      auto lit = std::find_if(t->by_label.begin(), t->by_label.end(),
                              [&](const LabelObjects& lo) { return lo.label == g.label; });
      if (lit == t->by_label.end()) {
        t->by_label.push_back(LabelObjects{g.label, {}});
        lit = t->by_label.end() - 1;
      }
      std::vector<History>& known = lit->objects;
      const int n_in = (int)g.dets.size(), n_ex = (int)known.size();
      in_c.resize(2 * (size_t)n_in);
      ex_c.resize(2 * (size_t)n_ex);
      for (int i = 0; i < n_in; ++i) centroid(*g.dets[i], &in_c[2 * i], &in_c[2 * i + 1]);
      for (int i = 0; i < n_ex; ++i) centroid(known[i].rows.front(), &ex_c[2 * i], &ex_c[2 * i + 1]);
      used_r.assign((size_t)n_ex, 0);
      used_c.assign((size_t)n_in, 0);
      int n_used_c = 0;
      if (n_ex > 0 && n_in > 0) {
        // rows = argsort(amin(dist, axis=1)); cols = argmin(dist, axis=1)[rows]   (track.py:63-72)
        row_min.resize((size_t)n_ex);
        row_arg.resize((size_t)n_ex);
        for (int r = 0; r < n_ex; ++r) {
          long long best = -1;
          int arg = 0;
          for (int c = 0; c < n_in; ++c) {
            const long long dx = ex_c[2 * r] - in_c[2 * c], dy = ex_c[2 * r + 1] - in_c[2 * c + 1];
            const long long d2 = dx * dx + dy * dy;
            if (best < 0 || d2 < best) {
              best = d2;
              arg = c;
            }
          }
          row_min[r] = best;
          row_arg[r] = arg;
        }
        numpy_argsort(row_min.data(), n_ex, &order);
        for (int r : order) {
          const int c = row_arg[r];
          if (used_r[r] || used_c[c]) continue;
          History& h = known[r];
          if ((int)h.rows.size() == t->history) h.rows.pop_front();
          h.rows.push_back(*g.dets[c]);
          used_r[r] = 1;
          used_c[c] = 1;
          ++n_used_c;
        }
      }
      // objects that are no longer detected (track.py:93-96)
      for (int r = n_ex - 1; r >= 0; --r)
        if (!used_r[r]) known.erase(known.begin() + r);
      // detections that were not present (track.py:98-101), in the iteration order of the reference's set
      unused_in_set_order(n_in, used_c, n_used_c, &fresh);
      for (int c : fresh) {
        History h;
        h.rows.push_back(*g.dets[c]);
        known.push_back(std::move(h));
      }
    }

    // the envelope of every object seen often enough (track.py:105-116)
    int k = 0;
    for (const LabelObjects& lo : t->by_label)
      for (const History& h : lo.objects) {
        if ((int)h.rows.size() < t->sensitivity) continue;
        if (k < out_cap && out != nullptr) out[k] = combine(h);
        ++k;
      }
    *n_out = k;
    return k > out_cap && out != nullptr ? 2 : 0;
  } catch (...) {
    return 1;  // no exception crosses the C boundary
  }
}

int wb_sieve_rows(wb_tracker* t, wb_detection* rows, int n_rows, const uint32_t* verdicts, int* suspicious_activity) {
  // sieve.py:21-27 with filters == [TrackFilter(...)] (main.py:293-299): clone, track, write back, zero-fill
  if (t == nullptr || rows == nullptr || n_rows < 0) return 1;
  try {
    std::vector<wb_detection> in(rows, rows + n_rows), out((size_t)n_rows);
    int n_out = 0;
    const int rc = wb_tracker_update(t, in.data(), n_rows, verdicts, out.data(), n_rows, &n_out, suspicious_activity);
    if (rc != 0) return rc;
    const int k = std::min(n_out, n_rows);
    if (k > 0) std::memcpy(rows, out.data(), (size_t)k * sizeof(wb_detection));
    if (n_rows > k) std::memset(rows + k, 0, (size_t)(n_rows - k) * sizeof(wb_detection));
    return 0;
  } catch (...) {
    return 1;
  }
}

int wb_debug_pyset_order(const int32_t* keys, int n, int32_t* out, int* n_out) {
  // the iteration order CPython gives a set after adding `keys` in order (test hook for PySmallIntSet)
  if ((keys == nullptr && n > 0) || out == nullptr || n_out == nullptr) return 1;
  PySmallIntSet s;
  for (int i = 0; i < n; ++i) {
    if (keys[i] < 0) return 1;
    s.add(keys[i]);
  }
  int k = 0;
  s.for_each([&](int key) { out[k++] = key; });
  *n_out = k;
  return 0;
}

int wb_debug_argsort(const int64_t* keys, int n, int32_t* out) {
  // np.argsort(keys) as the tracker restates it (test hook for numpy_argsort)
  if (n < 0 || (keys == nullptr && n > 0) || (out == nullptr && n > 0)) return 1;
  std::vector<long long> v(keys, keys + n);
  std::vector<int> order;
  numpy_argsort(v.data(), n, &order);
  for (int i = 0; i < n; ++i) out[i] = order[i];
  return 0;
}

int wb_debug_unused_order(int n, const uint8_t* used, int32_t* out, int* n_out) {
  // the iteration order of `set(range(n)).difference({i : used[i]})` (test hook for unused_in_set_order)
  if (n < 0 || (used == nullptr && n > 0) || out == nullptr || n_out == nullptr) return 1;
  std::vector<char> u((size_t)n);
  int n_used = 0;
  for (int i = 0; i < n; ++i) {
    u[i] = used[i] != 0;
    n_used += u[i];
  }
  std::vector<int> order;
  unused_in_set_order(n, u, n_used, &order);
  for (size_t i = 0; i < order.size(); ++i) out[i] = order[i];
  *n_out = (int)order.size();
  return 0;
}

}  // extern "C"
