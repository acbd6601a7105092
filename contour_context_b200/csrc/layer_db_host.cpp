// layer_db_host.cpp — see layer_db_host.h.  Plain host C++ (no CUDA); float compares only, no arithmetic on keys.
#include "layer_db_host.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <numeric>

namespace {

const float kMaxBucketVal = 1000.0f;  // MAX_BUCKET_VAL (contour_db.h:29)
const int kMinElemSplit = 100;        // LayerDB::min_elem_split_
const double kImbaDiffRatio = 0.2;    // LayerDB::imba_diff_ratio_

inline float k0(const C2gKeyRec &r) { return r.k[0]; }  // bucket_chann_ = 0

bool need_pop(const C2gBucket &b, double curr_ts, double max_elapse) {
  if (b.buffer.empty() || b.buffer[0].ts > curr_ts - max_elapse) return false;
  return true;
}

// `changed` is set when the searchable contents of a tree change (C2gHostDB::tree_version)
void pop_buffer_max(C2gBucket &b, double curr_ts, double min_elapse, bool &changed) {
  const double cutoff = curr_ts - min_elapse;
  size_t gap = 0;
  while (gap < b.buffer.size() && !(b.buffer[gap].ts >= cutoff)) ++gap;
  if (gap == 0) return;
  changed = true;
  for (size_t i = 0; i < gap; ++i) b.tree.push_back(b.buffer[i].key);
  b.buffer.erase(b.buffer.begin(), b.buffer.begin() + (long) gap);
  b.indexed = b.tree.size();  // rebuildTree()
}

// Moves `num` keys out of `from` (the larger tree) into `to`. `perm` is the reference's sort permutation of `from`
// (ascending key0 when the donor is the lower bucket, descending when it is the upper one); the moved keys are the last
// `num` entries of the permutation, appended to `to` starting from the very last one. The donor is compacted with the
// reference's in-place rotate loop so that the surviving keys keep the same order as in the reference.
void move_tail(C2gBucket &from, C2gBucket &to, const std::vector<int> &perm, int num, float split_val, bool donor_is_lower) {
  const int sz = (int) from.tree.size();
  for (int i = 0; i < num; ++i) to.tree.push_back(from.tree[perm[sz - i - 1]]);
  int p_dat = sz - 1;
  for (int p_perm = sz - 1; p_perm >= sz - num; --p_perm) {
    if (donor_is_lower) {
      while (k0(from.tree[p_dat]) >= split_val) --p_dat;
    } else {
      while (k0(from.tree[p_dat]) < split_val) --p_dat;
    }
    if (perm[p_perm] < p_dat) {
      std::swap(from.tree[p_dat], from.tree[perm[p_perm]]);
      --p_dat;
    }
  }
  from.tree.resize((size_t) (p_dat + 1));
}

// Splits the donor's buffer at split_val with the reference's two-pointer swap loop and appends the moved part to `to`.
void move_buffer(C2gBucket &from, C2gBucket &to, float split_val, bool donor_is_lower) {
  int p1 = 0, p2 = (int) from.buffer.size() - 1;
  auto stays = [&](const C2gBufRec &r) { return donor_is_lower ? (k0(r.key) < split_val) : (k0(r.key) >= split_val); };
  while (p1 <= p2) {
    if (!stays(from.buffer[p1]) && stays(from.buffer[p2])) {
      std::swap(from.buffer[p1], from.buffer[p2]);
      ++p1;
      --p2;
    } else {
      if (!stays(from.buffer[p2])) --p2;
      if (stays(from.buffer[p1])) ++p1;
    }
  }
  const int rem = p2 + 1;
  to.buffer.insert(to.buffer.end(), from.buffer.begin() + rem, from.buffer.end());
  from.buffer.resize((size_t) rem);
}

void rebuild_layer(C2gLayerHost &L, int idx_t1, double curr_ts, double max_elapse, double min_elapse, bool &changed) {
  C2gBucket &tr1 = L.buckets[idx_t1], &tr2 = L.buckets[idx_t1 + 1];
  const bool pb1 = need_pop(tr1, curr_ts, max_elapse), pb2 = need_pop(tr2, curr_ts, max_elapse);
  if (!pb1 && !pb2) return;
  const int sz1 = (int) tr1.tree.size(), sz2 = (int) tr2.tree.size();
  const double diff_ratio = 1.0 * std::abs(sz1 - sz2) / std::max(sz1, sz2);  // NaN for two empty trees, as in the reference
  const bool small = diff_ratio < kImbaDiffRatio || std::max(sz1, sz2) < kMinElemSplit;
  if (pb1 && !pb2 && small) {
    pop_buffer_max(tr1, curr_ts, min_elapse, changed);
    return;
  }
  if (!pb1 && pb2 && small) {
    pop_buffer_max(tr2, curr_ts, min_elapse, changed);
    return;
  }
  if (diff_ratio < 0.5 * kImbaDiffRatio) {
    if (pb1) pop_buffer_max(tr1, curr_ts, min_elapse, changed);
    if (pb2) pop_buffer_max(tr2, curr_ts, min_elapse, changed);
    return;
  }
  const bool donor_is_lower = sz1 > sz2;
  C2gBucket &big = donor_is_lower ? tr1 : tr2;
  C2gBucket &lit = donor_is_lower ? tr2 : tr1;
  const int szb = donor_is_lower ? sz1 : sz2, szl = donor_is_lower ? sz2 : sz1;
  if (szb == 0) {  // unreachable in the reference (it would index an empty vector); nothing to balance
    if (pb1) pop_buffer_max(tr1, curr_ts, min_elapse, changed);
    if (pb2) pop_buffer_max(tr2, curr_ts, min_elapse, changed);
    return;
  }
  const int to_move_max = int((szb - szl + kImbaDiffRatio * szl) / (2 - kImbaDiffRatio));
  const int to_move_mid = int((szb - szl) / 2.0);
  const int to_move_min = std::max(0, int((szb - szl - kImbaDiffRatio * szb) / (2 - kImbaDiffRatio)));
  std::vector<int> perm((size_t) szb);
  std::iota(perm.begin(), perm.end(), 0);
  if (donor_is_lower)
    std::sort(perm.begin(), perm.end(), [&](const int &a, const int &b) { return k0(big.tree[a]) < k0(big.tree[b]); });
  else
    std::sort(perm.begin(), perm.end(), [&](const int &a, const int &b) { return k0(big.tree[a]) > k0(big.tree[b]); });
  auto val = [&](int i_from_end) { return k0(big.tree[perm[szb - i_from_end]]); };  // i-th key counted from the moving end

  int num_to_move = 0;
  float split_val = tr1.end;
  // to_move_mid == 0 (sizes differ by one): the reference compares sort_permu[szb] - one past the end - with its neighbour;
  // whatever that read returns, nothing below can produce num_to_move > 0, so it ends in the "cannot split" branch
  if (to_move_mid == 0) {
  } else if (val(to_move_mid) != val(to_move_mid + 1)) {
    num_to_move = to_move_mid;
    // lower donor: the smallest moved key becomes the boundary; upper donor: the smallest key that stays
    split_val = donor_is_lower ? val(to_move_mid) : val(to_move_mid + 1);
  } else {
    const float contagious = val(to_move_mid);
    for (int i = to_move_mid - 1; i > to_move_min; --i)
      if (val(i) != contagious) {
        num_to_move = i;
        split_val = donor_is_lower ? val(i) : contagious;
        break;
      }
    if (num_to_move == 0)
      for (int i = to_move_mid + 1; i < to_move_max; ++i)
        if (val(i) != contagious) {
          num_to_move = i - 1;
          split_val = donor_is_lower ? contagious : val(i);
          break;
        }
  }
  if (num_to_move == 0) {  // a strip of equal bucket values prevents the split
    if (donor_is_lower) {
      pop_buffer_max(tr1, curr_ts, min_elapse, changed);
      if (pb2) pop_buffer_max(tr2, curr_ts, min_elapse, changed);
    } else {
      if (pb1) pop_buffer_max(tr1, curr_ts, min_elapse, changed);
      pop_buffer_max(tr2, curr_ts, min_elapse, changed);
    }
    return;
  }
  move_tail(big, lit, perm, num_to_move, split_val, donor_is_lower);  // `big` is permuted and cut, `lit` only grows
  big.restructured++;
  big.indexed = big.tree.size();  // (the reference's index of the donor is undefined from here until its next pop)
  changed = true;                 // `lit` keeps its index: the keys it received are searchable after its next pop
  move_buffer(big, lit, split_val, donor_is_lower);
  tr1.end = tr2.beg = split_val;
  L.ranges[idx_t1 + 1] = split_val;
  auto by_ts = [](const C2gBufRec &a, const C2gBufRec &b) { return a.ts < b.ts; };
  std::sort(tr1.buffer.begin(), tr1.buffer.end(), by_ts);
  std::sort(tr2.buffer.begin(), tr2.buffer.end(), by_ts);
  pop_buffer_max(tr1, curr_ts, min_elapse, changed);
  pop_buffer_max(tr2, curr_ts, min_elapse, changed);
}

}  // namespace

void c2g_hostdb_init(C2gHostDB &db, int n_layers, double max_elapse, double min_elapse) {
  db.n_layers = n_layers;
  db.max_elapse = max_elapse;
  db.min_elapse = min_elapse;
  db.n_scans = 0;
  for (int l = 0; l < C2G_NUM_Q_LEVELS_MAX; ++l) {
    C2gLayerHost &L = db.layers[l];
    for (int i = 0; i < C2G_NUM_BUCKETS; ++i) {
      L.buckets[i].tree.clear();
      L.buckets[i].indexed = 0;
      L.buckets[i].restructured++;
      L.buckets[i].buffer.clear();
      L.buckets[i].beg = L.buckets[i].end = kMaxBucketVal;
      L.ranges[i] = kMaxBucketVal;
    }
    L.buckets[0].beg = -kMaxBucketVal;
    L.ranges[0] = -kMaxBucketVal;
    L.ranges[C2G_NUM_BUCKETS] = kMaxBucketVal;
  }
}

void c2g_hostdb_push(C2gHostDB &db, int ll, const float *key, double ts, int gidx, int seq) {
  C2gLayerHost &L = db.layers[ll];
  for (int i = 0; i < C2G_NUM_BUCKETS; ++i) {
    if (L.ranges[i] <= key[0] && key[0] < L.ranges[i + 1]) {
      float sum = 0.0f;  // ArrayAsKey::sum
      for (int d = 0; d < C2G_KEY_DIM; ++d) sum += key[d];
      if (sum != 0) {
        C2gBufRec r;
        for (int d = 0; d < C2G_KEY_DIM; ++d) r.key.k[d] = key[d];
        r.key.gidx = gidx;
        r.key.seq = seq;
        r.ts = ts;
        L.buckets[i].buffer.push_back(r);
      }
      return;
    }
  }
}

void c2g_hostdb_push_and_balance(C2gHostDB &db, int seed, double ts) {
  int idx_t1 = std::abs(seed) % (2 * (C2G_NUM_BUCKETS - 2));
  if (idx_t1 > (C2G_NUM_BUCKETS - 2)) idx_t1 = 2 * (C2G_NUM_BUCKETS - 2) - idx_t1;
  bool changed = false;
  for (int l = 0; l < db.n_layers; ++l) rebuild_layer(db.layers[l], idx_t1, ts, db.max_elapse, db.min_elapse, changed);
  if (changed) db.tree_version++;
}
