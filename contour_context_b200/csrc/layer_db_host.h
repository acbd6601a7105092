// layer_db_host.h — host-side bookkeeping of ContourDB's retrieval-key database: which keys sit in which KD-tree bucket,
// which are still in a bucket's time-delay buffer, and the bucket boundaries.
//
// Restates (host C++, the reference keeps this on the CPU too and it costs 0.08 ms/scan):
//   TreeBucket::{pushBuffer,needPopBuffer,popBufferMax}   include/cont2/contour_db.h:98-143
//   LayerDB::{LayerDB,pushBuffer}                         include/cont2/contour_db.h:168-192
//   LayerDB::rebuild                                      src/cont2/contour_db.cpp:63-317
//   ContourDB::{addScan,pushAndBalance}                   include/cont2/contour_db.h:814-843
// The KD-tree itself is replaced by the flat device table (query.cu); this file decides WHICH keys are searchable and in
// which bucket, which is what the reference's results depend on.
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/c2g_types.h"

struct C2gKeyRec {
  float k[C2G_KEY_DIM];
  int gidx;
  int seq;
};
struct C2gBufRec {
  C2gKeyRec key;
  double ts;
};
struct C2gBucket {
  float beg, end;
  std::vector<C2gKeyRec> tree;    // data_tree_ + gkidx_tree_
  // The first `indexed` entries of `tree` are searchable: the reference searches a bucket through a KD index that it rebuilds
  // only when the bucket pops something from its buffer (TreeBucket::popBufferMax -> rebuildTree, contour_db.h:119-143).  A
  // bucket that RECEIVES keys in a rebalancing move (they are appended to its tree) without popping anything keeps its old
  // index: the moved keys are not found until its next pop (and a bucket that never popped has no index at all).  A bucket
  // that GIVES keys away without popping is left with an index into a permuted, shortened vector - undefined behaviour in
  // the reference; here all of its remaining keys stay searchable.
  size_t indexed = 0;
  std::vector<C2gBufRec> buffer;  // buffer_
  // bumped whenever existing tree entries move or disappear; between two bumps the tree only grows at its end, which is
  // what lets the device mirror (query.cu) be patched instead of rebuilt
  unsigned restructured = 0;
};
struct C2gLayerHost {
  C2gBucket buckets[C2G_NUM_BUCKETS];
  float ranges[C2G_NUM_BUCKETS + 1];
};
struct C2gHostDB {
  int n_layers;
  double max_elapse, min_elapse;
  C2gLayerHost layers[C2G_NUM_Q_LEVELS_MAX];
  int n_scans;  // all_bevs_.size()
  // bumped whenever the searchable contents of any tree change (keys popped from a buffer, keys moved between buckets):
  // consecutive scans of the online loop between two bumps see the same trees (c2g_online_commit's runs)
  unsigned long long tree_version = 0;
};

void c2g_hostdb_init(C2gHostDB &db, int n_layers, double max_elapse, double min_elapse);
// LayerDB::pushBuffer for one key of layer ll (all-zero keys are skipped exactly like the reference: key.sum() != 0)
void c2g_hostdb_push(C2gHostDB &db, int ll, const float *key, double ts, int gidx, int seq);
// ContourDB::pushAndBalance
void c2g_hostdb_push_and_balance(C2gHostDB &db, int seed, double ts);
