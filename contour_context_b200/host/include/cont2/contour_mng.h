// cont2/contour_mng.h (facade) — the reference's ContourManager surface (include/cont2/contour_mng.h:414-1314) kept
// name-for-name for what test/batch_bin_test.cpp and eval/evaluator.h call; all work is forwarded to the C-ABI (c2g.h).
#pragma once
#include <array>
#include <bitset>
#include <memory>
#include <string>
#include <vector>

#include "cont2/contour.h"
#include "c2g.h"

using KeyFloatType = float;
const int RET_KEY_DIM = C2G_KEY_DIM;

template <size_t sz>
struct ArrayAsKey {  // reference contour_mng.h:39-87
  enum { SizeAtCompileTime = sz };
  KeyFloatType array[sz]{};
  KeyFloatType *data() { return array; }
  KeyFloatType &operator()(size_t i) { return array[i]; }
  const KeyFloatType &operator()(size_t i) const { return array[i]; }
  KeyFloatType &operator[](size_t i) { return array[i]; }
  const KeyFloatType &operator[](size_t i) const { return array[i]; }
  void setZero() { for (auto &d : array) d = 0; }
  size_t size() const { return sz; }
  KeyFloatType sum() const {
    KeyFloatType ret(0);
    for (const auto &dat : array) ret += dat;
    return ret;
  }
};
using RetrievalKey = ArrayAsKey<RET_KEY_DIM>;

struct ContourManagerConfig {  // reference contour_mng.h:92-110
  std::vector<float> lv_grads_;
  float reso_row_ = 1.0f, reso_col_ = 1.0f;
  int n_row_ = 150, n_col_ = 150;
  float lidar_height_ = 2.0f;
  float blind_sq_ = 9.0f;
  int min_cont_key_cnt_ = 9;
  int min_cont_cell_cnt_ = 3;
  int piv_firsts_ = 6;
  int dist_firsts_ = 10;
  float roi_radius_ = 10.0f;
};

const int16_t BITS_PER_LAYER = C2G_BITS_PER_LAYER;
const int8_t DIST_BIN_LAYERS[] = {1, 2, 3, 4};
const float LAYER_AREA_WEIGHTS[] = {0.3, 0.3, 0.3, 0.1};
const int16_t NUM_BIN_KEY_LAYER = sizeof(DIST_BIN_LAYERS) / sizeof(int8_t);

union ScoreConstellSim {  // reference contour_mng.h:121-152
  enum { SizeAtCompileTime = 3 };
  int data[SizeAtCompileTime]{};
  struct {
    int i_ovlp_sum;
    int i_ovlp_max_one;
    int i_in_ang_rng;
  };
  inline const int &overall() const { return i_in_ang_rng; }
  inline int cnt() const { return i_in_ang_rng; }
  bool strictSmaller(const ScoreConstellSim &b) const {
    for (int i = 0; i < SizeAtCompileTime; i++)
      if (data[i] >= b.data[i]) return false;
    return true;
  }
};
union ScorePairwiseSim {  // reference contour_mng.h:154-186
  enum { SizeAtCompileTime = 2 };
  int data[SizeAtCompileTime]{};
  struct {
    int i_indiv_sim;
    int i_orie_sim;
  };
  inline const int &overall() const { return i_orie_sim; }
  inline int cnt() const { return i_orie_sim; }
  bool strictSmaller(const ScorePairwiseSim &b) const {
    for (int i = 0; i < SizeAtCompileTime; i++)
      if (data[i] >= b.data[i]) return false;
    return true;
  }
};
union ScorePostProc {  // reference contour_mng.h:188-219
  enum { SizeAtCompileTime = 3 };
  float data[SizeAtCompileTime]{};
  struct {
    float correlation;
    float area_perc;
    float neg_est_dist;
  };
  inline const float &overall() const { return correlation; }
  bool strictSmaller(const ScorePostProc &b) const {
    for (int i = 0; i < SizeAtCompileTime; i++)
      if (data[i] >= b.data[i]) return false;
    return true;
  }
};

struct BCI {  // reference contour_mng.h:243-280 (data members; the similarity check runs on the device)
  union RelativePoint {
    struct {
      int8_t level;
      int8_t seq;
      int16_t bit_pos;
      float r;
      float theta;
    };
    int data[3]{};
    RelativePoint(int8_t l, int8_t a, int16_t b, float f1, float f2) : level(l), seq(a), bit_pos(b), r(f1), theta(f2) {}
  };
  std::bitset<C2G_BITS_PER_LAYER * C2G_NUM_BIN_LAYERS> dist_bin_;
  std::vector<RelativePoint> nei_pts_;
  std::vector<uint16_t> nei_idx_segs_;
  int8_t piv_seq_, level_;
  explicit BCI(int8_t seq, int8_t lev) : dist_bin_(0), piv_seq_(seq), level_(lev) {}
};

namespace c2g_host {
// One CUDA context per process (c2g.h: "one context per (process, device)"), created on first use with the first
// ContourManagerConfig / ContourDBConfig seen.  Capacity via C2G_SCAN_CAPACITY (default 8192 scans).
struct Runtime;
Runtime &runtime();
}  // namespace c2g_host

class ContourDB;

class ContourManager {
  const ContourManagerConfig cfg_;
  const ContourViewStatConfig view_stat_cfg_;
  std::string str_id_;
  int int_id_;
  std::vector<float> pts_;  // staged x, y, z, 0 of the scan between makeBEV and makeContoursRecurs
  const float *ext_pts_ = nullptr;  // ... or the caller's pinned buffer (pinnedScanBuffer), not copied
  size_t ext_n_ = 0;
  int slot_ = -1;           // device slot of the finished descriptor
  bool owns_slot_ = false;
  c2g_scan_head head_;
  mutable std::vector<std::vector<std::shared_ptr<ContourView>>> cont_views_;
  mutable bool views_loaded_ = false;
  mutable std::vector<std::vector<RetrievalKey>> layer_keys_;
  mutable std::vector<std::vector<BCI>> layer_key_bcis_;
  friend class ContourDB;
  void loadViews() const;
  void adoptHead();  // head_ (read back from the device) -> layer_keys_ / layer_key_bcis_

 public:
  explicit ContourManager(const ContourManagerConfig &config, int int_id);
  ~ContourManager();

  template <typename PointType>
  void makeBEV(typename pcl::PointCloud<PointType>::ConstPtr &ptr_gapc, std::string str_id = "") {
    pts_.resize(ptr_gapc->size() * 4);
    size_t i = 0;
    for (const auto &pt : ptr_gapc->points) {
      pts_[i++] = pt.x;
      pts_[i++] = pt.y;
      pts_[i++] = pt.z;
      pts_[i++] = 0.0f;
    }
    str_id_ = str_id.empty() ? std::to_string(ptr_gapc->header.stamp) : std::move(str_id);
  }
  // makeBEV straight from a KITTI .bin buffer (N x 4 float32), skipping the PCL detour
  void makeBEVFromBin(const float *xyzi, size_t n_points, std::string str_id);
  // A reusable page-locked buffer of `n_floats` floats for readKITTIPointCloudBin-style loaders (tools/pointcloud_util.h:17:
  // the reference reads at most 1 000 000 floats per scan).  A scan handed to makeBEVFromBin from this buffer is not copied
  // on the host: it goes to the device with one asynchronous PCIe transfer.  The buffer must stay untouched until
  // makeContoursRecurs() of that scan has returned.
  static float *pinnedScanBuffer(size_t n_floats = 1000000);

  void makeContoursRecurs();  // BEV + contours + keys + BCI on the GPU (c2g_ingest)
  void clearImage() {}        // the BEV never leaves device scratch memory
  std::vector<float> getBevImage() const;

  const std::vector<RetrievalKey> &getLevRetrievalKey(int level) const { return layer_keys_[level]; }
  const RetrievalKey &getRetrievalKey(int level, int seq) const { return layer_keys_[level][seq]; }
  const std::vector<std::shared_ptr<ContourView>> &getLevContours(int level) const {
    loadViews();
    return cont_views_[level];
  }
  int getLevTotalPix(int level) const { return head_.layer_cell_cnt[level]; }
  const std::vector<BCI> &getLevBCI(int level) const { return layer_key_bcis_[level]; }
  const BCI &getBCI(int level, int seq) const { return layer_key_bcis_[level][seq]; }
  std::string getStrID() const { return str_id_; }
  int getIntID() const { return int_id_; }
  const ContourManagerConfig &getConfig() const { return cfg_; }
  float getAreaPerc(const int8_t &lev, const int8_t &seq) const {
    loadViews();
    return cont_views_[lev][seq]->cell_cnt_ * 1.0f / head_.layer_cell_cnt[lev];
  }
  int deviceSlot() const { return slot_; }
};
