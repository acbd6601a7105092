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

union ConstellationPair {  // reference contour_mng.h:221-240: a pair of "matched" contours of two scans at one level
  struct {
    int8_t level;
    int8_t seq_src;
    int8_t seq_tgt;
  };
  int data[1]{};
  ConstellationPair(int8_t l, int8_t s, int8_t t) : level(l), seq_src(s), seq_tgt(t) {}
  bool operator<(const ConstellationPair &a) const {
    return level < a.level || (level == a.level && seq_src < a.seq_src) || (level == a.level && seq_src == a.seq_src && seq_tgt < a.seq_tgt);
  }
  bool operator==(const ConstellationPair &a) const { return level == a.level && seq_src == a.seq_src && seq_tgt == a.seq_tgt; }
};

union Pixelf {  // reference contour_mng.h:392-411: 2.5-D continuous pixel (bev_pixfs_)
  struct {
    float row_f;
    float col_f;
    float elev;
  };
  int data[3]{};
  Pixelf(float r, float c, float e) : row_f(r), col_f(c), elev(e) {}
  Pixelf() {
    row_f = -1;
    col_f = -1;
    elev = -1;
  }
  bool operator<(const Pixelf &b) const { return row_f < b.row_f; }
};

struct BCI {  // reference contour_mng.h:243-389 (the hot path runs checkConstellSim on the device: csrc/query.cu)
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

  union DistSimPair {  // reference contour_mng.h:260-271
    struct {
      float orie_diff;
      int8_t seq_src;
      int8_t seq_tgt;
      int8_t level;
    };
    int data[2]{};
    DistSimPair(int8_t l, int8_t s, int8_t t, float o) : orie_diff(o), seq_src(s), seq_tgt(t), level(l) {}
  };
  // reference contour_mng.h:288-388 (host restatement, for callers outside the query path; defined in cont2_facade.cpp)
  static ScoreConstellSim checkConstellSim(const BCI &src, const BCI &tgt, const ScoreConstellSim &lb, std::vector<ConstellationPair> &constell_res);
};

namespace c2g_host {
// One CUDA context per process (c2g.h: "one context per (process, device)"), created on first use with the first
// ContourManagerConfig / ContourDBConfig seen.  Capacity via C2G_SCAN_CAPACITY (default 8192 scans).
struct Runtime;
Runtime &runtime();
c2g_ctx *context();  // the process-wide C-ABI context (created on first use); for callers that mix facade objects and c2g.h calls
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

  // Public statics of the reference (host restatements over the descriptors read back from the device; the query path itself runs
  // the same cascade in csrc/query.cu).  reference contour_mng.h:1124-1242, :1251-1277, :1279-1284.
  static ScorePairwiseSim checkConstellCorrespSim(const ContourManager &src, const ContourManager &tgt, const std::vector<ConstellationPair> &cstl_in,
                                                  const ScorePairwiseSim &lb, const ContourSimThresConfig &cont_sim,
                                                  std::vector<ConstellationPair> &cstl_out, std::vector<float> &area_perc);
  template <typename Iter>
  static Eigen::Isometry2d getTFFromConstell(const ContourManager &src, const ContourManager &tgt, Iter cstl_beg, Iter cstl_end) {
    std::vector<ConstellationPair> v(cstl_beg, cstl_end);
    return tfFromConstell(src, tgt, v.data(), (int) v.size());
  }
  static bool checkContPairSim(const ContourManager &src, const ContourManager &tgt, const ConstellationPair &cstl,
                               const ContourSimThresConfig &cont_sim) {
    return ContourView::checkSim(*src.getLevContours(cstl.level)[cstl.seq_src], *tgt.getLevContours(cstl.level)[cstl.seq_tgt], cont_sim);
  }

 private:
  static Eigen::Isometry2d tfFromConstell(const ContourManager &src, const ContourManager &tgt, const ConstellationPair *cstl, int n);
};
