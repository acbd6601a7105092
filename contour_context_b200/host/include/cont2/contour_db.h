// cont2/contour_db.h (facade) — the reference's ContourDB surface (include/cont2/contour_db.h:658-845) over the C-ABI.
#pragma once
#include <memory>
#include <vector>

#include "cont2/contour_mng.h"
#include "cont2/correlation.h"
#include "tools/bm_util.h"

extern SequentialTimeProfiler stp;  // defined by the executable, like the reference (contour_db.h:23)

struct TreeBucketConfig {  // reference contour_db.h:54-57
  double max_elapse_ = 25.0;
  double min_elapse_ = 15.0;
};

struct IndexOfKey {  // reference contour_db.h:59-65: where a retrieval key comes from: global scan index, level, sequence at that level
  size_t gidx{};
  int level{};
  int seq{};
  IndexOfKey(size_t g, int l, int s) : gidx(g), level(l), seq(s) {}
};

struct CandidateScoreEnsemble {  // reference contour_db.h:244-250
  ScoreConstellSim sim_constell;
  ScorePairwiseSim sim_pair;
  ScorePostProc sim_post;
};

struct ContourDBConfig {  // reference contour_db.h:658-669
  int nnk_ = 50;
  int max_fine_opt_ = 10;
  std::vector<int> q_levels_;
  ContourSimThresConfig cont_sim_cfg_;
  TreeBucketConfig tb_cfg_;
};

class ContourDB {
  const ContourDBConfig cfg_;
  std::vector<std::shared_ptr<const ContourManager>> all_bevs_;

 public:
  ContourDB(const ContourDBConfig &config);

  // reference contour_db.h:698-811.  fineOptimize's L-BFGS refinement runs on the device (csrc/refine.cu): the returned
  // correlation / transform are anch_props_[0].correlation_ / T_delta_ after the refinement (contour_db.h:626-627,642-643),
  // i.e. c2g_cand::corr_fine / T_fine.  Stage timers "KNN search", "Constell", "L2 opt" are booked in `stp` like the reference.
  void queryRangedKNN(const std::shared_ptr<const ContourManager> &q_ptr, const CandidateScoreEnsemble &thres_lb,
                      const CandidateScoreEnsemble &thres_ub, std::vector<std::shared_ptr<const ContourManager>> &cand_ptrs,
                      std::vector<double> &cand_corr, std::vector<Eigen::Isometry2d> &cand_tf) const;
  void addScan(const std::shared_ptr<ContourManager> &added, double curr_timestamp);  // reference contour_db.h:814-824
  void pushAndBalance(int seed, double curr_timestamp);                               // reference contour_db.h:827-843
  size_t size() const { return all_bevs_.size(); }

  // Extension (no reference counterpart; C-ABI c2g_online_window): W consecutive iterations of
  //     queryRangedKNN(scan_i) -> addScan(scan_i, ts[i]) -> pushAndBalance(seeds[i], ts[i])        (test/batch_bin_test.cpp:179,234,237)
  // in one call, with exactly the results of the scan-by-scan calls: the scans are ingested as one batch, the LayerDB bookkeeping of the
  // window is replayed on the host, every scan is searched against the trees as they stand after its predecessors.  The scans must
  // carry their points (makeBEV / makeBEVFromBin) but not have run makeContoursRecurs(); at most C2G_WINDOW scans (environment
  // variable read when the runtime starts).  out[i].found == false: no candidate (queryRangedKNN would return empty vectors).
  struct WindowResult {
    bool found = false;
    std::shared_ptr<const ContourManager> cand;
    double corr = 0.0;
    Eigen::Isometry2d tf;
  };
  void queryAddBalanceWindow(std::vector<std::shared_ptr<ContourManager>> &scans, const std::vector<double> &ts, const std::vector<int> &seeds,
                             const CandidateScoreEnsemble &thres_lb, const CandidateScoreEnsemble &thres_ub, std::vector<WindowResult> &out);
};
