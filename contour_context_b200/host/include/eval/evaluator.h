// eval/evaluator.h (facade) — ContLCDEvaluator with the reference's name, constructor arguments, call sequence and output file
// format (reference include/eval/evaluator.h:39-425), restated without Eigen 3-D types: ground-truth association of the
// laser scans, loop-closure positives (>= 15 s older pose within 5 m), TP/FP/TN/FN bookkeeping of every prediction, running
// translation / rotation error statistics, and the tab-separated outcome file that scripts/pr_mpe.py consumes.
// The Python twin is contour_context_b200/eval.py; tests compare the two.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <numeric>
#include <sstream>
#include <string>
#include <vector>

#include "cont2/contour_db.h"
#include "cont2/correlation.h"

template <int dim>
struct SimpleRMSE {  // evaluator.h:12-29
  double sum_sqs = 0, sum_abs = 0;
  int cnt_sqs = 0;
  void addOneErr(const double *d) {
    cnt_sqs++;
    double tmp = 0;
    for (int i = 0; i < dim; i++) tmp += d[i] * d[i];
    sum_sqs += tmp;
    sum_abs += std::sqrt(tmp);
  }
  double getRMSE() const { return cnt_sqs ? std::sqrt(sum_sqs / cnt_sqs) : -1; }
  double getMean() const { return cnt_sqs ? sum_abs / cnt_sqs : -1; }
};

struct PredictionOutcome {  // evaluator.h:31-41
  enum Res { TP, FP, TN, FN };
  int id_src = -1, id_tgt = -1;
  Res tfpn = Res::TN;
  double est_err[3] = {0, 0, 0};
  double correlation = 0;
};

// lookupNN (include/tools/algos.h:77-90): nearest value in a sorted vector within tol, else -1; ties go to the lower index
template <typename T>
inline int lookupNN(const T &q_val, const std::vector<T> &sorted_vec, const T &tol) {
  if (sorted_vec.empty()) return -1;
  auto it_low = std::lower_bound(sorted_vec.begin(), sorted_vec.end(), q_val);
  auto it = it_low;
  if (it_low == sorted_vec.end())
    it = it_low - 1;
  else if (it_low != sorted_vec.begin())
    it = std::abs(q_val - *it_low) < std::abs(q_val - *(it_low - 1)) ? it_low : it_low - 1;
  if (std::abs(*it - q_val) > tol) return -1;
  return (int) (it - sorted_vec.begin());
}

class ContLCDEvaluator {
 public:
  struct LaserScanInfo {
    bool has_gt_positive_lc = false;
    C2gPose3 sens_pose;
    int seq = 0;
    double ts = 0;
    std::string fpath;
  };

 private:
  std::vector<LaserScanInfo> laser_info_;
  std::vector<int> assigned_seqs_;
  const double ts_diff_tol = 10e-3;
  const double min_time_excl = 15.0;
  const double sim_thres;
  int p_lidar_curr = -1;
  SimpleRMSE<2> tp_trans_rmse, all_trans_rmse;
  SimpleRMSE<1> tp_rot_rmse, all_rot_rmse;
  std::vector<PredictionOutcome> pred_records;

  static void fail(const char *what) {
    std::fprintf(stderr, "CHECK failed: %s\n", what);
    std::abort();
  }

 public:
  // fpath_pose: "<ts> r00 r01 r02 tx r10 r11 r12 ty r20 r21 r22 tz" per line; fpath_laser: "<ts> <seq> <bin path>" per line
  ContLCDEvaluator(const std::string &fpath_pose, const std::string &fpath_laser, const double &bar) : sim_thres(bar) {
    std::vector<double> gt_tss;
    std::vector<C2gPose3> gt_poses;
    {
      std::ifstream in(fpath_pose);
      if (!in.good()) {
        std::cerr << "Error opening gt pose file: " << fpath_pose << std::endl;
        return;
      }
      std::string line;
      while (std::getline(in, line)) {
        std::istringstream iss(line);
        double v[13];
        for (int i = 0; i < 13; ++i)
          if (!(iss >> v[i])) fail("gt pose line has 13 numbers");
        gt_tss.push_back(v[0]);
        gt_poses.push_back(C2gPose3::fromRowMajor3x4(v + 1));
      }
    }
    std::printf("Added %lu stamped gt poses.\n", gt_poses.size());
    {  // sort by time stamp
      std::vector<int> perm(gt_tss.size());
      std::iota(perm.begin(), perm.end(), 0);
      std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return gt_tss[a] < gt_tss[b]; });
      std::vector<double> ts2;
      std::vector<C2gPose3> p2;
      for (int i : perm) {
        ts2.push_back(gt_tss[i]);
        p2.push_back(gt_poses[i]);
      }
      gt_tss.swap(ts2);
      gt_poses.swap(p2);
    }
    int n_bins = 0, cnt_valid_scans = 0;
    {
      std::ifstream in(fpath_laser);
      if (!in.good()) {
        std::cerr << "Error opening laser info file: " << fpath_laser << std::endl;
        return;
      }
      std::string line;
      while (std::getline(in, line)) {
        std::istringstream iss(line);
        double ts;
        int seq;
        std::string bin_path;
        if (!(iss >> ts)) continue;
        iss >> seq >> bin_path;
        n_bins++;
        const int gt_idx = lookupNN<double>(ts, gt_tss, ts_diff_tol);
        if (gt_idx < 0) continue;
        LaserScanInfo info;
        info.sens_pose = gt_poses[gt_idx];
        info.fpath = bin_path;
        info.ts = ts;
        info.seq = seq;
        cnt_valid_scans++;
        laser_info_.push_back(info);
        assigned_seqs_.push_back(seq);
      }
    }
    std::printf("Added %d laser bin paths.\n", n_bins);
    std::printf("Found %d laser scans with gt poses.\n", cnt_valid_scans);
    for (size_t i = 0; i + 1 < laser_info_.size(); ++i)
      if (!(laser_info_[i].seq < laser_info_[i + 1].seq && laser_info_[i].ts < laser_info_[i + 1].ts)) fail("laser scans ordered by seq and ts");
    std::printf("Ordering check passed\n");
    int cnt_gt_lc_p = 0, cnt_gt_lc = 0;
    for (auto &fast : laser_info_)
      for (auto &slow : laser_info_) {
        if (fast.ts < slow.ts + min_time_excl) break;
        const double dx = fast.sens_pose.t[0] - slow.sens_pose.t[0], dy = fast.sens_pose.t[1] - slow.sens_pose.t[1],
                     dz = fast.sens_pose.t[2] - slow.sens_pose.t[2];
        if (std::sqrt(dx * dx + dy * dy + dz * dz) < 5.0) {
          if (!fast.has_gt_positive_lc) {
            fast.has_gt_positive_lc = true;
            cnt_gt_lc_p++;
          }
          cnt_gt_lc++;
        }
      }
    std::printf("Found %d poses with %d gt loops.\n", cnt_gt_lc_p, cnt_gt_lc);
  }

  bool loadNewScan() {
    p_lidar_curr++;
    if (p_lidar_curr >= (int) laser_info_.size()) {
      std::printf("\n===\ncurrent addr %d exceeds boundary\n", p_lidar_curr);
      return false;
    }
    std::printf("\n===\nloaded scan addr %d, seq: %d, fpath: %s\n", p_lidar_curr, laser_info_[p_lidar_curr].seq,
                laser_info_[p_lidar_curr].fpath.c_str());
    return true;
  }

  const LaserScanInfo &getCurrScanInfo() const {
    if (p_lidar_curr < 0 || p_lidar_curr >= (int) laser_info_.size()) fail("current scan in range");
    return laser_info_[p_lidar_curr];
  }

  // KITTI .bin: N x (x, y, z, intensity) float32, at most 1e6 floats (tools/pointcloud_util.h:12-50)
  static std::vector<float> readKITTIPointCloudBinRaw(const std::string &path) {
    std::vector<float> buf;
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) {
      std::printf("Lidar bin file %s does not exist.\n", path.c_str());
      std::exit(-1);
    }
    buf.resize(1000000);
    const size_t n = std::fread(buf.data(), sizeof(float), buf.size(), f) / 4;
    std::fclose(f);
    buf.resize(n * 4);
    return buf;
  }

  std::shared_ptr<ContourManager> getCurrContourManager(const ContourManagerConfig &config) const {  // evaluator.h:285-302
    const LaserScanInfo &info = getCurrScanInfo();
    std::shared_ptr<ContourManager> cmng_ptr(new ContourManager(config, info.seq));
    // the .bin goes straight into the runtime's page-locked scan buffer (one asynchronous PCIe transfer, no host copy)
    float *bin = ContourManager::pinnedScanBuffer(1000000);
    FILE *f = std::fopen(info.fpath.c_str(), "rb");
    if (!f) {
      std::printf("Lidar bin file %s does not exist.\n", info.fpath.c_str());
      std::exit(-1);
    }
    const size_t n_points = std::fread(bin, sizeof(float), 1000000, f) / 4;
    std::fclose(f);
    std::string str_id = std::to_string(info.seq);
    str_id = "assigned_id_" + std::string(8 - std::min<size_t>(8, str_id.length()), '0') + str_id;
    cmng_ptr->makeBEVFromBin(bin, n_points, str_id);
    cmng_ptr->makeContoursRecurs();
    return cmng_ptr;
  }

  // Windowed loop (ContourDB::queryAddBalanceWindow): the current scan's .bin is read into `bin` (room for 1 000 000 floats, inside
  // the runtime's page-locked buffer) and handed to a new ContourManager WITHOUT running makeContoursRecurs - the window call
  // ingests all its scans as one batch.  Returns the number of points read.
  std::shared_ptr<ContourManager> loadCurrScanInto(const ContourManagerConfig &config, float *bin, size_t *n_points_out) const {
    const LaserScanInfo &info = getCurrScanInfo();
    std::shared_ptr<ContourManager> cmng_ptr(new ContourManager(config, info.seq));
    FILE *f = std::fopen(info.fpath.c_str(), "rb");
    if (!f) {
      std::printf("Lidar bin file %s does not exist.\n", info.fpath.c_str());
      std::exit(-1);
    }
    const size_t n_points = std::fread(bin, sizeof(float), 1000000, f) / 4;
    std::fclose(f);
    std::string str_id = std::to_string(info.seq);
    str_id = "assigned_id_" + std::string(8 - std::min<size_t>(8, str_id.length()), '0') + str_id;
    cmng_ptr->makeBEVFromBin(bin, n_points, str_id);
    if (n_points_out) *n_points_out = n_points;
    return cmng_ptr;
  }

  PredictionOutcome addPrediction(const std::shared_ptr<const ContourManager> &q_mng, double est_corr,
                                  const std::shared_ptr<const ContourManager> &cand_mng = nullptr,
                                  const Eigen::Isometry2d &T_est_delta_2d = Eigen::Isometry2d::Identity()) {  // evaluator.h:305-373
    const int id_tgt = q_mng->getIntID();
    const int addr_tgt = lookupNN<int>(id_tgt, assigned_seqs_, 0);
    if (addr_tgt < 0) fail("query id known");
    PredictionOutcome curr_res;
    curr_res.id_tgt = id_tgt;
    curr_res.correlation = est_corr;
    if (cand_mng) {
      const int id_src = cand_mng->getIntID();
      const int addr_src = lookupNN<int>(id_src, assigned_seqs_, 0);
      if (addr_src < 0) fail("candidate id known");
      curr_res.id_src = id_src;
      const ContourManagerConfig gen_bev_config = q_mng->getConfig();
      const Eigen::Isometry2d tf_err =
          ConstellCorrelation::evalMetricEst(T_est_delta_2d, laser_info_[addr_src].sens_pose, laser_info_[addr_tgt].sens_pose, gen_bev_config);
      const double est_trans_norm2d = ConstellCorrelation::getEstSensTF(T_est_delta_2d, gen_bev_config).translation().norm();
      const double *ts = laser_info_[addr_src].sens_pose.t, *tt = laser_info_[addr_tgt].sens_pose.t;
      const double gt_trans_norm3d = std::sqrt((ts[0] - tt[0]) * (ts[0] - tt[0]) + (ts[1] - tt[1]) * (ts[1] - tt[1]) + (ts[2] - tt[2]) * (ts[2] - tt[2]));
      std::printf(" Dist: Est2d: %.2f; GT3d: %.2f\n", est_trans_norm2d, gt_trans_norm3d);
      double err_vec[3] = {tf_err(0, 2), tf_err(1, 2), std::atan2(tf_err(1, 0), tf_err(0, 0))};
      std::printf(" Error: dx=%f, dy=%f, dtheta=%f\n", err_vec[0], err_vec[1], err_vec[2]);
      std::memcpy(curr_res.est_err, err_vec, sizeof(err_vec));
      if (est_corr >= sim_thres) {
        if (laser_info_[addr_tgt].has_gt_positive_lc && gt_trans_norm3d < 5.0) {
          curr_res.tfpn = PredictionOutcome::TP;
          tp_trans_rmse.addOneErr(err_vec);
          tp_rot_rmse.addOneErr(err_vec + 2);
        } else {
          curr_res.tfpn = PredictionOutcome::FP;
        }
      } else {
        curr_res.tfpn = laser_info_[addr_tgt].has_gt_positive_lc ? PredictionOutcome::FN : PredictionOutcome::TN;
      }
      all_trans_rmse.addOneErr(err_vec);
      all_rot_rmse.addOneErr(err_vec + 2);
    } else {
      curr_res.tfpn = laser_info_[addr_tgt].has_gt_positive_lc ? PredictionOutcome::FN : PredictionOutcome::TN;
    }
    pred_records.push_back(curr_res);
    return curr_res;
  }

  void savePredictionResults(const std::string &sav_path) const {  // evaluator.h:377-425
    std::fstream res_file(sav_path, std::ios::out);
    if (!res_file.good()) {
      std::cerr << "Error opening " << sav_path << std::endl;
      return;
    }
    for (const auto &rec : pred_records) {
      const int addr_tgt = lookupNN<int>(rec.id_tgt, assigned_seqs_, 0);
      if (addr_tgt < 0) fail("record id known");
      res_file << rec.tfpn << "\t";
      const std::string str_rep_tgt = laser_info_[addr_tgt].fpath;
      std::string str_rep_src;
      if (rec.id_src < 0) {
        res_file << rec.id_tgt << "-x" << "\t";
        str_rep_src = "x";
      } else {
        const int addr_src = lookupNN<int>(rec.id_src, assigned_seqs_, 0);
        if (addr_src < 0) fail("record src id known");
        res_file << rec.id_tgt << "-" << rec.id_src << "\t";
        str_rep_src = laser_info_[addr_src].fpath;
      }
      res_file << rec.correlation << "\t" << rec.est_err[0] << "\t" << rec.est_err[1] << "\t" << rec.est_err[2] << "\t";
      const int str_max_len = 32;
      const int beg_tgt = std::max(0, (int) str_rep_tgt.length() - str_max_len);
      const int beg_src = std::max(0, (int) str_rep_src.length() - str_max_len);
      res_file << str_rep_tgt.substr(beg_tgt) << "\t" << str_rep_src.substr(beg_src) << "\n";
    }
    res_file.close();
    std::printf("Outcome saved successfully.\n");
  }

  double getTPMeanTrans() const { return tp_trans_rmse.getMean(); }
  double getTPMeanRot() const { return tp_rot_rmse.getMean(); }
  double getTPRMSETrans() const { return tp_trans_rmse.getRMSE(); }
  double getTPRMSERot() const { return tp_rot_rmse.getRMSE(); }
};
