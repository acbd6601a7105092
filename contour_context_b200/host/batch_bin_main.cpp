// batch_bin_main.cpp — ROS-free equivalent of the reference's cont2_batch_bin_test main loop
// (test/batch_bin_test.cpp:105-247: load scan -> make bev -> query -> add -> balance), driving the facade classes with the
// same call sequence.  Input: a text file with one "<timestamp> <path/to/scan.bin>" per line (the layout of the
// reference's ts-lidar_bins-*.txt lists) and, optionally, MulRan/KITTI parameter selection by argv.
//   usage: cont2_batch_bin <list.txt> [kitti|mulran]
// With --eval the full harness of the reference runs (ContLCDEvaluator, test/batch_bin_test.cpp:131-237): ground-truth poses
// are associated to the scans, every prediction is classified TP/FP/TN/FN and the outcome file of scripts/pr_mpe.py is
// written.
//   usage: cont2_batch_bin --eval <ts-sens_pose.txt> <ts-lidar_bins.txt> <outcome.txt> [kitti|mulran] [correlation_thres]
// With --window W (first argument pair, either mode) the loop runs W scans at a time through ContourDB::queryAddBalanceWindow: the
// scans of a window are read back to back into one page-locked buffer, ingested as one batch and queried against exactly the
// database state the scan-by-scan loop would have shown each of them; the printed lines and the outcome file are the same.
//   usage: cont2_batch_bin --window 148 <list.txt> ...   |   cont2_batch_bin --window 148 --eval ...
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "cont2/contour_db.h"
#include "eval/evaluator.h"

SequentialTimeProfiler stp;  // the library's stage timers write here, like the reference executable

// tools/pointcloud_util.h:12-50 (format: N x 4 f32, at most 1 000 000 floats): reads straight into the runtime's page-locked
// scan buffer, returns the number of points
static size_t readKITTIBin(const std::string &path, float *buf, size_t cap_floats) {
  FILE *f = std::fopen(path.c_str(), "rb");
  if (!f) {
    std::printf("Lidar bin file %s does not exist.\n", path.c_str());
    std::exit(-1);
  }
  const size_t n = std::fread(buf, sizeof(float), cap_floats, f) / 4;
  std::fclose(f);
  return n;
}

static void printOutcome(const PredictionOutcome &pred_res, const ContLCDEvaluator &evaluator, int &cnt_tp, int &cnt_fn, int &cnt_fp) {
  switch (pred_res.tfpn) {
    case PredictionOutcome::TP: std::printf("Prediction outcome: TP\n"); cnt_tp++; break;
    case PredictionOutcome::FP: std::printf("Prediction outcome: FP\n"); cnt_fp++; break;
    case PredictionOutcome::TN: std::printf("Prediction outcome: TN\n"); break;
    case PredictionOutcome::FN: std::printf("Prediction outcome: FN\n"); cnt_fn++; break;
  }
  std::printf("TP Error mean: t:%7.4f m, r:%7.4f rad\n", evaluator.getTPMeanTrans(), evaluator.getTPMeanRot());
  std::printf("TP Error rmse: t:%7.4f m, r:%7.4f rad\n", evaluator.getTPRMSETrans(), evaluator.getTPRMSERot());
  std::printf("Accumulated tp poses: %d\nAccumulated fn poses: %d\nAccumulated fp poses: %d\n", cnt_tp, cnt_fn, cnt_fp);
}

static void printLC(int seq, const ContourManagerConfig &cm_config, bool found, const std::shared_ptr<const ContourManager> &cand, double corr,
                    const Eigen::Isometry2d &T, int &n_pos) {
  if (found) {
    n_pos++;
    const double est = ConstellCorrelation::getEstSensTF(T, cm_config).translation().norm();
    std::printf("LC %d -> %d corr %.6f  T(bev) = [%.4f %.4f %.4f]  est. dist %.3f m\n", seq, cand->getIntID(), corr, T(0, 2), T(1, 2),
                std::atan2(T(1, 0), T(0, 0)), est);
  } else {
    std::printf("LC %d -> none\n", seq);
  }
}

int main(int argc, char **argv) {
  int window = 1;
  if (argc >= 3 && std::string(argv[1]) == "--window") {
    window = std::atoi(argv[2]);
    if (window < 1) window = 1;
    setenv("C2G_WINDOW", argv[2], 1);  // read by the runtime when the context is created
    argv += 2;
    argc -= 2;
  }
  if (argc < 2) {
    std::printf("usage: %s <ts-lidar_bins.txt> [kitti|mulran]\n       %s --eval <ts-sens_pose.txt> <ts-lidar_bins.txt> <outcome.txt> [kitti|mulran] [corr_thres]\n",
                argv[0], argv[0]);
    return 1;
  }
  const bool eval_mode = std::string(argv[1]) == "--eval";
  if (eval_mode && argc < 5) {
    std::printf("--eval needs <ts-sens_pose.txt> <ts-lidar_bins.txt> <outcome.txt>\n");
    return 1;
  }
  const int dataset_arg = eval_mode ? 5 : 2;
  const bool mulran = argc > dataset_arg && std::string(argv[dataset_arg]) == "mulran";
  ContourManagerConfig cm_config;  // config/batch_bin_test_config.yaml:28-46
  cm_config.lv_grads_ = mulran ? std::vector<float>{1.0f, 2.5f, 4.0f, 5.5f, 7.0f, 8.5f} : std::vector<float>{1.5f, 2.f, 2.5f, 3.f, 3.5f, 4.f};
  ContourDBConfig db_config;       // config/batch_bin_test_config.yaml:6-23
  db_config.q_levels_ = {1, 2, 3};
  db_config.cont_sim_cfg_.ta_h_bar = mulran ? 0.75f : 0.3f;
  CandidateScoreEnsemble thres_lb_, thres_ub_;  // config/batch_bin_test_config.yaml:70-87
  thres_lb_.sim_constell.i_ovlp_sum = 3, thres_lb_.sim_constell.i_ovlp_max_one = 3, thres_lb_.sim_constell.i_in_ang_rng = 3;
  thres_lb_.sim_pair.i_indiv_sim = 3, thres_lb_.sim_pair.i_orie_sim = 4;
  thres_lb_.sim_post.correlation = 0.3f, thres_lb_.sim_post.area_perc = 0.03f, thres_lb_.sim_post.neg_est_dist = -5.01f;
  thres_ub_.sim_constell.i_ovlp_sum = 6, thres_ub_.sim_constell.i_ovlp_max_one = 6, thres_ub_.sim_constell.i_in_ang_rng = 6;
  thres_ub_.sim_pair.i_indiv_sim = 6, thres_ub_.sim_pair.i_orie_sim = 6;
  thres_ub_.sim_post.correlation = 0.75f, thres_ub_.sim_post.area_perc = 0.15f, thres_ub_.sim_post.neg_est_dist = -5.0f;

  ContourDB contour_db(db_config);
  if (eval_mode) {  // BatchBinSpinner::spinOnce (test/batch_bin_test.cpp:105-247) without ROS
    const double corr_thres = argc > 6 ? std::atof(argv[6]) : 0.0;  // correlation_thres of the yaml (config/batch_bin_test_config.yaml:2)
    ContLCDEvaluator evaluator(argv[2], argv[3], corr_thres);
    int cnt_tp = 0, cnt_fn = 0, cnt_fp = 0;
    if (window > 1) {
      float *bins = ContourManager::pinnedScanBuffer((size_t) window * 1000000);
      bool more = true;
      while (more) {
        std::vector<std::shared_ptr<ContourManager>> scans;
        std::vector<double> tss;
        std::vector<int> seeds;
        size_t used = 0;
        stp.lap();
        stp.start();
        while ((int) scans.size() < window && (more = evaluator.loadNewScan())) {
          const auto info = evaluator.getCurrScanInfo();
          size_t n_points = 0;
          scans.push_back(evaluator.loadCurrScanInto(cm_config, bins + used, &n_points));
          used += 4 * n_points;
          tss.push_back(info.ts);
          seeds.push_back(info.seq);
        }
        stp.record("read scans");
        if (scans.empty()) break;
        std::vector<ContourDB::WindowResult> res;
        contour_db.queryAddBalanceWindow(scans, tss, seeds, thres_lb_, thres_ub_, res);
        for (size_t i = 0; i < scans.size(); ++i) {
          const PredictionOutcome pred_res = res[i].found ? evaluator.addPrediction(scans[i], res[i].corr, res[i].cand, res[i].tf)
                                                          : evaluator.addPrediction(scans[i], 0.0);
          printOutcome(pred_res, evaluator, cnt_tp, cnt_fn, cnt_fp);
        }
      }
      evaluator.savePredictionResults(argv[4]);
      stp.printScreen();
      return 0;
    }
    while (evaluator.loadNewScan()) {
      const auto laser_info_tgt = evaluator.getCurrScanInfo();
      stp.lap();
      stp.start();
      std::shared_ptr<ContourManager> ptr_cm_tgt = evaluator.getCurrContourManager(cm_config);
      stp.record("make bev");
      ptr_cm_tgt->clearImage();
      std::vector<std::shared_ptr<const ContourManager>> ptr_cands;
      std::vector<double> cand_corr;
      std::vector<Eigen::Isometry2d> bev_tfs;
      contour_db.queryRangedKNN(ptr_cm_tgt, thres_lb_, thres_ub_, ptr_cands, cand_corr, bev_tfs);
      if (ptr_cands.size() >= 2) std::abort();  // CHECK(ptr_cands.size() < 2) (batch_bin_test.cpp:187)
      PredictionOutcome pred_res;
      if (ptr_cands.empty())
        pred_res = evaluator.addPrediction(ptr_cm_tgt, 0.0);
      else
        pred_res = evaluator.addPrediction(ptr_cm_tgt, cand_corr[0], ptr_cands[0], bev_tfs[0]);
      printOutcome(pred_res, evaluator, cnt_tp, cnt_fn, cnt_fp);
      stp.start();
      contour_db.addScan(ptr_cm_tgt, laser_info_tgt.ts);
      contour_db.pushAndBalance(laser_info_tgt.seq, laser_info_tgt.ts);
      stp.record("Update database");
    }
    evaluator.savePredictionResults(argv[4]);
    stp.printScreen();
    return 0;
  }
  std::ifstream list(argv[1]);
  std::string line;
  int seq = 0, n_pos = 0;
  if (window > 1) {
    float *bins = ContourManager::pinnedScanBuffer((size_t) window * 1000000);
    bool more = true;
    while (more) {
      std::vector<std::shared_ptr<ContourManager>> scans;
      std::vector<double> tss;
      std::vector<int> seeds;
      size_t used = 0;
      stp.lap();
      stp.start();
      while ((int) scans.size() < window && (more = (bool) std::getline(list, line))) {
        std::istringstream ss(line);
        double ts;
        std::string path;
        if (!(ss >> ts >> path)) continue;
        std::shared_ptr<ContourManager> cm(new ContourManager(cm_config, seq));
        const size_t n_points = readKITTIBin(path, bins + used, 1000000);
        cm->makeBEVFromBin(bins + used, n_points, "assigned_id_" + std::to_string(seq));
        used += 4 * n_points;
        scans.push_back(cm);
        tss.push_back(ts);
        seeds.push_back(seq);
        seq++;
      }
      stp.record("read scans");
      if (scans.empty()) break;
      std::vector<ContourDB::WindowResult> res;
      contour_db.queryAddBalanceWindow(scans, tss, seeds, thres_lb_, thres_ub_, res);
      for (size_t i = 0; i < scans.size(); ++i) printLC(scans[i]->getIntID(), cm_config, res[i].found, res[i].cand, res[i].corr, res[i].tf, n_pos);
    }
    std::printf("scans: %d, positive predictions: %d\n", seq, n_pos);
    stp.printScreen();
    return 0;
  }
  while (std::getline(list, line)) {
    std::istringstream ss(line);
    double ts;
    std::string path;
    if (!(ss >> ts >> path)) continue;
    stp.lap();
    stp.start();
    std::shared_ptr<ContourManager> ptr_cm_tgt(new ContourManager(cm_config, seq));
    float *bin = ContourManager::pinnedScanBuffer(1000000);
    const size_t n_points = readKITTIBin(path, bin, 1000000);
    ptr_cm_tgt->makeBEVFromBin(bin, n_points, "assigned_id_" + std::to_string(seq));
    ptr_cm_tgt->makeContoursRecurs();
    stp.record("make bev");
    ptr_cm_tgt->clearImage();

    std::vector<std::shared_ptr<const ContourManager>> ptr_cands;
    std::vector<double> cand_corr;
    std::vector<Eigen::Isometry2d> bev_tfs;
    contour_db.queryRangedKNN(ptr_cm_tgt, thres_lb_, thres_ub_, ptr_cands, cand_corr, bev_tfs);
    if (ptr_cands.size() >= 2) std::abort();  // CHECK(ptr_cands.size() < 2) (batch_bin_test.cpp:187)
    printLC(seq, cm_config, !ptr_cands.empty(), ptr_cands.empty() ? nullptr : ptr_cands[0], ptr_cands.empty() ? 0.0 : cand_corr[0],
            ptr_cands.empty() ? Eigen::Isometry2d() : bev_tfs[0], n_pos);
    stp.start();
    contour_db.addScan(ptr_cm_tgt, ts);
    contour_db.pushAndBalance(seq, ts);
    stp.record("Update database");
    seq++;
  }
  std::printf("scans: %d, positive predictions: %d\n", seq, n_pos);
  stp.printScreen();
  return 0;
}
