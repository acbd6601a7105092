// facade_statics_test.cpp — the host restatements of the reference's public statics (BCI::checkConstellSim,
// ContourManager::checkConstellCorrespSim / getTFFromConstell / checkContPairSim, ContourView::checkSim; facade headers) against the
// per-hint records of the device cascade (c2g_query trace), hint by hint, in the order CandidateManager::checkCandWithHint applies them
// (include/cont2/contour_db.h:374-437).  usage: facade_statics_test <list.txt> <n_db>   (lines "<ts> <scan.bin>"; the first n_db scans
// form the database, the rest are queries).  Prints "statics_ok <hints checked> <hints that reached addProposal>".
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "cont2/contour_db.h"

SequentialTimeProfiler stp;

#define REQUIRE(cond)                                                                       \
  do {                                                                                      \
    if (!(cond)) {                                                                          \
      std::printf("MISMATCH %s (query %d hint %d) at %s:%d\n", #cond, qi, h, __FILE__, __LINE__); \
      return 2;                                                                             \
    }                                                                                       \
  } while (0)

int main(int argc, char **argv) {
  if (argc < 3) return 1;
  const int n_db = std::atoi(argv[2]);
  ContourManagerConfig cm_config;
  cm_config.lv_grads_ = {1.5f, 2.f, 2.5f, 3.f, 3.5f, 4.f};
  ContourDBConfig db_config;
  db_config.q_levels_ = {1, 2, 3};
  CandidateScoreEnsemble lb, ub;
  lb.sim_constell.i_ovlp_sum = 3, lb.sim_constell.i_ovlp_max_one = 3, lb.sim_constell.i_in_ang_rng = 3;
  lb.sim_pair.i_indiv_sim = 3, lb.sim_pair.i_orie_sim = 4;
  lb.sim_post.correlation = 0.3f, lb.sim_post.area_perc = 0.03f, lb.sim_post.neg_est_dist = -5.01f;
  ub.sim_constell.i_ovlp_sum = 6, ub.sim_constell.i_ovlp_max_one = 6, ub.sim_constell.i_in_ang_rng = 6;
  ub.sim_pair.i_indiv_sim = 6, ub.sim_pair.i_orie_sim = 6;
  ub.sim_post.correlation = 0.75f, ub.sim_post.area_perc = 0.15f, ub.sim_post.neg_est_dist = -5.0f;
  ContourDB db(db_config);
  std::vector<std::shared_ptr<ContourManager>> scans;
  std::ifstream list(argv[1]);
  std::string line;
  while (std::getline(list, line)) {
    std::istringstream ss(line);
    double ts;
    std::string path;
    if (!(ss >> ts >> path)) continue;
    std::shared_ptr<ContourManager> cm(new ContourManager(cm_config, (int) scans.size()));
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) return 1;
    std::vector<float> buf(1000000);
    const size_t n = std::fread(buf.data(), sizeof(float), buf.size(), f) / 4;
    std::fclose(f);
    cm->makeBEVFromBin(buf.data(), n, "s");
    cm->makeContoursRecurs();
    if ((int) scans.size() < n_db) {
      db.addScan(cm, ts);
      db.pushAndBalance((int) scans.size(), ts);
    }
    scans.push_back(cm);
  }
  for (int k = 0; k < 16; ++k) db.pushAndBalance(k, 1.0e5 + k);  // every buffered key into its tree
  c2g_ctx *ctx = c2g_host::context();
  c2g_score_ensemble clb = {3, 3, 3, 3, 4, 0.3f, 0.03f, -5.01f}, cub = {6, 6, 6, 6, 6, 0.75f, 0.15f, -5.0f};
  const int per_q = 3 * C2G_MAX_PIV * db_config.nnk_;
  std::vector<c2g_hint> hints((size_t) per_q);
  std::vector<c2g_pair_score> scores((size_t) per_q);
  int n_checked = 0, n_passed = 0;
  for (int qi = n_db; qi < (int) scans.size(); ++qi) {
    const ContourManager &tgt = *scans[qi];
    c2g_query_result res;
    if (c2g_query(ctx, tgt.deviceSlot(), 1, &clb, &cub, &res, hints.data(), scores.data())) return 1;
    for (int h = 0; h < per_q; ++h) {
      if (hints[h].cand_gidx < 0) continue;
      const ContourManager &src = *scans[hints[h].cand_gidx];
      const c2g_pair_score &sc = scores[h];
      const ConstellationPair anchor(hints[h].level, hints[h].cand_seq, hints[h].q_seq);
      ++n_checked;
      if (!ContourManager::checkContPairSim(src, tgt, anchor, db_config.cont_sim_cfg_)) {
        REQUIRE(sc.passed == 0);
        continue;
      }
      std::vector<ConstellationPair> c1, c2;
      const ScoreConstellSim s1 =
          BCI::checkConstellSim(src.getBCI(anchor.level, anchor.seq_src), tgt.getBCI(anchor.level, anchor.seq_tgt), lb.sim_constell, c1);
      REQUIRE(s1.i_ovlp_sum == sc.constell[0] && s1.i_ovlp_max_one == sc.constell[1] && s1.i_in_ang_rng == sc.constell[2]);
      if (s1.i_ovlp_sum < lb.sim_constell.i_ovlp_sum || s1.i_ovlp_max_one < lb.sim_constell.i_ovlp_max_one || s1.i_in_ang_rng < lb.sim_constell.i_in_ang_rng) {
        REQUIRE(sc.passed == -1);
        continue;
      }
      std::vector<float> area;
      const ScorePairwiseSim s2 = ContourManager::checkConstellCorrespSim(src, tgt, c1, lb.sim_pair, db_config.cont_sim_cfg_, c2, area);
      REQUIRE(s2.i_indiv_sim == sc.pairwise[0] && s2.i_orie_sim == sc.pairwise[1]);
      if (s2.i_indiv_sim < lb.sim_pair.i_indiv_sim || s2.i_orie_sim < lb.sim_pair.i_orie_sim) {
        REQUIRE(sc.passed == -2);
        continue;
      }
      REQUIRE(sc.passed == 1 && sc.n_pairs == (int) c2.size() && area.size() == c2.size());
      unsigned long long bits[C2G_PAIR_WORDS] = {0};
      for (const auto &p : c2) {
        const int bit = (p.level - 1) * 100 + p.seq_src * 10 + p.seq_tgt;
        bits[bit >> 6] |= 1ull << (bit & 63);
      }
      for (int w = 0; w < C2G_PAIR_WORDS; ++w) REQUIRE(bits[w] == sc.pair_bits[w]);
      const Eigen::Isometry2d T = ContourManager::getTFFromConstell(src, tgt, c2.begin(), c2.end());
      REQUIRE(std::fabs(T(0, 0) - sc.T[0]) < 1e-9 && std::fabs(T(1, 0) - sc.T[1]) < 1e-9 && std::fabs(T(0, 2) - sc.T[2]) < 1e-9 &&
              std::fabs(T(1, 2) - sc.T[3]) < 1e-9);
      ++n_passed;
    }
  }
  std::printf("statics_ok %d %d\n", n_checked, n_passed);
  return 0;
}
