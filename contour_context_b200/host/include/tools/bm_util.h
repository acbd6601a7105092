// tools/bm_util.h (facade) — stage timers with the interface the reference's library code expects from the executable
// (`extern SequentialTimeProfiler stp`, include/cont2/contour_db.h:23; call sites :729,755,763,772,784,787): start(),
// record(name), lap(), printScreen().  Written for this repository; same stage names so tables line up with
// log/timing_cont2_paper.txt.
#pragma once
#include <chrono>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

class TicToc {
 public:
  TicToc() { tic(); }
  void tic() { t0_ = std::chrono::steady_clock::now(); }
  double toc() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count(); }
  double toctic() {
    double r = toc();
    tic();
    return r;
  }

 private:
  std::chrono::time_point<std::chrono::steady_clock> t0_;
};

class SequentialTimeProfiler {
  struct Entry {
    int order = 0, count = 0;
    double sum = 0, sumsq = 0;
  };
  TicToc clk_;
  std::map<std::string, Entry> logs_;
  int loops_ = 0;

 public:
  void start() { clk_.tic(); }
  void record(const std::string &name) {
    const double dt = clk_.toc();
    Entry &e = logs_[name];
    if (e.count == 0) e.order = (int) logs_.size();
    e.count++;
    e.sum += dt;
    e.sumsq += dt * dt;
    clk_.tic();
  }
  // a duration measured elsewhere (CUDA events of the kernels that replace a stage), booked under the reference's stage name
  void recordSeconds(const std::string &name, double dt) {
    Entry &e = logs_[name];
    if (e.count == 0) e.order = (int) logs_.size();
    e.count++;
    e.sum += dt;
    e.sumsq += dt * dt;
    clk_.tic();
  }
  void lap() { loops_++; }
  void printScreen() const {
    std::vector<std::pair<std::string, Entry>> v(logs_.begin(), logs_.end());
    std::printf("%-18s %10s %12s %12s\n", "Name", "Count", "Average(s)", "Per loop(s)");
    for (int k = 1; k <= (int) v.size(); ++k)
      for (auto &p : v)
        if (p.second.order == k)
          std::printf("%-18s %10d %12.6f %12.6f\n", p.first.c_str(), p.second.count, p.second.sum / p.second.count,
                      loops_ ? p.second.sum / loops_ : 0.0);
  }
};
