// ORACLE — TEST INFRASTRUCTURE ONLY (see c2o_ingest.hpp header).
//
// c2o_refine.hpp: CPU restatement of ConstellCorrelation::calcCorrelation (include/cont2/correlation.h:206-238): the
// GMM-L2 cost of GMMPair::operator() (correlation.h:125-152) minimised over (x, y, theta) by ceres::Solve on a
// ceres::GradientProblem with `max_num_iterations = 10` and every other GradientProblemSolver option at its default.
//
// PARITY UNPINNED.  Ceres (find_package(Ceres 2), CMakeLists.txt:52) is neither vendored under /root/reference nor
// installed in this image, and the reference ships no golden vector for this step.  What follows restates, from the
// published Ceres 2.x algorithm, the pieces that the default options select:
//   * AutoDiffFirstOrderFunction<GMMPair, 3>: forward-mode Jets with 3 partials (ceres/jet.h operator rules);
//   * LineSearchMinimizer (internal/ceres/line_search_minimizer.cc): LBFGS direction, rank 20, no eigenvalue scaling,
//     initial step min(1, 1/|g|_inf) then min(1, 2 (f_k - f_{k-1}) / phi'(0)), termination on gradient_tolerance 1e-10,
//     parameter_tolerance 1e-8, function_tolerance 1e-6, max_num_line_search_direction_restarts 5;
//   * WolfeLineSearch (internal/ceres/line_search.cc): bracketing + zoom, CUBIC interpolation, sufficient decrease 1e-4,
//     curvature 0.9, max step expansion 10, min step size 1e-9, at most 20 step-size iterations;
//   * polynomial.cc: interpolating polynomial through (value, gradient) samples solved by full-pivot LU, minimised
//     over an interval via closed-form roots of the derivative;
//   * LowRankInverseHessian two-loop recursion with the 1e-14 secant-condition skip.
// Jet evaluation order follows how Eigen evaluates the 2x2 expressions in GMMPair::operator() (coefficient-based lazy
// products, nested products through a temporary, 2x2 inverse = adjugate * (1/det)).
// tests/test_refine.py checks this file against finite differences (gradient) and scipy's BFGS (optimum), which is the
// strongest pin available here.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace c2o {
namespace refine {

// ---- ceres::Jet<double, 3> ------------------------------------------------------------------------------------------
struct J3 {
  double a;
  double v[3];
};
inline J3 jconst(double s) { return J3{s, {0.0, 0.0, 0.0}}; }
inline J3 jvar(double s, int k) {
  J3 r = jconst(s);
  r.v[k] = 1.0;
  return r;
}
inline J3 operator+(const J3 &f, const J3 &g) { return J3{f.a + g.a, {f.v[0] + g.v[0], f.v[1] + g.v[1], f.v[2] + g.v[2]}}; }
inline J3 operator-(const J3 &f, const J3 &g) { return J3{f.a - g.a, {f.v[0] - g.v[0], f.v[1] - g.v[1], f.v[2] - g.v[2]}}; }
inline J3 operator-(const J3 &f) { return J3{-f.a, {-f.v[0], -f.v[1], -f.v[2]}}; }
inline J3 operator+(const J3 &f, double s) { return J3{f.a + s, {f.v[0], f.v[1], f.v[2]}}; }
inline J3 operator-(const J3 &f, double s) { return J3{f.a - s, {f.v[0], f.v[1], f.v[2]}}; }
inline J3 operator*(const J3 &f, const J3 &g) {
  return J3{f.a * g.a, {f.a * g.v[0] + f.v[0] * g.a, f.a * g.v[1] + f.v[1] * g.a, f.a * g.v[2] + f.v[2] * g.a}};
}
inline J3 operator*(const J3 &f, double s) { return J3{f.a * s, {f.v[0] * s, f.v[1] * s, f.v[2] * s}}; }
inline J3 operator*(double s, const J3 &f) { return J3{f.a * s, {f.v[0] * s, f.v[1] * s, f.v[2] * s}}; }
inline J3 operator/(const J3 &f, const J3 &g) {
  const double g_a_inverse = 1.0 / g.a;
  const double f_a_by_g_a = f.a * g_a_inverse;
  return J3{f_a_by_g_a,
            {(f.v[0] - f_a_by_g_a * g.v[0]) * g_a_inverse, (f.v[1] - f_a_by_g_a * g.v[1]) * g_a_inverse,
             (f.v[2] - f_a_by_g_a * g.v[2]) * g_a_inverse}};
}
inline J3 operator/(double s, const J3 &g) {
  const double minus_s_g_a_inverse2 = -s / (g.a * g.a);
  return J3{s / g.a, {g.v[0] * minus_s_g_a_inverse2, g.v[1] * minus_s_g_a_inverse2, g.v[2] * minus_s_g_a_inverse2}};
}
inline J3 jsqrt(const J3 &f) {
  const double tmp = std::sqrt(f.a);
  const double two_a_inverse = 1.0 / (2.0 * tmp);
  return J3{tmp, {f.v[0] * two_a_inverse, f.v[1] * two_a_inverse, f.v[2] * two_a_inverse}};
}
inline J3 jexp(const J3 &f) {
  const double tmp = std::exp(f.a);
  return J3{tmp, {tmp * f.v[0], tmp * f.v[1], tmp * f.v[2]}};
}
inline J3 jcos(const J3 &f) {
  const double ms = -std::sin(f.a);
  return J3{std::cos(f.a), {ms * f.v[0], ms * f.v[1], ms * f.v[2]}};
}
inline J3 jsin(const J3 &f) {
  const double c = std::cos(f.a);
  return J3{std::sin(f.a), {c * f.v[0], c * f.v[1], c * f.v[2]}};
}

// ---- the problem: ellipses of both scans + the pairs pre-selected at T_init (correlation.h:84-96) -------------------
struct PairSel {
  int li, si, ti;
};

struct Problem {
  const GMMScanData *src = nullptr, *tgt = nullptr;
  std::vector<PairSel> pairs;
  int n_eval = 0;
};

inline Problem makeProblem(const GMMScanData &src, const GMMScanData &tgt, const Iso2 &T_init) {
  Problem P;
  P.src = &src;
  P.tgt = &tgt;
  for (int li = 0; li < 4; ++li)
    for (size_t si = 0; si < src.ell[li].size(); si++)
      for (size_t ti = 0; ti < tgt.ell[li].size(); ti++) {
        const GMMEllipse &a = src.ell[li][si], &b = tgt.ell[li][ti];
        double qx, qy;
        T_init.apply(a.mu[0], a.mu[1], qx, qy);
        const double dx = qx - b.mu[0], dy = qy - b.mu[1];
        if (std::sqrt(dx * dx + dy * dy) < 3.0 * (src.max_majax[li][si] + tgt.max_majax[li][ti]))
          P.pairs.push_back(PairSel{li, (int) si, (int) ti});
      }
  return P;
}

// GMMPair::operator()<Jet> (correlation.h:125-152): cost and gradient at p = (x, y, theta)
inline J3 evalJet(Problem &P, const double p[3]) {
  const double scale = 2.0;
  P.n_eval++;
  const J3 x = jvar(p[0], 0), y = jvar(p[1], 1), theta = jvar(p[2], 2);
  const J3 c = jcos(theta), s = jsin(theta);
  const J3 R00 = c, R01 = -s, R10 = s, R11 = c;
  J3 cost = jconst(0.0);
  for (const PairSel &pr : P.pairs) {
    const GMMEllipse &a = P.src->ell[pr.li][pr.si], &b = P.tgt->ell[pr.li][pr.ti];
    // tmp = R * cov_src  (cov column-major: cov[0]=(0,0) cov[1]=(1,0) cov[2]=(0,1) cov[3]=(1,1))
    const J3 t00 = R00 * a.cov[0] + R01 * a.cov[1], t01 = R00 * a.cov[2] + R01 * a.cov[3];
    const J3 t10 = R10 * a.cov[0] + R11 * a.cov[1], t11 = R10 * a.cov[2] + R11 * a.cov[3];
    // m = tmp * R^T
    const J3 m00 = t00 * R00 + t01 * R01, m01 = t00 * R10 + t01 * R11;
    const J3 m10 = t10 * R00 + t11 * R01, m11 = t10 * R10 + t11 * R11;
    const J3 c00 = scale * (m00 + b.cov[0]), c10 = scale * (m10 + b.cov[1]);
    const J3 c01 = scale * (m01 + b.cov[2]), c11 = scale * (m11 + b.cov[3]);
    const J3 mx = (R00 * a.mu[0] + R01 * a.mu[1]) + x - b.mu[0];
    const J3 my = (R10 * a.mu[0] + R11 * a.mu[1]) + y - b.mu[1];
    const J3 det = c00 * c11 - c10 * c01;
    const J3 invdet = jconst(1.0) / det;
    const J3 i00 = c11 * invdet, i10 = (-c10) * invdet, i01 = (-c01) * invdet, i11 = c00 * invdet;
    const J3 r0 = -0.5 * mx, r1 = -0.5 * my;
    const J3 q0 = r0 * i00 + r1 * i10, q1 = r0 * i01 + r1 * i11;
    const J3 qua = q0 * mx + q1 * my;
    cost = cost + ((-b.w * a.w * 1.0) / jsqrt(det)) * jexp(qua);
  }
  return cost;
}

// ---- polynomial.cc ----------------------------------------------------------------------------------------------------
struct Sample {  // FunctionSample
  double x = 0, value = 0, gradient = 0;
  bool value_is_valid = false, gradient_is_valid = false;
  double vx[3] = {0, 0, 0}, vg[3] = {0, 0, 0};  // vector_x, vector_gradient
  bool vector_x_is_valid = false, vector_gradient_is_valid = false;
};

inline double evalPoly(const std::vector<double> &poly, double x) {
  double v = 0.0;
  for (double cf : poly) v = v * x + cf;
  return v;
}

// x = A^-1 b through LU with complete pivoting (what Eigen's FullPivLU::solve does: column-major search for the
// biggest remaining coefficient, rank cut at eps * size * max pivot)
inline std::vector<double> fullPivLuSolve(std::vector<double> A, std::vector<double> b, int n) {  // A row-major n x n
  std::vector<int> rowT(n), colT(n);
  int nonzero = n;
  double maxpivot = 0.0;
  auto at = [&](int r, int c) -> double & { return A[r * n + c]; };
  for (int k = 0; k < n; ++k) {
    int br = k, bc = k;
    double best = -1.0;
    for (int c = k; c < n; ++c)
      for (int r = k; r < n; ++r)
        if (std::fabs(at(r, c)) > best) {
          best = std::fabs(at(r, c));
          br = r;
          bc = c;
        }
    if (best == 0.0) {
      nonzero = k;
      for (int i = k; i < n; ++i) rowT[i] = colT[i] = i;
      break;
    }
    if (best > maxpivot) maxpivot = best;
    rowT[k] = br;
    colT[k] = bc;
    if (br != k)
      for (int c = 0; c < n; ++c) std::swap(at(k, c), at(br, c));
    if (bc != k)
      for (int r = 0; r < n; ++r) std::swap(at(r, k), at(r, bc));
    if (k < n - 1)
      for (int r = k + 1; r < n; ++r) at(r, k) /= at(k, k);
    if (k < n - 1)
      for (int c = k + 1; c < n; ++c)
        for (int r = k + 1; r < n; ++r) at(r, c) -= at(r, k) * at(k, c);
  }
  // rank with the default threshold
  const double thr = 2.220446049250313e-16 * n;
  int rank = 0;
  for (int i = 0; i < nonzero; ++i) rank += (std::fabs(at(i, i)) > thr * maxpivot);
  std::vector<double> x(n, 0.0);
  if (rank == 0) return x;
  // c = P b  (row transpositions applied in order)
  std::vector<double> cvec = b;
  for (int k = 0; k < n; ++k)
    if (rowT[k] != k) std::swap(cvec[k], cvec[rowT[k]]);
  // L (unit lower) forward substitution on the first `n` rows (square)
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) cvec[i] -= at(i, j) * cvec[j];
  // U back substitution on the leading rank x rank block
  for (int i = rank - 1; i >= 0; --i) {
    for (int j = i + 1; j < rank; ++j) cvec[i] -= at(i, j) * cvec[j];
    cvec[i] /= at(i, i);
  }
  for (int i = rank; i < n; ++i) cvec[i] = 0.0;
  // undo the column permutation: Q = product of column transpositions; x = Q c
  std::vector<int> perm(n);
  for (int i = 0; i < n; ++i) perm[i] = i;
  for (int k = 0; k < n; ++k) std::swap(perm[k], perm[colT[k]]);
  for (int i = 0; i < n; ++i) x[perm[i]] = cvec[i];
  return x;
}

inline std::vector<double> findInterpolatingPolynomial(const std::vector<Sample> &samples) {
  int n = 0;
  for (const Sample &s : samples) n += (s.value_is_valid ? 1 : 0) + (s.gradient_is_valid ? 1 : 0);
  const int degree = n - 1;
  std::vector<double> lhs((size_t) n * n, 0.0), rhs(n, 0.0);
  int row = 0;
  for (const Sample &s : samples) {
    if (s.value_is_valid) {
      for (int j = 0; j <= degree; ++j) lhs[row * n + j] = std::pow(s.x, degree - j);
      rhs[row] = s.value;
      ++row;
    }
    if (s.gradient_is_valid) {
      for (int j = 0; j < degree; ++j) lhs[row * n + j] = (degree - j) * std::pow(s.x, degree - j - 1);
      rhs[row] = s.gradient;
      ++row;
    }
  }
  return fullPivLuSolve(lhs, rhs, n);
}

inline void minimizePolynomial(const std::vector<double> &poly, double x_min, double x_max, double *opt_x, double *opt_v) {
  *opt_x = (x_min + x_max) / 2.0;
  *opt_v = evalPoly(poly, *opt_x);
  const double vmin = evalPoly(poly, x_min);
  if (vmin < *opt_v) {
    *opt_v = vmin;
    *opt_x = x_min;
  }
  const double vmax = evalPoly(poly, x_max);
  if (vmax < *opt_v) {
    *opt_v = vmax;
    *opt_x = x_max;
  }
  if (poly.size() <= 2) return;
  // derivative
  const int degree = (int) poly.size() - 1;
  std::vector<double> d(degree);
  for (int i = 0; i < degree; ++i) d[i] = (degree - i) * poly[i];
  // remove leading zeros
  size_t lead = 0;
  while (lead + 1 < d.size() && d[lead] == 0.0) ++lead;
  d.erase(d.begin(), d.begin() + lead);
  std::vector<double> roots;
  const int dd = (int) d.size() - 1;
  if (dd == 1) {
    roots.push_back(-d[1] / d[0]);
  } else if (dd == 2) {
    const double a = d[0], b = d[1], c = d[2];
    const double D = b * b - 4 * a * c;
    const double sqrt_D = std::sqrt(std::fabs(D));
    if (D >= 0) {
      if (b >= 0) {
        roots.push_back((-b - sqrt_D) / (2.0 * a));
        roots.push_back((2.0 * c) / (-b - sqrt_D));
      } else {
        roots.push_back((2.0 * c) / (-b + sqrt_D));
        roots.push_back((-b + sqrt_D) / (2.0 * a));
      }
    } else {  // complex pair: the caller only looks at the real parts
      roots.push_back(-b / (2.0 * a));
      roots.push_back(-b / (2.0 * a));
    }
  }  // dd == 0: constant derivative, no roots; dd > 2 cannot happen with two (value, gradient) samples
  for (double root : roots) {
    if ((root < x_min) || (root > x_max)) continue;
    const double value = evalPoly(poly, root);
    if (value < *opt_v) {
      *opt_v = value;
      *opt_x = root;
    }
  }
}

// LineSearch::InterpolatingPolynomialMinimizingStepSize with CUBIC interpolation; `previous` is never valid here (the
// Wolfe search always passes an unused sample)
inline double interpStep(const Sample &lowerbound, const Sample &current, double min_step, double max_step) {
  if (!current.value_is_valid) return std::min(std::max(current.x * 0.5, min_step), max_step);
  std::vector<Sample> samples;
  samples.push_back(lowerbound);
  samples.push_back(current);
  const std::vector<double> poly = findInterpolatingPolynomial(samples);
  double step = 0.0, val = 0.0;
  minimizePolynomial(poly, min_step, max_step, &step, &val);
  for (const Sample &s : samples) {
    if ((s.x < min_step) || (s.x > max_step)) continue;
    const double v = evalPoly(poly, s.x);
    if (v < val) {
      step = s.x;
      val = v;
    }
  }
  return step;
}

// ---- line search -------------------------------------------------------------------------------------------------------
struct LsOptions {
  double sufficient_decrease = 1e-4, sufficient_curvature_decrease = 0.9, max_step_expansion = 10.0, min_step_size = 1e-9;
  int max_num_iterations = 20;
};

struct LsFunction {
  Problem *P;
  double pos[3], dir[3];
  double dirInf() const { return std::max(std::fabs(dir[0]), std::max(std::fabs(dir[1]), std::fabs(dir[2]))); }
  void evaluate(double x, Sample *out) const {
    *out = Sample();
    out->x = x;
    for (int k = 0; k < 3; ++k) out->vx[k] = pos[k] + x * dir[k];
    out->vector_x_is_valid = true;
    const J3 f = evalJet(*P, out->vx);
    out->value = f.a;
    for (int k = 0; k < 3; ++k) out->vg[k] = f.v[k];
    if (!std::isfinite(out->value)) return;
    out->value_is_valid = true;
    out->gradient = (dir[0] * out->vg[0] + dir[1] * out->vg[1]) + dir[2] * out->vg[2];
    if (!std::isfinite(out->gradient) || !std::isfinite(out->vg[0]) || !std::isfinite(out->vg[1]) || !std::isfinite(out->vg[2])) return;
    out->vector_gradient_is_valid = true;
    out->gradient_is_valid = true;
  }
};

struct LsSummary {
  bool success = false;
  Sample optimal_point;
  int num_iterations = 0;
};

inline bool bracketingPhase(const LsFunction &fn, const LsOptions &o, const Sample &initial, double step_estimate, Sample *low,
                            Sample *high, bool *do_zoom, LsSummary *sum) {
  Sample previous = initial, current;
  const double dmax = fn.dirInf();
  *do_zoom = false;
  *low = initial;
  fn.evaluate(step_estimate, &current);
  while (true) {
    ++sum->num_iterations;
    if (current.value_is_valid && (current.value > (initial.value + o.sufficient_decrease * initial.gradient * current.x) ||
                                   (previous.value_is_valid && current.value > previous.value))) {
      *do_zoom = true;
      *low = previous;
      *high = current;
      break;
    }
    if (current.value_is_valid && std::fabs(current.gradient) <= -o.sufficient_curvature_decrease * initial.gradient) {
      *low = current;
      *high = current;
      break;
    } else if (current.value_is_valid && current.gradient >= 0) {
      *do_zoom = true;
      *low = current;
      *high = previous;
      break;
    } else if (sum->num_iterations >= o.max_num_iterations) {
      *low = (current.value_is_valid && current.value < low->value) ? current : *low;
      break;
    }
    const double min_step = current.value_is_valid ? current.x : previous.x;
    const double max_step = current.value_is_valid ? (current.x * o.max_step_expansion) : current.x;
    const double step = interpStep(previous, current, min_step, max_step);
    if (step * dmax < o.min_step_size) return false;
    previous = current.value_is_valid ? current : previous;
    fn.evaluate(step, &current);
  }
  if (*do_zoom && std::fabs(high->x - low->x) * dmax < o.min_step_size) *do_zoom = false;
  return true;
}

inline bool zoomPhase(const LsFunction &fn, const LsOptions &o, const Sample &initial, Sample low, Sample high, Sample *solution,
                      LsSummary *sum) {
  if (low.gradient * (high.x - low.x) >= 0) {
    solution->value_is_valid = false;
    return false;
  }
  const double dmax = fn.dirInf();
  while (true) {
    *solution = low;
    if (sum->num_iterations >= o.max_num_iterations) return false;
    if (std::fabs(high.x - low.x) * dmax < o.min_step_size) return false;
    ++sum->num_iterations;
    const Sample &lower_bound_step = low.x < high.x ? low : high;
    const Sample &upper_bound_step = low.x < high.x ? high : low;
    const double step = interpStep(lower_bound_step, upper_bound_step, lower_bound_step.x, upper_bound_step.x);
    fn.evaluate(step, solution);
    if (!solution->value_is_valid || !solution->gradient_is_valid) return false;
    if ((solution->value > (initial.value + o.sufficient_decrease * initial.gradient * solution->x)) ||
        (solution->value >= low.value)) {
      high = *solution;
      continue;
    }
    if (std::fabs(solution->gradient) <= -o.sufficient_curvature_decrease * initial.gradient) {
      break;
    } else if (solution->gradient * (high.x - low.x) >= 0) {
      high = low;
    }
    low = *solution;
  }
  return true;
}

inline void wolfeSearch(const LsFunction &fn, const LsOptions &o, double step_estimate, double initial_cost, double initial_gradient,
                        LsSummary *sum) {
  Sample initial;
  initial.x = 0.0;
  initial.value = initial_cost;
  initial.gradient = initial_gradient;
  initial.value_is_valid = initial.gradient_is_valid = true;
  for (int k = 0; k < 3; ++k) initial.vx[k] = fn.pos[k];
  initial.vector_x_is_valid = true;
  bool do_zoom = false;
  Sample solution, low, high;
  if (!bracketingPhase(fn, o, initial, step_estimate, &low, &high, &do_zoom, sum)) return;
  if (!do_zoom) {
    sum->optimal_point = low;
    sum->success = true;
    return;
  }
  if (!zoomPhase(fn, o, initial, low, high, &solution, sum) && !solution.value_is_valid) return;
  // zoomPhase keeps its own copy of the bracket; the comparison below is against the bracket's low end at the time the
  // zoom was entered, as in WolfeLineSearch::DoSearch
  if (!solution.value_is_valid || solution.value > low.value)
    sum->optimal_point = low;
  else
    sum->optimal_point = solution;
  sum->success = true;
}

// ---- L-BFGS direction + line-search minimizer -----------------------------------------------------------------------
struct Lbfgs {
  static constexpr int kRank = 20;
  double s[kRank][3], y[kRank][3], sy[kRank];
  std::vector<int> indices;  // circular buffer order, oldest first
  void reset() { indices.clear(); }
  void update(const double dx[3], const double dg[3]) {
    const double d = (dx[0] * dg[0] + dx[1] * dg[1]) + dx[2] * dg[2];
    if (d <= 1e-14) return;
    int next = (int) indices.size();
    if (next == kRank) {
      next = indices.front();
      indices.erase(indices.begin());
    }
    indices.push_back(next);
    for (int k = 0; k < 3; ++k) {
      s[next][k] = dx[k];
      y[next][k] = dg[k];
    }
    sy[next] = d;
  }
  void rightMultiply(const double g[3], double out[3]) const {
    double alpha[kRank];
    for (int k = 0; k < 3; ++k) out[k] = g[k];
    for (auto it = indices.rbegin(); it != indices.rend(); ++it) {
      const int i = *it;
      const double a = ((s[i][0] * out[0] + s[i][1] * out[1]) + s[i][2] * out[2]) / sy[i];
      for (int k = 0; k < 3; ++k) out[k] -= a * y[i][k];
      alpha[i] = a;
    }
    for (int i : indices) {
      const double beta = ((y[i][0] * out[0] + y[i][1] * out[1]) + y[i][2] * out[2]) / sy[i];
      for (int k = 0; k < 3; ++k) out[k] += s[i][k] * (alpha[i] - beta);
    }
  }
};

struct Result {
  double x[3];
  double initial_cost = 0, final_cost = -1.0;  // GradientProblemSolver::Summary defaults to -1 until a usable solution exists
  int iterations = 0;                          // successful line-search iterations
  int termination = 0;                         // 0 NO_CONVERGENCE (iteration cap), 1 CONVERGENCE, 2 FAILURE
  int n_eval = 0;
};

struct State {
  double cost = 0, g[3] = {0, 0, 0}, gmax = 0, dir[3] = {0, 0, 0}, dd = 0, step = 0;
};

inline Result minimize(Problem &P, const double x0[3], int max_num_iterations = 10) {
  const double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
  const int max_restarts = 5;
  Result R;
  double x[3] = {x0[0], x0[1], x0[2]};
  for (int k = 0; k < 3; ++k) R.x[k] = x0[k];
  State cur, prev;
  {
    const J3 f = evalJet(P, x);
    cur.cost = f.a;
    for (int k = 0; k < 3; ++k) cur.g[k] = f.v[k];
    cur.gmax = std::max(std::fabs(cur.g[0]), std::max(std::fabs(cur.g[1]), std::fabs(cur.g[2])));
  }
  R.initial_cost = cur.cost;
  double min_cost = cur.cost;
  bool usable = true;
  if (!(std::isfinite(cur.cost))) {
    R.termination = 2;
    usable = false;
  } else if (cur.gmax <= gradient_tolerance) {
    R.termination = 1;
  } else {
    Lbfgs lbfgs;
    lbfgs.reset();
    LsOptions lo;
    int restarts = 0, iteration = 0;
    while (true) {
      if (iteration >= max_num_iterations) {
        R.termination = 0;
        break;
      }
      ++iteration;
      bool ok = true;
      if (iteration == 1) {
        for (int k = 0; k < 3; ++k) cur.dir[k] = -cur.g[k];
      } else {
        const double dx[3] = {prev.dir[0] * prev.step, prev.dir[1] * prev.step, prev.dir[2] * prev.step};
        const double dg[3] = {cur.g[0] - prev.g[0], cur.g[1] - prev.g[1], cur.g[2] - prev.g[2]};
        lbfgs.update(dx, dg);
        double d[3];
        lbfgs.rightMultiply(cur.g, d);
        for (int k = 0; k < 3; ++k) cur.dir[k] = d[k] * -1.0;
        if ((cur.dir[0] * cur.g[0] + cur.dir[1] * cur.g[1]) + cur.dir[2] * cur.g[2] >= 0.0) ok = false;
      }
      if (!ok && restarts >= max_restarts) {
        R.termination = 2;
        usable = false;
        break;
      } else if (!ok) {
        ++restarts;
        lbfgs.reset();
        for (int k = 0; k < 3; ++k) cur.dir[k] = -cur.g[k];
      }
      LsFunction fn;
      fn.P = &P;
      for (int k = 0; k < 3; ++k) {
        fn.pos[k] = x[k];
        fn.dir[k] = cur.dir[k];
      }
      cur.dd = (cur.g[0] * cur.dir[0] + cur.g[1] * cur.dir[1]) + cur.g[2] * cur.dir[2];
      const double initial_step = (iteration == 1 || !ok) ? std::min(1.0, 1.0 / cur.gmax)
                                                           : std::min(1.0, 2.0 * (cur.cost - prev.cost) / cur.dd);
      if (initial_step < 0.0) {
        R.termination = 2;
        usable = false;
        break;
      }
      LsSummary ls;
      wolfeSearch(fn, lo, initial_step, cur.cost, cur.dd, &ls);
      if (!ls.success) {
        R.termination = 2;
        usable = false;
        break;
      }
      const Sample &opt = ls.optimal_point;
      cur.step = opt.x;
      prev = cur;
      if (opt.vector_gradient_is_valid) {
        cur.cost = opt.value;
        for (int k = 0; k < 3; ++k) cur.g[k] = opt.vg[k];
      } else {
        const J3 f = evalJet(P, opt.vx);
        cur.cost = f.a;
        for (int k = 0; k < 3; ++k) cur.g[k] = f.v[k];
      }
      cur.gmax = std::max(std::fabs(cur.g[0]), std::max(std::fabs(cur.g[1]), std::fabs(cur.g[2])));
      const double ex = opt.vx[0] - x[0], ey = opt.vx[1] - x[1], ez = opt.vx[2] - x[2];
      const double step_norm = std::sqrt((ex * ex + ey * ey) + ez * ez);
      const double x_norm = std::sqrt((x[0] * x[0] + x[1] * x[1]) + x[2] * x[2]);
      for (int k = 0; k < 3; ++k) x[k] = opt.vx[k];
      R.iterations = iteration;
      if (cur.cost < min_cost) min_cost = cur.cost;
      if (cur.gmax <= gradient_tolerance) {
        R.termination = 1;
        break;
      }
      const double cost_change = prev.cost - cur.cost;
      if (step_norm <= parameter_tolerance * (x_norm + parameter_tolerance)) {
        R.termination = 1;
        break;
      }
      if (std::fabs(cost_change) <= function_tolerance * std::fabs(prev.cost)) {
        R.termination = 1;
        break;
      }
    }
  }
  if (usable) {  // Summary::IsSolutionUsable(): parameters and final_cost are only written back then
    for (int k = 0; k < 3; ++k) R.x[k] = x[k];
    R.final_cost = min_cost;
  }
  R.n_eval = P.n_eval;
  return R;
}

// ConstellCorrelation::calcCorrelation (correlation.h:206-238): returns (correlation, T_best)
struct CorrResult {
  double correlation;
  Iso2 T;
  Result opt;
};

inline CorrResult calcCorrelation(const GMMScanData &src, const GMMScanData &tgt, const Iso2 &T_init) {
  Problem P = makeProblem(src, tgt, T_init);
  const double p0[3] = {T_init.tx, T_init.ty, std::atan2(T_init.m10, T_init.m00)};
  CorrResult out;
  out.opt = minimize(P, p0, 10);
  out.T = Iso2::fromAngTrans(out.opt.x[2], out.opt.x[0], out.opt.x[1]);
  out.correlation = -out.opt.final_cost / std::sqrt(src.auto_corr * tgt.auto_corr);
  return out;
}

}  // namespace refine
}  // namespace c2o
