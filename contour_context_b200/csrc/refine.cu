// refine.cu — ConstellCorrelation::calcCorrelation on the device (include/cont2/correlation.h:206-238) and the second half
// of CandidateManager::fineOptimize (include/cont2/contour_db.h:604-648).
//
//   refine_kernel  one CTA (4 warps) per (query scan, pre-selected candidate): builds the pair list GMMPair's constructor selects at
//                  T_init (correlation.h:84-96), then minimises GMMPair::operator() (correlation.h:125-152) over
//                  (x, y, theta) with the solver ceres::Solve runs for a GradientProblem with default options and
//                  max_num_iterations = 10: L-BFGS direction + strong-Wolfe line search with cubic interpolation.
//                  The cost value follows the reference's expression operation by operation; its gradient (the reference
//                  gets it from AutoDiffFirstOrderFunction<GMMPair, 3>) is the closed-form derivative.  The lanes split
//                  the pairs, the scalar solver logic runs redundantly (and identically) on every lane.
//   rank_kernel    the final std::sort of the refined candidates (contour_db.h:630-636) with the libstdc++ replay.
//
// Ceres is not part of the reference tree; the solver below follows the published Ceres 2.x line-search minimizer
// (line_search_minimizer.cc, line_search.cc, line_search_direction.cc, low_rank_inverse_hessian.cc, polynomial.cc) with the
// option values GradientProblemSolver::Options defaults to.  Parity tests: tests/test_query_gpu.py.
#include <cuda_runtime.h>
#include <math.h>

#include "c2g_ctx.cuh"
#include "c2g_libm.cuh"
#include "stdsort.cuh"

namespace {

__device__ const uint64_t rf_exp_tab[256] = C2G_EXP_TAB_INIT;

constexpr int RF_MAX_WARPS = 4;
constexpr int RF_TGT_CAP = 128;  // target ellipses of one level staged in shared memory by the pair pre-selection
constexpr int RF_WARPS = 2;  // default warps per candidate (measured best of 1..4: 1.41 / 1.18 / 1.27 / 1.41 ms per 592-query batch) (one CTA each): the pair terms of one evaluation are split over 128 lanes, which
                             // shortens the sequential evaluation chain of the long problems that set the kernel's makespan

// value + gradient of the cost at one point
struct D3 {
  double a, v0, v1, v2;
};
#define RF_FN __device__ __forceinline__

struct Prob {
  const c2g_ell *se, *te;  // ellipse tables of the candidate (src) and the query (tgt) scan, indexed like their views
  const uint32_t *pairs;   // this warp's share of the pre-selected pairs: (src view index << 16) | tgt view index
  double (*red)[4];        // [RF_WARPS][4] shared-memory slots of the block reduction
  int n_pairs, exp_mode, lane, warp, n_warps;
  int *n_eval;             // evaluations so far (work counter, per thread)
};

// GMMPair::operator() (correlation.h:125-152) and its gradient; all threads of the CTA call, all get the same result
// (fixed reduction order: butterfly inside a warp, then warps 0..RF_WARPS-1).
// The VALUE of every pair term follows the reference's expression (2x2 products coefficient by coefficient, inverse =
// adjugate / det, -0.5 mu^T Sigma^-1 mu, K det^-1/2 exp(.)) with det^-1/2 evaluated once.
// The GRADIENT is the closed form of what the reference obtains by automatic differentiation:
//   Sigma = 2 (R A R^T + B), mu = R a + t - b, f = K det(Sigma)^-1/2 exp(-1/2 mu^T Sigma^-1 mu), w = Sigma^-1 mu
//   df/dt     = -f w
//   df/dtheta = f (-1/2 tr(Sigma^-1 Sigma') - mu'^T w + 1/2 w^T Sigma' w),  Sigma' = 2 (J C + C J^T), mu' = J R a, J = [0 -1; 1 0]
// everything of one pair term that does not depend on exp(): the term is coef * exp(qua)
struct PairPre {
  double qua, coef, q0, q1, gfac;
};
RF_FN PairPre rf_pair_pre(const c2g_ell &ea, const c2g_ell &eb, double c, double s, double ns, double x, double y) {
  const double a00 = ea.c00, a10 = ea.c10, a01 = ea.c01, a11 = ea.c11, ax = ea.mx, ay = ea.my;
  const double t00 = c * a00 + ns * a10, t01 = c * a01 + ns * a11;
  const double t10 = s * a00 + c * a10, t11 = s * a01 + c * a11;
  const double m00 = t00 * c + t01 * ns, m01 = t00 * s + t01 * c;
  const double m10 = t10 * c + t11 * ns, m11 = t10 * s + t11 * c;
  const double c00 = 2.0 * (m00 + (double) eb.c00), c10 = 2.0 * (m10 + (double) eb.c10);
  const double c01 = 2.0 * (m01 + (double) eb.c01), c11 = 2.0 * (m11 + (double) eb.c11);
  const double rax = c * ax + ns * ay, ray = s * ax + c * ay;
  const double mux = rax + x - (double) eb.mx, muy = ray + y - (double) eb.my;
  const double det = c00 * c11 - c10 * c01;
  // det^-1/2 once (MUFU.RSQ64H + Newton) instead of the reference's 1/det, sqrt(det) and K/sqrt(det): three long-latency
  // subroutines on the critical path of every pair term; the few-ulp difference is far below the solver's tolerances
  const double rs = rsqrt(det);
  const double invdet = rs * rs;
  const double i00 = c11 * invdet, i10 = -c10 * invdet, i01 = -c01 * invdet, i11 = c00 * invdet;
  const double r0 = -0.5 * mux, r1 = -0.5 * muy;
  PairPre o;
  o.q0 = r0 * i00 + r1 * i10;  // -1/2 mu^T Sigma^-1
  o.q1 = r0 * i01 + r1 * i11;
  o.qua = o.q0 * mux + o.q1 * muy;
  o.coef = (-(double) eb.w * (double) ea.w) * rs;
  const double wx = -2.0 * o.q0, wy = -2.0 * o.q1;
  const double s00 = -2.0 * (m10 + m01), s01 = 2.0 * (m00 - m11);  // Sigma' = [s00 s01; s01 -s00]
  const double tr = (i00 - i11) * s00 + (i01 + i10) * s01;
  const double mpw = rax * wy - ray * wx;
  const double quad = s00 * (wx * wx - wy * wy) + 2.0 * s01 * (wx * wy);
  o.gfac = 0.5 * (quad - tr) - mpw;
  return o;
}

// Two pairs per lane and step: the two terms are independent straight-line code up to their exp() (branch-free main path of the
// glibc restatement, the rare special arguments are redone through the full function), so their FP64 latency chains overlap.
// They are ADDED in the order of the one-pair-per-step loop, so the sums are bit-identical to it.
template <int MODE>
__device__ __forceinline__ void rf_eval_pairs(const Prob &P, double c, double s, double ns, double x, double y, double &fa, double &g0,
                                              double &g1, double &g2) {
  int i = P.lane;
  for (; i + 32 < P.n_pairs; i += 64) {
    const uint32_t pr0 = P.pairs[i], pr1 = P.pairs[i + 32];
    const c2g_ell ea0 = P.se[pr0 >> 16], eb0 = P.te[pr0 & 0xFFFFu], ea1 = P.se[pr1 >> 16], eb1 = P.te[pr1 & 0xFFFFu];
    const PairPre t0 = rf_pair_pre(ea0, eb0, c, s, ns, x, y), t1 = rf_pair_pre(ea1, eb1, c, s, ns, x, y);
    double e0, e1;
    if (MODE != 0) {
      bool special = false;
      e0 = c2g_exp_glibc_main<MODE == 2>(t0.qua, rf_exp_tab, &special);
      e1 = c2g_exp_glibc_main<MODE == 2>(t1.qua, rf_exp_tab, &special);
      if (special) {
        e0 = c2g_exp(t0.qua, MODE, rf_exp_tab);
        e1 = c2g_exp(t1.qua, MODE, rf_exp_tab);
      }
    } else {
      e0 = exp(t0.qua);
      e1 = exp(t1.qua);
    }
    const double f0 = t0.coef * e0, f1 = t1.coef * e1;
    fa += f0;
    g0 += (2.0 * f0) * t0.q0;  // -f w_x, w = -2 q
    g1 += (2.0 * f0) * t0.q1;
    g2 += f0 * t0.gfac;
    fa += f1;
    g0 += (2.0 * f1) * t1.q0;
    g1 += (2.0 * f1) * t1.q1;
    g2 += f1 * t1.gfac;
  }
  for (; i < P.n_pairs; i += 32) {
    const uint32_t pr = P.pairs[i];
    const c2g_ell ea = P.se[pr >> 16], eb = P.te[pr & 0xFFFFu];
    const PairPre t = rf_pair_pre(ea, eb, c, s, ns, x, y);
    const double f = t.coef * c2g_exp(t.qua, MODE, rf_exp_tab);
    fa += f;
    g0 += (2.0 * f) * t.q0;
    g1 += (2.0 * f) * t.q1;
    g2 += f * t.gfac;
  }
}

__device__ __noinline__ D3 rf_eval(const Prob &P, const double p[3]) {
  const double c = cos(p[2]), s = sin(p[2]), ns = -s;
  const double x = p[0], y = p[1];
  ++*P.n_eval;
  double fa = 0.0, g0 = 0.0, g1 = 0.0, g2 = 0.0;
  if (P.exp_mode == 2)
    rf_eval_pairs<2>(P, c, s, ns, x, y, fa, g0, g1, g2);
  else if (P.exp_mode == 1)
    rf_eval_pairs<1>(P, c, s, ns, x, y, fa, g0, g1, g2);
  else
    rf_eval_pairs<0>(P, c, s, ns, x, y, fa, g0, g1, g2);
  for (int o = 16; o > 0; o >>= 1) {
    fa += __shfl_xor_sync(0xFFFFFFFFu, fa, o);
    g0 += __shfl_xor_sync(0xFFFFFFFFu, g0, o);
    g1 += __shfl_xor_sync(0xFFFFFFFFu, g1, o);
    g2 += __shfl_xor_sync(0xFFFFFFFFu, g2, o);
  }
  if (P.lane == 0) {
    P.red[P.warp][0] = fa;
    P.red[P.warp][1] = g0;
    P.red[P.warp][2] = g1;
    P.red[P.warp][3] = g2;
  }
  __syncthreads();
  fa = P.red[0][0];
  g0 = P.red[0][1];
  g1 = P.red[0][2];
  g2 = P.red[0][3];
  for (int w = 1; w < P.n_warps; ++w) {
    fa += P.red[w][0];
    g0 += P.red[w][1];
    g1 += P.red[w][2];
    g2 += P.red[w][3];
  }
  __syncthreads();  // the slots are rewritten by the next evaluation
  return D3{fa, g0, g1, g2};
}

// ---- polynomial interpolation (polynomial.cc) -----------------------------------------------------------------------------
struct Smp {  // FunctionSample
  double x, value, gradient;
  double vx[3], vg[3];
  bool value_ok, grad_ok;
};

RF_FN Smp smp_empty() {
  Smp s;
  s.x = s.value = s.gradient = 0.0;
  s.vx[0] = s.vx[1] = s.vx[2] = 0.0;
  s.vg[0] = s.vg[1] = s.vg[2] = 0.0;
  s.value_ok = s.grad_ok = false;
  return s;
}

RF_FN double poly_eval(const double *poly, int n, double x) {
  double v = 0.0;
  for (int i = 0; i < n; ++i) v = v * x + poly[i];
  return v;
}

RF_FN double ipow(double x, int e) {
  double r = 1.0;
  for (int i = 0; i < e; ++i) r *= x;
  return r;
}

// n x n (n <= 4) solve through LU with complete pivoting, rank cut at eps * n * max pivot (Eigen FullPivLU::solve).
// N is a template parameter and every loop is unrolled, so the matrix lives in REGISTERS: the data-dependent pivot position only
// selects which (compile-time) rows / columns are exchanged.  Same operations in the same order as the array version of round 1
// (which kept A in local memory and cost 18 % of the refinement kernel's stall samples).
template <int N>
RF_FN void lu_solve_n(const double *Ain, const double *bin, double *x) {
  double A[N][N], b[N];
#pragma unroll
  for (int r = 0; r < N; ++r) {
    b[r] = bin[r];
#pragma unroll
    for (int cc = 0; cc < N; ++cc) A[r][cc] = Ain[r * 4 + cc];
  }
  int rowT[N], colT[N];
  int nonzero = N;
  double maxpivot = 0.0;
  bool stopped = false;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    if (!stopped) {
      int br = k, bc = k;
      double best = -1.0;
#pragma unroll
      for (int cc = k; cc < N; ++cc)
#pragma unroll
        for (int r = k; r < N; ++r)
          if (fabs(A[r][cc]) > best) {
            best = fabs(A[r][cc]);
            br = r;
            bc = cc;
          }
      if (best == 0.0) {
        nonzero = k;
#pragma unroll
        for (int i = k; i < N; ++i) rowT[i] = colT[i] = i;
        stopped = true;
      } else {
        if (best > maxpivot) maxpivot = best;
        rowT[k] = br;
        colT[k] = bc;
#pragma unroll
        for (int r2 = k + 1; r2 < N; ++r2)
          if (r2 == br) {
#pragma unroll
            for (int cc = 0; cc < N; ++cc) {
              const double t = A[k][cc];
              A[k][cc] = A[r2][cc];
              A[r2][cc] = t;
            }
          }
#pragma unroll
        for (int c2 = k + 1; c2 < N; ++c2)
          if (c2 == bc) {
#pragma unroll
            for (int r = 0; r < N; ++r) {
              const double t = A[r][k];
              A[r][k] = A[r][c2];
              A[r][c2] = t;
            }
          }
#pragma unroll
        for (int r = k + 1; r < N; ++r) A[r][k] /= A[k][k];
#pragma unroll
        for (int cc = k + 1; cc < N; ++cc)
#pragma unroll
          for (int r = k + 1; r < N; ++r) A[r][cc] -= A[r][k] * A[k][cc];
      }
    }
  }
  const double thr = 2.220446049250313e-16 * N;
  int rank = 0;
#pragma unroll
  for (int i = 0; i < N; ++i)
    if (i < nonzero) rank += (fabs(A[i][i]) > thr * maxpivot);
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = 0.0;
  if (rank == 0) return;
#pragma unroll
  for (int k = 0; k < N; ++k) {  // b = P b: exchange b[k] and b[rowT[k]] (rowT[k] >= k)
#pragma unroll
    for (int r2 = k + 1; r2 < N; ++r2)
      if (rowT[k] == r2) {
        const double t = b[k];
        b[k] = b[r2];
        b[r2] = t;
      }
  }
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) b[i] -= A[i][j] * b[j];
#pragma unroll
  for (int i = N - 1; i >= 0; --i)
    if (i < rank) {
#pragma unroll
      for (int j = i + 1; j < N; ++j)
        if (j < rank) b[i] -= A[i][j] * b[j];
      b[i] /= A[i][i];
    }
#pragma unroll
  for (int i = 0; i < N; ++i)
    if (i >= rank) b[i] = 0.0;
  int perm[N];
#pragma unroll
  for (int i = 0; i < N; ++i) perm[i] = i;
#pragma unroll
  for (int k = 0; k < N; ++k) {  // perm[k] <-> perm[colT[k]] (colT[k] >= k)
#pragma unroll
    for (int c2 = k + 1; c2 < N; ++c2)
      if (colT[k] == c2) {
        const int t = perm[k];
        perm[k] = perm[c2];
        perm[c2] = t;
      }
  }
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (perm[i] == j) x[j] = b[i];
}

__device__ __noinline__ void lu_solve(double *A, double *b, int n, double *x) {
  if (n == 4)
    lu_solve_n<4>(A, b, x);
  else if (n == 3)
    lu_solve_n<3>(A, b, x);
  else if (n == 2)
    lu_solve_n<2>(A, b, x);
  else if (n == 1)
    lu_solve_n<1>(A, b, x);
}

// LineSearch::InterpolatingPolynomialMinimizingStepSize, CUBIC, two samples
__device__ __noinline__ double interp_step(const Smp &lower, const Smp &current, double min_step, double max_step) {
  if (!current.value_ok) return fmin(fmax(current.x * 0.5, min_step), max_step);
  const Smp *ss[2] = {&lower, &current};
  int n = 0;
  for (int k = 0; k < 2; ++k) n += (ss[k]->value_ok ? 1 : 0) + (ss[k]->grad_ok ? 1 : 0);
  const int degree = n - 1;
  double A[16], rhs[4], poly[4];
  for (int i = 0; i < 16; ++i) A[i] = 0.0;
  int row = 0;
  for (int k = 0; k < 2; ++k) {
    const Smp &s = *ss[k];
    if (s.value_ok) {
      for (int j = 0; j <= degree; ++j) A[row * 4 + j] = ipow(s.x, degree - j);
      rhs[row] = s.value;
      ++row;
    }
    if (s.grad_ok) {
      for (int j = 0; j < degree; ++j) A[row * 4 + j] = (double) (degree - j) * ipow(s.x, degree - j - 1);
      rhs[row] = s.gradient;
      ++row;
    }
  }
  lu_solve(A, rhs, n, poly);
  // MinimizePolynomial over [min_step, max_step]
  double opt_x = (min_step + max_step) / 2.0;
  double opt_v = poly_eval(poly, n, opt_x);
  const double vmin = poly_eval(poly, n, min_step);
  if (vmin < opt_v) {
    opt_v = vmin;
    opt_x = min_step;
  }
  const double vmax = poly_eval(poly, n, max_step);
  if (vmax < opt_v) {
    opt_v = vmax;
    opt_x = max_step;
  }
  if (n > 2) {
    double d[3];
    int nd = degree;
    for (int i = 0; i < degree; ++i) d[i] = (double) (degree - i) * poly[i];
    int lead = 0;
    while (lead + 1 < nd && d[lead] == 0.0) ++lead;
    const int dd = nd - lead - 1;  // degree of the derivative after stripping leading zeros
    double roots[2];
    int nr = 0;
    if (dd == 1) {
      roots[nr++] = -d[lead + 1] / d[lead];
    } else if (dd == 2) {
      const double a = d[lead], b = d[lead + 1], cq = d[lead + 2];
      const double D = b * b - 4 * a * cq;
      const double sq = sqrt(fabs(D));
      if (D >= 0) {
        if (b >= 0) {
          roots[nr++] = (-b - sq) / (2.0 * a);
          roots[nr++] = (2.0 * cq) / (-b - sq);
        } else {
          roots[nr++] = (2.0 * cq) / (-b + sq);
          roots[nr++] = (-b + sq) / (2.0 * a);
        }
      } else {
        roots[nr++] = -b / (2.0 * a);
        roots[nr++] = -b / (2.0 * a);
      }
    }
    for (int i = 0; i < nr; ++i) {
      const double r = roots[i];
      if ((r < min_step) || (r > max_step)) continue;
      const double v = poly_eval(poly, n, r);
      if (v < opt_v) {
        opt_v = v;
        opt_x = r;
      }
    }
  }
  for (int k = 0; k < 2; ++k) {
    const double sx = ss[k]->x;
    if ((sx < min_step) || (sx > max_step)) continue;
    const double v = poly_eval(poly, n, sx);
    if (v < opt_v) {
      opt_x = sx;
      opt_v = v;
    }
  }
  return opt_x;
}

// ---- strong-Wolfe line search (line_search.cc) ------------------------------------------------------------------------------
struct LsFn {
  double pos[3], dir[3];
};
RF_FN double dir_inf(const LsFn &f) { return fmax(fabs(f.dir[0]), fmax(fabs(f.dir[1]), fabs(f.dir[2]))); }

__device__ __noinline__ void ls_eval(const Prob &P, const LsFn &fn, double x, Smp *out) {
  *out = smp_empty();
  out->x = x;
  for (int k = 0; k < 3; ++k) out->vx[k] = fn.pos[k] + x * fn.dir[k];
  const D3 f = rf_eval(P, out->vx);
  out->value = f.a;
  out->vg[0] = f.v0;
  out->vg[1] = f.v1;
  out->vg[2] = f.v2;
  if (!isfinite(out->value)) return;
  out->value_ok = true;
  out->gradient = (fn.dir[0] * out->vg[0] + fn.dir[1] * out->vg[1]) + fn.dir[2] * out->vg[2];
  if (!isfinite(out->gradient) || !isfinite(out->vg[0]) || !isfinite(out->vg[1]) || !isfinite(out->vg[2])) return;
  out->grad_ok = true;
}

constexpr double LS_DECREASE = 1e-4, LS_CURVATURE = 0.9, LS_EXPANSION = 10.0, LS_MIN_STEP = 1e-9;
constexpr int LS_MAX_ITER = 20;

// returns false if the search failed; *opt is the accepted step otherwise
__device__ __noinline__ bool wolfe_search(const Prob &P, const LsFn &fn, double step_estimate, double cost0, double grad0, Smp *opt) {
  Smp initial = smp_empty();
  initial.value = cost0;
  initial.gradient = grad0;
  initial.value_ok = initial.grad_ok = true;
  for (int k = 0; k < 3; ++k) initial.vx[k] = fn.pos[k];
  const double dmax = dir_inf(fn);
  int iters = 0;
  bool do_zoom = false;
  Smp low = initial, high = initial;
  {  // bracketing phase
    Smp previous = initial, current;
    ls_eval(P, fn, step_estimate, &current);
    while (true) {
      ++iters;
      if (current.value_ok &&
          (current.value > (initial.value + LS_DECREASE * initial.gradient * current.x) || (previous.value_ok && current.value > previous.value))) {
        do_zoom = true;
        low = previous;
        high = current;
        break;
      }
      if (current.value_ok && fabs(current.gradient) <= -LS_CURVATURE * initial.gradient) {
        low = current;
        high = current;
        break;
      } else if (current.value_ok && current.gradient >= 0) {
        do_zoom = true;
        low = current;
        high = previous;
        break;
      } else if (iters >= LS_MAX_ITER) {
        if (current.value_ok && current.value < low.value) low = current;
        break;
      }
      const double min_step = current.value_ok ? current.x : previous.x;
      const double max_step = current.value_ok ? (current.x * LS_EXPANSION) : current.x;
      const double step = interp_step(previous, current, min_step, max_step);
      if (step * dmax < LS_MIN_STEP) return false;
      if (current.value_ok) previous = current;
      ls_eval(P, fn, step, &current);
    }
    if (do_zoom && fabs(high.x - low.x) * dmax < LS_MIN_STEP) do_zoom = false;
  }
  if (!do_zoom) {
    *opt = low;
    return true;
  }
  // zoom phase
  const Smp entry_low = low;
  Smp solution = smp_empty();
  bool zoom_ok = true;
  if (low.gradient * (high.x - low.x) >= 0) {
    solution.value_ok = false;
    zoom_ok = false;
  } else {
    while (true) {
      solution = low;
      if (iters >= LS_MAX_ITER) {
        zoom_ok = false;
        break;
      }
      if (fabs(high.x - low.x) * dmax < LS_MIN_STEP) {
        zoom_ok = false;
        break;
      }
      ++iters;
      const bool low_first = low.x < high.x;
      const double step = low_first ? interp_step(low, high, low.x, high.x) : interp_step(high, low, high.x, low.x);
      ls_eval(P, fn, step, &solution);
      if (!solution.value_ok || !solution.grad_ok) {
        zoom_ok = false;
        break;
      }
      if ((solution.value > (initial.value + LS_DECREASE * initial.gradient * solution.x)) || (solution.value >= low.value)) {
        high = solution;
        continue;
      }
      if (fabs(solution.gradient) <= -LS_CURVATURE * initial.gradient) {
        break;
      } else if (solution.gradient * (high.x - low.x) >= 0) {
        high = low;
      }
      low = solution;
    }
  }
  if (!zoom_ok && !solution.value_ok) return false;
  if (!solution.value_ok || solution.value > entry_low.value)
    *opt = entry_low;
  else
    *opt = solution;
  return true;
}

// ---- L-BFGS (low_rank_inverse_hessian.cc) + LineSearchMinimizer::Minimize --------------------------------------------------
constexpr int RF_MAX_ITER = 10;  // options.max_num_iterations (correlation.h:215); the rank-20 history never wraps

struct RfOut {
  double x[3], final_cost;
  int iterations, termination;
};

__device__ RfOut rf_minimize(const Prob &P, const double x0[3]) {
  const double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
  RfOut R;
  for (int k = 0; k < 3; ++k) R.x[k] = x0[k];
  R.final_cost = -1.0;
  R.iterations = 0;
  R.termination = 0;
  double x[3] = {x0[0], x0[1], x0[2]};
  double cost, g[3], gmax, dir[3], step = 0.0;
  double p_cost = 0.0, p_g[3] = {0, 0, 0}, p_dir[3] = {0, 0, 0}, p_step = 0.0;
  {
    const D3 f = rf_eval(P, x);
    cost = f.a;
    g[0] = f.v0;
    g[1] = f.v1;
    g[2] = f.v2;
    gmax = fmax(fabs(g[0]), fmax(fabs(g[1]), fabs(g[2])));
  }
  double min_cost = cost;
  bool usable = true;
  if (!isfinite(cost)) {
    R.termination = 2;
    usable = false;
  } else if (gmax <= gradient_tolerance) {
    R.termination = 1;
  } else {
    double hs[RF_MAX_ITER][3], hy[RF_MAX_ITER][3], hsy[RF_MAX_ITER];
    int nh = 0, restarts = 0, iteration = 0;
    while (true) {
      if (iteration >= RF_MAX_ITER) {
        R.termination = 0;
        break;
      }
      ++iteration;
      bool ok = true;
      if (iteration == 1) {
        for (int k = 0; k < 3; ++k) dir[k] = -g[k];
      } else {
        const double dx[3] = {p_dir[0] * p_step, p_dir[1] * p_step, p_dir[2] * p_step};
        const double dg[3] = {g[0] - p_g[0], g[1] - p_g[1], g[2] - p_g[2]};
        const double sy = (dx[0] * dg[0] + dx[1] * dg[1]) + dx[2] * dg[2];
        if (!(sy <= 1e-14) && nh < RF_MAX_ITER) {
          for (int k = 0; k < 3; ++k) {
            hs[nh][k] = dx[k];
            hy[nh][k] = dg[k];
          }
          hsy[nh] = sy;
          ++nh;
        }
        double alpha[RF_MAX_ITER], d[3] = {g[0], g[1], g[2]};
        for (int i = nh - 1; i >= 0; --i) {
          const double a = ((hs[i][0] * d[0] + hs[i][1] * d[1]) + hs[i][2] * d[2]) / hsy[i];
          for (int k = 0; k < 3; ++k) d[k] -= a * hy[i][k];
          alpha[i] = a;
        }
        for (int i = 0; i < nh; ++i) {
          const double beta = ((hy[i][0] * d[0] + hy[i][1] * d[1]) + hy[i][2] * d[2]) / hsy[i];
          for (int k = 0; k < 3; ++k) d[k] += hs[i][k] * (alpha[i] - beta);
        }
        for (int k = 0; k < 3; ++k) dir[k] = d[k] * -1.0;
        if ((dir[0] * g[0] + dir[1] * g[1]) + dir[2] * g[2] >= 0.0) ok = false;
      }
      if (!ok && restarts >= 5) {
        R.termination = 2;
        usable = false;
        break;
      } else if (!ok) {
        ++restarts;
        nh = 0;
        for (int k = 0; k < 3; ++k) dir[k] = -g[k];
      }
      LsFn fn;
      for (int k = 0; k < 3; ++k) {
        fn.pos[k] = x[k];
        fn.dir[k] = dir[k];
      }
      const double ddv = (g[0] * dir[0] + g[1] * dir[1]) + g[2] * dir[2];
      const double initial_step = (iteration == 1 || !ok) ? fmin(1.0, 1.0 / gmax) : fmin(1.0, 2.0 * (cost - p_cost) / ddv);
      if (initial_step < 0.0) {
        R.termination = 2;
        usable = false;
        break;
      }
      Smp opt;
      if (!wolfe_search(P, fn, initial_step, cost, ddv, &opt)) {
        R.termination = 2;
        usable = false;
        break;
      }
      step = opt.x;
      p_cost = cost;
      p_step = step;
      for (int k = 0; k < 3; ++k) {
        p_g[k] = g[k];
        p_dir[k] = dir[k];
      }
      if (opt.grad_ok) {
        cost = opt.value;
        for (int k = 0; k < 3; ++k) g[k] = opt.vg[k];
      } else {  // the zero step (initial position): gradient vector not carried by the sample
        const D3 f = rf_eval(P, opt.vx);
        cost = f.a;
        g[0] = f.v0;
        g[1] = f.v1;
        g[2] = f.v2;
      }
      gmax = fmax(fabs(g[0]), fmax(fabs(g[1]), fabs(g[2])));
      const double ex = opt.vx[0] - x[0], ey = opt.vx[1] - x[1], ez = opt.vx[2] - x[2];
      const double step_norm = sqrt((ex * ex + ey * ey) + ez * ez);
      const double x_norm = sqrt((x[0] * x[0] + x[1] * x[1]) + x[2] * x[2]);
      for (int k = 0; k < 3; ++k) x[k] = opt.vx[k];
      R.iterations = iteration;
      if (cost < min_cost) min_cost = cost;
      if (gmax <= gradient_tolerance) {
        R.termination = 1;
        break;
      }
      if (step_norm <= parameter_tolerance * (x_norm + parameter_tolerance)) {
        R.termination = 1;
        break;
      }
      if (fabs(p_cost - cost) <= function_tolerance * fabs(p_cost)) {
        R.termination = 1;
        break;
      }
    }
  }
  if (usable) {
    for (int k = 0; k < 3; ++k) R.x[k] = x[k];
    R.final_cost = min_cost;
  }
  return R;
}

__global__ void __launch_bounds__(RF_MAX_WARPS * 32, 6)
refine_kernel(const c2g_scan_head *__restrict__ heads, const c2g_ell *__restrict__ ells, int first_slot, int q0, int B, int max_fine_opt,
              int exp_mode, uint32_t *__restrict__ pair_scratch, int pair_cap, c2g_query_result *__restrict__ results,
              unsigned long long *__restrict__ work) {
  __shared__ double red[RF_MAX_WARPS][4];
  __shared__ float4 tgt_s[RF_TGT_CAP];
  const int n_warps = blockDim.x >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wg = q0 * max_fine_opt + blockIdx.x;  // one (query, candidate rank) per CTA; the exits below are CTA-uniform
  const int q = wg / max_fine_opt, ci = wg % max_fine_opt;
  if (q >= q0 + B) return;
  c2g_query_result &R = results[q];
  const int pre = min(max_fine_opt, R.n_cand);
  if (ci >= pre) return;
  c2g_cand &C = R.cand[ci];
  const int src = C.cand_gidx, tgt = first_slot + q;
  const double T[4] = {C.T[0], C.T[1], C.T[2], C.T[3]};
  const int warp_cap = pair_cap / n_warps;
  uint32_t *pairs = pair_scratch + (size_t) wg * pair_cap + (size_t) warp * warp_cap;
  // (staging both ellipse tables in shared memory was measured 30 % slower: 32-byte records at random indices conflict on the
  // banks, while the 32-byte sectors are served well by L1)
  const c2g_ell *se = ells + (size_t) src * C2G_VIEW_CAP, *te = ells + (size_t) tgt * C2G_VIEW_CAP;
  // pre-selection at T_init (correlation.h:84-96): |T_init * mu_s - mu_t| < 3 (sqrt(eig_s) + sqrt(eig_t)).  Warp w takes the
  // source ellipses w, w + RF_WARPS, ... of every level and keeps its own pair list; lanes hold 32 target ellipses of the
  // level in registers while the sources stream by.
  int n_pairs = 0, overflow = 0, n_eval = 0;
  long long n_tests = 0;
  for (int li = 0; li < C2G_NUM_BIN_LAYERS; ++li) {
    const int lev = li + 1;
    const int ns = heads[src].n_ell[li], nt = heads[tgt].n_ell[li];
    n_tests += (long long) nt * ((ns - warp + n_warps - 1) / n_warps);
    const int so = heads[src].view_off[lev], to = heads[tgt].view_off[lev];
    if (nt <= 32) {
      double bx = 0.0, by = 0.0;
      float bmaj = 0.f, bxf = 0.f, byf = 0.f;
      if (lane < nt) {
        const c2g_ell b = te[to + lane];
        bx = (double) b.mx;
        by = (double) b.my;
        bxf = b.mx;
        byf = b.my;
        bmaj = b.maj;
      }
      for (int si = warp; si < ns; si += n_warps) {
        const c2g_ell a = se[so + si];
        const double ax = (double) a.mx, ay = (double) a.my;
        const double qx = (T[0] * ax + (-T[1]) * ay) + T[2], qy = (T[1] * ax + T[0] * ay) + T[3];
        bool sel = false;
        if (lane < nt) {
          const float fx = (float) qx - bxf, fy = (float) qy - byf;
          const float d2 = fx * fx + fy * fy, yy = 3.0f * (a.maj + bmaj), y2 = yy * yy;
          if (d2 < y2 * 0.9999f)
            sel = true;
          else if (!(d2 > y2 * 1.0001f)) {  // borderline (or NaN): the exact double test
            const double ddx = qx - bx, ddy = qy - by;
            sel = c2g_sqrt_lt(ddx * ddx + ddy * ddy, 3.0 * (double) (a.maj + bmaj));
          }
        }
        const unsigned m = __ballot_sync(0xFFFFFFFFu, sel);
        if (sel) {
          const int pos = n_pairs + __popc(m & ((1u << lane) - 1u));
          if (pos < warp_cap) pairs[pos] = ((uint32_t) (so + si) << 16) | (uint32_t) (to + lane);
        }
        n_pairs += __popc(m);
      }
    } else if (nt <= RF_TGT_CAP) {
      // the level's target ellipses (mean, sigma) staged once per CTA as floats; the float test with a 1e-4 guard band decides
      // all but borderline pairs, those take the exact double test.  Pair order as before: own sources ascending, targets ascending.
      __syncthreads();  // the previous level's readers are done
      for (int i = threadIdx.x; i < nt; i += blockDim.x) {
        const c2g_ell b = te[to + i];
        tgt_s[i] = make_float4(b.mx, b.my, b.maj, 0.f);
      }
      __syncthreads();
      c2g_ell a_next;
      if (warp < ns) a_next = se[so + warp];
      for (int si = warp; si < ns; si += n_warps) {
        const c2g_ell a = a_next;
        if (si + n_warps < ns) a_next = se[so + si + n_warps];  // the next source's record travels while this one is tested
        const double ax = (double) a.mx, ay = (double) a.my;
        const double qx = (T[0] * ax + (-T[1]) * ay) + T[2], qy = (T[1] * ax + T[0] * ay) + T[3];
        const float qxf = (float) qx, qyf = (float) qy;
        for (int t0 = 0; t0 < nt; t0 += 32) {
          const int ti = t0 + lane;
          bool sel = false;
          if (ti < nt) {
            const float4 b = tgt_s[ti];
            const float fx = qxf - b.x, fy = qyf - b.y;
            const float d2 = fx * fx + fy * fy, yy = 3.0f * (a.maj + b.z), y2 = yy * yy;
            if (d2 < y2 * 0.9999f)
              sel = true;
            else if (!(d2 > y2 * 1.0001f)) {  // borderline (or NaN): the exact double test
              const double ddx = qx - (double) b.x, ddy = qy - (double) b.y;
              sel = c2g_sqrt_lt(ddx * ddx + ddy * ddy, 3.0 * (double) (a.maj + b.z));
            }
          }
          const unsigned m = __ballot_sync(0xFFFFFFFFu, sel);
          if (sel) {
            const int pos = n_pairs + __popc(m & ((1u << lane) - 1u));
            if (pos < warp_cap) pairs[pos] = ((uint32_t) (so + si) << 16) | (uint32_t) (to + ti);
          }
          n_pairs += __popc(m);
        }
      }
    } else {
      for (int si = warp; si < ns; si += n_warps) {
        const c2g_ell a = se[so + si];
        const double ax = (double) a.mx, ay = (double) a.my;
        const double qx = (T[0] * ax + (-T[1]) * ay) + T[2], qy = (T[1] * ax + T[0] * ay) + T[3];
        for (int t0 = 0; t0 < nt; t0 += 32) {
          const int ti = t0 + lane;
          bool sel = false;
          if (ti < nt) {
            const c2g_ell b = te[to + ti];
            const double ddx = qx - (double) b.mx, ddy = qy - (double) b.my;
            sel = c2g_sqrt_lt(ddx * ddx + ddy * ddy, 3.0 * (double) (a.maj + b.maj));
          }
          const unsigned m = __ballot_sync(0xFFFFFFFFu, sel);
          if (sel) {
            const int pos = n_pairs + __popc(m & ((1u << lane) - 1u));
            if (pos < warp_cap) pairs[pos] = ((uint32_t) (so + si) << 16) | (uint32_t) (to + ti);
          }
          n_pairs += __popc(m);
        }
      }
    }
  }
  if (n_pairs > warp_cap) {
    overflow = 1;
    n_pairs = warp_cap;
  }
  overflow = __syncthreads_or(overflow);  // also orders the pair lists before the first evaluation
  Prob P;
  P.red = red;
  P.warp = warp;
  P.n_warps = n_warps;
  P.se = se;
  P.te = te;
  P.pairs = pairs;
  P.n_pairs = n_pairs;
  P.exp_mode = exp_mode;
  P.lane = lane;
  P.n_eval = &n_eval;
  const double p0[3] = {T[2], T[3], atan2(T[1], T[0])};
  const RfOut o = rf_minimize(P, p0);
  if (work && lane == 0) {
    atomicAdd(work + 4, (unsigned long long) n_tests);
    atomicAdd(work + 5, (unsigned long long) n_pairs * (unsigned long long) n_eval);
    if (warp == 0) atomicAdd(work + 6, (unsigned long long) n_eval);
  }
  if (threadIdx.x == 0) {
    const double corr = -o.final_cost / sqrt(heads[src].gmm_auto_corr * heads[tgt].gmm_auto_corr);
    C.corr_fine = (float) corr;  // CandidateAnchorProp::correlation_ is a float (contour_db.h:270)
    C.fine_iters = (int16_t) o.iterations;
    C.fine_term = (int8_t) o.termination;
    C.fine_flags = (int8_t) overflow;
    C.T_fine[0] = cos(o.x[2]);
    C.T_fine[1] = sin(o.x[2]);
    C.T_fine[2] = o.x[0];
    C.T_fine[3] = o.x[1];
  }
}

// second std::sort of fineOptimize (contour_db.h:630-636) over the first min(max_fine_opt, n_cand) candidates
__global__ void rank_kernel(int q0, int B, int max_fine_opt, c2g_query_result *__restrict__ results) {
  const int q = q0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= q0 + B) return;
  c2g_query_result &R = results[q];
  const int pre = min(max_fine_opt, R.n_cand);
  if (pre <= 1) return;
  uint32_t ord[C2G_MAX_CAND];
  float corr[C2G_MAX_CAND];
  for (int i = 0; i < pre; ++i) {
    ord[i] = (uint32_t) i;
    corr[i] = R.cand[i].corr_fine;
  }
  const float *cp = corr;
  c2g_sort::std_sort(ord, (long) pre, [cp](uint32_t a, uint32_t b) { return cp[a] > cp[b]; });
  // apply the permutation in place, cycle by cycle
  for (int i = 0; i < pre; ++i) {
    if ((int) ord[i] == i || ord[i] == 0xFFFFFFFFu) continue;
    const c2g_cand first = R.cand[i];
    int j = i;
    while (true) {
      const int from = (int) ord[j];
      ord[j] = 0xFFFFFFFFu;
      if (from == i) {
        R.cand[j] = first;
        break;
      }
      R.cand[j] = R.cand[from];
      j = from;
    }
  }
}

}  // namespace

int c2g_refine_alloc(c2g_ctx *ctx) {
  ctx->pair_cap = 8192;  // per candidate, split evenly over the warps of its CTA
  const size_t n = (size_t) ctx->max_batch * (size_t) (ctx->db.max_fine_opt > 0 ? ctx->db.max_fine_opt : 1) * (size_t) ctx->pair_cap;
  C2G_CUDA_TRY(cudaMalloc((void **) &ctx->d_pair_scratch, sizeof(uint32_t) * n));
  return 0;
}

void c2g_refine_free(c2g_ctx *ctx) { cudaFree(ctx->d_pair_scratch); }

int c2g_launch_refine(c2g_ctx *ctx, int first_slot, int q0, int B, cudaStream_t st) {
  const int mfo = ctx->db.max_fine_opt;
  if (mfo <= 0) return 0;
  static const int rf_warps = getenv("C2G_REFINE_WARPS") ? max(1, min(RF_MAX_WARPS, atoi(getenv("C2G_REFINE_WARPS")))) : RF_WARPS;
  refine_kernel<<<B * mfo, rf_warps * 32, 0, st>>>(ctx->d_heads, ctx->d_ells, first_slot, q0, B, mfo, ctx->P.exp_mode, ctx->d_pair_scratch, ctx->pair_cap,
                                                   ctx->d_results, ctx->count_work ? ctx->d_work : nullptr);
  C2G_CUDA_TRY(cudaGetLastError());
  if (ctx->prof_on) cudaEventRecord(ctx->prof_ev[7], st);
  rank_kernel<<<(B + 127) / 128, 128, 0, st>>>(q0, B, mfo, ctx->d_results);
  C2G_CUDA_TRY(cudaGetLastError());
  if (ctx->prof_on) cudaEventRecord(ctx->prof_ev[8], st);
  ctx->launches += 2;
  return 0;
}
