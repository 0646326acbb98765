// Quadrature and interpolation tables for the Esde kernels.
//
// The reference integrates every interval with scipy.integrate.quad_vec
// (free_energy.py:405-417): adaptive Gauss-Kronrod 21.  After its mandatory
// first bisection the rule is GK21 on [ti,c] + GK21 on [c,tj] = 42 nodes.
// The node/weight values are the published 21-point Kronrod / 10-point Gauss
// constants (same table as scipy/_quad_vec.py::_quadrature_gk21).
//
// Interpolation: cubic Lagrange for m(t) on knots ti + k*dt/3, quadratic for
// s(t) on knots ti + k*dt/2 (free_energy.py:386-398, symbolics.py:64-90).
// In the local coordinate tau = (t - ti)/dt the basis values at the 42 nodes
// are universal constants; d/dt = (1/dt) d/dtau.
#pragma once

#include <cmath>

namespace mfvgp {

constexpr int kNodesPerPanel = 21;
constexpr int kNodes = 42;          // two panels
constexpr int kTabStride = 16;      // doubles per node in the device table

// table columns
constexpr int T_B3 = 0;    // [0..3]   cubic basis value
constexpr int T_DB3 = 4;   // [4..7]   cubic basis d/dtau
constexpr int T_B2 = 8;    // [8..10]  quadratic basis value
constexpr int T_DB2 = 11;  // [11..13] quadratic basis d/dtau
constexpr int T_V = 14;    // Kronrod weight v_r
constexpr int T_C = 15;    // v_r - w_(r-1)/2 on odd r (Kronrod minus Gauss), else v_r

// positive half of the Kronrod abscissae, descending (r = 0..9), then 0
static const long double kGK21_xh[10] = {
    0.995657163025808080735527280689003L, 0.973906528517171720077964012084452L,
    0.930157491355708226001207180059508L, 0.865063366688984510732096688423493L,
    0.780817726586416897063717578345042L, 0.679409568299024406234327365114874L,
    0.562757134668604683339000099272694L, 0.433395394129247190799265943165784L,
    0.294392862701460198131126603103866L, 0.148874338981631210884826001129720L};
// Kronrod weights for r = 0..10 (symmetric)
static const long double kGK21_vh[11] = {
    0.011694638867371874278064396062192L, 0.032558162307964727478818972459390L,
    0.054755896574351996031381300244580L, 0.075039674810919952767043140916190L,
    0.093125454583697605535065465083366L, 0.109387158802297641899210590325805L,
    0.123491976262065851077958109831074L, 0.134709217311473325928054001771707L,
    0.142775938577060080797094273138717L, 0.147739104901338491374841515972068L,
    0.149445554002916905664936468389821L};
// Gauss-10 weights for i = 0..4 (symmetric); they sit on odd Kronrod nodes
static const long double kGK21_wh[5] = {
    0.066671344308688137593568809893332L, 0.149451349150580593145776339657697L,
    0.219086362515982043995534934228163L, 0.269266719309996355091226921569469L,
    0.295524224714752870173892994651338L};

inline long double gk21_x(int r) {
    return r < 10 ? kGK21_xh[r] : (r == 10 ? 0.0L : -kGK21_xh[20 - r]);
}
inline long double gk21_v(int r) { return r <= 10 ? kGK21_vh[r] : kGK21_vh[20 - r]; }
inline long double gk21_w(int i) { return i < 5 ? kGK21_wh[i] : kGK21_wh[9 - i]; }

// Lagrange basis k of order K-1 on equispaced knots j/(K-1) at tau.
inline void lagrange_unit(int K, int k, long double tau, long double *val,
                          long double *der) {
    long double v = 1.0L, d = 0.0L;
    const long double step = 1.0L / (K - 1);
    for (int l = 0; l < K; ++l)
        if (l != k) v *= (tau - l * step) / ((k - l) * step);
    for (int j = 0; j < K; ++j) {
        if (j == k) continue;
        long double term = 1.0L / ((k - j) * step);
        for (int l = 0; l < K; ++l)
            if (l != k && l != j) term *= (tau - l * step) / ((k - l) * step);
        d += term;
    }
    *val = v;
    *der = d;
}

// Fills one row of the table for a node at local coordinate tau with
// Kronrod weight v and Kronrod-minus-Gauss weight c.
inline void fill_row(double *row, long double tau, long double v, long double c) {
    for (int k = 0; k < 4; ++k) {
        long double b, d;
        lagrange_unit(4, k, tau, &b, &d);
        row[T_B3 + k] = (double)b;
        row[T_DB3 + k] = (double)d;
    }
    for (int k = 0; k < 3; ++k) {
        long double b, d;
        lagrange_unit(3, k, tau, &b, &d);
        row[T_B2 + k] = (double)b;
        row[T_DB2 + k] = (double)d;
    }
    row[T_V] = (double)v;
    row[T_C] = (double)c;
}

// 42-node table of the fixed rule: panel p in {0,1}, node r in 0..20,
// tau = 0.25 + 0.5 p + 0.25 x_r ; integral over the interval of length dt is
// (dt/4) * sum_q v_q f(t_q).
inline void build_fixed_table(double *tab /*[42][16]*/) {
    for (int p = 0; p < 2; ++p)
        for (int r = 0; r < kNodesPerPanel; ++r) {
            long double tau = 0.25L + 0.5L * p + 0.25L * gk21_x(r);
            long double v = gk21_v(r);
            long double c = (r & 1) ? v - gk21_w((r - 1) / 2) : v;
            fill_row(tab + (p * kNodesPerPanel + r) * kTabStride, tau, v, c);
        }
}

// ---------------------------------------------------------------------------
// Split rule (kernels.cuh, esde_l96_split_kernel).  Every integrand of the
// path is  polynomial(t) + terms in q/s and (q/s)^2  with q linear and s
// quadratic in t.  GK21 (degree 31) and its embedded Gauss-10
// (degree 19) integrate the polynomial part (degree <= 12 for L96) exactly, so
// that part may be integrated by ANY exact rule -- here 7-point Gauss-Legendre
// on the whole interval (exact to degree 13) -- while the rational part keeps
// scipy's 42 Kronrod nodes, which is what fixes the result to the reference's.
// ---------------------------------------------------------------------------
constexpr int kGL = 7;

// table of the rational part, per GK node: tau, v, v*B2[3], v*dB2[3], and -- on the
// odd nodes, which carry the embedded Gauss-10 rule -- w, w*B2[3], w*dB2[3]
// [15]: unused.  A second small table carries v*x*B2_0 and v*x*dB2_0 (x = node
// position in [-1,1] within its panel), used for the lower bound of scipy's
// "dabs" (sum_r v_r |f_r - mean| >= |sum_r v_r x_r f_r| because sum v_r x_r = 0).
constexpr int R_TAU = 0, R_V = 1, R_VB = 2, R_VD = 5, R_W = 8, R_WB = 9, R_WD = 12;

inline void build_moment_table(double *vxB /*[42]*/, double *vxD /*[42]*/,
                               double *poly /*[2][2]: per panel sum vxB, sum vxD*/) {
    for (int p = 0; p < 2; ++p) {
        long double s1 = 0.0L, s2 = 0.0L;
        for (int r = 0; r < kNodesPerPanel; ++r) {
            long double x = gk21_x(r), tau = 0.25L + 0.5L * p + 0.25L * x, b, d;
            lagrange_unit(3, 0, tau, &b, &d);
            long double v = gk21_v(r);
            vxB[p * kNodesPerPanel + r] = (double)(v * x * b);
            vxD[p * kNodesPerPanel + r] = (double)(v * x * d);
            s1 += v * x * b;
            s2 += v * x * d;
        }
        poly[2 * p] = (double)s1;
        poly[2 * p + 1] = (double)s2;
    }
}

inline void build_rational_table(double *tab /*[42][16]*/) {
    for (int p = 0; p < 2; ++p)
        for (int r = 0; r < kNodesPerPanel; ++r) {
            long double tau = 0.25L + 0.5L * p + 0.25L * gk21_x(r);
            long double v = gk21_v(r);
            long double w = (r & 1) ? gk21_w((r - 1) / 2) : 0.0L;
            double *row = tab + (p * kNodesPerPanel + r) * kTabStride;
            row[R_TAU] = (double)tau;
            row[R_V] = (double)v;
            row[R_W] = (double)w;
            for (int k = 0; k < 3; ++k) {
                long double b, d;
                lagrange_unit(3, k, tau, &b, &d);
                row[R_VB + k] = (double)(v * b);
                row[R_VD + k] = (double)(v * d);
                row[R_WB + k] = (double)(w * b);
                row[R_WD + k] = (double)(w * d);
            }
            row[15] = 0.0;
        }
}

// ---------------------------------------------------------------------------
// Pair table of the rational part in PANEL-LOCAL coordinates (l96_walk.cuh).
// GK21 nodes are symmetric, x and -x; with s and q re-expanded about the panel
// centre the node constants are the same for both panels, and every weighted
// basis sum becomes a moment sum  sum_r v_r x_r^j f(x_r): even moments see
// f(x)+f(-x), odd ones f(x)-f(-x).  Row i = 0..9 (x descending, i odd = node of
// the embedded Gauss-10 rule): x, v, v x, v x^2, v x^3, w, w x, w x^2.
// ---------------------------------------------------------------------------
constexpr int kPairs = 10;
constexpr int kPairStride = 8;
constexpr int PR_X = 0, PR_V = 1, PR_VX = 2, PR_VX2 = 3, PR_VX3 = 4, PR_W = 5, PR_WX = 6,
              PR_WX2 = 7;

inline void build_pair_table(double *tab /*[10][8]*/, double *v_centre) {
    for (int i = 0; i < kPairs; ++i) {
        const long double x = gk21_x(i), v = gk21_v(i);
        const long double w = (i & 1) ? gk21_w((i - 1) / 2) : 0.0L;
        double *row = tab + i * kPairStride;
        row[PR_X] = (double)x;
        row[PR_V] = (double)v;
        row[PR_VX] = (double)(v * x);
        row[PR_VX2] = (double)(v * x * x);
        row[PR_VX3] = (double)(v * x * x * x);
        row[PR_W] = (double)w;
        row[PR_WX] = (double)(w * x);
        row[PR_WX2] = (double)(w * x * x);
    }
    *v_centre = (double)gk21_v(10);
}

// n-point Gauss-Legendre on [-1,1] by Newton iteration in long double
inline void gauss_legendre(int n, long double *x, long double *w) {
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int i = 0; i < n; ++i) {
        long double z = cosl(pi * (i + 0.75L) / (n + 0.5L));
        long double pp = 0.0L;
        for (int it = 0; it < 100; ++it) {
            long double p1 = 1.0L, p2 = 0.0L;
            for (int j = 1; j <= n; ++j) {
                long double p3 = p2;
                p2 = p1;
                p1 = ((2.0L * j - 1.0L) * z * p2 - (j - 1.0L) * p3) / j;
            }
            pp = n * (z * p1 - p2) / (z * z - 1.0L);
            long double dz = p1 / pp;
            z -= dz;
            if (fabsl(dz) < 1e-19L) break;
        }
        x[i] = z;
        w[i] = 2.0L / ((1.0L - z * z) * pp * pp);
    }
}

// table of the polynomial part: same columns as the fixed table, T_V = weight on [0,1]
inline void build_gl_table(double *tab /*[kGL][16]*/) {
    long double x[kGL], w[kGL];
    gauss_legendre(kGL, x, w);
    for (int g = 0; g < kGL; ++g)
        fill_row(tab + g * kTabStride, 0.5L + 0.5L * x[g], 0.5L * w[g], 0.0L);
}

}  // namespace mfvgp
