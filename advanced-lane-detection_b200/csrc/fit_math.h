// fit_math.h -- the fp64 / double-double arithmetic of the lane fit, shared by the CUDA kernels
// (fit_kernels.cuh, search_kernels.cuh) and by the CPU-only fuzz harness under tests/hostsim (which exists so this
// arithmetic can be checked against numpy in a container without a GPU; the package never loads it).
//
// What it reproduces: np.polyfit(y, x, 2) as called at Project.py:118 and :133
//   lhs = vander(y, 3); scale = sqrt(sum(lhs*lhs, axis=0)); lhs /= scale
//   c = lstsq(lhs, x, rcond)        (LAPACK dgelsd: SVD, singular values <= rcond*s_max dropped,
//                                    minimum-norm solution in the kept subspace)
//   c /= scale
// from the exact power sums of the selected pixels instead of the pixel lists:
//   S_k = sum y^k (k=0..4),  T_k = sum x*y^k (k=0..2).
// The 3x3 scaled Gram matrix is diagonalised by cyclic Jacobi in double-double, so the result
// is the exactly-rounded truncated-SVD solution to ~1e-25 -- the only difference left against
// numpy is numpy's own rounding error (cond * 1e-16).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define LD_HD __host__ __device__ __forceinline__
#define LD_MUL(a, b) __dmul_rn((a), (b))
#define LD_ADD(a, b) __dadd_rn((a), (b))
#define LD_SUB(a, b) __dsub_rn((a), (b))
#define LD_FMA(a, b, c) __fma_rn((a), (b), (c))
#define LD_DIV(a, b) __ddiv_rn((a), (b))
#define LD_SQRT(a) __dsqrt_rn((a))
#elif defined(__CUDACC__)
#define LD_HD __host__ __device__ __forceinline__
#define LD_MUL(a, b) ((a) * (b))
#define LD_ADD(a, b) ((a) + (b))
#define LD_SUB(a, b) ((a) - (b))
#define LD_FMA(a, b, c) fma((a), (b), (c))
#define LD_DIV(a, b) ((a) / (b))
#define LD_SQRT(a) sqrt((a))
#else   // plain g++ (tests/hostsim), built with -ffp-contract=off
#define LD_HD static inline
#define LD_MUL(a, b) ((a) * (b))
#define LD_ADD(a, b) ((a) + (b))
#define LD_SUB(a, b) ((a) - (b))
#define LD_FMA(a, b, c) fma((a), (b), (c))
#define LD_DIV(a, b) ((a) / (b))
#define LD_SQRT(a) sqrt((a))
#endif

namespace lanefit {

struct dd { double hi, lo; };

LD_HD dd dd_make(double h, double l = 0.0) { dd r; r.hi = h; r.lo = l; return r; }
LD_HD dd two_sum(double a, double b) {
    double s = LD_ADD(a, b), bb = LD_SUB(s, a);
    return dd_make(s, LD_ADD(LD_SUB(a, LD_SUB(s, bb)), LD_SUB(b, bb)));
}
LD_HD dd quick_two_sum(double a, double b) {
    double s = LD_ADD(a, b);
    return dd_make(s, LD_SUB(b, LD_SUB(s, a)));
}
LD_HD dd two_prod(double a, double b) {
    double p = LD_MUL(a, b);
    return dd_make(p, LD_FMA(a, b, -p));
}
LD_HD dd dd_add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi), t = two_sum(a.lo, b.lo);
    s.lo = LD_ADD(s.lo, t.hi);
    s = quick_two_sum(s.hi, s.lo);
    s.lo = LD_ADD(s.lo, t.lo);
    return quick_two_sum(s.hi, s.lo);
}
LD_HD dd dd_neg(dd a) { return dd_make(-a.hi, -a.lo); }
LD_HD dd dd_sub(dd a, dd b) { return dd_add(a, dd_neg(b)); }
LD_HD dd dd_mul(dd a, dd b) {
    dd p = two_prod(a.hi, b.hi);
    p.lo = LD_ADD(p.lo, LD_ADD(LD_MUL(a.hi, b.lo), LD_MUL(a.lo, b.hi)));
    return quick_two_sum(p.hi, p.lo);
}
LD_HD dd dd_mul_d(dd a, double b) {
    dd p = two_prod(a.hi, b);
    p.lo = LD_ADD(p.lo, LD_MUL(a.lo, b));
    return quick_two_sum(p.hi, p.lo);
}
LD_HD dd dd_div(dd a, dd b) {
    double q1 = LD_DIV(a.hi, b.hi);
    dd r = dd_sub(a, dd_mul_d(b, q1));
    double q2 = LD_DIV(r.hi, b.hi);
    r = dd_sub(r, dd_mul_d(b, q2));
    double q3 = LD_DIV(r.hi, b.hi);
    dd q = quick_two_sum(q1, q2);
    return dd_add(q, dd_make(q3));
}
LD_HD dd dd_sqrt(dd a) {
    if (a.hi <= 0.0) return dd_make(0.0);
    double x = LD_DIV(1.0, LD_SQRT(a.hi));
    double ax = LD_MUL(a.hi, x);
    dd r = dd_sub(a, two_prod(ax, ax));
    return dd_add(dd_make(ax), dd_make(LD_MUL(LD_MUL(r.hi, x), 0.5)));
}
LD_HD dd dd_from_u64(uint64_t v) {           // exact for v < 2^62 (moments of a 1280x720 frame stay below 2^58)
    double h = (double)v;
    return dd_make(h, (double)((int64_t)v - (int64_t)h));
}
LD_HD double dd_to_double(dd a) { return LD_ADD(a.hi, a.lo); }
LD_HD bool dd_abs_lt(dd a, dd b) {           // |a| < |b|
    dd x = a.hi < 0 ? dd_neg(a) : a, y = b.hi < 0 ? dd_neg(b) : b;
    return x.hi < y.hi || (x.hi == y.hi && x.lo < y.lo);
}

// The point set one polyfit call sees, as power sums.
struct PowerSums {
    dd S[5];        // sum y^k
    dd T[3];        // sum x*y^k
    double n;       // number of points (len(x) in np.polyfit's rcond)
};

struct FitResult {
    double c[3];    // A, B, C  (x = A*y^2 + B*y + C)
    int rank;       // singular values kept
};

// np.polyfit(y, x, 2) with rcond = n * eps.
LD_HD FitResult polyfit2(const PowerSums& ps, double eps) {
    FitResult out;
    // scale_i = sqrt(sum lhs_i^2), lhs columns = y^2, y, 1
    dd sc[3];
    sc[0] = dd_sqrt(ps.S[4]);
    sc[1] = dd_sqrt(ps.S[2]);
    sc[2] = dd_sqrt(ps.S[0]);
    for (int i = 0; i < 3; ++i)
        if (sc[i].hi == 0.0) sc[i] = dd_make(1.0);     // a zero column stays zero (numpy would emit nan; never hit with y=720 anchors or n>0 rows>0)
    dd G[3][3], r[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = i; j < 3; ++j) {
            dd g = (i == j && ps.S[4 - 2 * i].hi != 0.0) ? dd_make(1.0)
                                                         : dd_div(ps.S[4 - i - j], dd_mul(sc[i], sc[j]));
            G[i][j] = g;
            G[j][i] = g;
        }
        r[i] = dd_div(ps.T[2 - i], sc[i]);
    }
    // cyclic Jacobi, V accumulates the eigenvectors (columns)
    dd V[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) V[i][j] = dd_make(i == j ? 1.0 : 0.0);
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = fabs(G[0][1].hi) + fabs(G[0][2].hi) + fabs(G[1][2].hi);
        if (off < 1e-33) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                dd apq = G[p][q];
                if (fabs(apq.hi) < 1e-300) continue;
                // t = sgn(theta) / (|theta| + sqrt(theta^2 + 1)), theta = (aqq - app) / (2 apq)
                dd theta = dd_div(dd_sub(G[q][q], G[p][p]), dd_mul_d(apq, 2.0));
                dd at = theta.hi < 0 ? dd_neg(theta) : theta;
                dd t = dd_div(dd_make(1.0), dd_add(at, dd_sqrt(dd_add(dd_mul(theta, theta), dd_make(1.0)))));
                if (theta.hi < 0) t = dd_neg(t);
                dd c = dd_div(dd_make(1.0), dd_sqrt(dd_add(dd_mul(t, t), dd_make(1.0))));
                dd s = dd_mul(t, c);
                dd app = G[p][p], aqq = G[q][q];
                G[p][p] = dd_sub(app, dd_mul(t, apq));
                G[q][q] = dd_add(aqq, dd_mul(t, apq));
                G[p][q] = dd_make(0.0);
                G[q][p] = dd_make(0.0);
                int k = 3 - p - q;                      // the third index
                dd akp = G[k][p], akq = G[k][q];
                G[k][p] = dd_sub(dd_mul(c, akp), dd_mul(s, akq));
                G[p][k] = G[k][p];
                G[k][q] = dd_add(dd_mul(s, akp), dd_mul(c, akq));
                G[q][k] = G[k][q];
                for (int i = 0; i < 3; ++i) {
                    dd vip = V[i][p], viq = V[i][q];
                    V[i][p] = dd_sub(dd_mul(c, vip), dd_mul(s, viq));
                    V[i][q] = dd_add(dd_mul(s, vip), dd_mul(c, viq));
                }
            }
    }
    double lmax = fmax(G[0][0].hi, fmax(G[1][1].hi, G[2][2].hi));
    double smax = lmax > 0 ? sqrt(lmax) : 0.0;
    double cutoff = ps.n * eps * smax;
    dd cs[3] = {dd_make(0.0), dd_make(0.0), dd_make(0.0)};
    out.rank = 0;
    for (int e = 0; e < 3; ++e) {
        double lam = G[e][e].hi;
        if (!(lam > 0.0) || !(sqrt(lam) > cutoff)) continue;
        out.rank++;
        dd proj = dd_add(dd_add(dd_mul(V[0][e], r[0]), dd_mul(V[1][e], r[1])), dd_mul(V[2][e], r[2]));
        dd coef = dd_div(proj, G[e][e]);
        for (int i = 0; i < 3; ++i) cs[i] = dd_add(cs[i], dd_mul(V[i][e], coef));
    }
    for (int i = 0; i < 3; ++i) out.c[i] = dd_to_double(dd_div(cs[i], sc[i]));
    return out;
}

// power sums of an integer pixel set from its exact u64 moments
//   m = {n, Sy, Sy2, Sy3, Sy4, Sx, Sxy, Sxy2}
LD_HD PowerSums sums_from_moments(const uint64_t* m) {
    PowerSums ps;
    for (int k = 0; k < 5; ++k) ps.S[k] = dd_from_u64(m[k]);
    for (int k = 0; k < 3; ++k) ps.T[k] = dd_from_u64(m[5 + k]);
    ps.n = (double)m[0];
    return ps;
}
// add `count` points (x_i, y) whose x sum is `xsum` (y integer-valued)
LD_HD void sums_add_level(PowerSums& ps, double y, double count, dd xsum) {
    dd yk = dd_make(1.0);
    for (int k = 0; k < 5; ++k) {
        ps.S[k] = dd_add(ps.S[k], dd_mul_d(yk, count));
        if (k < 3) ps.T[k] = dd_add(ps.T[k], dd_mul(yk, xsum));
        yk = dd_mul_d(yk, y);
    }
    ps.n += count;
}

// x(y) exactly as numpy evaluates  fit[0]*y**2 + fit[1]*y + fit[2]  (no FMA contraction)
LD_HD double eval_fit(const double* c, double y) {
    return LD_ADD(LD_ADD(LD_MUL(c[0], LD_MUL(y, y)), LD_MUL(c[1], y)), c[2]);
}

// np.median of m <= 16 values (copy is sorted in place)
LD_HD double median_small(double* v, int m) {
    for (int i = 1; i < m; ++i) {
        double key = v[i];
        int j = i - 1;
        while (j >= 0 && v[j] > key) { v[j + 1] = v[j]; --j; }
        v[j + 1] = key;
    }
    if (m & 1) return v[m / 2];
    return LD_MUL(LD_ADD(v[m / 2 - 1], v[m / 2]), 0.5);
}

}  // namespace lanefit
