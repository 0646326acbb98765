// Nodal SDE energies and their partial derivatives for the four drifts.
//
// The reference keeps these as dill-pickled SymPy expressions
// (src/dynamical_systems/energy_functions/*.sym, gradient_functions/*.sym,
// loaded at double_well.py:111-127, ornstein_uhlenbeck.py:105-121,
// lorenz_63.py:198-224, lorenz_96.py:396-415).  They all follow one template
// (examples/how_to_compute_energies.ipynb; notes/PRE_Note_01.jpg):
//
//   E_i = (E[f_i] - dm_i)^2 + Var[f_i] + q_i^2/(4 s_i) - q_i E[df_i/dx_i],
//   q_i = ds_i - Sigma_i,
//
// with Gaussian moments of the drift f under N(m, diag(s)).  Here they are
// written in closed form; m, dm, s, ds are the interpolated mean, its time
// derivative, the variance and its time derivative at one quadrature node.
#pragma once

namespace mfvgp {

// 1/x to ~1 ulp: hardware seed + two Newton steps.  Variances are exp() of
// finite numbers so x is a positive normal number.
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// rational part shared by every system: rat = q^2/(4s) and its derivatives
struct Rational {
    double q, rat, rat_s, rat_ds;
    __device__ __forceinline__ Rational(double s, double ds, double Sig) {
        q = ds - Sig;
        const double qi = q * fast_rcp(s);     // q / s
        rat = 0.25 * q * qi;                   // q^2 / (4 s)
        rat_s = -0.25 * qi * qi;               // d rat / d s
        rat_ds = 0.5 * qi;                     // d rat / d ds
    }
};

template <int SYS> struct Drift;

// ---- Ornstein-Uhlenbeck: f = theta (mu - x), ornstein_uhlenbeck.py:77 -------
template <> struct Drift<1> {
    static constexpr int D = 1;
    static constexpr int NTHETA = 2;
    // E[i], Em[i][j] = dE_i/dm_j, Edm[i] = dE_i/d(dm_i), Es[i][j], Eds[i]
    __device__ static __forceinline__ void nodal(
        const double *m, const double *dm, const double *s, const double *ds,
        const double *Sig, const double *th, double *E, double (*Em)[1],
        double *Edm, double (*Es)[1], double *Eds) {
        const double t = th[0], mu = th[1];
        const Rational r(s[0], ds[0], Sig[0]);
        const double R = t * (mu - m[0]) - dm[0];
        E[0] = R * R + t * t * s[0] + r.rat + t * r.q;
        Em[0][0] = -2.0 * t * R;
        Edm[0] = -2.0 * R;
        Es[0][0] = t * t + r.rat_s;
        Eds[0] = r.rat_ds + t;
    }
};

// ---- Double well: f = 4 x (theta - x^2), double_well.py:82-83 ---------------
template <> struct Drift<0> {
    static constexpr int D = 1;
    static constexpr int NTHETA = 1;
    __device__ static __forceinline__ void nodal(
        const double *m, const double *dm, const double *s, const double *ds,
        const double *Sig, const double *th, double *E, double (*Em)[1],
        double *Edm, double (*Es)[1], double *Eds) {
        const double t = th[0];
        const Rational r(s[0], ds[0], Sig[0]);
        const double m1 = m[0], s1 = s[0], d1 = dm[0];
        const double m2 = m1 * m1, m3 = m2 * m1, m4 = m2 * m2, s2 = s1 * s1;
        const double Ef = 4.0 * (t * m1 - m3 - 3.0 * m1 * s1);
        const double Ef2 = 16.0 * (t * t * (m2 + s1)
                                   - 2.0 * t * (m4 + 6.0 * m2 * s1 + 3.0 * s2)
                                   + m4 * m2 + 15.0 * m4 * s1 + 45.0 * m2 * s2
                                   + 15.0 * s2 * s1);
        const double Edf = 4.0 * t - 12.0 * (m2 + s1);
        E[0] = Ef2 - 2.0 * Ef * d1 + d1 * d1 + r.rat - r.q * Edf;
        const double dEf_dm = 4.0 * (t - 3.0 * m2 - 3.0 * s1);
        const double dEf_ds = -12.0 * m1;
        const double dEf2_dm = 16.0 * (2.0 * t * t * m1
                                       - 2.0 * t * (4.0 * m3 + 12.0 * m1 * s1)
                                       + 6.0 * m4 * m1 + 60.0 * m3 * s1
                                       + 90.0 * m1 * s2);
        const double dEf2_ds = 16.0 * (t * t - 2.0 * t * (6.0 * m2 + 6.0 * s1)
                                       + 15.0 * m4 + 90.0 * m2 * s1 + 45.0 * s2);
        Em[0][0] = dEf2_dm - 2.0 * d1 * dEf_dm + 24.0 * r.q * m1;
        Edm[0] = -2.0 * Ef + 2.0 * d1;
        Es[0][0] = dEf2_ds - 2.0 * d1 * dEf_ds + r.rat_s + 12.0 * r.q;
        Eds[0] = r.rat_ds - Edf;
    }
};

// ---- Lorenz 63: theta = (sigma, rho, beta), lorenz_63.py:30-32 --------------
template <> struct Drift<2> {
    static constexpr int D = 3;
    static constexpr int NTHETA = 3;
    __device__ static __forceinline__ void nodal(
        const double *m, const double *dm, const double *s, const double *ds,
        const double *Sig, const double *th, double *E, double (*Em)[3],
        double *Edm, double (*Es)[3], double *Eds) {
        const double a = th[0], rho = th[1], b = th[2];
        const Rational r0(s[0], ds[0], Sig[0]);
        const Rational r1(s[1], ds[1], Sig[1]);
        const Rational r2(s[2], ds[2], Sig[2]);
        const double rz = rho - m[2];
        const double R0 = a * (m[1] - m[0]) - dm[0];
        const double R1 = m[0] * rz - m[1] - dm[1];
        const double R2 = m[0] * m[1] - b * m[2] - dm[2];
        const double m00 = m[0] * m[0], m11 = m[1] * m[1];
        E[0] = R0 * R0 + a * a * (s[0] + s[1]) + r0.rat + a * r0.q;
        E[1] = R1 * R1 + m00 * s[2] + rz * rz * s[0] + s[0] * s[2] + s[1] + r1.rat + r1.q;
        E[2] = R2 * R2 + m00 * s[1] + m11 * s[0] + s[0] * s[1] + b * b * s[2] + r2.rat
               + b * r2.q;
        Em[0][0] = -2.0 * a * R0;  Em[0][1] = 2.0 * a * R0;  Em[0][2] = 0.0;
        Em[1][0] = 2.0 * R1 * rz + 2.0 * m[0] * s[2];
        Em[1][1] = -2.0 * R1;
        Em[1][2] = -2.0 * R1 * m[0] - 2.0 * rz * s[0];
        Em[2][0] = 2.0 * R2 * m[1] + 2.0 * m[0] * s[1];
        Em[2][1] = 2.0 * R2 * m[0] + 2.0 * m[1] * s[0];
        Em[2][2] = -2.0 * b * R2;
        Edm[0] = -2.0 * R0;  Edm[1] = -2.0 * R1;  Edm[2] = -2.0 * R2;
        Es[0][0] = a * a + r0.rat_s;  Es[0][1] = a * a;  Es[0][2] = 0.0;
        Es[1][0] = rz * rz + s[2];  Es[1][1] = 1.0 + r1.rat_s;  Es[1][2] = m00 + s[0];
        Es[2][0] = m11 + s[1];  Es[2][1] = m00 + s[0];  Es[2][2] = b * b + r2.rat_s;
        Eds[0] = r0.rat_ds + a;  Eds[1] = r1.rat_ds + 1.0;  Eds[2] = r2.rat_ds + b;
    }
};

}  // namespace mfvgp
