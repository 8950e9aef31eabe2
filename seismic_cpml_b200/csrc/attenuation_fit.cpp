// Relaxation times of the generalized Zener body from (N, Q, f0, f_min, f_max): driver-side set-up of
// the viscoelastic programs, host code (the reference runs it once, on the CPU, before the time loop:
// 3D-visco and the 2-D viscoelastic programs call compute_attenuation_coeffs twice, for Qp and Qs).
//
// Restates attenuation_model_with_SolvOpt.f90:
//   compute_attenuation_coeffs        :122-169
//   classical linear least squares    :175-485  (LU solve, lfit_zener, func_zener, remplit_point)
//   SolvOpt (Shor's r-algorithm, Kappel & Kuntsevich 2000), constrained branch with user gradients,
//   the only one nonlinear_optimization ever takes  :489-1751
//   objective / gradient / constraint :1777-1950, nonlinear_optimization :1952-1999
//
// Arithmetic notes kept from the Fortran so that the iteration follows the same path:
//   * exponents written with default-real literals, `(l-1.)/(N-1.)`, are evaluated in single precision
//     and then widened (frac32 below);
//   * `eps = 1.e-20` in the LU decomposition is a single-precision literal;
//   * `x ** 4.` etc. are pow() calls, `ajb ** integer` is repeated multiplication (powi below).
// The unconstrained / finite-difference-gradient branches of SolvOpt are dead code in the reference
// (flg = flfc = flgc = .true. at :1967-1969, apprgrdn commented out) and are not restated.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "../../include/cpml_b200.h"

namespace {

constexpr double PI = 3.141592653589793;   // :127 (16-digit literal, same double as the 25-digit one)
constexpr double TWO_PI = 2.0 * PI;

// (a - 1.) / (b - 1.) with default-real literals: integer -> real(4), single-precision arithmetic
inline double frac32(int a, int b) {
    volatile float num = static_cast<float>(a) - 1.0f;
    volatile float den = static_cast<float>(b) - 1.0f;
    volatile float q = num / den;
    return static_cast<double>(q);
}

// x ** m for an integer m the way the compiler's runtime does it (square-and-multiply, reciprocal last)
inline double powi(double x, long m) {
    unsigned long n = m < 0 ? static_cast<unsigned long>(-m) : static_cast<unsigned long>(m);
    double y = (n & 1ul) ? x : 1.0;
    while (n >>= 1) {
        x = x * x;
        if (n & 1ul) y *= x;
    }
    return m < 0 ? 1.0 / y : y;
}

// ---------------------------------------------------------------- classical linear least squares

// Crout LU with implicit scaling and partial pivoting, :175-251; a is n x n, a[i*n+j] = a(i,j)
void lu_decompose(std::vector<double> &a, int n, std::vector<int> &indx) {
    const double eps = static_cast<double>(1.e-20f);
    std::vector<double> vv(n);
    int imax = 0;
    for (int i = 0; i < n; ++i) {
        double big = 0.0;
        for (int j = 0; j < n; ++j)
            if (std::fabs(a[i * n + j]) > big) big = std::fabs(a[i * n + j]);
        vv[i] = 1.0 / big;   // a singular matrix only prints a message in the reference (:201-203)
    }
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < j; ++i) {
            double s = a[i * n + j];
            for (int k = 0; k < i; ++k) s -= a[i * n + k] * a[k * n + j];
            a[i * n + j] = s;
        }
        double big = 0.0;
        for (int i = j; i < n; ++i) {
            double s = a[i * n + j];
            for (int k = 0; k < j; ++k) s -= a[i * n + k] * a[k * n + j];
            a[i * n + j] = s;
            const double dum = vv[i] * std::fabs(s);
            if (dum >= big) { big = dum; imax = i; }
        }
        if (j != imax) {
            for (int k = 0; k < n; ++k) {
                const double dum = a[imax * n + k];
                a[imax * n + k] = a[j * n + k];
                a[j * n + k] = dum;
            }
            vv[imax] = vv[j];
        }
        indx[j] = imax;
        if (a[j * n + j] == 0.0) a[j * n + j] = eps;
        if (j != n - 1) {
            const double dum = 1.0 / a[j * n + j];
            for (int i = j + 1; i < n; ++i) a[i * n + j] *= dum;
        }
    }
}

// forward / back substitution for one right-hand side, :253-292
void lu_solve(const std::vector<double> &a, int n, const std::vector<int> &indx, std::vector<double> &b) {
    int ii = -1;
    for (int i = 0; i < n; ++i) {
        const int ip = indx[i];
        double s = b[ip];
        b[ip] = b[i];
        if (ii != -1) {
            for (int j = ii; j < i; ++j) s -= a[i * n + j] * b[j];
        } else if (s != 0.0) {
            ii = i;
        }
        b[i] = s;
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (int j = i + 1; j < n; ++j) s -= a[i * n + j] * b[j];
        b[i] = s / a[i * n + i];
    }
}

// basis functions of the linear fit, :401-419
void zener_basis(double x, double Qref, const std::vector<double> &point, std::vector<double> &afunc) {
    for (size_t k = 0; k < point.size(); ++k) {
        const double num = x * (point[k] - x / Qref);
        const double deno = point[k] * point[k] + x * x;
        afunc[k] = num / deno;
    }
}

// relaxation frequencies log-spaced over [fmin, fmax], :421-444 (N = 1: no 2 pi, as in the reference)
void fill_points(double fmin, double fmax, int N, std::vector<double> &point) {
    if (N == 1) {
        point[0] = std::sqrt(fmin * fmax);
        return;
    }
    for (int l = 1; l <= N; ++l) {
        const double p = std::pow(fmax / fmin, frac32(l, N));
        point[l - 1] = TWO_PI * p * fmin;
    }
}

// weights of the classical fit of 1/Q at 2N-1 log-spaced frequencies, :446-485 with lfit_zener :319-399
// (all parameters fitted, unit sigmas)
void classical_least_squares(double Qref, int N, double fmin, double fmax,
                             std::vector<double> &point, std::vector<double> &weight) {
    const int m = 2 * N - 1;
    fill_points(fmin, fmax, N, point);
    const double ref = 1.0 / Qref;
    std::vector<double> x(m), afunc(N), covar(static_cast<size_t>(N) * N, 0.0), beta(N, 0.0);
    for (int k = 1; k <= m; ++k) {
        double freq = std::pow(fmax / fmin, frac32(k, m));
        freq = TWO_PI * fmin * freq;
        x[k - 1] = freq;
    }
    for (int i = 0; i < m; ++i) {
        zener_basis(x[i], Qref, point, afunc);
        const double ym = ref;
        const double sig2i = 1.0;
        for (int l = 0; l < N; ++l) {
            const double wt = afunc[l] * sig2i;
            for (int k = 0; k <= l; ++k) covar[l * N + k] = covar[l * N + k] + wt * afunc[k];
            beta[l] = beta[l] + ym * wt;
        }
    }
    for (int j = 1; j < N; ++j)
        for (int k = 0; k < j; ++k) covar[k * N + j] = covar[j * N + k];
    if (N == 1) {
        weight[0] = beta[0] / covar[0];
    } else {
        std::vector<int> indx(N, 0);
        lu_decompose(covar, N, indx);
        lu_solve(covar, N, indx, beta);
        for (int j = 0; j < N; ++j) weight[j] = beta[j];
    }
}

// ---------------------------------------------------------------- the nonlinear problem

struct ZenerProblem {
    int N;           // mechanisms; unknowns x = (sqrt(theta_l - theta_min), sqrt(kappa_l)), 2N of them
    int K;           // sampling frequencies, 4N (:1971)
    double Qref, f_min, f_max, theta_min, theta_max;

    double freq(int i) const {   // i = 1..K, :1817 / :1845
        return TWO_PI * f_min * std::pow(f_max / f_min, frac32(i, K));
    }
    // Q_ref * (1/Q of the model) at one angular frequency, :1777-1796
    double model(const double *x, double w) const {
        double res = 0.0;
        for (int i = 0; i < N; ++i) {
            const double num = x[N + i] * x[N + i] * w * Qref * (x[i] * x[i] - w / Qref);
            const double deno = std::pow(x[i], 4.0) + w * w;
            res = res + num / deno;
        }
        return res;
    }
    // misfit, :1798-1824
    double fun(const double *x) const {
        double res = 0.0;
        for (int i = 1; i <= K; ++i) {
            const double d = model(x, freq(i)) - 1.0;
            res = res + d * d;
        }
        return res;
    }
    // its gradient, :1826-1881
    void grad(const double *x, double *g) const {
        std::vector<double> w(K);
        for (int i = 1; i <= K; ++i) w[i - 1] = freq(i);
        for (int l = 0; l < N; ++l) {
            const double p = x[l], q = x[N + l];
            g[l] = 0.0;
            g[N + l] = 0.0;
            for (int i = 0; i < K; ++i) {
                const double R = model(x, w[i]);
                const double temp3 = R - 1.0;
                const double temp0 = w[i] * Qref;
                double temp1 = temp0 * (p * p - w[i] / Qref);
                temp1 = temp1 * 2.0 * q;
                const double temp2 = std::pow(p, 4.0) + w[i] * w[i];
                temp1 = temp1 / temp2;
                double tamp = 2.0 * temp3 * temp1;
                g[N + l] = g[N + l] + tamp;
                const double aux1 = -2.0 * std::pow(p, 5.0) + 2.0 * p * w[i] * w[i]
                                    + 4.0 * std::pow(p, 3.0) * w[i] / Qref;
                const double aux3 = temp2 * temp2;
                double aux4 = aux1 / aux3;
                aux4 = aux4 * temp0;
                const double aux2 = aux4 * q * q;
                tamp = 2.0 * temp3 * aux2;
                g[l] = g[l] + tamp;
            }
        }
    }
    // maximal residual of the constraints theta_l <= theta_max, :1883-1904
    double func(const double *x) const {
        double res = 0.0;
        for (int l = 0; l < N; ++l) {
            const double aux = res;
            const double temp = std::fmax(0.0, x[l] * x[l] - (theta_max - theta_min));
            res = std::fmax(temp, aux);
        }
        return res;
    }
    // gradient of the constraint with the maximal residual, :1906-1950
    void gradc(const double *x, double *g) const {
        double res = 0.0;
        int l0 = 0;
        for (int l = 0; l < N; ++l) {
            const double aux = res;
            const double temp = std::fmax(0.0, x[l] * x[l] - (theta_max - theta_min));
            res = std::fmax(temp, aux);
            if (temp > aux) l0 = l;
        }
        for (int l = 0; l < N; ++l) {
            g[N + l] = 0.0;
            if (l != l0) g[l] = 0.0;
            else g[l0] = (func(x) == 0.0) ? 0.0 : 2.0 * x[l0];
        }
    }
};

struct SolvOptResult {
    double f = 0.0;
    double status = 0.0;   // options(9): iterations if > 0, stop code if < 0
    long nfun = 0, ngrad = 0, nfunc = 0, ngradc = 0;
};

inline double norm2(const std::vector<double> &v) {
    double s = 0.0;
    for (double e : v) s = s + e * e;
    return std::sqrt(s);
}

// SolvOpt, constrained minimisation with analytic gradients (:489-1751), default options (:1753-1775)
SolvOptResult solvopt(const ZenerProblem &P, std::vector<double> &x, bool warn) {
    const int n = static_cast<int>(x.size());
    SolvOptResult R;
    const double infty = 1e100, epsnorm = 1e-15, epsnorm2 = 1e-30, powerm12 = 1e-12;
    // options(2), (3), (4), (6), (7) — defaults; options(3) is never changed on the constrained path
    const double opt2 = 1e-4, opt3 = 1e-6, opt6 = 1e-8, opt7 = 2.5;
    const long iterlimit = 15000;
    const double n_float = static_cast<double>(n);
    std::vector<double> B(static_cast<size_t>(n) * n), g(n), g0(n), g1(n), gt(n), gc(n, 0.0), z(n), x1(n),
        xopt(n), xrec(n), grec(n), xx(n, 0.0);
    std::vector<int> idx(n);
    auto msg = [&](const char *s) { if (warn) std::fprintf(stderr, "SolvOpt: %s\n", s); };

    const double h1 = -1.0;                 // constrained problems are minimised
    const double cnteps = opt6;
    long k = 0;
    const double wdef = 1.0 / opt7 - 1.0;
    const double ajb = 1.0 + 1.0e-1 / (n_float * n_float);
    long ajp = 20;
    const long ajpp = ajp;
    const double ajs = 1.15;
    int knorms = 0;
    double gnorms[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const double dq = 5.1, du20 = 2.0, du10 = 1.5, du03 = 1.05;
    const int kstore = 3;
    double nsteps[3] = {0, 0, 0};
    const double des = 3.3;
    const int mxtc = 3;
    int termx = 0;
    const int limxterm = 50;
    const double low_bound = -1.0 + 1e-4;
    const double ZeroGrad = n_float * 1e-16;
    const double lowxbound = std::fmax(opt2, 1e-3);
    const double lowfbound = opt3 * opt3;
    int warnno = 0, stepvanish = 0;
    bool stopf = false;

    // first function value, :819-862
    double f = P.fun(x.data());
    R.nfun++;
    // (NaN added to the reference's test: a NaN misfit would never leave the 1-D search)
    if (std::fabs(f) >= infty || std::isnan(f)) { msg("function is not finite at the starting point"); R.status = -3; R.f = f; return R; }
    xrec = x;
    double frec = f;
    int kless = 0;
    double fp = f;
    double fc = P.func(x.data());
    R.nfunc++;
    if (std::fabs(fc) >= infty) { msg("constraint function is infinite at the starting point"); R.status = -5; R.f = f; return R; }
    double PenCoef = 1.0;
    bool FsbPnt;
    if (fc <= cnteps) { FsbPnt = true; fc = 0.0; } else { FsbPnt = false; }
    f = f + PenCoef * fc;

    // first gradient, :863-947
    P.grad(x.data(), g.data());
    R.ngrad++;
    double ng = norm2(g);
    if (ng >= infty) { msg("gradient is infinite at the starting point"); R.status = -4; R.f = f; return R; }
    if (ng < ZeroGrad) { msg("gradient is zero at the starting point"); R.status = -4; R.f = f; return R; }
    if (!FsbPnt) {
        P.gradc(x.data(), gc.data());
        // the reference tests ng (not the constraint gradient norm) here, :917-932
        for (int i = 0; i < n; ++i) g[i] = g[i] + PenCoef * gc[i];
        ng = norm2(g);
    }
    grec = g;
    double nng = ng;

    // initial step size, :949-959 (options(1) = -1: computed)
    double d = 0.0;
    for (int i = 0; i < n; ++i) if (d < std::fabs(x[i])) d = std::fabs(x[i]);
    double h = h1 * std::sqrt(opt2) * d;
    h = h1 * std::fmax(1.0 / std::log(ng + 1.1), std::fabs(h));

    double dx = 0.0, laststep = 0.0, fst = f, fopt = f, hp = h, f1 = f, fp1 = fp;
    bool FsbPnt1 = FsbPnt, Reset = false;
    int kd = 4;

    for (;;) {   // resetting loop, :961
        long kcheck = 0;
        int kg = 0, kj = 0;
        for (int i = 0; i < n; ++i) {
            for (int j = 0; j < n; ++j) B[i * n + j] = 0.0;
            B[i * n + i] = 1.0;
            g1[i] = g[i];
        }
        fst = f;
        dx = 0.0;

        for (;;) {   // main iterations, :978
            k = k + 1;
            kcheck = kcheck + 1;
            laststep = dx;
            // gamma, :983-985
            double gamma = 1.0 + std::fmax(powi(ajb, (ajp - kcheck) * n), 2.0 * opt3);
            gamma = std::fmin(gamma, std::pow(ajs, std::fmax(1.0, std::log10(nng + 1.0))));
            double ngt = 0.0, ng1 = 0.0, dd = 0.0;
            for (int i = 0; i < n; ++i) {
                d = 0.0;
                for (int j = 0; j < n; ++j) d = d + B[j * n + i] * g[j];
                gt[i] = d;
                dd = dd + d * g1[i];
                ngt = ngt + d * d;
                ng1 = ng1 + g1[i] * g1[i];
            }
            ngt = std::sqrt(ngt);
            ng1 = std::sqrt(ng1);
            dd = dd / ngt / ng1;

            double w = wdef;
            // jumping over a ravine, :1005-1030
            if (dd < low_bound) {
                if (kj == 2) xx = x;
                if (kj == 0) kd = 4;
                kj = kj + 1;
                w = -0.9;
                h = h * 2.0;
                if (kj > 2 * kd) {
                    kd = kd + 1;
                    warnno = 1;
                    for (int i = 0; i < n; ++i)
                        if (std::fabs(x[i] - xx[i]) < epsnorm * std::fabs(x[i])) msg("ravine with a flat bottom is detected");
                }
            } else {
                kj = 0;
            }
            // dilation, :1032-1084
            double nrmz = 0.0;
            for (int i = 0; i < n; ++i) {
                z[i] = gt[i] - g1[i];
                nrmz = nrmz + z[i] * z[i];
            }
            nrmz = std::sqrt(nrmz);
            if (nrmz > epsnorm * ngt) {
                for (int i = 0; i < n; ++i) z[i] = z[i] / nrmz;
                d = 0.0;
                for (int i = 0; i < n; ++i) d = d + z[i] * gt[i];
                ng1 = 0.0;
                d = d * w;
                for (int i = 0; i < n; ++i) {
                    dd = 0.0;
                    g1[i] = gt[i] + d * z[i];
                    ng1 = ng1 + g1[i] * g1[i];
                    for (int j = 0; j < n; ++j) dd = dd + B[i * n + j] * z[j];
                    dd = w * dd;
                    for (int j = 0; j < n; ++j) B[i * n + j] = B[i * n + j] + dd * z[j];
                }
                ng1 = std::sqrt(ng1);
            } else {
                for (int i = 0; i < n; ++i) { z[i] = 0.0; g1[i] = gt[i]; }
                nrmz = 0.0;
            }
            for (int i = 0; i < n; ++i) gt[i] = g1[i] / ng1;
            for (int i = 0; i < n; ++i) {
                d = 0.0;
                for (int j = 0; j < n; ++j) d = d + B[i * n + j] * gt[j];
                g0[i] = d;
            }
            // resetting of the transformation matrix, :1086-1117
            if (kcheck > 1) {
                int numelem = 0;
                for (int i = 0; i < n; ++i)
                    if (std::fabs(g[i]) > ZeroGrad) idx[numelem++] = i;
                if (numelem > 0) {
                    const double grbnd = epsnorm * static_cast<double>(numelem * numelem);
                    int ii = 0;
                    for (int i = 0; i < numelem; ++i) {
                        const int j = idx[i];
                        if (std::fabs(g1[j]) <= std::fabs(g[j]) * grbnd) ii = ii + 1;
                    }
                    if (ii == n || nrmz == 0.0) {
                        msg("normal re-setting of a transformation matrix");
                        if (std::fabs(fst - f) < std::fabs(f) * 1e-2) ajp = ajp - 10 * n;
                        else ajp = ajpp;
                        h = h1 * dx / 3.0;
                        k = k - 1;
                        break;
                    }
                }
            }
            // one-dimensional search along g0, :1118-1282
            xopt = x;
            fopt = f;
            int k1 = 0, k2 = 0, kc = 0;
            bool ksm = false, knan = false;
            hp = h;
            Reset = false;
            for (;;) {
                x1 = x;
                f1 = f;
                FsbPnt1 = FsbPnt;
                fp1 = fp;
                for (int i = 0; i < n; ++i) x[i] = x[i] + hp * g0[i];
                int ii = 0;
                for (int i = 0; i < n; ++i)
                    if (std::fabs(x[i] - x1[i]) < std::fabs(x[i]) * epsnorm) ii = ii + 1;
                f = P.fun(x.data());
                R.nfun++;
                if (h1 * f >= infty || std::isnan(f)) { msg("function is unbounded"); R.status = -7; R.f = f; return R; }
                fp = f;
                fc = P.func(x.data());
                R.nfunc++;
                if (std::fabs(fc) >= infty) { msg("constraint function is infinite"); R.status = -5; R.f = f; return R; }
                if (fc <= cnteps) {
                    FsbPnt = true;
                    fc = 0.0;
                } else {
                    FsbPnt = false;
                    const double fp_rate = fp - fp1;
                    if (fp_rate < -epsnorm && !FsbPnt1) {
                        d = 0.0;
                        for (int i = 0; i < n; ++i) d = d + (x[i] - x1[i]) * (x[i] - x1[i]);
                        d = std::sqrt(d);
                        const double PenCoefNew = -1.5e1 * fp_rate / d;
                        if (PenCoefNew > 1.2 * PenCoef) {
                            PenCoef = PenCoefNew;
                            Reset = true;
                            kless = 0;
                            f = f + PenCoef * fc;
                            break;
                        }
                    }
                }
                f = f + PenCoef * fc;
                if (std::fabs(f) >= infty) {
                    msg("function equals infinity at the point");
                    if (ksm || kc >= mxtc) { R.status = -3; R.f = f; return R; }
                    k2 = k2 + 1;
                    k1 = 0;
                    hp = hp / dq;
                    x = x1;
                    f = f1;
                    knan = true;
                    FsbPnt = FsbPnt1;
                    fp = fp1;
                } else if (ii == n) {   // the step vanished to the extent of epsnorm
                    stepvanish = stepvanish + 1;
                    if (stepvanish >= 5) {
                        msg("stopping criteria are not fulfilled: the function is very steep at the solution");
                        R.status = -14; R.f = f; return R;
                    }
                    x = x1;
                    f = f1;
                    hp = hp * 10.0;
                    ksm = true;
                    FsbPnt = FsbPnt1;
                    fp = fp1;
                } else if (h1 * f < h1 * powi(gamma, f1 >= 0.0 ? 1 : -1) * f1) {   // use a smaller step
                    if (ksm) break;
                    k2 = k2 + 1;
                    k1 = 0;
                    hp = hp / dq;
                    x = x1;
                    f = f1;
                    FsbPnt = FsbPnt1;
                    fp = fp1;
                    if (kc >= mxtc) break;
                } else {
                    if (h1 * f <= h1 * f1) break;   // the 1-D optimizer is left behind
                    k1 = k1 + 1;                    // use a larger step
                    if (k2 > 0) kc = kc + 1;
                    k2 = 0;
                    if (k1 >= 20) hp = du20 * hp;
                    else if (k1 >= 10) hp = du10 * hp;
                    else if (k1 >= 3) hp = du03 * hp;
                }
            }
            // trial step size, :1283-1316
            dx = 0.0;
            for (int i = 0; i < n; ++i) dx = dx + (xopt[i] - x[i]) * (xopt[i] - x[i]);
            dx = std::sqrt(dx);
            if (kg < kstore) kg = kg + 1;
            if (kg >= 2)
                for (int i = kg; i >= 2; --i) nsteps[i - 1] = nsteps[i - 2];
            d = norm2(g0);
            nsteps[0] = dx / (std::fabs(h) * d);
            double kk = 0.0;
            d = 0.0;
            for (int i = 1; i <= kg; ++i) {
                dd = static_cast<double>(kg - i + 1);
                d = d + dd;
                kk = kk + nsteps[i - 1] * dd;
            }
            kk = kk / d;
            if (kk > des) {
                if (kg == 1) h = h * (kk - des + 1.0);
                else h = h * std::sqrt(kk - des + 1.0);
            } else if (kk < des) {
                h = h * std::sqrt(kk / des);
            }
            if (ksm) stepvanish = stepvanish + 1;
            // gradient at the new point, :1318-1420
            P.grad(x.data(), g.data());
            R.ngrad++;
            ng = norm2(g);
            if (ng >= infty) { msg("gradient is infinite"); R.status = -4; R.f = f; return R; }
            if (ng < ZeroGrad) { msg("gradient is zero, but stopping criteria are not fulfilled"); ng = ZeroGrad; }
            if (!FsbPnt) {
                if (ng < 1e-2 * PenCoef) {
                    kless = kless + 1;
                    if (kless >= 20) { PenCoef = PenCoef / 10.0; Reset = true; kless = 0; }
                } else {
                    kless = 0;
                }
                P.gradc(x.data(), gc.data());
                R.ngradc++;
                const double ngc = norm2(gc);
                if (ngc >= infty) { msg("constraint gradient is infinite"); R.status = -6; R.f = f; return R; }
                if (ngc < ZeroGrad) { msg("constraint gradient is zero at an infeasible point"); R.status = -6; R.f = f; return R; }
                for (int i = 0; i < n; ++i) g[i] = g[i] + PenCoef * gc[i];
                ng = norm2(g);
                if (Reset) {
                    msg("re-setting due to the use of a new penalty coefficient");
                    h = h1 * dx / 3.0;
                    k = k - 1;
                    nng = ng;
                    break;
                }
            }
            if (h1 * f > h1 * frec) {
                frec = f;
                xrec = x;
                grec = g;
            }
            if (ng > ZeroGrad) {   // geometric mean of the last gradient norms, :1422-1436
                if (knorms < 10) knorms = knorms + 1;
                if (knorms >= 2)
                    for (int i = knorms; i >= 2; --i) gnorms[i - 1] = gnorms[i - 2];
                gnorms[0] = ng;
                nng = 1.0;
                for (int i = 0; i < knorms; ++i) nng = nng * gnorms[i];
                nng = std::pow(nng, 1.0 / static_cast<double>(knorms));
            }
            const double nx = norm2(x);

            // stopping criteria, :1449-1566
            bool termflag = true;
            if (!FsbPnt) termflag = false;
            if (kcheck <= 5 || (kcheck <= 12 && ng > 1.0)) termflag = false;
            if (kc >= mxtc || knan) termflag = false;
            if (termflag) {
                int ii = 0;
                bool stopping = true;
                for (int i = 0; i < n; ++i) {
                    if (std::fabs(x[i]) >= lowxbound) {
                        idx[ii++] = i;
                        if (std::fabs(xopt[i] - x[i]) > opt2 * std::fabs(x[i])) stopping = false;
                    }
                }
                if (ii == 0 || stopping) {
                    termx = termx + 1;
                    d = 0.0;
                    for (int i = 0; i < n; ++i) d = d + (x[i] - xrec[i]) * (x[i] - xrec[i]);
                    d = std::sqrt(d);
                    // (the "re-run from recorded point" branch needs an unconstrained problem, :1479-1512)
                    if (std::fabs(f - frec) > opt3 * std::fabs(f) && d < opt2 * nx) {
                        // constrained: keep iterating, :1513-1515
                    } else if (std::fabs(f - fopt) <= opt3 * std::fabs(f) || std::fabs(f) <= lowfbound ||
                               (std::fabs(f - fopt) <= opt3 && termx >= limxterm)) {
                        if (stopf) {
                            if (dx <= laststep) {
                                if (warnno == 1 && ng < std::sqrt(opt3)) warnno = 0;
                                for (int i = 0; i < n; ++i)
                                    if (std::fabs(g[i]) <= epsnorm2) { warnno = 3; break; }
                                if (warnno != 0) {
                                    R.status = -static_cast<double>(warnno) - 10.0;
                                    msg(warnno == 1 ? "termination warning: premature stop is possible"
                                                    : "termination warning: the function is flat at the solution");
                                } else {
                                    R.status = static_cast<double>(k);
                                }
                                R.f = f;
                                return R;
                            }
                        } else {
                            stopf = true;
                        }
                    } else if (dx < powerm12 * std::fmax(nx, 1.0) && termx >= limxterm) {
                        msg("termination warning: the function is very steep at the solution");
                        f = frec;   // (inside `if (dispwarn)` in the reference, which is on by default, :1553-1561)
                        x = xrec;
                        R.status = -14; R.f = f; return R;
                    }
                }
            }
            if (k == iterlimit) { msg("iterations limit exceeded"); R.status = -9; R.f = f; return R; }
            if (ng <= ZeroGrad) {   // constrained problems stop on a zero gradient, :1579-1587
                msg("gradient is zero, but stopping criteria are not fulfilled");
                R.status = -8; R.f = f; return R;
            }
        }
    }
}

}  // namespace

// compute_attenuation_coeffs, attenuation_model_with_SolvOpt.f90:122-169
extern "C" int32_t cpml_host_attenuation_fit(int32_t n_sls, double qref, double f0, double f_min,
                                             double f_max, double *tau_epsilon, double *tau_sigma,
                                             double *info) {
    // N = 1 is not usable in the reference either: its first guess samples 2N-1 = 1 frequency at
    // (k-1.)/(m-1.) = 0./0. (:469)
    if (n_sls < 2 || !(qref > 0.0) || !(f0 > 0.0) || !(f_min > 0.0) || !(f_max > f_min) ||
        tau_epsilon == nullptr || tau_sigma == nullptr)
        return CPML_EINVAL;
    const int N = n_sls;
    std::vector<double> point(N), weight(N);
    ZenerProblem P;
    P.N = N;
    P.K = 4 * N;
    P.Qref = qref;
    P.f_min = f_min;
    P.f_max = f_max;
    P.theta_min = TWO_PI * 0.0;
    P.theta_max = TWO_PI * 100.0 * f0;
    classical_least_squares(qref, N, f_min, f_max, point, weight);   // first guess, :1976
    std::vector<double> x(2 * N);
    for (int i = 0; i < N; ++i) {
        x[i] = std::sqrt(std::fabs(point[i]) - P.theta_min);
        x[N + i] = std::sqrt(std::fabs(weight[i]));
    }
    const SolvOptResult R = solvopt(P, x, false);
    for (int i = 0; i < N; ++i) {
        point[i] = P.theta_min + x[i] * x[i];
        weight[i] = x[N + i] * x[N + i];
    }
    for (int i = 0; i < N; ++i) {
        tau_sigma[i] = 1.0 / point[i];
        tau_epsilon[i] = tau_sigma[i] * (1.0 + N * weight[i]);
    }
    if (info != nullptr) {
        info[0] = R.status;
        info[1] = R.f;
        info[2] = static_cast<double>(R.nfun);
        info[3] = static_cast<double>(R.ngrad);
    }
    for (int i = 0; i < N; ++i)
        if (!std::isfinite(tau_sigma[i]) || !std::isfinite(tau_epsilon[i])) return CPML_EINVAL;
    return CPML_OK;
}

// the classical linear least-squares fit alone (the first guess of the nonlinear one, :446-485)
extern "C" int32_t cpml_host_attenuation_fit_linear(int32_t n_sls, double qref, double f_min, double f_max,
                                                    double *tau_epsilon, double *tau_sigma) {
    if (n_sls < 2 || !(qref > 0.0) || !(f_min > 0.0) || !(f_max > f_min) || tau_epsilon == nullptr ||
        tau_sigma == nullptr)
        return CPML_EINVAL;
    const int N = n_sls;
    std::vector<double> point(N), weight(N);
    classical_least_squares(qref, N, f_min, f_max, point, weight);
    for (int i = 0; i < N; ++i) {
        tau_sigma[i] = 1.0 / point[i];
        tau_epsilon[i] = tau_sigma[i] * (1.0 + N * weight[i]);
    }
    return CPML_OK;
}
