// cl2_stats.cuh -- fp64 tail probabilities used by estLoopSig / estAnchorSig, usable from host and device.
//
// The reference calls scipy.stats.hypergeom.sf / poisson.sf / binom.sf (callCisLoops.py:246,310,329,332;
// callTransLoops.py:153,178,181).  scipy is a third-party dependency the reference does not pin; parity is defined
// against scipy 1.18.1 (Boost hypergeometric / binomial, cephes pdtrc) to 1e-9 relative, tails down to 1e-300.
// Everything here is built on Catherine Loader's saddle-point form of the binomial / Poisson mass functions
// ("Fast and accurate computation of binomial probabilities", 2000): the mass at the starting point is accurate to
// a few ulps even for N ~ 1e8, and the tail is summed with the exact term ratio, all terms positive.
#pragma once
#include <math.h>
#include <float.h>

#ifdef __CUDACC__
#define CL2_HD __host__ __device__ __forceinline__
#define CL2_HD_OUTLINE __attribute__((unused)) static __host__ __device__ __noinline__      // the three tail functions: called rarely, keep code small
#else
#define CL2_HD static inline
#define CL2_HD_OUTLINE static
#endif

#define CL2_LN_SQRT_2PI 0.918938533204672741780329736406
#define CL2_2PI 6.283185307179586476925286766559

// stirlerr(n) for n = 0, 0.5, 1, ..., 15 (exact to double precision)
#define CL2_SFERR_TABLE {0.0, 0.1534264097200273452913848, 0.0810614667953272582196702, 0.0548141210519176538961390, 0.0413406959554092940938221, 0.03316287351993628748511048, 0.02767792568499833914878929, 0.02374616365629749597132920, 0.02079067210376509311152277, 0.01848845053267318523077934, 0.01664469118982119216319487, 0.01513497322191737887351255, 0.01387612882307074799874573, 0.01281046524292022692424986, 0.01189670994589177009505572, 0.01110455975820691732662991, 0.010411265261972096497478567, 0.009799416126158803298389475, 0.009255462182712732917728637, 0.008768700134139385462952823, 0.008330563433362871256469318, 0.007934114564314020547248100, 0.007573675487951840794972024, 0.007244554301320383179543912, 0.006942840107209529865664152, 0.006665247032707682442354394, 0.006408994188004207068439631, 0.006171712263039457647532867, 0.005951370112758847735624416, 0.005746216513010115682023589, 0.005554733551962801371038690}
#ifdef __CUDACC__
static __constant__ double cl2_sferr_dev[31] = CL2_SFERR_TABLE;
#endif
static const double cl2_sferr_host[31] = CL2_SFERR_TABLE;

// error of Stirling's formula: log(n!) - log(sqrt(2 pi n) (n/e)^n)
CL2_HD double cl2_stirlerr(double n) {
    const double S0 = 0.083333333333333333333;        // 1/12
    const double S1 = 0.00277777777777777777778;      // 1/360
    const double S2 = 0.00079365079365079365079365;   // 1/1260
    const double S3 = 0.000595238095238095238095238;  // 1/1680
    const double S4 = 0.0008417508417508417508417508; // 1/1188
    if (n <= 15.0) {
        double nn = n + n;
#ifdef __CUDA_ARCH__
        if (nn == floor(nn)) return cl2_sferr_dev[(int)nn];
#else
        if (nn == floor(nn)) return cl2_sferr_host[(int)nn];
#endif
        return lgamma(n + 1.0) - (n + 0.5) * log(n) + n - CL2_LN_SQRT_2PI;
    }
    double nn = n * n;
    if (n > 500) return (S0 - S1 / nn) / n;
    if (n > 80) return (S0 - (S1 - S2 / nn) / nn) / n;
    if (n > 35) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
    return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}

// deviance term bd0(x, np) = x log(x/np) + np - x, evaluated without cancellation near x == np
CL2_HD double cl2_bd0(double x, double np) {
    if (fabs(x - np) < 0.1 * (x + np)) {
        double v = (x - np) / (x + np);
        double s = (x - np) * v;
        if (fabs(s) < DBL_MIN) return s;
        double ej = 2 * x * v;
        v = v * v;
        for (int j = 1; j < 1000; j++) {
            ej *= v;
            double s1 = s + ej / ((j << 1) + 1);
            if (s1 == s) return s1;
            s = s1;
        }
    }
    return x * log(x / np) + np - x;
}

// log of the binomial mass C(n,x) p^x q^(n-x)
CL2_HD double cl2_log_dbinom(double x, double n, double p, double q) {
    if (p == 0) return x == 0 ? 0.0 : -INFINITY;
    if (q == 0) return x == n ? 0.0 : -INFINITY;
    if (x == 0) {
        if (n == 0) return 0.0;
        return p < 0.1 ? -cl2_bd0(n, n * q) - n * p : n * log(q);
    }
    if (x == n) return q < 0.1 ? -cl2_bd0(n, n * p) - n * q : n * log(p);
    if (x < 0 || x > n) return -INFINITY;
    double lc = cl2_stirlerr(n) - cl2_stirlerr(x) - cl2_stirlerr(n - x) - cl2_bd0(x, n * p) - cl2_bd0(n - x, n * q);
    double lf = CL2_LN_SQRT_2PI + 0.5 * (log(x) + log1p(-x / n));   // 0.5*log(2 pi x (n-x)/n)
    return lc - lf;
}

// log of the Poisson mass
CL2_HD double cl2_log_dpois(double x, double lambda) {
    if (lambda == 0) return x == 0 ? 0.0 : -INFINITY;
    if (x < 0) return -INFINITY;
    if (x == 0) return -lambda;
    return -cl2_stirlerr(x) - cl2_bd0(x, lambda) - 0.5 * log(CL2_2PI * x);
}

// upper tail sum pmf(k0) * (1 + r(k0) + r(k0) r(k0+1) + ...) with term ratio given by a functor-like switch
// kind 0: binomial(n,p): r(j) = (n-j)/(j+1) * p/q ; kind 1: Poisson(mu): r(j) = mu/(j+1) ;
// kind 2: hypergeometric: r(j) = (a-j)(b-j)/((j+1)(N-a-b+j+1))
CL2_HD double cl2_ratio_up(int kind, double j, double a, double b, double c) {
    if (kind == 0) return (a - j) / (j + 1.0) * b;                 // a = n, b = p/q
    if (kind == 1) return a / (j + 1.0);                           // a = mu
    return (a - j) * (b - j) / ((j + 1.0) * (c - a - b + j + 1.0)); // a, b = successes / draws, c = N
}
CL2_HD double cl2_ratio_down(int kind, double j, double a, double b, double c) {   // pmf(j-1)/pmf(j)
    if (kind == 0) return j / (a - j + 1.0) / b;
    if (kind == 1) return j / a;
    return j * (c - a - b + j) / ((a - j + 1.0) * (b - j + 1.0));
}

// P(X >= k) given log pmf(k) (sum upward) -- all terms positive, stops when the term no longer changes the sum
CL2_HD double cl2_tail_up(int kind, double k, double kmax, double logpk, double a, double b, double c) {
    double term = 1.0, sum = 1.0;
    for (double j = k; j < kmax; j += 1.0) {
        term *= cl2_ratio_up(kind, j, a, b, c);
        double s1 = sum + term;
        if (s1 == sum || term < sum * 1e-18) break;
        sum = s1;
    }
    return exp(logpk + log(sum));
}
// P(X <= k) given log pmf(k) (sum downward to kmin)
CL2_HD double cl2_tail_down(int kind, double k, double kmin, double logpk, double a, double b, double c) {
    double term = 1.0, sum = 1.0;
    for (double j = k; j > kmin; j -= 1.0) {
        term *= cl2_ratio_down(kind, j, a, b, c);
        double s1 = sum + term;
        if (s1 == sum || term < sum * 1e-18) break;
        sum = s1;
    }
    return exp(logpk + log(sum));
}

// scipy.stats.poisson.sf(k - 1, mu) = P(X >= k), k integer >= 0 as double
CL2_HD_OUTLINE double cl2_poisson_sf_ge(double k, double mu) {
    if (!(mu >= 0)) return NAN;
    if (k <= 0) return 1.0;
    if (mu == 0) return 0.0;
    if (k > mu) return cl2_tail_up(1, k, INFINITY, cl2_log_dpois(k, mu), mu, 0, 0);
    double lower = cl2_tail_down(1, k - 1.0, 0.0, cl2_log_dpois(k - 1.0, mu), mu, 0, 0);   // P(X <= k-1)
    return 1.0 - lower;
}

// scipy.stats.binom.sf(k - 1, n, p) = P(X >= k)
CL2_HD_OUTLINE double cl2_binom_sf_ge(double k, double n, double p) {
    if (!(p >= 0 && p <= 1) || !(n >= 0)) return NAN;
    if (k <= 0) return 1.0;
    if (k > n) return 0.0;
    if (p == 0) return 0.0;
    if (p == 1) return 1.0;
    double q = 1.0 - p;
    if (k > n * p) return cl2_tail_up(0, k, n, cl2_log_dbinom(k, n, p, q), n, p / q, 0);
    double lower = cl2_tail_down(0, k - 1.0, 0.0, cl2_log_dbinom(k - 1.0, n, p, q), n, p / q, 0);
    return 1.0 - lower;
}

// scipy.stats.hypergeom.sf(k - 1, M, n, N) = P(X >= k): population M, n successes, N draws
CL2_HD double cl2_log_dhyper(double x, double M, double n, double N) {
    // R's dhyper: p = N/M ; dbinom(x; n, p) dbinom(N-x; M-n, p) / dbinom(N; M, p)
    double p = N / M, q = (M - N) / M;
    return cl2_log_dbinom(x, n, p, q) + cl2_log_dbinom(N - x, M - n, p, q) - cl2_log_dbinom(N, M, p, q);
}
CL2_HD_OUTLINE double cl2_hypergeom_sf_ge(double k, double M, double n, double N) {
    double lo = fmax(0.0, N + n - M), hi = fmin(n, N);
    if (k <= lo) return 1.0;
    if (k > hi) return 0.0;
    double mode = floor((N + 1.0) * (n + 1.0) / (M + 2.0));
    if (k > mode) return cl2_tail_up(2, k, hi, cl2_log_dhyper(k, M, n, N), n, N, M);
    double lower = cl2_tail_down(2, k - 1.0, lo, cl2_log_dhyper(k - 1.0, M, n, N), n, N, M);
    return 1.0 - lower;
}
