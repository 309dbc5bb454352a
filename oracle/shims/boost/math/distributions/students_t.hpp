// oracle/shims/boost/math/distributions/students_t.hpp -- TEST INFRASTRUCTURE.
// Minimal stand-in for the one Boost.Math feature the reference uses (utilities.cpp:330-331):
//   boost::math::students_t dist(nu);  cdf(complement(dist, t))
// The upper tail is computed with the regularised incomplete beta function (Lentz continued
// fraction).  The reference's vamp.cpp also relies on this header pulling in <iomanip>.
#pragma once
#include <cmath>
#include <iomanip>
#include <limits>

namespace boost { namespace math {

namespace shim_detail {
inline double betacf(double a, double b, double x) {
    const int maxit = 500;
    const double eps = 1e-16, fpmin = 1e-300;
    double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (std::fabs(d) < fpmin) d = fpmin;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= maxit; m++) {
        int m2 = 2 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d; if (std::fabs(d) < fpmin) d = fpmin;
        c = 1.0 + aa / c; if (std::fabs(c) < fpmin) c = fpmin;
        d = 1.0 / d; h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d; if (std::fabs(d) < fpmin) d = fpmin;
        c = 1.0 + aa / c; if (std::fabs(c) < fpmin) c = fpmin;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (std::fabs(del - 1.0) < eps) break;
    }
    return h;
}
inline double ibeta(double a, double b, double x) {
    if (x <= 0.0) return 0.0;
    if (x >= 1.0) return 1.0;
    double bt = std::exp(std::lgamma(a + b) - std::lgamma(a) - std::lgamma(b) + a * std::log(x) + b * std::log1p(-x));
    if (x < (a + 1.0) / (a + b + 2.0)) return bt * betacf(a, b, x) / a;
    return 1.0 - bt * betacf(b, a, 1.0 - x) / b;
}
}  // namespace shim_detail

class students_t {
public:
    explicit students_t(double nu) : nu_(nu) {}
    double degrees_of_freedom() const { return nu_; }
private:
    double nu_;
};

template <class Dist, class Real>
struct shim_complement { const Dist& dist; Real param; };

template <class Dist, class Real>
inline shim_complement<Dist, Real> complement(const Dist& d, Real r) { return shim_complement<Dist, Real>{d, r}; }

inline double cdf(const students_t& d, double t) {
    double nu = d.degrees_of_freedom();
    double x = nu / (nu + t * t);
    double tail = 0.5 * shim_detail::ibeta(0.5 * nu, 0.5, x);
    return t > 0 ? 1.0 - tail : tail;
}
template <class Real>
inline double cdf(const shim_complement<students_t, Real>& c) {
    double nu = c.dist.degrees_of_freedom();
    double t = (double)c.param;
    double x = nu / (nu + t * t);
    double tail = 0.5 * shim_detail::ibeta(0.5 * nu, 0.5, x);
    return t > 0 ? tail : 1.0 - tail;
}

}}  // namespace boost::math
