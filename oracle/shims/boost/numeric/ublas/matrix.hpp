// oracle/shims/boost/numeric/ublas -- TEST INFRASTRUCTURE.
// Just enough of uBLAS for vamp_probit.cpp:936-1000 (Newton_method_cov): dense row-major
// matrix<double>, vector<double>, prod(), permutation_matrix, lu_factorize, lu_substitute.
#pragma once
#include <cstddef>
#include <cmath>
#include <vector>

namespace boost { namespace numeric { namespace ublas {

template <class T>
class vector {
public:
    vector() {}
    explicit vector(std::size_t n) : d_(n, T()) {}
    std::size_t size() const { return d_.size(); }
    T& operator()(std::size_t i) { return d_[i]; }
    const T& operator()(std::size_t i) const { return d_[i]; }
    T& operator[](std::size_t i) { return d_[i]; }
    const T& operator[](std::size_t i) const { return d_[i]; }
private:
    std::vector<T> d_;
};

template <class T>
class matrix {
public:
    matrix() : r_(0), c_(0) {}
    matrix(std::size_t r, std::size_t c) : r_(r), c_(c), d_(r * c, T()) {}
    std::size_t size1() const { return r_; }
    std::size_t size2() const { return c_; }
    T& operator()(std::size_t i, std::size_t j) { return d_[i * c_ + j]; }
    const T& operator()(std::size_t i, std::size_t j) const { return d_[i * c_ + j]; }
private:
    std::size_t r_, c_;
    std::vector<T> d_;
};

template <class T>
inline matrix<T> prod(const matrix<T>& a, const matrix<T>& b) {
    matrix<T> out(a.size1(), b.size2());
    for (std::size_t i = 0; i < a.size1(); i++)
        for (std::size_t k = 0; k < a.size2(); k++) {
            T aik = a(i, k);
            for (std::size_t j = 0; j < b.size2(); j++) out(i, j) += aik * b(k, j);
        }
    return out;
}
template <class T>
inline vector<T> prod(const matrix<T>& a, const vector<T>& x) {
    vector<T> out(a.size1());
    for (std::size_t i = 0; i < a.size1(); i++) {
        T s = T();
        for (std::size_t k = 0; k < a.size2(); k++) s += a(i, k) * x(k);
        out(i) = s;
    }
    return out;
}

template <class T = std::size_t>
class permutation_matrix {
public:
    explicit permutation_matrix(std::size_t n) : p_(n) { for (std::size_t i = 0; i < n; i++) p_[i] = i; }
    std::size_t size() const { return p_.size(); }
    std::size_t& operator()(std::size_t i) { return p_[i]; }
    const std::size_t& operator()(std::size_t i) const { return p_[i]; }
private:
    std::vector<std::size_t> p_;
};

// LU with partial pivoting, in place; returns 0 when non-singular (uBLAS convention: index+1 of
// the first zero pivot otherwise).
template <class T, class P>
inline int lu_factorize(matrix<T>& m, permutation_matrix<P>& pm) {
    std::size_t n = m.size1();
    int singular = 0;
    for (std::size_t i = 0; i < n; i++) {
        std::size_t piv = i;
        T best = std::fabs(m(i, i));
        for (std::size_t r = i + 1; r < n; r++)
            if (std::fabs(m(r, i)) > best) { best = std::fabs(m(r, i)); piv = r; }
        pm(i) = piv;
        if (m(piv, i) != T()) {
            if (piv != i)
                for (std::size_t c = 0; c < n; c++) { T t = m(i, c); m(i, c) = m(piv, c); m(piv, c) = t; }
            T inv = T(1) / m(i, i);
            for (std::size_t r = i + 1; r < n; r++) m(r, i) *= inv;
        } else if (singular == 0) {
            singular = (int)i + 1;
        }
        for (std::size_t r = i + 1; r < n; r++) {
            T f = m(r, i);
            for (std::size_t c = i + 1; c < n; c++) m(r, c) -= f * m(i, c);
        }
    }
    return singular;
}

template <class T, class P>
inline void lu_substitute(const matrix<T>& m, const permutation_matrix<P>& pm, vector<T>& b) {
    std::size_t n = m.size1();
    for (std::size_t i = 0; i < n; i++)
        if (pm(i) != i) { T t = b(i); b(i) = b(pm(i)); b(pm(i)) = t; }
    for (std::size_t i = 0; i < n; i++) {
        T s = b(i);
        for (std::size_t k = 0; k < i; k++) s -= m(i, k) * b(k);
        b(i) = s;
    }
    for (std::size_t ii = n; ii-- > 0;) {
        T s = b(ii);
        for (std::size_t k = ii + 1; k < n; k++) s -= m(ii, k) * b(k);
        b(ii) = s / m(ii, ii);
    }
}

}}}  // namespace boost::numeric::ublas
