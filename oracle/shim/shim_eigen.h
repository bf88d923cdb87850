// shim_eigen.h -- TEST INFRASTRUCTURE (our own code, not a copy of Eigen).
// A minimal, eagerly evaluated stand-in for the small part of Eigen 3 / RcppEigen that the reference's five source
// files use, so that those files can be compiled UNMODIFIED from /root/reference/src into oracle/_ref/ and the C
// restatement (bayesrr_oracle.c) can be pinned against them.  Eigen itself, Rcpp and R are absent from this image.
// Semantics kept: column-major storage, coefficient-wise .array() algebra, segment/col/row views that alias their
// parent, comma initialisation, IOFormat printing through ostream << double.  Not kept: expression templates,
// Eigen's packet reduction order (results agree with real Eigen to rounding, not bitwise).
// Vectors are zero-initialised (a zeroed heap) and carry two slack elements so that the reference's one-past-the-end
// reads stay defined.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <iostream>
#include <random>
#include <string>
#include <vector>

#define EIGEN_EMPTY_STRUCT_CTOR(X)

namespace Eigen {

typedef std::ptrdiff_t Index;
enum { StreamPrecision = -1, FullPrecision = -2 };
enum { DontAlignCols = 1 };
enum { Lower = 1, Upper = 2 };

struct IOFormat {
    int precision, flags;
    std::string coeffSeparator, rowSeparator, rowPrefix, rowSuffix, matPrefix, matSuffix;
    IOFormat(int p = StreamPrecision, int f = 0, const std::string &cs = " ", const std::string &rs = "\n",
             const std::string &rp = "", const std::string &rsuf = "", const std::string &mp = "", const std::string &ms = "")
        : precision(p), flags(f), coeffSeparator(cs), rowSeparator(rs), rowPrefix(rp), rowSuffix(rsuf), matPrefix(mp), matSuffix(ms) {}
};

struct BoolArr {
    std::vector<char> b;
    bool any() const { for (char c : b) if (c) return true; return false; }
    bool all() const { for (char c : b) if (!c) return false; return true; }
};

class Mat;
class View;

// ---- coefficient-wise value type (what .array() yields)
class Arr {
public:
    std::vector<double> d;
    Arr() {}
    explicit Arr(size_t n, double v = 0.0) : d(n, v) {}
    size_t size() const { return d.size(); }
    Arr array() const { return *this; }
    double sum() const { double s = 0; for (double x : d) s += x; return s; }
    Arr log() const { Arr r(d.size()); for (size_t i = 0; i < d.size(); ++i) r.d[i] = std::log(d[i]); return r; }
    Arr exp() const { Arr r(d.size()); for (size_t i = 0; i < d.size(); ++i) r.d[i] = std::exp(d[i]); return r; }
    Arr abs() const { Arr r(d.size()); for (size_t i = 0; i < d.size(); ++i) r.d[i] = std::fabs(d[i]); return r; }
    Arr pow(double p) const { Arr r(d.size()); for (size_t i = 0; i < d.size(); ++i) r.d[i] = std::pow(d[i], p); return r; }
    Arr cwiseInverse() const { Arr r(d.size()); for (size_t i = 0; i < d.size(); ++i) r.d[i] = 1.0 / d[i]; return r; }
    template <class F> Arr unaryExpr(const F &f) const { Arr r(d.size()); for (size_t i = 0; i < d.size(); ++i) r.d[i] = f(d[i]); return r; }
    BoolArr operator>(double v) const { BoolArr r; r.b.resize(d.size()); for (size_t i = 0; i < d.size(); ++i) r.b[i] = d[i] > v; return r; }
    BoolArr operator<(double v) const { BoolArr r; r.b.resize(d.size()); for (size_t i = 0; i < d.size(); ++i) r.b[i] = d[i] < v; return r; }
    BoolArr operator==(double v) const { BoolArr r; r.b.resize(d.size()); for (size_t i = 0; i < d.size(); ++i) r.b[i] = d[i] == v; return r; }
};
#define SHIM_AA(op)                                                                                            \
    inline Arr operator op(const Arr &a, const Arr &b) { Arr r(a.size()); for (size_t i = 0; i < a.size(); ++i) r.d[i] = a.d[i] op b.d[i]; return r; } \
    inline Arr operator op(const Arr &a, double b) { Arr r(a.size()); for (size_t i = 0; i < a.size(); ++i) r.d[i] = a.d[i] op b; return r; }         \
    inline Arr operator op(double a, const Arr &b) { Arr r(b.size()); for (size_t i = 0; i < b.size(); ++i) r.d[i] = a op b.d[i]; return r; }
SHIM_AA(+) SHIM_AA(-) SHIM_AA(*) SHIM_AA(/)
#undef SHIM_AA

// ---- strided view into a Mat (segment / col / row); aliases the parent's storage
class View {
public:
    double *p; Index n, stride;
    View(double *p_, Index n_, Index s_) : p(p_), n(n_), stride(s_) {}
    Index size() const { return n; }
    double &operator()(Index i) { return p[i * stride]; }
    double operator()(Index i) const { return p[i * stride]; }
    double &operator[](Index i) { return p[i * stride]; }
    double operator[](Index i) const { return p[i * stride]; }
    Arr array() const { Arr r((size_t)n); for (Index i = 0; i < n; ++i) r.d[(size_t)i] = p[i * stride]; return r; }
    double sum() const { return array().sum(); }
    double squaredNorm() const { double s = 0; for (Index i = 0; i < n; ++i) s += p[i * stride] * p[i * stride]; return s; }
    View segment(Index i, Index len) const { return View(p + i * stride, len, stride); }
    Arr cwiseInverse() const { return array().cwiseInverse(); }
    inline Mat cwiseProduct(const Mat &o) const;
    View &operator=(const Arr &a) { for (Index i = 0; i < n; ++i) p[i * stride] = a.d[(size_t)i]; return *this; }
    View &operator=(const View &o) { Arr a = o.array(); return *this = a; }
    inline View &operator=(const Mat &m);
};

// ---- dense column-major matrix; VectorXd and MatrixXd are both this type (a vector is n x 1)
class Mat {
public:
    Index r, c;
    std::vector<double> d;     // r * c coefficients + 2 slack elements (zero)
    Mat() : r(0), c(0), d(2, 0.0) {}
    Mat(Index n) : r(n), c(1), d((size_t)n + 2, 0.0) {}
    Mat(Index rows_, Index cols_) : r(rows_), c(cols_), d((size_t)(rows_ * cols_) + 2, 0.0) {}
    Mat(const Arr &a) : r((Index)a.size()), c(1), d(a.d) { d.resize(a.size() + 2, 0.0); }
    Mat(const View &v) : r(v.n), c(1), d((size_t)v.n + 2, 0.0) { for (Index i = 0; i < v.n; ++i) d[(size_t)i] = v(i); }
    Index size() const { return r * c; }
    Index rows() const { return r; }
    Index cols() const { return c; }
    double *data() { return d.data(); }
    const double *data() const { return d.data(); }
    double &operator[](Index i) { return d[(size_t)i]; }
    double operator[](Index i) const { return d[(size_t)i]; }
    double &operator()(Index i) { return d[(size_t)i]; }
    double operator()(Index i) const { return d[(size_t)i]; }
    double &operator()(Index i, Index j) { return d[(size_t)(i + j * r)]; }
    double operator()(Index i, Index j) const { return d[(size_t)(i + j * r)]; }
    double &coeffRef(Index i) { return d[(size_t)i]; }
    Mat &setZero() { std::fill(d.begin(), d.end(), 0.0); return *this; }
    Mat &setOnes() { std::fill(d.begin(), d.begin() + (size_t)size(), 1.0); return *this; }
    Arr array() const { Arr a((size_t)size()); std::copy(d.begin(), d.begin() + (size_t)size(), a.d.begin()); return a; }
    double sum() const { double s = 0; for (Index i = 0; i < size(); ++i) s += d[(size_t)i]; return s; }
    double squaredNorm() const { double s = 0; for (Index i = 0; i < size(); ++i) s += d[(size_t)i] * d[(size_t)i]; return s; }
    View segment(Index i, Index len) { return View(d.data() + i, len, 1); }
    View col(Index j) { return View(d.data() + j * r, r, 1); }
    View row(Index i) { return View(d.data() + i, c, r); }
    Mat cwiseInverse() const { return Mat(array().cwiseInverse()); }
    Mat cwiseProduct(const Mat &o) const { return Mat(array() * o.array()); }
    template <class F> Mat unaryExpr(const F &f) const { return Mat(array().unaryExpr(f)); }
    Mat &operator=(const Arr &a)
    {
        if ((Index)a.size() != size()) { r = (Index)a.size(); c = 1; }
        d.assign(a.d.begin(), a.d.end()); d.resize(a.size() + 2, 0.0); return *this;
    }
    Mat &operator=(const View &v) { return *this = v.array(); }
    Mat &operator/=(double s) { for (Index i = 0; i < size(); ++i) d[(size_t)i] /= s; return *this; }
    struct ColwiseProxy {
        const Mat &m;
        Mat squaredNorm() const { Mat o(m.c); for (Index j = 0; j < m.c; ++j) { double s = 0; for (Index i = 0; i < m.r; ++i) s += m(i, j) * m(i, j); o[j] = s; } return o; }
    };
    ColwiseProxy colwise() const { return ColwiseProxy{ *this }; }
    // printing: row.transpose().format(fmt)
    struct Formatted { const Mat &m; const IOFormat &f; bool transposed; };
    struct Transposed {
        const Mat &m;
        Formatted format(const IOFormat &f) const { return Formatted{ m, f, true }; }
    };
    Transposed transpose() const { return Transposed{ *this }; }
    Mat adjoint() const { Mat o(c, r); for (Index i = 0; i < r; ++i) for (Index j = 0; j < c; ++j) o(j, i) = (*this)(i, j); return o; }
    template <int UpLo> struct SelfAdj {
        Mat &m;
        Mat rankUpdate(const Mat &u) const { Mat o(m.r, m.c); for (Index i = 0; i < m.r; ++i) for (Index j = 0; j < m.c; ++j) { double s = 0; for (Index k = 0; k < u.c; ++k) s += u(i, k) * u(j, k); o(i, j) = s; } return o; }
    };
    template <int UpLo> SelfAdj<UpLo> selfadjointView() { return SelfAdj<UpLo>{ *this }; }
    // comma initialisation: v << a, b, mat, vec ...   (the reference packs every sample row this way)
    struct Comma {
        Mat &m; Index pos;
        Comma &put(double x) { m.d[(size_t)pos++] = x; return *this; }
        Comma &operator,(double x) { return put(x); }
        Comma &operator,(int x) { return put((double)x); }
        Comma &operator,(const Mat &o) { for (Index i = 0; i < o.size(); ++i) put(o[i]); return *this; }
        ~Comma();
    };
    Comma operator<<(double x) { Comma cm{ *this, 0 }; cm.put(x); return cm; }
    Comma operator<<(int x) { return *this << (double)x; }
};

extern void (*shim_row_hook)(const double *row, long len);   // set by ref_glue.cpp: receives every comma-initialised row
inline Mat::Comma::~Comma() { if (shim_row_hook && pos == m.size()) shim_row_hook(m.d.data(), (long)m.size()); }

inline Mat View::cwiseProduct(const Mat &o) const { return Mat(array() * o.array()); }
inline View &View::operator=(const Mat &m) { Arr a = m.array(); return *this = a; }

inline std::ostream &operator<<(std::ostream &os, const Mat::Formatted &fm)
{
    // a transposed column vector is one row: coefficients joined by coeffSeparator, printed by ostream << double
    os << fm.f.matPrefix << fm.f.rowPrefix;
    for (Index i = 0; i < fm.m.size(); ++i) { if (i) os << fm.f.coeffSeparator; os << fm.m[i]; }
    os << fm.f.rowSuffix << fm.f.matSuffix;
    return os;
}

inline Mat operator+(const Mat &a, const Mat &b) { return Mat(a.array() + b.array()); }
inline Mat operator-(const Mat &a, const Mat &b) { return Mat(a.array() - b.array()); }
inline Mat operator*(const Mat &a, double s) { Mat o(a); for (Index i = 0; i < o.size(); ++i) o[i] *= s; return o; }
inline Mat operator*(double s, const Mat &a) { return a * s; }
inline Mat operator*(const View &v, double s) { return Mat(v) * s; }
inline Mat operator*(double s, const View &v) { return Mat(v) * s; }
inline Mat operator-(const Mat &a, const View &b) { return a - Mat(b); }
inline Mat operator*(const Mat &a, const Mat &b)   // matrix product (the reference uses it once: X * beta)
{
    Mat o(a.r, b.c);
    for (Index j = 0; j < b.c; ++j) for (Index k = 0; k < a.c; ++k) { const double bk = b(k, j); if (bk != 0.0) for (Index i = 0; i < a.r; ++i) o(i, j) += a(i, k) * bk; }
    return o;
}

typedef Mat MatrixXd;
typedef Mat VectorXd;

class VectorXi {
public:
    std::vector<int> d;
    VectorXi() {}
    VectorXi(Index n) : d((size_t)n, 0) {}
    Index size() const { return (Index)d.size(); }
    int *data() { return d.data(); }
    const int *data() const { return d.data(); }
    int &operator()(Index i) { return d[(size_t)i]; }
    int operator()(Index i) const { return d[(size_t)i]; }
    int &operator[](Index i) { return d[(size_t)i]; }
    int operator[](Index i) const { return d[(size_t)i]; }
};

template <class T> class Map : public T { public: Map() {} };
template <class S, int R = -1, int C = -1> class Matrix {};
template <class S> class SparseVector {};
template <class M> class LLT {};

inline void initParallel() {}
inline void setNbThreads(int) {}

}  // namespace Eigen
