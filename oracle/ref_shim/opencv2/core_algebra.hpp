// ref_shim/opencv2/core_algebra.hpp — TEST INFRASTRUCTURE (oracle/_ref build only; never part of the product).
// The small dense CV_32F algebra core/operators/mapInit/OP_2ViewReconstruction.cpp spells out with cv::Mat, so that the
// reference's translation unit compiles UNCHANGED in a container without OpenCV C++: products, sums, scalar factors,
// t(), inv(), eye / diag, norm, determinant, dot and cv::SVD (one-sided Jacobi in double).  Written for this shim; it
// is NOT OpenCV's gemm / LU / SVD and does not reproduce their last-bit rounding.  That is irrelevant to what the build
// pins: CheckHomography / CheckFundamental (:447-610) take their matrices element by element with at<float>() and are
// plain float code of the reference itself; the matrices they score are INPUTS of the parity tests.
#pragma once

namespace cv {

// The value of an expression.  OpenCV's MatExpr is lazy; here it is evaluated eagerly and only keeps the one semantic
// the reference relies on: assigning it to an existing matrix or VIEW of the same size and type writes into that
// storage (`A.row(0) = kp1.pt.x*P1.row(2)-P1.row(0)`, OP_2ViewReconstruction.cpp:1184).
class MatExpr : public Mat {
public:
    MatExpr() {}
    explicit MatExpr(const Mat& m) : Mat(m) {}
};

inline Mat& Mat::operator=(const MatExpr& e) {
    if (data && rows == e.rows && cols == e.cols && type() == e.type()) { e.copyTo(std::move(*this)); return *this; }
    return *this = static_cast<const Mat&>(e);
}

namespace shim_alg {
inline void need32f(const Mat& m) { assert(m.type() == CV_32F); (void)m; }
inline MatExpr make(int r, int c) { Mat m(r, c, CV_32F); return MatExpr(m); }
template <class F> inline MatExpr zip(const Mat& a, const Mat& b, F f) {
    need32f(a); need32f(b); assert(a.rows == b.rows && a.cols == b.cols);
    MatExpr o = make(a.rows, a.cols);
    for (int y = 0; y < a.rows; ++y) for (int x = 0; x < a.cols; ++x) o.at<float>(y, x) = f(a.at<float>(y, x), b.at<float>(y, x));
    return o;
}
template <class F> inline MatExpr map(const Mat& a, F f) {
    need32f(a);
    MatExpr o = make(a.rows, a.cols);
    for (int y = 0; y < a.rows; ++y) for (int x = 0; x < a.cols; ++x) o.at<float>(y, x) = f(a.at<float>(y, x));
    return o;
}
inline std::vector<double> to_double(const Mat& a) {
    need32f(a);
    std::vector<double> v((size_t)a.rows * a.cols);
    for (int y = 0; y < a.rows; ++y) for (int x = 0; x < a.cols; ++x) v[(size_t)y * a.cols + x] = a.at<float>(y, x);
    return v;
}
// LU with partial pivoting on an n x n double matrix; returns the sign of the permutation (0: singular)
inline int lu(std::vector<double>& a, int n, std::vector<int>& piv) {
    int sign = 1;
    piv.resize(n);
    for (int k = 0; k < n; ++k) {
        int p = k;
        for (int i = k + 1; i < n; ++i) if (std::fabs(a[(size_t)i * n + k]) > std::fabs(a[(size_t)p * n + k])) p = i;
        piv[k] = p;
        if (a[(size_t)p * n + k] == 0.0) return 0;
        if (p != k) { for (int j = 0; j < n; ++j) std::swap(a[(size_t)p * n + j], a[(size_t)k * n + j]); sign = -sign; }
        for (int i = k + 1; i < n; ++i) {
            const double f = a[(size_t)i * n + k] /= a[(size_t)k * n + k];
            for (int j = k + 1; j < n; ++j) a[(size_t)i * n + j] -= f * a[(size_t)k * n + j];
        }
    }
    return sign;
}
}  // namespace shim_alg

inline MatExpr operator*(const Mat& a, const Mat& b) {
    shim_alg::need32f(a); shim_alg::need32f(b); assert(a.cols == b.rows);
    MatExpr o = shim_alg::make(a.rows, b.cols);
    for (int y = 0; y < a.rows; ++y) for (int x = 0; x < b.cols; ++x) {
        double s = 0;
        for (int k = 0; k < a.cols; ++k) s += (double)a.at<float>(y, k) * (double)b.at<float>(k, x);
        o.at<float>(y, x) = (float)s;
    }
    return o;
}
inline MatExpr operator+(const Mat& a, const Mat& b) { return shim_alg::zip(a, b, [](float p, float q) { return p + q; }); }
inline MatExpr operator-(const Mat& a, const Mat& b) { return shim_alg::zip(a, b, [](float p, float q) { return p - q; }); }
inline MatExpr operator-(const Mat& a) { return shim_alg::map(a, [](float p) { return -p; }); }
inline MatExpr operator*(double s, const Mat& a) { return shim_alg::map(a, [s](float p) { return (float)(s * p); }); }
inline MatExpr operator*(const Mat& a, double s) { return s * a; }
inline Mat& operator*=(Mat& a, double s) { shim_alg::need32f(a); for (int y = 0; y < a.rows; ++y) for (int x = 0; x < a.cols; ++x) a.at<float>(y, x) = (float)(a.at<float>(y, x) * s); return a; }
inline MatExpr operator/(const Mat& a, double s) { return shim_alg::map(a, [s](float p) { return (float)(p / s); }); }

inline Mat Mat::eye(int r, int c, int type) { Mat m = Mat::zeros(r, c, type); assert(type == CV_32F); for (int i = 0; i < r && i < c; ++i) m.at<float>(i, i) = 1.f; return m; }
inline Mat Mat::diag(const Mat& d) {
    shim_alg::need32f(d);
    const int n = (int)d.total();
    Mat m = Mat::zeros(n, n, CV_32F);
    for (int i = 0; i < n; ++i) m.at<float>(i, i) = d.at<float>(i);
    return m;
}
inline MatExpr Mat::t() const {
    shim_alg::need32f(*this);
    MatExpr o = shim_alg::make(cols, rows);
    for (int y = 0; y < rows; ++y) for (int x = 0; x < cols; ++x) o.at<float>(x, y) = at<float>(y, x);
    return o;
}
inline MatExpr Mat::inv() const {      // DECOMP_LU: a singular matrix gives zeros, like cv::invert
    assert(rows == cols);
    const int n = rows;
    std::vector<double> a = shim_alg::to_double(*this);
    std::vector<int> piv;
    MatExpr o = shim_alg::make(n, n);
    o.fill(0);
    if (shim_alg::lu(a, n, piv) == 0) return o;
    for (int c = 0; c < n; ++c) {
        std::vector<double> x(n, 0.0);
        x[c] = 1.0;
        for (int k = 0; k < n; ++k) { std::swap(x[k], x[piv[k]]); for (int i = k + 1; i < n; ++i) x[i] -= a[(size_t)i * n + k] * x[k]; }
        for (int i = n - 1; i >= 0; --i) { for (int j = i + 1; j < n; ++j) x[i] -= a[(size_t)i * n + j] * x[j]; x[i] /= a[(size_t)i * n + i]; }
        for (int i = 0; i < n; ++i) o.at<float>(i, c) = (float)x[i];
    }
    return o;
}
inline double Mat::dot(const Mat& m) const {
    shim_alg::need32f(*this); shim_alg::need32f(m); assert(total() == m.total());
    double s = 0;
    for (int y = 0; y < rows; ++y) for (int x = 0; x < cols; ++x) s += (double)at<float>(y, x) * (double)m.at<float>(y, x);
    return s;
}
inline double norm(const Mat& a) { return std::sqrt(a.dot(a)); }
inline double determinant(const Mat& m) {
    assert(m.rows == m.cols);
    std::vector<double> a = shim_alg::to_double(m);
    std::vector<int> piv;
    double d = shim_alg::lu(a, m.rows, piv);
    for (int i = 0; i < m.rows && d != 0.0; ++i) d *= a[(size_t)i * m.rows + i];
    return d;
}

// Singular value decomposition A = u * diag(w) * vt, singular values in descending order, FULL_UV always (u is m x m and
// vt is n x n: the reference takes vt.row(8) of an 8 x 9 system, OP_2ViewReconstruction.cpp:438).  One-sided Jacobi
// (Hestenes) in double on the taller orientation, orthonormal completion by Gram-Schmidt.
class SVD {
public:
    enum { MODIFY_A = 1, NO_UV = 2, FULL_UV = 4 };
    static void compute(const Mat& A, Mat& w, Mat& u, Mat& vt, int /*flags*/ = 0) {
        shim_alg::need32f(A);
        const int m = A.rows, n = A.cols;
        const bool flip = m < n;                       // work on B (p x q, p >= q) = A or A^T
        const int p = flip ? n : m, q = flip ? m : n;
        std::vector<double> B((size_t)p * q), V((size_t)q * q, 0.0);
        for (int y = 0; y < m; ++y) for (int x = 0; x < n; ++x) (flip ? B[(size_t)x * q + y] : B[(size_t)y * q + x]) = A.at<float>(y, x);
        for (int i = 0; i < q; ++i) V[(size_t)i * q + i] = 1.0;
        for (int sweep = 0; sweep < 60; ++sweep) {
            bool rotated = false;
            for (int i = 0; i < q - 1; ++i) for (int j = i + 1; j < q; ++j) {
                double a = 0, b = 0, c = 0;
                for (int k = 0; k < p; ++k) { const double bi = B[(size_t)k * q + i], bj = B[(size_t)k * q + j]; a += bi * bi; b += bj * bj; c += bi * bj; }
                if (std::fabs(c) <= 1e-15 * std::sqrt(a * b) || c == 0.0) continue;
                rotated = true;
                const double zeta = (b - a) / (2.0 * c);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
                for (int k = 0; k < p; ++k) { double& bi = B[(size_t)k * q + i]; double& bj = B[(size_t)k * q + j]; const double x = bi, y = bj; bi = cs * x - sn * y; bj = sn * x + cs * y; }
                for (int k = 0; k < q; ++k) { double& vi = V[(size_t)k * q + i]; double& vj = V[(size_t)k * q + j]; const double x = vi, y = vj; vi = cs * x - sn * y; vj = sn * x + cs * y; }
            }
            if (!rotated) break;
        }
        std::vector<double> sv(q);
        std::vector<int> ord(q);
        for (int i = 0; i < q; ++i) { double s = 0; for (int k = 0; k < p; ++k) s += B[(size_t)k * q + i] * B[(size_t)k * q + i]; sv[i] = std::sqrt(s); ord[i] = i; }
        std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return sv[x] > sv[y]; });
        // P: p x p orthonormal, first columns = normalised columns of B (in order), the rest completed
        std::vector<double> P((size_t)p * p, 0.0);
        int filled = 0;
        auto add_column = [&](const std::vector<double>& cand) {
            std::vector<double> v = cand;
            for (int pass = 0; pass < 2; ++pass) for (int c = 0; c < filled; ++c) {
                double d = 0; for (int k = 0; k < p; ++k) d += v[k] * P[(size_t)k * p + c];
                for (int k = 0; k < p; ++k) v[k] -= d * P[(size_t)k * p + c];
            }
            double nn = 0; for (int k = 0; k < p; ++k) nn += v[k] * v[k];
            nn = std::sqrt(nn);
            if (nn < 1e-9) return false;
            for (int k = 0; k < p; ++k) P[(size_t)k * p + filled] = v[k] / nn;
            ++filled;
            return true;
        };
        const double tiny = (q ? sv[ord[0]] : 0.0) * 1e-12;
        std::vector<char> placed(q, 0);
        for (int c = 0; c < q; ++c) {
            const int i = ord[c];
            std::vector<double> col(p);
            for (int k = 0; k < p; ++k) col[k] = B[(size_t)k * q + i];
            if (sv[i] > tiny && add_column(col)) { placed[c] = 1; continue; }
            for (int e = 0; e < p; ++e) { std::vector<double> unit(p, 0.0); unit[e] = 1.0; if (add_column(unit)) break; }
        }
        for (int e = 0; e < p && filled < p; ++e) { std::vector<double> unit(p, 0.0); unit[e] = 1.0; add_column(unit); }
        assert(filled == p);
        // B = P(:, :q) diag(sv) Vs^T with Vs = V(:, ord)
        w.create(q, 1, CV_32F);
        for (int c = 0; c < q; ++c) w.at<float>(c, 0) = (float)sv[ord[c]];
        Mat big(p, p, CV_32F), small(q, q, CV_32F);      // big = P, small = Vs^T
        for (int y = 0; y < p; ++y) for (int x = 0; x < p; ++x) big.at<float>(y, x) = (float)P[(size_t)y * p + x];
        for (int c = 0; c < q; ++c) for (int k = 0; k < q; ++k) small.at<float>(c, k) = (float)V[(size_t)k * q + ord[c]];
        if (!flip) { u = big; vt = small; }                 // A = P S Vs^T
        else { u = Mat(small.t()); vt = Mat(big.t()); }      // A = B^T = Vs S P^T
    }
    static void compute(const Mat& A, Mat& w) { Mat u, vt; compute(A, w, u, vt); }
};
inline void SVDecomp(const Mat& A, Mat& w, Mat& u, Mat& vt, int flags = 0) { SVD::compute(A, w, u, vt, flags); }

}  // namespace cv
