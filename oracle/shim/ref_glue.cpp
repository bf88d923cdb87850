// ref_glue.cpp -- TEST INFRASTRUCTURE.  C entry points around the reference's OWN sampler functions (compiled from
// /root/reference/src by `make ref`), with R's nmath draws re-routed to the sequential draw source of the oracle
// library so that the C restatement and the real sources consume identical uniform / normal / gamma streams
// (R::rnorm(m, s) = m + s z, R::rgamma(a, scale) = scale g(a), R::rbeta(1,1) = R::runif(0,1) = u).
#include "Rcpp.h"
#include <cstring>
#include <sstream>
extern "C" {
#include "bayesrr_oracle.h"
}

namespace {
struct NullBuf : std::streambuf { int overflow(int c) override { return c; } };
NullBuf g_nullbuf;
std::ostream g_null(&g_nullbuf);
orc_seq g_seq;
double *g_rows = nullptr; long g_max_rows = 0, g_n_rows = 0, g_row_len = 0;
void row_hook(const double *row, long len)
{
    if (len != g_row_len) return;
    if (g_rows && g_n_rows < g_max_rows) std::memcpy(g_rows + g_n_rows * len, row, (size_t)len * 8);
    ++g_n_rows;
}
void begin(uint64_t seed, double *rows, long max_rows, long row_len)
{
    orc_seq_init(&g_seq, seed);          // also srand(seed): std::random_shuffle in the reference uses rand()
    g_rows = rows; g_max_rows = max_rows; g_n_rows = 0; g_row_len = row_len;
    Eigen::shim_row_hook = row_hook;
}
Eigen::Mat mat(const double *p, long r, long c)
{
    Eigen::Mat m(r, c);
    if (p) std::memcpy(m.data(), p, (size_t)(r * c) * 8);
    return m;
}
Eigen::VectorXi ivec(const int32_t *p, long n)
{
    Eigen::VectorXi v(n);
    for (long i = 0; i < n; ++i) v(i) = p[i];
    return v;
}
}  // namespace

namespace Eigen { void (*shim_row_hook)(const double *, long) = nullptr; }
namespace Rcpp { std::ostream &Rcout = g_null; std::ostream &Rcerr = g_null; }
namespace R {
double rgamma(double shape, double scale) { return scale * orc_seq_next_gamma(&g_seq, shape); }
double rnorm(double mu, double sigma) { return mu + sigma * orc_seq_next_normal(&g_seq); }
double rbeta(double a, double b) { (void)a; (void)b; return orc_seq_next_uniform(&g_seq); }   // only rbeta(1,1) is used
double runif(double a, double b) { return a + (b - a) * orc_seq_next_uniform(&g_seq); }
double rexp(double scale) { return -scale * std::log(orc_seq_next_uniform(&g_seq)); }
}

// the reference's entry points (src/BayesRv2.cpp:60, src/BayesRv2Groups.cpp:75, src/BRv2Grstart.cpp:77, src/HorseshoeR.cpp:109)
void BayesRSamplerV2(std::string outputFile, int seed, int max_iterations, int burn_in, int thinning, Eigen::MatrixXd X, Eigen::VectorXd Y,
                     double sigma0, double v0E, double s02E, double v0G, double s02G, Eigen::VectorXd cva);
void BayesRSamplerV2Groups(std::string outputFile, int seed, int max_iterations, int burn_in, int thinning, Eigen::MatrixXd X, Eigen::VectorXd Y,
                           double sigma0, double v0E, double s02E, double v0G, double s02G, Eigen::MatrixXd cva, int groups,
                           Eigen::VectorXi gAssign, Eigen::MatrixXd fixed);
void BRV2Grstart(std::string outputFile, int seed, int max_iterations, int burn_in, int thinning, double mu, Eigen::MatrixXd beta, double sigmaE,
                 Eigen::VectorXd sigmaGG, Eigen::MatrixXd X, Eigen::VectorXd epsilon, Eigen::VectorXd components, double sigma0, double v0E,
                 double s02E, double v0G, double s02G, Eigen::MatrixXd cva, int groups, Eigen::VectorXi gAssign);
void HorseshoeR(std::string outputFile, int seed, int max_iterations, int burn_in, int thinning, Eigen::MatrixXd X, Eigen::VectorXd Y, double A,
                double v0E, double s02E, double vL, double vT, double c2, double vC, double sC);

extern "C" {

long ref_v2(const char *out, uint64_t seed, int max_it, int burn_in, int thinning, const double *X, long N, long M, const double *Y,
            double sigma0, double v0E, double s02E, double v0G, double s02G, const double *cva, int ncva, double *rows, long max_rows)
{
    begin(seed, rows, max_rows, 2 * M + 4 + N);
    BayesRSamplerV2(out, (int)seed, max_it, burn_in, thinning, mat(X, N, M), mat(Y, N, 1), sigma0, v0E, s02E, v0G, s02G, mat(cva, ncva, 1));
    return g_n_rows;
}
long ref_groups(const char *out, uint64_t seed, int max_it, int burn_in, int thinning, const double *X, long N, long M, const double *Y,
                double sigma0, double v0E, double s02E, double v0G, double s02G, const double *cva, int ncva, int groups,
                const int32_t *gAssign, const double *fixed, long F, double *rows, long max_rows)
{
    begin(seed, rows, max_rows, 2 * M + 3 + groups + N + F + 1);
    BayesRSamplerV2Groups(out, (int)seed, max_it, burn_in, thinning, mat(X, N, M), mat(Y, N, 1), sigma0, v0E, s02E, v0G, s02G,
                          mat(cva, groups, ncva), groups, ivec(gAssign, M), mat(fixed, N, F));
    return g_n_rows;
}
long ref_grstart(const char *out, uint64_t seed, int max_it, int burn_in, int thinning, double mu, const double *beta, double sigmaE,
                 const double *sigmaGG, const double *X, long N, long M, const double *epsilon, const double *components,
                 double sigma0, double v0E, double s02E, double v0G, double s02G, const double *cva, int ncva, int groups,
                 const int32_t *gAssign, double *rows, long max_rows)
{
    begin(seed, rows, max_rows, 2 * M + 3 + groups + N);
    BRV2Grstart(out, (int)seed, max_it, burn_in, thinning, mu, mat(beta, M, 1), sigmaE, mat(sigmaGG, groups, 1), mat(X, N, M),
                mat(epsilon, N, 1), mat(components, M, 1), sigma0, v0E, s02E, v0G, s02G, mat(cva, groups, ncva), groups, ivec(gAssign, M));
    return g_n_rows;
}
long ref_horseshoe(const char *out, uint64_t seed, int max_it, int burn_in, int thinning, const double *X, long N, long M, const double *Y,
                   double A, double v0E, double s02E, double vL, double vT, double c2, double vC, double sC, double *rows, long max_rows)
{
    begin(seed, rows, max_rows, 2 * M + 4 + N);
    HorseshoeR(out, (int)seed, max_it, burn_in, thinning, mat(X, N, M), mat(Y, N, 1), A, v0E, s02E, vL, vT, c2, vC, sC);
    return g_n_rows;
}

}  // extern "C"
