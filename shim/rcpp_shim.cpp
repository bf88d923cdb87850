// rcpp_shim.cpp -- the Rcpp side of the drop-in: the reference's four C++ entry points with their EXACT signatures
// (reference src/BayesRv2.cpp:60, src/BayesRv2Groups.cpp:75, src/BRv2Grstart.cpp:77, src/HorseshoeR.cpp:109; the generated
// glue src/RcppExports.cpp:10,39,61,86 calls them by name and R/RcppExports.R:25,49,70,74 reaches that glue through .Call), each
// body one call into the C ABI of include/bayesrr_b200.h.  A maintainer drops this file into the package's src/ in place of
// BayesRv2.cpp, BayesRv2Groups.cpp, BRv2Grstart.cpp and HorseshoeR.cpp; RcppExports.cpp / RcppExports.R / NAMESPACE stay as
// they are, so the package keeps exporting _BayesRRcpp_BayesRSamplerV2, _BayesRRcpp_BayesRSamplerV2Groups,
// _BayesRRcpp_BRV2Grstart, _BayesRRcpp_HorseshoeR and R_init_BayesRRcpp (src/RcppExports.cpp:110-121).
//
// R, Rcpp and Eigen are not in this build image: tests/test_abi.py compiles this file against the minimal Rcpp / Eigen stand-ins
// of the oracle's reference build, type-checks the four signatures against the reference's and links it to libbayesrr_b200.so.
// [[Rcpp::depends(RcppEigen)]]
#include <RcppEigen.h>
#include <string>
#include "bayesrr_b200.h"

namespace {
// the reference prints its validation errors and returns normally; nothing is thrown at R (src/BayesRv2.cpp:76-95)
void report(int rc) { if (rc != BRR_OK) Rcpp::Rcerr << brr_last_error() << "\n"; }
// "iteration: <n>" / "duration: <s>s" (src/BayesRv2.cpp:173-175,276-278) go where the reference sends them
void to_rcout(void *, const char *text) { Rcpp::Rcout << text; }
struct Messages { Messages() { brr_set_message_handler(to_rcout, nullptr); } ~Messages() { brr_set_message_handler(nullptr, nullptr); } };
}  // namespace

// [[Rcpp::export]]
void BayesRSamplerV2(std::string outputFile, int seed, int max_iterations, int burn_in, int thinning,
                     Eigen::MatrixXd X, Eigen::VectorXd Y, double sigma0, double v0E, double s02E,
                     double v0G, double s02G, Eigen::VectorXd cva)
{
    Messages m;
    report(brr_BayesRSamplerV2(outputFile.c_str(), seed, max_iterations, burn_in, thinning,
                               X.data(), X.rows(), X.cols(), Y.data(),
                               sigma0, v0E, s02E, v0G, s02G, cva.data(), (int)cva.size()));
}

// [[Rcpp::export]]
void BayesRSamplerV2Groups(std::string outputFile, int seed, int max_iterations, int burn_in, int thinning,
                           Eigen::MatrixXd X, Eigen::VectorXd Y, double sigma0, double v0E, double s02E,
                           double v0G, double s02G, Eigen::MatrixXd cva, int groups,
                           Eigen::VectorXi gAssign, Eigen::MatrixXd fixed)
{
    Messages m;
    report(brr_BayesRSamplerV2Groups(outputFile.c_str(), seed, max_iterations, burn_in, thinning,
                                     X.data(), X.rows(), X.cols(), Y.data(),
                                     sigma0, v0E, s02E, v0G, s02G,
                                     cva.data(), (int)cva.cols(), groups, gAssign.data(),   // cva: groups x (K-1), column-major
                                     fixed.data(), fixed.cols()));
}

// [[Rcpp::export]]
void BRV2Grstart(std::string outputFile, int seed, int max_iterations, int burn_in, int thinning,
                 double mu, Eigen::MatrixXd beta, double sigmaE, Eigen::VectorXd sigmaGG,
                 Eigen::MatrixXd X, Eigen::VectorXd epsilon, Eigen::VectorXd components,
                 double sigma0, double v0E, double s02E, double v0G, double s02G,
                 Eigen::MatrixXd cva, int groups, Eigen::VectorXi gAssign)
{
    Messages m;
    report(brr_BRV2Grstart(outputFile.c_str(), seed, max_iterations, burn_in, thinning,
                           mu, beta.data(), sigmaE, sigmaGG.data(),
                           X.data(), X.rows(), X.cols(), epsilon.data(), components.data(),
                           sigma0, v0E, s02E, v0G, s02G, cva.data(), (int)cva.cols(), groups, gAssign.data()));
}

// [[Rcpp::export]]
void HorseshoeR(std::string outputFile, int seed, int max_iterations, int burn_in, int thinning,
                Eigen::MatrixXd X, Eigen::VectorXd Y, double A, double v0E, double s02E,
                double vL, double vT, double c2, double vC, double sC)
{
    Messages m;
    report(brr_HorseshoeR(outputFile.c_str(), seed, max_iterations, burn_in, thinning,
                          X.data(), X.rows(), X.cols(), Y.data(), A, v0E, s02E, vL, vT, c2, vC, sC));
}
