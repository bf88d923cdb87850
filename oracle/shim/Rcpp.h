// TEST INFRASTRUCTURE: stand-in for <Rcpp.h> (see shim_eigen.h).  Only what the reference's sources touch.
#pragma once
#include "shim_eigen.h"
namespace Rcpp {
extern std::ostream &Rcout;
extern std::ostream &Rcerr;
}
// R nmath entry points used by src/distributions.cpp, re-routed to the shared sequential draw source (ref_glue.cpp)
namespace R {
double rgamma(double shape, double scale);
double rnorm(double mu, double sigma);
double rbeta(double a, double b);
double runif(double a, double b);
double rexp(double scale);
}
