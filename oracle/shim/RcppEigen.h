// TEST INFRASTRUCTURE: stand-in for <RcppEigen.h> (see shim_eigen.h).
#pragma once
#include "Rcpp.h"
namespace RcppEigen {}
