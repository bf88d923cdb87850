// Parameters of the post-sweep hyper-parameter kernels (hyper.cu).
#pragma once
#include "sweep.cuh"

namespace brr {

struct HyperParams {
    int kind; int64_t it;            // iteration just swept; its gamma slots are keyed by `it`, the next intercept by it + 1
    double n_total; int64_t M; int K, G; int64_t F;
    double v0E, s02E, v0G, s02G;
    IterScalars *sc;
    double *sigmaG, *pi;             // G, G x K
    const double *vcount, *betaAcum; // from the sweep
    const double *beta, *alpha;
    const double *fin; int nW;       // per-worker sum eps, sum eps^2
    PhiloxKey key;
    const double *tbl_gam;           // replay: this iteration's unit-scale gamma variates by slot, or null
    const double *tbl_gam_next;      // replay: next iteration's slots (horseshoe eta), or null
    const double *tbl_mu_z_next;     // replay: pointer to the next iteration's intercept normal, or null
    // horseshoe
    double A, vL, vT, vC, sC;
    double *lambda, *nu;
    const double *tbl_hs_lam, *tbl_hs_nu_next;
    double *hs_part;                 // ceil(M/256) x 2 partial sums
};

void launch_hyper(const HyperParams &h, cudaStream_t stream);
int hyper_launch_count(int kind);

}  // namespace brr
