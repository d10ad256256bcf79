// Device side of the aggregation AMG: hierarchy upload and the V-cycle (amg.cu).
#pragma once
#include "amg_setup.h"
#include "sell.cuh"

struct AmgLevelDev {
    int n = 0;
    double rho = 0.0;
    SellMat A, P, R;
    double *dinv = nullptr;
    double *x = nullptr, *b = nullptr;   // owned on levels >= 1 (level 0 uses the caller's)
    double *r = nullptr, *t0 = nullptr, *t1 = nullptr;  // residual / Chebyshev work vectors
    double *Ainv = nullptr;              // dense inverse on the last level
};

struct AmgHierarchyDev {
    AmgParams params;
    std::vector<AmgLevelHost> host;      // kept for introspection (ctl_amg_get_csr)
    std::vector<AmgLevelDev> dev;
    FusedProgram fused;                  // recorded sub-cycle of levels >= fused_from (0 = none)
    int fused_from = 0;
    int64_t bytes_per_cycle = 0;         // algorithmic bytes of one V-cycle (byte model)
    double *acc_r = nullptr, *acc_z = nullptr, *acc_p = nullptr;   // level-0 work vectors of the accelerated solve
};

// A0 on level 0 may share an existing SELL pattern (fine-level matrices all have the
// mesh pattern); pass nullptr to build one.
int amg_build(ctl_handle_s *h, const HostCSR &A0, const AmgParams &p,
              const std::shared_ptr<SellPattern> &fine_pattern, AmgHierarchyDev &H);
void amg_free(AmgHierarchyDev &H);
// x = `cycles` V-cycles from a zero guess for A x = b (level-0 vectors supplied by the caller)
int amg_solve(ctl_handle_s *h, AmgHierarchyDev &H, const double *b, double *x);
