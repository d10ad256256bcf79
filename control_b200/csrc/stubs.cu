// Entry points not implemented yet (replaced unit by unit).
#include "common.cuh"
#define NOTIMPL(name) { ctl_set_error(h, name ": not implemented in this build"); return CTL_ERR_STATE; }
void ctl_pc_free(ctl_handle_s *) {}
void ctl_krylov_free(ctl_handle_s *) {}
void ctl_comm_free(ctl_handle_s *) {}
int ctl_pc_invalidate(ctl_handle_s *) { return CTL_OK; }
int ctl_halo_exchange(ctl_handle_s *h, const double *) NOTIMPL("halo exchange")
int ctl_allreduce_sum(ctl_handle_s *h, double *, int) NOTIMPL("allreduce")
extern "C" {
int ctl_pc_default_options(ctl_pc_options *) { return CTL_ERR_STATE; }
int ctl_pc_setup(ctl_handle h, const ctl_pc_options *) NOTIMPL("ctl_pc_setup")
int ctl_pc_apply(ctl_handle h, const double *, double *, int) NOTIMPL("ctl_pc_apply")
int ctl_pc_fn(ctl_handle h, const double *, double *, int) NOTIMPL("ctl_pc_fn")
int ctl_set_pc_callback(ctl_handle h, ctl_pc_callback, void *) NOTIMPL("ctl_set_pc_callback")
int ctl_krylov_default_options(ctl_krylov_options *) { return CTL_ERR_STATE; }
int ctl_solve(ctl_handle h, const double *, double *, int, const ctl_krylov_options *, ctl_solve_result *) NOTIMPL("ctl_solve")
int ctl_solve_host(ctl_handle h, const double *, double *, const ctl_krylov_options *, ctl_solve_result *) NOTIMPL("ctl_solve_host")
int ctl_kkt_residual_norm(ctl_handle h, const double *, const double *, int, double *) NOTIMPL("ctl_kkt_residual_norm")
int ctl_objective_host(ctl_handle h, const double *, const double *, const double *, double *) NOTIMPL("ctl_objective_host")
int32_t ctl_amg_num_hierarchies(ctl_handle) { return 0; }
int32_t ctl_amg_num_levels(ctl_handle, int32_t) { return 0; }
int ctl_amg_level_size(ctl_handle h, int32_t, int32_t, int32_t *, int64_t *, int64_t *) NOTIMPL("ctl_amg_level_size")
int ctl_amg_get_csr(ctl_handle h, int32_t, int32_t, int, int32_t *, int32_t *, double *) NOTIMPL("ctl_amg_get_csr")
int ctl_amg_get_aggregates(ctl_handle h, int32_t, int32_t, int32_t *) NOTIMPL("ctl_amg_get_aggregates")
int ctl_amg_solve(ctl_handle h, int32_t, const double *, double *) NOTIMPL("ctl_amg_solve")
int ctl_comm_unique_id(void *) { return CTL_ERR_STATE; }
int ctl_comm_init(ctl_handle h, const void *) NOTIMPL("ctl_comm_init")
}
