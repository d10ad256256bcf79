// Single-process build of the communication unit: halo exchange / all-reduce entry points
// (replaced by comm.cu once the NCCL path is in).
#include "common.cuh"
void ctl_comm_free(ctl_handle_s *) {}
int ctl_halo_exchange(ctl_handle_s *h, const double *)
{
    ctl_set_error(h, "halo exchange: not available in this build");
    return CTL_ERR_STATE;
}
int ctl_allreduce_sum(ctl_handle_s *h, double *, int)
{
    ctl_set_error(h, "allreduce: not available in this build");
    return CTL_ERR_STATE;
}
extern "C" {
int ctl_comm_unique_id(void *) { return CTL_ERR_STATE; }
int ctl_comm_init(ctl_handle h, const void *)
{
    ctl_set_error(h, "ctl_comm_init: not available in this build");
    return CTL_ERR_STATE;
}
}
