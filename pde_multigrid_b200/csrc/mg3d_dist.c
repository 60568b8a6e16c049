/*
 * mg3d_dist.c -- multi-GPU (z-slab) entry points of the 3D driver.  Placeholder until the slab
 * partition lands: the calls fail loudly instead of silently running on one GPU.
 */
#include "mg_host_common.h"

int mg_comm_unique_id(void* out128)
{
    (void)out128;
    return mg_fail(MG_ERR_COMM, "multi-GPU slabs are not built into this version of libmg_b200");
}

int mg3d_create_dist(mg3d_t** out, const int finest_size_xyz[3], const double range[6], int dtype, int residual_mode,
                     int rank, int nranks, const void* nccl_unique_id)
{
    (void)nccl_unique_id; (void)rank;
    if (nranks == 1) return mg3d_create(out, finest_size_xyz, range, dtype, residual_mode);
    return mg_fail(MG_ERR_COMM, "multi-GPU slabs are not built into this version of libmg_b200");
}
