// Multi-GPU plumbing of the C ABI: one process per GPU, NCCL over NVLink 5 / NVSwitch.
// The unique id is created on rank 0 and broadcast by the caller (torch.distributed in bench.py).
#include "ctx.cuh"
#include "nccl_dyn.hpp"

#include <cstring>

static_assert(sizeof(ncclUniqueId) <= ACVD_NCCL_ID_BYTES, "ACVD_NCCL_ID_BYTES too small");

extern "C" int acvd_dist_unique_id(void* id_out) {
    if (!id_out) return ACVD_EINVAL;
    if (!nccl().load()) return ACVD_ENCCL;
    ncclUniqueId id;
    if (nccl().GetUniqueId(&id) != ncclSuccess) return ACVD_ENCCL;
    memset(id_out, 0, ACVD_NCCL_ID_BYTES);
    memcpy(id_out, &id, sizeof id);
    return ACVD_OK;
}

extern "C" int acvd_dist_init(acvd_ctx* c, int rank, int world, const void* id_bytes) {
    if (!c || !id_bytes || world < 1 || rank < 0 || rank >= world) return fail(c, ACVD_EINVAL, "acvd_dist_init: bad arguments");
    if (cudaSetDevice(c->device) != cudaSuccess) return fail(c, ACVD_ECUDA, "acvd_dist_init: cudaSetDevice failed");
    if (!nccl().load()) return fail(c, ACVD_ENCCL, "acvd_dist_init: " + nccl().error);
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof id);
    if (c->comm) { nccl().CommDestroy(c->comm); c->comm = nullptr; }
    ncclResult_t r = nccl().CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) return fail(c, ACVD_ENCCL, std::string("ncclCommInitRank: ") + nccl().GetErrorString(r));
    c->rank = rank;
    c->world = world;
    return ACVD_OK;
}
