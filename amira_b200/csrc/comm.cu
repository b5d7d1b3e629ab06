// comm.cu -- multi-GPU plumbing of libamira_gmg.so (NCCL over NVLink 5 / NVSwitch).
#include "common.cuh"

extern "C" {

int amira_gmg_nccl_unique_id(void *out_128_bytes) {
    (void)out_128_bytes;
    amira::set_error("multi-GPU support is not built into this library yet");
    return AMIRA_E_STATE;
}

int amira_gmg_comm_init(amira_gmg *h, const void *nccl_unique_id, int rank, int world) {
    (void)h; (void)nccl_unique_id; (void)rank; (void)world;
    amira::set_error("multi-GPU support is not built into this library yet");
    return AMIRA_E_STATE;
}

int amira_gmg_set_shard(amira_gmg *h, int64_t first_read_global, int64_t first_call_global) {
    (void)h; (void)first_read_global; (void)first_call_global;
    amira::set_error("multi-GPU support is not built into this library yet");
    return AMIRA_E_STATE;
}

}  // extern "C"
