// tcgen05 in-batch BPR (modes 1 and 2) — placeholder until the tensor-core kernel lands.
#include "common.cuh"

size_t rb_inbatch_tc_workspace_bytes(int B, int D) { (void)B; (void)D; return 0; }

int rb_inbatch_tc(const float* U, const float* I, int B, int D, int mode, float* loss, float* dU, float* dI,
                  float grad_scale, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    (void)U; (void)I; (void)B; (void)D; (void)loss; (void)dU; (void)dI; (void)grad_scale; (void)workspace;
    (void)workspace_bytes; (void)st;
    return rb_set_error(RB200_ERR_INVALID, "bpr_inbatch: mode %d (tcgen05) is not built in this version", mode);
}
