// fc_sweep.cu — FC model (PMP_FC.py:21-44) log-target sweep.  Placeholder until the tcgen05 path lands:
// the entry points exist so the ABI is complete, and fail loudly (no fallback).
#include "common.cuh"

extern "C" {
int pmp_fc_destroy(pmp_ctx*) { return PMP_OK; }
int pmp_fc_loglik(pmp_ctx*) { pmp::set_error("PMP_TARGET_FC sweep not built yet"); return PMP_ERR_UNSUPPORTED; }
int pmp_set_data_fc(pmp_ctx*, const float*, const int64_t*, int64_t, int64_t, int64_t) {
    pmp::set_error("PMP_TARGET_FC sweep not built yet"); return PMP_ERR_UNSUPPORTED;
}
}
