// chains.cu — batched independent chains on analytic targets.  Placeholder; filled in next.
#include "common.cuh"
extern "C" {
int pmp_chains_create(pmp_ctx*, int64_t, const float*) { pmp::set_error("chains not built yet"); return PMP_ERR_UNSUPPORTED; }
int pmp_chains_run(pmp_ctx*, int64_t, int) { pmp::set_error("chains not built yet"); return PMP_ERR_UNSUPPORTED; }
int pmp_chains_read_states(pmp_ctx*, float*) { pmp::set_error("chains not built yet"); return PMP_ERR_UNSUPPORTED; }
int pmp_chains_read_samples(pmp_ctx*, float*, int64_t) { pmp::set_error("chains not built yet"); return PMP_ERR_UNSUPPORTED; }
int pmp_chains_run_timed(pmp_ctx*, int64_t, int, float*) { pmp::set_error("chains not built yet"); return PMP_ERR_UNSUPPORTED; }
}
