#include "dpp_internal.cuh"
namespace dpp {
int csr_assemble(dpp_context* ctx, int64_t*) { ctx->set_error("csr: not built yet"); return DPP_ERR_INVALID; }
int csr_export(dpp_context* ctx, int64_t*, int32_t*, double*) { ctx->set_error("csr: not built yet"); return DPP_ERR_INVALID; }
int csr_spmv(dpp_context* ctx, const double*, double*, double*, int*) { ctx->set_error("csr: not built yet"); return DPP_ERR_INVALID; }
void csr_destroy(dpp_context*) {}
}
