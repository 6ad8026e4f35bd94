#include "dpp_internal.cuh"
namespace dpp {
int structured_apply_q2(dpp_context* ctx, const OpArgs&, int*) { ctx->set_error("structured Q2: not built yet"); return DPP_ERR_INVALID; }
}
