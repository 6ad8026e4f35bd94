// Coefficient blocks of the DPP bilinear form and dispatch to the kernel family.
//   A = (1/mu) [[k1 K + beta M, -beta M], [-beta M, k2 K + beta M]]      (forms/dpp.py:27,57,89)
#include "dpp_internal.cuh"

namespace dpp {

Coef dpp_coef(const dpp_context* ctx) {
  Coef c{};
  const double bm = ctx->beta / ctx->mu;
  c.cK[0][0] = ctx->k1 / ctx->mu;
  c.cK[1][1] = ctx->k2 / ctx->mu;
  c.cM[0][0] = bm;
  c.cM[0][1] = -bm;
  c.cM[1][0] = -bm;
  c.cM[1][1] = bm;
  return c;
}

Coef block_coef(const dpp_context* ctx, int row, int col) {
  const Coef full = dpp_coef(ctx);
  Coef c{};
  c.cK[0][0] = full.cK[row][col];
  c.cM[0][0] = full.cM[row][col];
  return c;
}

int op_apply(dpp_context* ctx, const OpArgs& a, int* n_partial_blocks) {
  if (ctx->family == DPP_KERNEL_STRUCTURED) return structured_apply(ctx, a, n_partial_blocks);
  return general_apply(ctx, a, n_partial_blocks);
}

int op_diagonal(dpp_context* ctx) {
  if (!ctx->d_diag) DPP_CHECK(dev_alloc(ctx, &ctx->d_diag, 2 * ctx->n_nodes));
  const Coef c = dpp_coef(ctx);
  if (ctx->family == DPP_KERNEL_STRUCTURED) return structured_diagonal(ctx, c, ctx->d_diag);
  return general_diagonal(ctx, c, ctx->d_diag);
}

}  // namespace dpp
