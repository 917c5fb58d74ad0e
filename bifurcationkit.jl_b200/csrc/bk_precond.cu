// bk_precond.cu -- K6 preconditioners (placeholder: filled in below)
#include "bk_common.cuh"
extern "C" int32_t bk_precond_setup(bk_ctx* c, int32_t kind, double a0, double a1) {
  if (!c) return BK_ERR_ARG;
  if (kind == BK_PC_NONE) { c->pc.kind = BK_PC_NONE; return BK_OK; }
  return bk_fail(c, BK_ERR_ARG, "preconditioner kind not implemented", __FILE__, __LINE__);
}
int bk_precond_apply_dev(bk_ctx* c, const double* in, double* out, long long n) {
  return bk_fail(c, BK_ERR_STATE, "no preconditioner set up", __FILE__, __LINE__);
}
extern "C" int32_t bk_precond_apply(bk_ctx* c, const double* in, double* out) {
  if (!c) return BK_ERR_ARG;
  return bk_fail(c, BK_ERR_STATE, "no preconditioner set up", __FILE__, __LINE__);
}
