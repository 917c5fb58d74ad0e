// bk_eigs.cu -- S10 shift-invert Arnoldi (placeholder: filled in below)
#include "bk_common.cuh"
extern "C" int32_t bk_eigs_shift_invert(bk_ctx* c, double sigma, int32_t nev, int32_t krylovdim, double tol, int32_t maxrestart,
                             const bk_gmres_opts* inner, const double* v0, double* vals_re, double* vals_im, double* vecs,
                             int32_t* nconv, int32_t* nops) {
  if (!c) return BK_ERR_ARG;
  return bk_fail(c, BK_ERR_STATE, "eigensolver not implemented yet", __FILE__, __LINE__);
}
