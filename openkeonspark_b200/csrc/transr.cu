// TransR (TransR.py:16-87): per-relation projection e' = e . M_r — placeholder entry points
// until the grouped-GEMM path lands; they fail loudly rather than fall back.
#include "okb_internal.h"

extern "C" {
int okb_transr_grad_sizes(okb_ctx *c, const okb_model *, INT, INT, INT, INT *, INT *, INT *, INT *) { OKB_FAIL(c, OKB_ERR_ARG, "TransR train path not built yet"); }
int okb_transr_grad(okb_ctx *c, const okb_model *, const okb_hyper *, INT, INT, INT, float *, float *, float *, void *) { OKB_FAIL(c, OKB_ERR_ARG, "TransR train path not built yet"); }
int okb_transr_update(okb_ctx *c, const okb_model *, const okb_hyper *, INT, const float *, const float *, const float *, float *, void *) { OKB_FAIL(c, OKB_ERR_ARG, "TransR train path not built yet"); }
}
