// query.cu — kNN over the mirrored key tables, hint scoring cascade, candidate replay, GMM-L2 (placeholder until the
// kernels land; the entry points report C2G_ERR_STATE).
#include "../../include/c2g.h"
#include "c2g_ctx.cuh"

int c2g_query_alloc(c2g_ctx *ctx) { (void) ctx; return 0; }
void c2g_query_free(c2g_ctx *ctx) { (void) ctx; }

extern "C" {
int c2g_db_set_layer(c2g_ctx *, int, int, const float *, const int *, const signed char *, const unsigned char *, const float *) { return C2G_ERR_STATE; }
int c2g_query(c2g_ctx *, int, int, const c2g_score_ensemble *, const c2g_score_ensemble *, c2g_query_result *, c2g_hint *, c2g_pair_score *) { return C2G_ERR_STATE; }
int c2g_query_async(c2g_ctx *, int, int, const c2g_score_ensemble *, const c2g_score_ensemble *) { return C2G_ERR_STATE; }
int c2g_query_buffers(c2g_ctx *, void **, void **, void **, long long *) { return C2G_ERR_STATE; }
int c2g_finish_from_scores(c2g_ctx *, int, int, const c2g_score_ensemble *, const void *, const void *, c2g_query_result *) { return C2G_ERR_STATE; }
}
