// temporary stubs until append / nlml / nccl land
#include "../../include/abo.h"
#include "abo_internal.h"
void abo_nccl_teardown(abo_ctx*) {}
extern "C" int32_t abo_gp_append(abo_gp*, const double*, const double*, int64_t*) { return abo_fail(1, "not implemented"); }
extern "C" int32_t abo_nlml_batch(abo_gp*, const double*, const double*, int64_t, const double*, int64_t, double*, double*, int32_t*) { return abo_fail(1, "not implemented"); }
extern "C" int32_t abo_nccl_unique_id(uint8_t*) { return abo_fail(1, "not implemented"); }
extern "C" int32_t abo_ctx_init_rank(abo_ctx*, int32_t, int32_t, const uint8_t*) { return abo_fail(1, "not implemented"); }
extern "C" int32_t abo_gp_sync(abo_gp*, int32_t) { return abo_fail(1, "not implemented"); }
extern "C" int32_t abo_topk_allgather(abo_ctx*, int64_t, int64_t, int64_t*, double*, int64_t*) { return abo_fail(1, "not implemented"); }
