// esvo_b200 product code -- tracking (RegProblemLM); placeholder until the kernels land.
#include "common.cuh"
namespace esvo {
struct TrackState { int dummy; };
int track_alloc(Ctx* c) { c->trk = nullptr; return ESVO_OK; }
void track_free(Ctx*) {}
}
using namespace esvo;
extern "C" {
ESVO_API int esvo_track_reset(esvo_ctx* c, float*, size_t, const double*, const double*, const uint8_t*) { if (c) c->set_error("tracking not built yet"); return ESVO_ERR_UNSUPPORTED; }
ESVO_API int esvo_track_solve(esvo_ctx* c, int, double*, esvo_lm_stats*) { if (c) c->set_error("tracking not built yet"); return ESVO_ERR_UNSUPPORTED; }
ESVO_API int esvo_track_srand(esvo_ctx*, unsigned) { return ESVO_ERR_UNSUPPORTED; }
ESVO_API int esvo_track_get_negative_ts(esvo_ctx*, double*, double*, double*) { return ESVO_ERR_UNSUPPORTED; }
}
