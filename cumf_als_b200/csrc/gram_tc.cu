// gram_tc.cu -- fused TMA-gather + tcgen05 Gram + in-register CG (placeholder until the
// kernel lands: the planner then never selects this path).
#include "common.cuh"

namespace cumf {
bool tc_path_supports(int) { return false; }
struct TcWork {};
int tc_plan_create(TcWork** out, const std::vector<Chunk>&, const std::vector<SplitRow>&, int, int) {
    *out = nullptr;
    set_last_error("fused tcgen05 path not built");
    return CUMF_EUNSUPPORTED;
}
void tc_plan_destroy(TcWork*) {}
int tc_update_factor(TcWork*, const Chunk*, int, const int*, const float*, const float*, float*, int, float, float,
                     float*, float*, cudaStream_t, int*) {
    set_last_error("fused tcgen05 path not built");
    return CUMF_EUNSUPPORTED;
}
}  // namespace cumf
