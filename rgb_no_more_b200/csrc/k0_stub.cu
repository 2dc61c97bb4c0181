#include "../../include/rgbnm_b200.h"
extern "C" {
int rgbnm_k0_dcstats(const int16_t*, const int16_t*, const int16_t*, const rgbnm_plan*, const rgbnm_k0_tables*, float*, int, int, int, void*) { return RGBNM_ERR_UNSUPPORTED; }
int rgbnm_k0_fused(const int16_t*, const int16_t*, const int16_t*, const rgbnm_plan*, const rgbnm_k0_tables*, const float*, void*, int, int, int, int, void*) { return RGBNM_ERR_UNSUPPORTED; }
int rgbnm_k0_launch_count(void) { return 0; }
}
