// K1+K4, variant 2 (tcgen05 3xTF32) -- placeholder until the tensor-core kernel lands.
#include "gb_common.cuh"
extern "C" int gb_ffae_tc_supported(const gb_ffnet*) {
  gb::set_error("tcgen05 variant not built yet");
  return GB_E_SHAPE;
}
extern "C" int gb_ffae_infer_score_tc(const gb_ffnet*, const float*, const gb_job*, int32_t, int32_t, const float*,
                                      const float*, const float*, const float*, const float*, float*, float*, float*,
                                      float*, float*, float*, float*, void*) {
  gb::set_error("tcgen05 variant not built yet");
  return GB_E_SHAPE;
}
