// Test infrastructure: serial HOST driver around super_gradients_b200/csrc/preprocess_math.cuh (the arithmetic of the CUDA
// pre-processing kernel), compiled with g++ by tests/host_preprocess.py.  Output: bf16 bit patterns of the NHWC slot.
#include <cstring>

#include "preprocess_math.cuh"

static inline uint16_t f32_to_bf16_rn(float f) {  // round to nearest even, as __float2bfloat16_rn (no NaNs here)
  uint32_t u;
  std::memcpy(&u, &f, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

extern "C" int preprocess_host(const SgbPreprocDesc* d, const uint8_t* src, uint16_t* out) {
  for (int oy = 0; oy < d->out_h; ++oy)
    for (int ox = 0; ox < d->out_w; ++ox) {
      uint16_t* o = out + ((int64_t)oy * d->out_w + ox) * d->out_pitch;
      for (int c = 0; c < d->out_pitch; ++c) o[c] = f32_to_bf16_rn(c < d->src_c ? sgb_prep::sample_pixel(*d, src, oy, ox, c) : 0.f);
    }
  return 0;
}
