// DetectionMetrics matching (row (f)-N4): which NMS outputs are true positives / ignored, for every IoU threshold, one launch per
// validation batch instead of the reference's per-image Python loop over (prediction, target) pairs with a dozen small tensor ops
// per pair (detection_utils.py:937-958).  One CTA per image; its predictions, targets and crowd targets live in shared memory;
// warp j runs threshold j's greedy assignment (the thresholds never interact), lanes stride over the targets.
// Latency-bound integer / compare work on a few KB per image -- no roofline to speak of; the point is removing ~1e4 launches and a
// device->host sync per validation batch.  The arithmetic is in detection_match_math.cuh (shared with the CPU test build).
#include "common.cuh"
#include "detection_match_math.cuh"

namespace {

using sgb_match::Best;
using sgb_match::Box;

struct Smem {
  Box* pbox;
  float *parea, *pscore, *pcls;
  int* order;       // used predictions in confidence order
  uint8_t* used;
  Box* tbox;
  float *tarea, *tcls;
  Box* cbox;
  float* ccls;
  uint8_t* taken;   // [n_thresholds][max_targets]
};

__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

__host__ __device__ inline size_t carve(const SgbMatchDesc& d, char* base, Smem* s) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align16(bytes);
    return p;
  };
  Smem t;
  t.pbox = (Box*)take(sizeof(Box) * d.max_preds);
  t.parea = (float*)take(4 * (size_t)d.max_preds);
  t.pscore = (float*)take(4 * (size_t)d.max_preds);
  t.pcls = (float*)take(4 * (size_t)d.max_preds);
  t.order = (int*)take(4 * (size_t)d.max_preds);
  t.used = (uint8_t*)take((size_t)d.max_preds);
  t.tbox = (Box*)take(sizeof(Box) * d.max_targets);
  t.tarea = (float*)take(4 * (size_t)d.max_targets);
  t.tcls = (float*)take(4 * (size_t)d.max_targets);
  t.cbox = (Box*)take(sizeof(Box) * (d.max_crowd > 0 ? d.max_crowd : 1));
  t.ccls = (float*)take(4 * (size_t)(d.max_crowd > 0 ? d.max_crowd : 1));
  t.taken = (uint8_t*)take((size_t)d.n_thresholds * d.max_targets);
  if (s) *s = t;
  return off;
}

__global__ void detection_match_kernel(const SgbMatchDesc d, const float* __restrict__ preds, const int32_t* __restrict__ pred_count,
                                       const float* __restrict__ targets, const int32_t* __restrict__ target_count,
                                       const float* __restrict__ crowd, const int32_t* __restrict__ crowd_count,
                                       const float* __restrict__ thresholds, uint8_t* __restrict__ matched, uint8_t* __restrict__ ignore) {
  extern __shared__ __align__(16) char smem_raw[];
  __shared__ int n_used_s;
  Smem s;
  carve(d, smem_raw, &s);
  const int b = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x, T = d.n_thresholds;
  const int P = min(max(pred_count[b], 0), d.max_preds);
  const int M = min(max(target_count[b], 0), d.max_targets);
  const int C = d.max_crowd > 0 ? min(max(crowd_count[b], 0), d.max_crowd) : 0;
  const float* pr = preds + (int64_t)b * d.max_preds * 6;
  uint8_t* mt = matched + (int64_t)b * d.max_preds * T;
  uint8_t* ig = ignore + (int64_t)b * d.max_preds * T;

  for (int i = tid; i < P; i += nthr) {
    const float* r = pr + i * 6;
    const Box bx = sgb_match::clip_box(Box{r[0], r[1], r[2], r[3]}, d.height, d.width);
    s.pbox[i] = bx;
    s.parea[i] = sgb_match::area(bx);
    s.pscore[i] = r[4];
    s.pcls[i] = r[5];
  }
  for (int i = tid; i < M; i += nthr) {
    const float* r = targets + ((int64_t)b * d.max_targets + i) * 5;
    const Box bx = sgb_match::target_xyxy(r[1], r[2], r[3], r[4], d.denormalize_targets != 0, d.height, d.width);
    s.tbox[i] = bx;
    s.tarea[i] = sgb_match::area(bx);
    s.tcls[i] = r[0];
  }
  for (int i = tid; i < C; i += nthr) {
    const float* r = crowd + ((int64_t)b * d.max_crowd + i) * 5;
    s.cbox[i] = sgb_match::target_xyxy(r[1], r[2], r[3], r[4], d.denormalize_targets != 0, d.height, d.width);
    s.ccls[i] = r[0];
  }
  for (int i = tid; i < T * d.max_targets; i += nthr) s.taken[i] = 0;
  if (tid == 0) n_used_s = 0;
  __syncthreads();

  // get_top_k_idx_per_cls: non-zero score and fewer than top_k same-class predictions ahead in confidence order
  for (int i = tid; i < P; i += nthr) {
    const float sc = s.pscore[i], cl = s.pcls[i];
    int rank = 0;
    for (int j = 0; j < P; ++j) rank += (s.pcls[j] == cl && sgb_match::before(s.pscore[j], j, sc, i)) ? 1 : 0;
    s.used[i] = (rank < d.top_k && sc != 0.f) ? 1 : 0;
  }
  __syncthreads();
  for (int i = tid; i < P; i += nthr) {
    const uint8_t u = s.used[i];
    if (u) {
      const float sc = s.pscore[i];
      int pos = 0;
      for (int j = 0; j < P; ++j) pos += (s.used[j] && sgb_match::before(s.pscore[j], j, sc, i)) ? 1 : 0;
      s.order[pos] = i;
      atomicAdd(&n_used_s, 1);
    }
    for (int j = 0; j < T; ++j) {
      mt[i * T + j] = 0;
      ig[i * T + j] = u ? 0 : 1;
    }
  }
  for (int i = P * T + tid; i < d.max_preds * T; i += nthr) {
    mt[i] = 0;
    ig[i] = 0;
  }
  __syncthreads();
  const int n_used = n_used_s;

  // IoUMatching.compute_targets: warp j owns threshold j
  const int warp = tid >> 5, lane = tid & 31, n_warps = nthr >> 5;
  if (M > 0) {
    for (int j = warp; j < T; j += n_warps) {
      const float thr = thresholds[j];
      uint8_t* taken = s.taken + (size_t)j * d.max_targets;
      for (int k = 0; k < n_used; ++k) {
        const int p = s.order[k];
        Best best = sgb_match::best_free_target(s.pbox[p], s.parea[p], s.pcls[p], thr, s.tbox, s.tarea, s.tcls, taken, M, lane, 32);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          Best other;
          other.v = __shfl_xor_sync(0xffffffffu, best.v, o);
          other.t = __shfl_xor_sync(0xffffffffu, best.t, o);
          best = sgb_match::better(best, other);
        }
        if (best.t >= 0 && lane == 0) {
          taken[best.t] = 1;
          mt[p * T + j] = 1;
        }
        __syncwarp();
      }
    }
  }
  __syncthreads();

  // IoUMatching.compute_crowd_targets
  if (C > 0) {
    for (int k = tid; k < n_used; k += nthr) {
      const int p = s.order[k];
      const float best = sgb_match::best_crowd_ioa(s.pbox[p], s.parea[p], s.pcls[p], s.cbox, s.ccls, C);
      for (int j = 0; j < T; ++j)
        if (best > thresholds[j]) ig[p * T + j] = 1;
    }
  }
}

}  // namespace

extern "C" int sgb_detection_matching(const SgbMatchDesc* d, const float* preds, const int32_t* pred_count, const float* targets,
                                      const int32_t* target_count, const float* crowd, const int32_t* crowd_count,
                                      const float* thresholds, uint8_t* matched, uint8_t* ignore, void* stream) {
  SGB_REQUIRE(d && preds && pred_count && targets && target_count && thresholds && matched && ignore, "null pointer");
  SGB_REQUIRE(d->B > 0 && d->max_preds > 0 && d->max_targets > 0 && d->max_crowd >= 0, "bad shape");
  SGB_REQUIRE(d->n_thresholds > 0 && d->n_thresholds <= SGB_MATCH_MAX_THRESHOLDS, "1..32 IoU thresholds");
  SGB_REQUIRE(d->max_crowd == 0 || (crowd && crowd_count), "crowd targets missing");
  SGB_REQUIRE(d->top_k > 0, "top_k");
  const size_t bytes = carve(*d, nullptr, nullptr);
  SGB_REQUIRE(bytes <= 200 * 1024, "predictions + targets of one image exceed shared memory");
  if (bytes > 48 * 1024) cudaFuncSetAttribute(detection_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  int warps = d->n_thresholds < 4 ? 4 : d->n_thresholds;
  detection_match_kernel<<<d->B, warps * 32, bytes, (cudaStream_t)stream>>>(*d, preds, pred_count, targets, target_count, crowd, crowd_count,
                                                                            thresholds, matched, ignore);
  SGB_LAUNCH_CHECK("detection_match_kernel");
  return SGB_OK;
}
