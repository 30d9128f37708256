// Batched, class-aware NMS: one CTA (1024 threads) per image does
//   threshold -> (radix-select top-k if needed) -> ordered compaction -> bitonic sort -> IoU bit-matrix -> greedy sweep
// entirely in shared memory, reproducing index-for-index what the reference does per image in Python:
//   PPYoloEPostPredictionCallback.forward  (training/models/detection_models/pp_yolo_e/post_prediction_callback.py:42-98)
//   YoloNASPosePostPredictionCallback      (…/yolo_nas_pose/yolo_nas_pose_post_prediction_callback.py:38-94)
//   torchvision.ops.boxes.batched_nms / nms (torchvision 0.26: ops/boxes.py:43-120 and csrc/ops/cpu/nms_kernel.cpp)
//
// Bit-exactness notes (all mirrored here):
//   * candidates are enumerated row-major (anchor, class), `score > thr` in fp32 (multi-label) or `>=` (single-label);
//   * if there are more than top_k candidates, torch.topk(sorted) picks them (ties: we take the lowest index);
//   * torchvision sorts by score with a STABLE descending sort, so ties keep candidate-list order;
//   * batched_nms uses the coordinate trick when 4*n <= 4000 on CPU, otherwise per-class NMS on the raw boxes;
//   * IoU arithmetic is fp32 with no fused multiply-add, and `ovr > iou_threshold` is evaluated in double.
#include "common.cuh"

namespace {

constexpr int NT = 1024;      // threads per CTA
constexpr int KMAX = 1024;    // max candidates entering NMS
constexpr int NBIN = 2048;

struct NmsSmem {
  unsigned long long mask[KMAX * (KMAX / 64)];  // 128 KB
  unsigned long long keys[KMAX];                // sort keys
  int flat[KMAX];                               // candidate flat index by candidate position
  float bx[4][KMAX];                            // (offset) boxes in sorted order
  float area[KMAX];
  int label[KMAX];
  int sflat[KMAX];                              // flat index in sorted order
  int hist[NBIN];
  int wcnt[32], wtie[32];
  int kept[KMAX];
  int misc[16];
  float fmisc[4];
};

__device__ __forceinline__ unsigned okey(float f) {
  unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ bool passes(float s, float thr, int incl) { return incl ? (s >= thr) : (s > thr); }

// argmax pre-pass for single-label mode (torch.max(dim=1): first maximal index)
__global__ void nms_argmax_kernel(const float* __restrict__ scores, int64_t rows, int C, float* conf, int* lab) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < rows; i += (int64_t)gridDim.x * blockDim.x) {
    float best = scores[i * C];
    int bi = 0;
    for (int c = 1; c < C; ++c) {
      float v = scores[i * C + c];
      if (v > best) {
        best = v;
        bi = c;
      }
    }
    conf[i] = best;
    lab[i] = bi;
  }
}

// histogram of `shift`-ed key digits over this warp's segment, restricted to keys whose higher bits equal `prefix`
template <int BITS>
__device__ void hist_pass(const float* __restrict__ sc, int seg0, int seg1, float thr, int incl, unsigned prefix_mask,
                          unsigned prefix, int shift, int* hist, int lane) {
  for (int base = seg0; base < seg1; base += 128) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int idx = base + u * 32 + lane;
      v[u] = idx < seg1 ? sc[idx] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int idx = base + u * 32 + lane;
      if (idx < seg1 && passes(v[u], thr, incl)) {
        unsigned k = okey(v[u]);
        if ((k & prefix_mask) == prefix) atomicAdd(&hist[(k >> shift) & ((1u << BITS) - 1)], 1);
      }
    }
  }
}

// warp 0: find the bin holding the `need`-th largest element; returns bin, writes remaining need
__device__ int find_bin(const int* hist, int nbins, int need, int lane, int* need_out) {
  const int per = nbins / 32;
  int s = 0;
  for (int j = 0; j < per; ++j) s += hist[lane * per + j];
  // suffix sum over lanes (lane 31 holds the largest keys)
  int suf = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_down_sync(0xffffffffu, suf, o);
    if (lane + o < 32) suf += t;
  }
  int above = suf - s;  // elements in lanes above mine
  bool mine = above < need && suf >= need;
  unsigned ball = __ballot_sync(0xffffffffu, mine);
  int owner = 31 - __clz(ball);  // exactly one lane satisfies it when total >= need
  int bin = 0, rem = 0;
  if (lane == owner) {
    int cum = above;
    for (int j = per - 1; j >= 0; --j) {
      int h = hist[lane * per + j];
      if (cum + h >= need) {
        bin = lane * per + j;
        rem = need - cum;
        break;
      }
      cum += h;
    }
  }
  bin = __shfl_sync(0xffffffffu, bin, owner);
  rem = __shfl_sync(0xffffffffu, rem, owner);
  *need_out = rem;
  return bin;
}

__global__ void __launch_bounds__(NT, 1) nms_kernel(SgbNmsDesc d, const float* __restrict__ boxes,
                                                    const float* __restrict__ scores, const float* __restrict__ conf,
                                                    const int* __restrict__ lab, float* out, int* out_idx,
                                                    int* out_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  NmsSmem& S = *reinterpret_cast<NmsSmem*>(smem_raw);
  const int b = blockIdx.x;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const bool multi = d.multi_label != 0;
  const int n_items = multi ? d.L * d.ncls : d.L;
  const float* sc = multi ? scores + (int64_t)b * n_items : conf + (int64_t)b * d.L;
  const float thr = d.score_thr;
  const int incl = d.thr_inclusive;
  // contiguous segment per warp (multiple of 128 items so that the unrolled loop stays in order)
  int seg_len = (n_items + 31) / 32;
  seg_len = ((seg_len + 127) / 128) * 128;
  const int seg0 = min(warp * seg_len, n_items), seg1 = min(seg0 + seg_len, n_items);

  // ---- pass 1: level-1 histogram (top 11 key bits) + per-warp candidate counts
  for (int i = t; i < NBIN; i += NT) S.hist[i] = 0;
  __syncthreads();
  hist_pass<11>(sc, seg0, seg1, thr, incl, 0u, 0u, 21, S.hist, lane);
  __syncthreads();
  if (warp == 0) {
    int s = 0;
    for (int j = 0; j < NBIN / 32; ++j) s += S.hist[lane * (NBIN / 32) + j];
    int tot = s;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == 0) S.misc[0] = tot;
  }
  __syncthreads();
  const int total = S.misc[0];
  const int top_k = d.top_k < KMAX ? d.top_k : KMAX;
  unsigned Tkey = 0;  // select keys > Tkey, plus `tie_need` keys == Tkey (lowest index first)
  int tie_need = 0;
  bool select_all = total <= top_k;
  if (!select_all) {
    // 3-level radix select of the top_k-th largest key
    if (warp == 0) {
      int need;
      int b1 = find_bin(S.hist, NBIN, top_k, lane, &need);
      if (lane == 0) {
        S.misc[1] = b1;
        S.misc[2] = need;
      }
    }
    __syncthreads();
    unsigned prefix = (unsigned)S.misc[1] << 21;
    int need = S.misc[2];
    for (int i = t; i < NBIN; i += NT) S.hist[i] = 0;
    __syncthreads();
    hist_pass<11>(sc, seg0, seg1, thr, incl, 0xffe00000u, prefix, 10, S.hist, lane);
    __syncthreads();
    if (warp == 0) {
      int need2;
      int b2 = find_bin(S.hist, NBIN, need, lane, &need2);
      if (lane == 0) {
        S.misc[1] = b2;
        S.misc[2] = need2;
      }
    }
    __syncthreads();
    prefix |= (unsigned)S.misc[1] << 10;
    need = S.misc[2];
    for (int i = t; i < NBIN; i += NT) S.hist[i] = 0;
    __syncthreads();
    hist_pass<10>(sc, seg0, seg1, thr, incl, 0xfffffc00u, prefix, 0, S.hist, lane);
    __syncthreads();
    if (warp == 0) {
      int need3;
      int b3 = find_bin(S.hist, 1024, need, lane, &need3);
      if (lane == 0) {
        S.misc[1] = b3;
        S.misc[2] = need3;
      }
    }
    __syncthreads();
    Tkey = prefix | (unsigned)S.misc[1];
    tie_need = S.misc[2];
  }

  // ---- count pass: per-warp (# selected strictly above, # ties) in index order
  {
    int cg = 0, ct = 0;
    for (int base = seg0; base < seg1; base += 128) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int idx = base + u * 32 + lane;
        if (idx < seg1) {
          float v = sc[idx];
          if (passes(v, thr, incl)) {
            unsigned k = okey(v);
            if (select_all || k > Tkey) ++cg;
            else if (k == Tkey) ++ct;
          }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cg += __shfl_xor_sync(0xffffffffu, cg, o);
      ct += __shfl_xor_sync(0xffffffffu, ct, o);
    }
    if (lane == 0) {
      S.wcnt[warp] = cg;
      S.wtie[warp] = ct;
    }
  }
  __syncthreads();
  if (t == 0) {
    // exclusive scan over warps; ties are granted in index order until tie_need is exhausted
    int pos = 0, ties_before = 0;
    for (int w = 0; w < 32; ++w) {
      int cg = S.wcnt[w], ct = S.wtie[w];
      int grant = 0;
      if (!select_all) {
        int left = tie_need - ties_before;
        grant = left > 0 ? (ct < left ? ct : left) : 0;
      }
      S.wcnt[w] = pos;           // output base of this warp
      S.wtie[w] = ties_before;   // ties preceding this warp
      pos += cg + grant;
      ties_before += ct;
    }
    S.misc[3] = pos;  // number of selected candidates
  }
  __syncthreads();
  const int nsel = S.misc[3];

  // ---- write pass: ordered compaction into S.flat / S.keys
  {
    int pos = S.wcnt[warp];
    int ties_seen = S.wtie[warp];
    for (int base = seg0; base < seg1; base += 128) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int idx = base + u * 32 + lane;
        bool valid = false, tie = false;
        unsigned k = 0;
        if (idx < seg1) {
          float v = sc[idx];
          if (passes(v, thr, incl)) {
            k = okey(v);
            if (select_all || k > Tkey) valid = true;
            else if (k == Tkey) tie = true;
          }
        }
        unsigned tb = __ballot_sync(0xffffffffu, tie);
        if (tie) {
          int rank = ties_seen + __popc(tb & ((1u << lane) - 1));
          if (rank < tie_need) valid = true;
        }
        ties_seen += __popc(tb);
        unsigned vb = __ballot_sync(0xffffffffu, valid);
        if (valid) {
          int p = pos + __popc(vb & ((1u << lane) - 1));
          if (p < KMAX) {
            S.flat[p] = idx;
            S.keys[p] = ((unsigned long long)(~k) << 32) | (unsigned)p;
          }
        }
        pos += __popc(vb);
      }
    }
  }
  __syncthreads();

  // ---- bitonic sort of (score desc, position asc)
  int np2 = 1;
  while (np2 < nsel) np2 <<= 1;
  for (int i = nsel + t; i < np2; i += NT) S.keys[i] = ~0ull;
  __syncthreads();
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = t; i < np2; i += NT) {
        int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long a = S.keys[i], c = S.keys[ixj];
          bool up = (i & k) == 0;
          if ((a > c) == up) {
            S.keys[i] = c;
            S.keys[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }

  // ---- gather boxes / labels in sorted order
  float lmax = -INFINITY;
  for (int i = t; i < nsel; i += NT) {
    int p = (int)(S.keys[i] & 0xffffffffu);
    int f = S.flat[p];
    int anchor = multi ? f / d.ncls : f;
    int label = multi ? f - anchor * d.ncls : (lab ? lab[(int64_t)b * d.L + f] : 0);
    const float* bp = boxes + ((int64_t)b * d.L + anchor) * 4;
    float x1 = bp[0], y1 = bp[1], x2 = bp[2], y2 = bp[3];
    S.bx[0][i] = x1;
    S.bx[1][i] = y1;
    S.bx[2][i] = x2;
    S.bx[3][i] = y2;
    S.label[i] = label;
    S.sflat[i] = f;
    lmax = fmaxf(lmax, fmaxf(fmaxf(x1, y1), fmaxf(x2, y2)));
  }
  // batched_nms path selection exactly as torchvision (CPU): coordinate trick iff boxes.numel() <= 4000
  const bool trick = !d.class_agnostic && (4 * nsel <= 4000);
  if (trick) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    float* wmax = reinterpret_cast<float*>(S.hist);
    if (lane == 0) wmax[warp] = lmax;
    __syncthreads();
    if (t == 0) {
      float m = -INFINITY;
      for (int w = 0; w < 32; ++w) m = fmaxf(m, wmax[w]);
      S.fmisc[1] = __fadd_rn(m, 1.0f);  // max_coordinate + 1
    }
    __syncthreads();
    const float step = S.fmisc[1];
    for (int i = t; i < nsel; i += NT) {
      float off = __fmul_rn((float)S.label[i], step);
      S.bx[0][i] = __fadd_rn(S.bx[0][i], off);
      S.bx[1][i] = __fadd_rn(S.bx[1][i], off);
      S.bx[2][i] = __fadd_rn(S.bx[2][i], off);
      S.bx[3][i] = __fadd_rn(S.bx[3][i], off);
    }
  }
  __syncthreads();
  for (int i = t; i < nsel; i += NT)
    S.area[i] = __fmul_rn(__fsub_rn(S.bx[2][i], S.bx[0][i]), __fsub_rn(S.bx[3][i], S.bx[1][i]));
  __syncthreads();

  // ---- IoU bit matrix: mask[i][w] bit j: j > i, suppressed by i
  const int nw = (nsel + 63) / 64;
  const bool same_class_only = !d.class_agnostic && !trick;
  for (int pair = t; pair < nsel * nw; pair += NT) {
    int i = pair / nw, w = pair - i * nw;
    unsigned long long bits = 0ull;
    if (w * 64 + 63 > i) {
      float ix1 = S.bx[0][i], iy1 = S.bx[1][i], ix2 = S.bx[2][i], iy2 = S.bx[3][i], ia = S.area[i];
      int li = S.label[i];
      int j0 = w * 64;
      int jend = min(j0 + 64, nsel);
      for (int j = max(j0, i + 1); j < jend; ++j) {
        if (same_class_only && S.label[j] != li) continue;
        float xx1 = fmaxf(ix1, S.bx[0][j]), yy1 = fmaxf(iy1, S.bx[1][j]);
        float xx2 = fminf(ix2, S.bx[2][j]), yy2 = fminf(iy2, S.bx[3][j]);
        float ww = fmaxf(0.f, __fsub_rn(xx2, xx1)), hh = fmaxf(0.f, __fsub_rn(yy2, yy1));
        float inter = __fmul_rn(ww, hh);
        float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ia, S.area[j]), inter));
        if ((double)ovr > d.iou_thr) bits |= 1ull << (j - j0);
      }
    }
    S.mask[i * (KMAX / 64) + w] = bits;
  }
  __syncthreads();

  // ---- greedy sweep by warp 0: lane w owns word w of the "removed" bitset
  if (warp == 0) {
    unsigned long long remv = 0ull;  // lanes 0..15
    int nkeep = 0;
    for (int i = 0; i < nsel; ++i) {
      unsigned long long wv = __shfl_sync(0xffffffffu, remv, i >> 6);
      if (!((wv >> (i & 63)) & 1ull)) {
        if (lane == 0) S.kept[nkeep] = i;
        ++nkeep;
        if (nkeep >= d.max_out) break;
        if (lane < nw) remv |= S.mask[i * (KMAX / 64) + lane];
      }
    }
    if (lane == 0) S.misc[4] = nkeep;
  }
  __syncthreads();
  const int nkeep = S.misc[4];
  if (t == 0) out_count[b] = nkeep;
  for (int r = t; r < d.max_out; r += NT) {
    float* o = out + ((int64_t)b * d.max_out + r) * 6;
    if (r < nkeep) {
      int i = S.kept[r];
      int f = S.sflat[i];
      int anchor = multi ? f / d.ncls : f;
      const float* bp = boxes + ((int64_t)b * d.L + anchor) * 4;
      o[0] = bp[0];
      o[1] = bp[1];
      o[2] = bp[2];
      o[3] = bp[3];
      o[4] = sc[f];
      o[5] = (float)S.label[i];
      out_idx[(int64_t)b * d.max_out + r] = multi ? f : f * d.ncls + S.label[i];
    } else {
      o[0] = o[1] = o[2] = o[3] = o[4] = o[5] = 0.f;
      out_idx[(int64_t)b * d.max_out + r] = -1;
    }
  }
}

}  // namespace

extern "C" int64_t sgb_nms_workspace_bytes(const SgbNmsDesc* d) {
  if (!d) return 0;
  return (int64_t)d->B * d->L * 8 + 256;
}

extern "C" int sgb_batched_nms(const SgbNmsDesc* d, const float* boxes, const float* scores, float* out,
                               int32_t* out_idx, int32_t* out_count, void* workspace, int64_t workspace_bytes,
                               void* stream) {
  SGB_REQUIRE(d && boxes && scores && out && out_idx && out_count, "null pointer");
  SGB_REQUIRE(d->B > 0 && d->L > 0 && d->ncls > 0, "bad dims");
  SGB_REQUIRE(d->max_out > 0 && d->max_out <= KMAX, "max_out must be in [1, 1024]");
  SGB_REQUIRE(d->top_k > 0, "top_k must be positive");
  if (d->top_k > KMAX) {
    sgb_set_error("sgb_batched_nms: top_k %d > %d is not supported by the shared-memory IoU matrix", d->top_k, KMAX);
    return SGB_E_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* conf = nullptr;
  int* lab = nullptr;
  if (!d->multi_label) {
    SGB_REQUIRE(workspace && workspace_bytes >= sgb_nms_workspace_bytes(d), "workspace too small");
    conf = reinterpret_cast<float*>(workspace);
    lab = reinterpret_cast<int*>(conf + (int64_t)d->B * d->L);
    int64_t rows = (int64_t)d->B * d->L;
    int grid = (int)((rows + 255) / 256 > 148 * 8 ? 148 * 8 : (rows + 255) / 256);
    nms_argmax_kernel<<<grid, 256, 0, st>>>(scores, rows, d->ncls, conf, lab);
    SGB_LAUNCH_CHECK("nms_argmax_kernel");
  }
  static bool attr = false;
  if (!attr) {
    if (int rc = sgb_cuda_check(
            cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NmsSmem)),
            "cudaFuncSetAttribute(nms_kernel)"))
      return rc;
    attr = true;
  }
  nms_kernel<<<d->B, NT, sizeof(NmsSmem), st>>>(*d, boxes, scores, conf, lab, out, out_idx, out_count);
  SGB_LAUNCH_CHECK("nms_kernel");
  return SGB_OK;
}
