// 3x3 stride-1 "same" convolution on tcgen05 with ONE shared-memory halo tile per output tile.
//
// The im2col kernel (conv_sm100.cu) re-fetches every activation nine times through L2 and pays one mbarrier round trip per
// (tap, channel chunk); for the narrow layers of YOLO-NAS (32..96 channels on 160x160 .. 40x40 maps) both costs dominate.
// Here a CTA owns a 16 x 16 tile of output pixels:
//   * TMA (tiled mode, 4-D box {KC, 18, 18, 1}) brings the 18 x 18 halo of every channel chunk into shared memory ONCE;
//     out-of-image rows / columns are zero-filled by the TMA unit, which is exactly the convolution's zero padding.
//   * the A operand of tap (dh, dw) is the SAME tile read through a matrix descriptor whose start address is advanced by
//     (dh * 18 + dw) pixels and whose 8-row-group stride (SBO) is the halo row pitch: M row i = (i / 8, i % 8) of a
//     16 x 8 sub-tile.  tcgen05 applies the 128B/64B/32B swizzle to absolute shared-memory address bits, so a descriptor
//     may start at any 16-byte aligned row of a TMA-written tile (verified on B200 by tools/umma_probe.cu).
//   * two sub-tiles (columns 0-7 and 8-15) share each B (filter) tile, accumulating into two TMEM column ranges.
//   * the nine filter taps stay resident in shared memory for the CTA's lifetime when they fit, else stream per tap.
// Per tile the producer / MMA / epilogue warps exchange O(1) barrier phases instead of O(taps * chunks); several CTAs are
// co-resident per SM so that one CTA's loads overlap another's MMAs and epilogue.
//
// Serves sgb_conv_fprop and the stride-1 sgb_conv_dgrad (flipped taps over the CRSK filter) for C in {32,48,64,96,128,192}.
// Reference arithmetic replaced: nn.Conv2d(k=3, s=1, p=1) forward / input gradient of modules/qarepvgg_block.py:184-204,
// modules/conv_bn_act_block.py:92-93, training/models/detection_models/yolo_nas/dfl_heads.py:65-83.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "conv_sm100.h"
#include "sm100_host.h"
#include "sm100_ptx.cuh"

namespace sm100 {

namespace {

constexpr int TILE_H = 16, NSUB = 2, TILE_W = 8 * NSUB;
constexpr int HALO_H = TILE_H + 2, HALO_W = TILE_W + 2, HALO_PX = HALO_H * HALO_W;
constexpr int HALO_THREADS = 192;

struct HaloParams {
  int N, H, W;
  int K;   // valid output channels
  int BN;  // GEMM N (K rounded up to 16)
  int tiles_h, tiles_w, total_tiles;
  int tap_b[9];  // column block of the B matrix used by tap (dh, dw) = (t / 3, t % 3)
  int b_cols_per_tap;
  int a_stages;              // halo tiles in flight (1..3)
  int acc_bufs;              // TMEM accumulator sets (1 or 2)
  int b_resident, nb_tiles;  // B ring of (tap, chunk) tiles; resident: nb_tiles == 9 * CHUNKS, loaded once
  uint32_t b_tile_bytes;
  long long y_pitch;
  int y_off;
  int wide_store;  // 32-byte aligned output rows: one 256-bit store per 16 channels
  bf16* y;
  const bf16* residual;  // same geometry as y (dgrad accumulation)
  double* stats;
  int stats_repl;
  int tmem_cols;
  int dbg;
  long long* trace;  // [12][512] clock stamps of CTA 0 when dbg & 16
};

constexpr int MAX_A_STAGES = 3, MAX_B_TILES = 27;

__device__ __forceinline__ uint64_t desc_hi(int row_bytes, uint32_t sbo_bytes) {
  const uint64_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
  return ((uint64_t)1 << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46) | (layout << 61);
}

__device__ __forceinline__ void st_global_256(void* ptr, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// KC: channels per chunk (row of KC bf16 = swizzle span), CHUNKS * KC = C, SNCH: 0 = no statistics (any BN), else BN / 16
// with per-channel sum / sum-of-squares kept in registers across the CTA's tiles.
//
// Pipeline of one CTA (all phases tracked with running stage / parity counters, no divisions):
//   producer : a_empty[s] -> TMA halo tile of the NEXT tile into stage s -> a_full[s];  B tiles once (resident) or per tile
//   MMA warp : acc_empty[b], a_full[s] (, b_full[t]) -> 9 taps x CHUNKS x KC/16 x 2 sub-tiles of UMMAs -> commit
//              a_empty[s] (halo stage reusable) and acc_full[b] (accumulators complete)
//   epilogue : acc_full[b] -> tcgen05.ld -> bf16 -> global (+ statistics) -> acc_empty[b]
template <int KC, int CHUNKS, int SNCH>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const HaloParams p) {
  constexpr int ROWB = KC * 2;
  constexpr int KSTEPS = KC / 16;
  constexpr uint32_t A_REGION = ((uint32_t)(HALO_PX * ROWB) + 1023u) & ~1023u;
  constexpr uint32_t A_BYTES = CHUNKS * A_REGION;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base, b_base = smem_base + (uint32_t)p.a_stages * A_BYTES;
  const uint32_t ctrl = b_base + (uint32_t)p.nb_tiles * p.b_tile_bytes;
  auto a_full = [&](int s) { return ctrl + 8u * s; };
  auto a_empty = [&](int s) { return ctrl + 8u * (MAX_A_STAGES + s); };
  auto acc_full = [&](int b) { return ctrl + 8u * (2 * MAX_A_STAGES + b); };
  auto acc_empty = [&](int b) { return ctrl + 8u * (2 * MAX_A_STAGES + 2 + b); };
  auto b_full = [&](int t) { return ctrl + 8u * (2 * MAX_A_STAGES + 4 + t); };
  auto b_empty = [&](int t) { return ctrl + 8u * (2 * MAX_A_STAGES + 4 + MAX_B_TILES + t); };
  constexpr uint32_t N_BARS = 2 * MAX_A_STAGES + 4 + 2 * MAX_B_TILES;
  const uint32_t tmem_slot = ctrl + 8u * N_BARS;
  float* s_stats = reinterpret_cast<float*>(smem_raw + (tmem_slot + 16u - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_A_STAGES; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), 4);
    }
    for (int t = 0; t < MAX_B_TILES; ++t) {
      mbar_init(b_full(t), 1);
      mbar_init(b_empty(t), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (SNCH > 0)
    for (int i = threadIdx.x; i < 2 * p.BN; i += HALO_THREADS) s_stats[i] = 0.f;
  if (warp == 1) tcgen05_alloc(tmem_slot, p.tmem_cols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const bool resident = p.b_resident != 0;

  if (warp == 0) {
    // ===================================================================================== TMA producer
    if (elect_one()) {
      int astg = 0, bt = 0;
      uint32_t aph = 1, bph = 1;  // parities awaited on the empty barriers: the first pass through each ring is free
      auto load_a = [&](int tile, int tc) {
        const int n = tile / tiles_per_img, rem = tile - n * tiles_per_img;
        const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
        const int h0 = th * TILE_H, w0 = tw * TILE_W;
        const bool tr = (p.dbg & 16) && blockIdx.x == 0 && tc < 512;
        if (tr) p.trace[7 * 512 + tc] = clock64();
        mbar_wait(a_empty(astg), aph);
        if (tr) p.trace[0 * 512 + tc] = clock64();
        mbar_expect_tx(a_full(astg), (uint32_t)(CHUNKS * HALO_PX * ROWB));
        const uint32_t sa = a_base + (uint32_t)astg * A_BYTES;
#pragma unroll
        for (int ck = 0; ck < CHUNKS; ++ck) tma_load_tiled_4d(sa + ck * A_REGION, &map_a, a_full(astg), ck * KC, w0 - 1, h0 - 1, n);
        if (tr) p.trace[1 * 512 + tc] = clock64();
        if (++astg == p.a_stages) {
          astg = 0;
          aph ^= 1;
        }
      };
      int tile = blockIdx.x, tcount = 0;
      if (tile < p.total_tiles) load_a(tile, 0);
      for (; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
        // keep the halo ring ahead of the MMAs: the NEXT tile's halo goes out before this tile's filter tiles stream
        if (p.a_stages > 1 && tile + (int)gridDim.x < p.total_tiles) load_a(tile + gridDim.x, tcount + 1);
        if (!(resident && tcount > 0)) {
          for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
            for (int ck = 0; ck < CHUNKS; ++ck) {
              if (resident) bt = tap * CHUNKS + ck;
              mbar_wait(b_empty(bt), bph);
              mbar_expect_tx(b_full(bt), (uint32_t)(p.BN * ROWB));
              tma_load_2d(b_base + (uint32_t)bt * p.b_tile_bytes, &map_b, b_full(bt), p.tap_b[tap] * p.b_cols_per_tap + ck * KC, 0);
              if (!resident && ++bt == p.nb_tiles) {
                bt = 0;
                bph ^= 1;
              }
            }
          }
        }
        if (p.a_stages == 1 && tile + (int)gridDim.x < p.total_tiles) load_a(tile + gridDim.x, tcount + 1);
      }
    }
  } else if (warp == 1) {
    // ===================================================================================== MMA issuer
    // The whole warp runs this loop converged (uniform registers hold the descriptors); only the elected lane issues.
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t a_hi = desc_hi(ROWB, HALO_W * ROWB), b_hi = desc_hi(ROWB, 8 * ROWB);
    const uint32_t bn = (uint32_t)p.BN, tile_b = p.b_tile_bytes >> 4;
    int tcount = 0, astg = 0, ab = 0, bt = 0;
    uint32_t aph = 0, accph = 1, bph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
      const bool tr0 = (p.dbg & 16) && blockIdx.x == 0 && tcount < 512 && lane == 0;
      if (tr0) p.trace[8 * 512 + tcount] = clock64();
      mbar_wait(acc_empty(ab), accph);
      if (tr0) p.trace[2 * 512 + tcount] = clock64();
      mbar_wait(a_full(astg), aph);
      if (tr0) p.trace[3 * 512 + tcount] = clock64();
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)ab * (NSUB * bn);
      uint32_t a_row = (a_base + (uint32_t)astg * A_BYTES) >> 4;  // descriptor address field of tap (dh, 0)
#pragma unroll 1
      for (int dh = 0; dh < 3; ++dh, a_row += (uint32_t)(HALO_W * ROWB) >> 4) {
#pragma unroll
        for (int dw = 0; dw < 3; ++dw) {
          const uint32_t a_tap = a_row + (uint32_t)(dw * ROWB >> 4);
#pragma unroll
          for (int ck = 0; ck < CHUNKS; ++ck) {
            if (resident) bt = (dh * 3 + dw) * CHUNKS + ck;
            if (!(resident && tcount > 0)) {
              mbar_wait(b_full(bt), resident ? 0u : bph);
              tcgen05_fence_after();
            }
            const uint32_t sb = (b_base >> 4) + (uint32_t)bt * tile_b;
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
              const uint64_t db = b_hi | (uint64_t)((sb + ks * 2) & 0x3fff);
#pragma unroll
              for (int j = 0; j < NSUB; ++j) {
                const uint64_t da = a_hi | (uint64_t)((a_tap + ck * (A_REGION >> 4) + j * (8 * ROWB >> 4) + ks * 2) & 0x3fff);
                if (!(p.dbg & 8)) umma_bf16_if(leader, d_tmem + j * bn, da, db, idesc, (dh | dw | ck | ks) != 0);
              }
            }
            if (!resident) {
              umma_commit_if(leader, b_empty(bt));
              if (++bt == p.nb_tiles) {
                bt = 0;
                bph ^= 1;
              }
            }
          }
        }
      }
      umma_commit_if(leader, a_empty(astg));
      umma_commit_if(leader, acc_full(ab));
      if (tr0) p.trace[4 * 512 + tcount] = clock64();
      if (++astg == p.a_stages) {
        astg = 0;
        aph ^= 1;
      }
      if (++ab == p.acc_bufs) {
        ab = 0;
        accph ^= 1;
      }
    }
  } else {
    // ===================================================================================== epilogue
    const int quarter = warp & 3;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int i = quarter * 32 + lane, r = i >> 3, c = i & 7;
    float a1[SNCH > 0 ? SNCH * 16 : 1], a2[SNCH > 0 ? SNCH * 16 : 1];
    if constexpr (SNCH > 0) {
#pragma unroll
      for (int k = 0; k < SNCH * 16; ++k) a1[k] = a2[k] = 0.f;
    }
    int tcount = 0, ab = 0;
    uint32_t accph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
      const int n = tile / tiles_per_img, rem = tile - n * tiles_per_img;
      const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
      const int h = th * TILE_H + r;
      const bool tre = (p.dbg & 16) && blockIdx.x == 0 && tcount < 512 && threadIdx.x == 64;
      if (tre) p.trace[9 * 512 + tcount] = clock64();
      mbar_wait(acc_full(ab), accph);
      if (tre) p.trace[5 * 512 + tcount] = clock64();
      tcgen05_fence_after();
      const uint32_t t_acc = tmem_base + lane_base + (uint32_t)ab * (uint32_t)(NSUB * p.BN);
#pragma unroll
      for (int j = 0; j < NSUB; ++j) {
        const int w = tw * TILE_W + j * 8 + c;
        const bool ok = h < p.H && w < p.W;
        const long long pix = ((long long)n * p.H + h) * p.W + w;
        bf16* yrow = p.y + pix * p.y_pitch + p.y_off;
        const bf16* rrow = p.residual ? p.residual + pix * p.y_pitch + p.y_off : nullptr;
        auto chunk = [&](int c16, float* s1, float* s2) {
          float v[16];
          tmem_ld16(t_acc + (uint32_t)(j * p.BN + c16 * 16), v);
          const bool full = c16 * 16 + 16 <= p.K;  // else exactly 8 valid channels (K % 8 == 0)
          if (rrow && ok) {
            const uint4 r0 = *reinterpret_cast<const uint4*>(rrow + c16 * 16);
            const uint4 r1 = full ? *reinterpret_cast<const uint4*>(rrow + c16 * 16 + 8) : make_uint4(0, 0, 0, 0);
            const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              v[2 * k] += __uint_as_float(rr[k] << 16);
              v[2 * k + 1] += __uint_as_float(rr[k] & 0xffff0000u);
            }
          }
          uint32_t pk[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
            pk[k] = *reinterpret_cast<uint32_t*>(&hh);
          }
          if (ok && !(p.dbg & 1)) {
            if (full && p.wide_store) {
              st_global_256(yrow + c16 * 16, pk);
            } else {
              *reinterpret_cast<uint4*>(yrow + c16 * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              if (full) *reinterpret_cast<uint4*>(yrow + c16 * 16 + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
          if constexpr (SNCH > 0) {
            if (ok) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const float lo = __uint_as_float(pk[k] << 16), hi = __uint_as_float(pk[k] & 0xffff0000u);
                s1[2 * k] += lo;
                s2[2 * k] = fmaf(lo, lo, s2[2 * k]);
                s1[2 * k + 1] += hi;
                s2[2 * k + 1] = fmaf(hi, hi, s2[2 * k + 1]);
              }
            }
          }
        };
        if constexpr (SNCH > 0) {
#pragma unroll
          for (int c16 = 0; c16 < SNCH; ++c16) chunk(c16, a1 + c16 * 16, a2 + c16 * 16);
        } else {
          for (int c16 = 0; c16 < p.BN / 16; ++c16) chunk(c16, nullptr, nullptr);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(ab));
      if (tre) p.trace[6 * 512 + tcount] = clock64();
      if (++ab == p.acc_bufs) {
        ab = 0;
        accph ^= 1;
      }
    }
    if constexpr (SNCH > 0) {
#pragma unroll
      for (int c16 = 0; c16 < SNCH; ++c16) {
        float t1[16], t2[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          t1[k] = a1[c16 * 16 + k];
          t2[k] = a2[c16 * 16 + k];
        }
        const float s1 = butterfly_colsum(t1, lane), s2 = butterfly_colsum(t2, lane);
        const int col = c16 * 16 + col_of_lane(lane);
        if ((lane & 1) == 0 && col < p.K) {
          atomicAdd(&s_stats[col], s1);
          atomicAdd(&s_stats[p.BN + col], s2);
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 1) tcgen05_dealloc(tmem_base, p.tmem_cols);
  if constexpr (SNCH > 0) {
    // stats layout in global memory: [repl][2][K]
    double* st = p.stats + (long long)(blockIdx.x & (p.stats_repl - 1)) * 2 * p.K;
    for (int k = threadIdx.x; k < p.K; k += HALO_THREADS) {
      if (s_stats[k] != 0.f) atomicAdd(&st[k], (double)s_stats[k]);
      if (s_stats[p.BN + k] != 0.f) atomicAdd(&st[p.K + k], (double)s_stats[p.BN + k]);
    }
  }
}

typedef void (*HaloFn)(const CUtensorMap, const CUtensorMap, const HaloParams);
struct HaloVariant {
  int kc, chunks, snch;
  HaloFn fn;
  int regs;
  bool ready;
};
#define SGB_HALO(KC_, CH_) \
  {KC_, CH_, 0, conv3x3_halo_kernel<KC_, CH_, 0>, 0, false}, {KC_, CH_, (KC_ * CH_) / 16, conv3x3_halo_kernel<KC_, CH_, (KC_ * CH_) / 16>, 0, false}
HaloVariant g_variants[] = {SGB_HALO(32, 1), SGB_HALO(16, 3), SGB_HALO(64, 1), SGB_HALO(32, 3),
                            {64, 2, 0, conv3x3_halo_kernel<64, 2, 0>, 0, false}, {64, 3, 0, conv3x3_halo_kernel<64, 3, 0>, 0, false}};
#undef SGB_HALO

HaloVariant* find_variant(int C, int bn, bool stats) {
  for (HaloVariant& v : g_variants)
    if (v.kc * v.chunks == C && (stats ? v.snch == bn / 16 : v.snch == 0)) return &v;
  return nullptr;
}

bool halo_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SGB_DISABLE_HALO");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

struct HaloPlan {
  HaloVariant* var;
  int bn, tmem_cols, ctas, a_stages, acc_bufs, nb_tiles, resident;
  uint32_t a_bytes, b_tile_bytes, ctrl_bytes;
  size_t smem;
};

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && atoi(e) > 0) ? atoi(e) : dflt;
}

// Variant + shared-memory / TMEM plan, or false when the shape has to stay on the im2col kernel.
bool make_plan(const Problem& q, HaloPlan& pl) {
  if (!halo_enabled()) return false;
  if (q.R != 3 || q.S != 3 || q.stride != 1 || q.pad != 1 || q.ntaps > 0 || q.out_mode != 0) return false;
  if (q.P != q.H || q.Q != q.W) return false;
  if (q.scale || q.shift || q.act != SGB_ACT_NONE) return false;
  if (q.b_rows % 8 != 0 || q.b_rows > 256) return false;
  if (q.b_cols_per_tap < q.C) return false;
  if (q.stats && q.residual) return false;
  pl.bn = ((q.b_rows + 15) / 16) * 16;
  pl.var = find_variant(q.C, pl.bn, q.stats != nullptr);
  if (!pl.var) return false;
  HaloVariant* var = pl.var;
  if (!var->ready) {
    if (cudaFuncSetAttribute(var->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return false;
    cudaFuncAttributes fa{};
    if (cudaFuncGetAttributes(&fa, var->fn) != cudaSuccess) return false;
    var->regs = fa.numRegs;
    var->ready = true;
  }
  const int rowb = var->kc * 2, chunks = var->chunks;
  pl.a_bytes = (uint32_t)chunks * ((HALO_PX * rowb + 1023u) & ~1023u);
  pl.b_tile_bytes = ((uint32_t)(pl.bn * rowb) + 1023u) & ~1023u;
  pl.ctrl_bytes = 8 * (2 * MAX_A_STAGES + 4 + 2 * MAX_B_TILES) + 16 + 2 * pl.bn * 4 + 64;
  const int regs_alloc = ((var->regs + 7) / 8) * 8 * HALO_THREADS;
  const int max_ctas_regs = 65536 / regs_alloc > 0 ? 65536 / regs_alloc : 1;
  auto tmem_cols = [&](int bufs) {
    int tc = 32;
    while (tc < bufs * NSUB * pl.bn) tc <<= 1;
    return tc;
  };
  if (tmem_cols(1) > 512) return false;
  auto smem_bytes = [&](int stages, int tiles) { return (size_t)1024 + (size_t)stages * pl.a_bytes + (size_t)tiles * pl.b_tile_bytes + pl.ctrl_bytes; };
  auto budget = [&](int ctas) { return (size_t)(ctas == 1 ? 226 : 227 / ctas - 1) * 1024; };
  // Candidates in order of preference: a fully pipelined CTA (two halo stages, two accumulator sets, resident filters),
  // as many of them per SM as fit; then the same with streamed filters; then single-stage CTAs that rely on co-residency.
  const int force_ctas = env_int("SGB_HALO_CTAS", 0), force_stages = env_int("SGB_HALO_ASTAGES", 0), force_tiles = env_int("SGB_HALO_BTILES", 0);
  const int all_tiles = 9 * chunks;
  const int stream_tiles = chunks * 3 > 6 ? chunks * 3 : 6;
  struct Cand { int stages, tiles; };
  const Cand cands[] = {{2, all_tiles}, {2, stream_tiles}, {1, all_tiles}, {1, stream_tiles}, {1, 2 * chunks}};
  for (const Cand& cd : cands) {
    const int stages = force_stages ? force_stages : cd.stages;
    int tiles = force_tiles ? force_tiles : cd.tiles;
    if (tiles > all_tiles) tiles = all_tiles;
    for (int ctas = force_ctas ? force_ctas : 3; ctas >= 1; --ctas) {
      if (ctas > max_ctas_regs || smem_bytes(stages, tiles) > budget(ctas)) { if (force_ctas) break; continue; }
      int bufs = 2;
      if (tmem_cols(2) * ctas > 512) bufs = 1;
      if (tmem_cols(bufs) * ctas > 512) { if (force_ctas) break; continue; }
      pl.ctas = ctas; pl.a_stages = stages; pl.acc_bufs = bufs; pl.nb_tiles = tiles; pl.resident = tiles == all_tiles;
      pl.tmem_cols = tmem_cols(bufs);
      pl.smem = smem_bytes(stages, tiles);
      return true;
    }
  }
  return false;
}

}  // namespace

long long g_halo_launches = 0;
long long halo_launch_count() { return g_halo_launches; }

bool halo_supported(const Problem& q) {
  HaloPlan pl{};
  return make_plan(q, pl);
}

int halo_launch(const Problem& q, cudaStream_t st) {
  if (int rc = init_driver()) return rc;
  HaloPlan pl{};
  if (!make_plan(q, pl)) return SGB_E_UNSUPPORTED;
  HaloVariant* var = pl.var;
  const int bn = pl.bn, kc = var->kc, ctas = pl.ctas;
  HaloParams p{};
  p.N = q.N; p.H = q.H; p.W = q.W; p.K = q.b_rows; p.BN = bn;
  p.tiles_h = (q.H + TILE_H - 1) / TILE_H;
  p.tiles_w = (q.W + TILE_W - 1) / TILE_W;
  const long long total = (long long)q.N * p.tiles_h * p.tiles_w;
  if (total >= (1ll << 31)) return SGB_E_UNSUPPORTED;
  p.total_tiles = (int)total;
  for (int t = 0; t < 9; ++t) p.tap_b[t] = q.flip ? 8 - t : t;
  p.b_cols_per_tap = q.b_cols_per_tap;
  p.y = (bf16*)q.y; p.y_pitch = q.y_pitch; p.y_off = q.y_off;
  p.residual = (const bf16*)q.residual;
  p.stats = q.stats; p.stats_repl = q.stats_repl > 0 ? q.stats_repl : 1;
  p.dbg = debug_skip_mask();
  p.trace = (p.dbg & 16) ? trace_buffer() : nullptr;
  p.tmem_cols = pl.tmem_cols;
  p.b_tile_bytes = pl.b_tile_bytes;
  p.nb_tiles = pl.nb_tiles;
  p.b_resident = pl.resident;
  p.a_stages = pl.a_stages;
  p.acc_bufs = pl.acc_bufs;
  p.wide_store = (q.y_pitch % 16 == 0 && q.y_off % 16 == 0 && ((uintptr_t)q.y & 31) == 0) ? 1 : 0;
  if (p.residual && !p.wide_store) p.wide_store = 0;
  const size_t smem = pl.smem;

  alignas(64) CUtensorMap map_a, map_b;
  {
    cuuint64_t dims[4] = {(cuuint64_t)q.C, (cuuint64_t)q.W, (cuuint64_t)q.H, (cuuint64_t)q.N};
    cuuint64_t strides[3] = {(cuuint64_t)q.a_pitch * 2, (cuuint64_t)q.W * q.a_pitch * 2, (cuuint64_t)q.H * q.W * q.a_pitch * 2};
    cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)HALO_W, (cuuint32_t)HALO_H, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_tiled(&map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(q.a), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(kc), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      sgb_set_error("cuTensorMapEncodeTiled(halo A) failed with %d (C=%d W=%d H=%d N=%d pitch=%d KC=%d)", (int)r, q.C, q.W, q.H, q.N,
                    q.a_pitch, kc);
      return SGB_E_CUDA;
    }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)q.b_cols, (cuuint64_t)q.b_rows};
    cuuint64_t strides[1] = {(cuuint64_t)q.b_cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)bn};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_tiled(&map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(q.b), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(kc), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      sgb_set_error("cuTensorMapEncodeTiled(halo B) failed with %d (cols=%d rows=%d KC=%d BN=%d)", (int)r, q.b_cols, q.b_rows, kc, bn);
      return SGB_E_CUDA;
    }
  }
  int grid = p.total_tiles;
  if (grid > g_num_sms * ctas) grid = g_num_sms * ctas;
  var->fn<<<grid, HALO_THREADS, smem, st>>>(map_a, map_b, p);
  ++g_launches;
  ++g_halo_launches;
  return sgb_cuda_check(cudaGetLastError(), "conv3x3_halo_kernel");
}

}  // namespace sm100
