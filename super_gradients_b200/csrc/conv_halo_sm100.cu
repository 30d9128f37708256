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
#define HB_KERNEL conv3x3_halo_kernel
#define HB_R 3
#define HB_TAPS 9
#define HB_PAD 1
#define HB_W HALO_W
#define HB_PX HALO_PX
#include "conv_halo_body.inc"
#undef HB_KERNEL
#undef HB_R
#undef HB_TAPS
#undef HB_PAD
#undef HB_W
#undef HB_PX

#ifdef SGB_HALO_1X1
// Experiment: the same pipeline on a plain 16 x 16 tile with one tap = a 1x1 convolution with resident filters, 256-pixel tiles,
// the 256-bit store epilogue and register-held statistics, instead of the im2col kernel's 128-pixel tiles.
#define HB_KERNEL conv1x1_tile_kernel
#define HB_R 1
#define HB_TAPS 1
#define HB_PAD 0
#define HB_W TILE_W
#define HB_PX (TILE_H * TILE_W)
#include "conv_halo_body.inc"
#undef HB_KERNEL
#undef HB_R
#undef HB_TAPS
#undef HB_PAD
#undef HB_W
#undef HB_PX
#endif

// ------------------------------------------------------------------------------------------------ wgrad from halo tiles
// dW[k][dh][dw][c] += sum over the 16 x 16 tile of dy[pix][k] * x[pix + (dh-1, dw-1)][c]       (3x3, stride 1, pad 1)
// GEMM view per filter row dh: M = out-channels (TMEM lanes), N = (dw, c) (TMEM columns), K = the 16 pixels of one tile row.
//   * A = the dy tile [16][16][K] exactly as TMA lands it: MN-major, one MMA per tile row r.  Out-channel atoms that do not
//     exist (K < 128) alias atom 0 (LBO = 0); the duplicated accumulator rows are never read.
//   * B = the x halo tile [18][18][c]: the pixels (r + dh, 0..15) are 16 consecutive rows of the tile, and the three
//     horizontal taps dw = 0, 1, 2 are the SAME rows shifted by one pixel -- expressed as three overlapping N atoms with
//     LBO = one pixel.  One MMA therefore covers a whole filter row for KCB channels: 3 x fewer MMAs than per-tap issue and
//     one activation load per tile instead of nine.
// A CTA owns (128 out-channels) x (Cs = KCB * CHB in-channels) x (a contiguous range of tiles) and keeps all nine taps of
// its slice in TMEM (9 * Cs fp32 columns) across its tiles; one round of fp32 reductions to global memory at the end.
struct WHaloParams {
  int N, H, W;
  int K, C, cpad;
  int tiles_h, tiles_w, total_tiles, tiles_per_cta;
  int n_mblocks, n_slices;
  int stages;
  int a_atoms;            // 64- or 32-channel dy boxes loaded per stage (1 or 2)
  int row_a;              // bytes per pixel row of a dy atom (64 or 128)
  uint32_t a_lbo_bytes;   // 0: every M atom aliases atom 0
  uint32_t stage_bytes, a_bytes;
  float* dw;
  int tmem_cols;
  int dbg;
};

template <int KCB, int CHB>
__global__ void __launch_bounds__(HALO_THREADS, 1)
wgrad3x3_halo_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, const WHaloParams p) {
  constexpr int ROWB = KCB * 2;
  constexpr uint32_t B_REGION = ((uint32_t)(HALO_PX * ROWB) + 1023u) & ~1023u;
  constexpr int NB = 3 * KCB;  // GEMM N of one MMA: three horizontal taps
  constexpr int WG_MAX_STAGES = 8;
  SGB_GRID_DEP_LAUNCH();
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ctrl = smem_base + (uint32_t)p.stages * p.stage_bytes;
  auto full_bar = [&](int s) { return ctrl + 8u * s; };
  auto empty_bar = [&](int s) { return ctrl + 8u * (WG_MAX_STAGES + s); };
  const uint32_t done_bar = ctrl + 8u * (2 * WG_MAX_STAGES);
  const uint32_t tmem_slot = ctrl + 8u * (2 * WG_MAX_STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int wi = blockIdx.x;
  const int mblock = wi % p.n_mblocks; wi /= p.n_mblocks;
  const int slice = wi % p.n_slices; wi /= p.n_slices;
  const int t0 = wi * p.tiles_per_cta;
  const int t1 = min(t0 + p.tiles_per_cta, p.total_tiles);
  const int n_tiles = t1 - t0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_MAX_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) tcgen05_alloc(tmem_slot, p.tmem_cols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  SGB_GRID_DEP_WAIT();  // everything above touches only shared memory / TMEM
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (n_tiles > 0) {
    if (warp == 0) {
      if (elect_one()) {
        const int tiles_per_img = p.tiles_h * p.tiles_w;
        int stg = 0;
        uint32_t par = 1;
        const uint32_t tx = (uint32_t)p.a_atoms * 256u * (uint32_t)p.row_a + (uint32_t)(CHB * HALO_PX * ROWB);
        for (int tile = t0; tile < t1; ++tile) {
          const int n = tile / tiles_per_img, rem = tile - n * tiles_per_img;
          const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
          const int h0 = th * TILE_H, w0 = tw * TILE_W;
          mbar_wait(empty_bar(stg), par);
          mbar_expect_tx(full_bar(stg), tx);
          const uint32_t sa = smem_base + (uint32_t)stg * p.stage_bytes, sb = sa + p.a_bytes;
          for (int a = 0; a < p.a_atoms; ++a)
            tma_load_tiled_4d(sa + (uint32_t)a * 256u * (uint32_t)p.row_a, &map_dy, full_bar(stg), mblock * 128 + a * (p.row_a / 2), w0, h0, n);
#pragma unroll
          for (int ck = 0; ck < CHB; ++ck)
            tma_load_tiled_4d(sb + ck * B_REGION, &map_x, full_bar(stg), slice * (KCB * CHB) + ck * KCB, w0 - 1, h0 - 1, n);
          if (++stg == p.stages) {
            stg = 0;
            par ^= 1;
          }
        }
      }
    } else if (warp == 1) {
      const uint32_t leader = elect_one() ? 1u : 0u;
      // A and B MN-major (bits 15, 16); M = 128, N = NB
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t row_a = (uint32_t)p.row_a;
      const uint64_t a_layout = row_a == 128 ? 2u : 4u;
      const uint64_t a_hi = ((uint64_t)((p.a_lbo_bytes >> 4) & 0x3fff) << 16) | ((uint64_t)(((8 * row_a) >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46) | (a_layout << 61);
      constexpr uint64_t b_layout = ROWB == 128 ? 2u : (ROWB == 64 ? 4u : 6u);
      constexpr uint64_t b_hi = ((uint64_t)((ROWB >> 4) & 0x3fff) << 16) | ((uint64_t)(((8 * ROWB) >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46) | (b_layout << 61);
      int stg = 0;
      uint32_t par = 0;
      for (int it = 0; it < n_tiles; ++it) {
        mbar_wait(full_bar(stg), par);
        tcgen05_fence_after();
        const uint32_t sa = (smem_base + (uint32_t)stg * p.stage_bytes) >> 4, sb = sa + (p.a_bytes >> 4);
#pragma unroll 1
        for (int r = 0; r < TILE_H; ++r) {
          const uint64_t da = a_hi | (uint64_t)((sa + (uint32_t)r * ((16u * row_a) >> 4)) & 0x3fff);
#pragma unroll
          for (int dh = 0; dh < 3; ++dh) {
#pragma unroll
            for (int ck = 0; ck < CHB; ++ck) {
              const uint64_t db = b_hi | (uint64_t)((sb + ck * (B_REGION >> 4) + (uint32_t)(r + dh) * ((HALO_W * ROWB) >> 4)) & 0x3fff);
              if (!(p.dbg & 8)) umma_bf16_if(leader, tmem_base + (uint32_t)((dh * CHB + ck) * NB), da, db, idesc, (it | r) != 0);
            }
          }
        }
        umma_commit_if(leader, empty_bar(stg));
        if (it == n_tiles - 1) umma_commit_if(leader, done_bar);
        if (++stg == p.stages) {
          stg = 0;
          par ^= 1;
        }
      }
    } else {
      const int quarter = warp & 3;
      mbar_wait(done_bar, 0);
      tcgen05_fence_after();
      const int ko = mblock * 128 + quarter * 32 + lane;
      float* drow = p.dw + (long long)ko * 9 * p.cpad + slice * (KCB * CHB);
#pragma unroll 1
      for (int dh = 0; dh < 3; ++dh)
#pragma unroll 1
        for (int ck = 0; ck < CHB; ++ck)
#pragma unroll 1
          for (int dw = 0; dw < 3; ++dw)
#pragma unroll
            for (int c0 = 0; c0 < KCB; c0 += 16) {
              float v[16];
              tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((dh * CHB + ck) * NB + dw * KCB + c0), v);
              if (ko < p.K && !(p.dbg & 1)) {
                float* dst = drow + (dh * 3 + dw) * p.cpad + ck * KCB + c0;
#pragma unroll
                for (int k = 0; k < 16; k += 4)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + k), "f"(v[k]), "f"(v[k + 1]), "f"(v[k + 2]), "f"(v[k + 3])
                               : "memory");
              }
            }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 1) tcgen05_dealloc(tmem_base, p.tmem_cols);
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

#ifdef SGB_HALO_1X1
#define SGB_TILE1(KC_, CH_) \
  {KC_, CH_, 0, conv1x1_tile_kernel<KC_, CH_, 0>, 0, false}, {KC_, CH_, (KC_ * CH_) / 16, conv1x1_tile_kernel<KC_, CH_, (KC_ * CH_) / 16>, 0, false}
HaloVariant g_variants_1x1[] = {SGB_TILE1(32, 1), SGB_TILE1(16, 3), SGB_TILE1(64, 1), SGB_TILE1(32, 3),
                                {64, 2, 0, conv1x1_tile_kernel<64, 2, 0>, 0, false}, {64, 3, 0, conv1x1_tile_kernel<64, 3, 0>, 0, false},
                                // ConvBNAct 1x1 layers of the CSP stages change the channel count: statistics for K != C
                                {32, 3, 2, conv1x1_tile_kernel<32, 3, 2>, 0, false}, {64, 1, 6, conv1x1_tile_kernel<64, 1, 6>, 0, false},
                                {64, 3, 4, conv1x1_tile_kernel<64, 3, 4>, 0, false}, {32, 3, 3, conv1x1_tile_kernel<32, 3, 3>, 0, false}};
#undef SGB_TILE1
HaloVariant* find_variant_1x1(int C, int bn, bool stats) {
  for (HaloVariant& v : g_variants_1x1)
    if (v.kc * v.chunks == C && (stats ? v.snch == bn / 16 : v.snch == 0)) return &v;
  return nullptr;
}
bool tile_1x1_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SGB_HALO_1X1");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
#endif

// filter size served by a plan: 3 (halo tile) or, in the -DSGB_HALO_1X1 experiment build, 1 (plain 16 x 16 tile)
int plan_filter_size(const Problem& q) {
  if (q.R == 3 && q.S == 3 && q.pad == 1) return 3;
#ifdef SGB_HALO_1X1
  if (q.R == 1 && q.S == 1 && q.pad == 0 && tile_1x1_enabled()) return 1;
#endif
  return 0;
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
  const int fs = plan_filter_size(q);
  if (fs == 0 || q.stride != 1 || q.ntaps > 0 || q.out_mode != 0) return false;
  if (q.P != q.H || q.Q != q.W) return false;
  if (q.scale || q.shift || q.act != SGB_ACT_NONE) return false;
  if (q.b_rows % 8 != 0 || q.b_rows > 256) return false;
  if (q.b_cols_per_tap < q.C) return false;
  if (q.stats && q.residual) return false;
  pl.bn = ((q.b_rows + 15) / 16) * 16;
  pl.var = find_variant(q.C, pl.bn, q.stats != nullptr);
#ifdef SGB_HALO_1X1
  if (fs == 1) pl.var = find_variant_1x1(q.C, pl.bn, q.stats != nullptr);
#endif
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
  const uint32_t tile_px = fs == 3 ? (uint32_t)HALO_PX : (uint32_t)(TILE_H * TILE_W);
  pl.a_bytes = (uint32_t)chunks * ((tile_px * rowb + 1023u) & ~1023u);
  pl.b_tile_bytes = ((uint32_t)(pl.bn * rowb) + 1023u) & ~1023u;
  pl.ctrl_bytes = 8 * (2 * MAX_A_STAGES + 4 + 2 * MAX_B_TILES) + 16 + SGB_STATS_SLOTS * 2 * pl.bn * 4 + 64;
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
  const int all_tiles = fs * fs * chunks;
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
  const int fs = plan_filter_size(q);
  for (int t = 0; t < 9; ++t) p.tap_b[t] = fs == 1 ? 0 : (q.flip ? 8 - t : t);
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
    cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)(fs == 3 ? HALO_W : TILE_W), (cuuint32_t)(fs == 3 ? HALO_H : TILE_H), 1};
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
  SGB_LAUNCH(var->fn, grid, HALO_THREADS, smem, st, map_a, map_b, p);
  ++g_launches;
  ++g_halo_launches;
  return sgb_cuda_check(cudaGetLastError(), "conv3x3_halo_kernel");
}

// ---- wgrad
namespace {
typedef void (*WHaloFn)(const CUtensorMap, const CUtensorMap, const WHaloParams);
struct WHaloPlan {
  WHaloFn fn;
  int kcb, chb, cs;
  WHaloParams p;
  size_t smem;
  int grid;
};

bool make_wplan(const WgradProblem& q, WHaloPlan& pl) {
  if (!halo_enabled()) return false;
  {
    const char* e = getenv("SGB_DISABLE_HALO_WGRAD");
    if (e && e[0] == '1') return false;
  }
  if (q.R != 3 || q.S != 3 || q.stride != 1 || q.pad != 1 || q.P != q.H || q.Q != q.W) return false;
  if (q.C % 16 != 0 || q.K % 8 != 0) return false;
  if (q.x_pitch % 8 != 0 || q.y_pitch % 8 != 0) return false;
  if (((uintptr_t)q.x & 15) || ((uintptr_t)q.dy & 15) || ((uintptr_t)q.dw & 15)) return false;
  if (q.C % 32 == 0) { pl.kcb = 32; pl.chb = 1; pl.fn = wgrad3x3_halo_kernel<32, 1>; }
  else if (q.C % 48 == 0) { pl.kcb = 16; pl.chb = 3; pl.fn = wgrad3x3_halo_kernel<16, 3>; }
  else { pl.kcb = 16; pl.chb = 1; pl.fn = wgrad3x3_halo_kernel<16, 1>; }
  pl.cs = pl.kcb * pl.chb;
  WHaloParams& p = pl.p;
  p = WHaloParams{};
  p.N = q.N; p.H = q.H; p.W = q.W; p.K = q.K; p.C = q.C; p.cpad = q.C;
  p.tiles_h = (q.H + TILE_H - 1) / TILE_H;
  p.tiles_w = (q.W + TILE_W - 1) / TILE_W;
  const long long total = (long long)q.N * p.tiles_h * p.tiles_w;
  if (total >= (1ll << 31)) return false;
  p.total_tiles = (int)total;
  p.n_mblocks = (q.K + 127) / 128;
  p.n_slices = q.C / pl.cs;
  const int kb = q.K < 128 ? q.K : 128;  // widest M block
  if (kb <= 32) { p.row_a = 64; p.a_atoms = 1; p.a_lbo_bytes = 0; }
  else if (kb <= 64) { p.row_a = 128; p.a_atoms = 1; p.a_lbo_bytes = 0; }
  else { p.row_a = 128; p.a_atoms = 2; p.a_lbo_bytes = 256u * 128u; }
  p.a_bytes = (uint32_t)p.a_atoms * 256u * (uint32_t)p.row_a;  // multiples of 16 KB
  const uint32_t b_bytes = (uint32_t)pl.chb * (((uint32_t)(HALO_PX * pl.kcb * 2) + 1023u) & ~1023u);
  p.stage_bytes = p.a_bytes + b_bytes;
  const uint32_t ctrl_bytes = 8 * (2 * 8 + 1) + 16 + 64;
  int stages = (int)((226u * 1024u - 1024u - ctrl_bytes) / p.stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) return false;
  p.stages = stages;
  int tc = 32;
  while (tc < 9 * pl.cs) tc <<= 1;
  if (tc > 512) return false;
  p.tmem_cols = tc;
  p.dw = q.dw;
  p.dbg = debug_skip_mask();
  pl.smem = 1024 + (size_t)stages * p.stage_bytes + ctrl_bytes;
  const int base = p.n_mblocks * p.n_slices;
  int splits = g_num_sms > 0 ? g_num_sms / base : 1;
  if (splits < 1) splits = 1;
  if (splits > p.total_tiles) splits = p.total_tiles;
  p.tiles_per_cta = (p.total_tiles + splits - 1) / splits;
  splits = (p.total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  pl.grid = base * splits;
  return true;
}
}  // namespace

bool wgrad_halo_supported(const WgradProblem& q) {
  if (init_driver() != SGB_OK) return false;
  WHaloPlan pl{};
  return make_wplan(q, pl);
}

int wgrad_halo_launch(const WgradProblem& q, cudaStream_t st) {
  if (int rc = init_driver()) return rc;
  WHaloPlan pl{};
  if (!make_wplan(q, pl)) return SGB_E_UNSUPPORTED;
  static bool attr[3] = {false, false, false};
  const int vi = pl.kcb == 32 ? 0 : (pl.chb == 3 ? 1 : 2);
  if (!attr[vi]) {
    if (int rc = sgb_cuda_check(cudaFuncSetAttribute(pl.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024),
                                "cudaFuncSetAttribute(wgrad3x3_halo_kernel)"))
      return rc;
    attr[vi] = true;
  }
  const WHaloParams& p = pl.p;
  alignas(64) CUtensorMap map_dy, map_x;
  {
    cuuint64_t dims[4] = {(cuuint64_t)q.K, (cuuint64_t)q.W, (cuuint64_t)q.H, (cuuint64_t)q.N};
    cuuint64_t strides[3] = {(cuuint64_t)q.y_pitch * 2, (cuuint64_t)q.W * q.y_pitch * 2, (cuuint64_t)q.H * q.W * q.y_pitch * 2};
    cuuint32_t box[4] = {(cuuint32_t)(p.row_a / 2), (cuuint32_t)TILE_W, (cuuint32_t)TILE_H, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_tiled(&map_dy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(q.dy), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(p.row_a / 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { sgb_set_error("cuTensorMapEncodeTiled(halo wgrad dy) failed with %d", (int)r); return SGB_E_CUDA; }
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)q.C, (cuuint64_t)q.W, (cuuint64_t)q.H, (cuuint64_t)q.N};
    cuuint64_t strides[3] = {(cuuint64_t)q.x_pitch * 2, (cuuint64_t)q.W * q.x_pitch * 2, (cuuint64_t)q.H * q.W * q.x_pitch * 2};
    cuuint32_t box[4] = {(cuuint32_t)pl.kcb, (cuuint32_t)HALO_W, (cuuint32_t)HALO_H, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_tiled(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(q.x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(pl.kcb), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { sgb_set_error("cuTensorMapEncodeTiled(halo wgrad x) failed with %d", (int)r); return SGB_E_CUDA; }
  }
  SGB_LAUNCH(pl.fn, pl.grid, HALO_THREADS, pl.smem, st, map_dy, map_x, p);
  ++g_launches;
  ++g_halo_launches;
  return sgb_cuda_check(cudaGetLastError(), "wgrad3x3_halo_kernel");
}

}  // namespace sm100
