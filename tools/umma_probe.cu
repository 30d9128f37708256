// Hardware probe (development tool, not part of the library): does a tcgen05 shared-memory matrix descriptor whose
// start address points INSIDE a TMA-written swizzle pattern (an arbitrary row of a 1024-byte aligned tile) address the
// rows TMA put there?  This decides whether a 3x3 convolution can read its nine taps out of ONE halo tile in shared
// memory (start address = tile + (dh * pitch + dw) * row_bytes, SBO = halo row pitch) instead of nine im2col loads.
//
//   mode 0: A is K-major, 128 rows taken as 16 groups of 8 consecutive rows, group stride sbo_rows, first row `off`.
//   mode 1: B is MN-major (rows = GEMM-K = pixels, columns = channels), 16 consecutive rows starting at row `off`.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

struct Cfg {
  int mode, row_bytes, off_rows, sbo_rows, kstep, base_mode, s_row_bytes;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t addr, int row_bytes, uint32_t sbo_bytes, int base_mode) {
  const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  if (base_mode == 1) d |= (uint64_t)((addr >> 7) & 7) << 49;
  d |= (uint64_t)layout << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_s, const Cfg c, float* out) {
  extern __shared__ __align__(1024) unsigned char raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t g = base, s = base + 32768, bar0 = base + 49152, bar1 = bar0 + 8, slot = bar0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  const int s_rows = (c.mode == 0 || c.mode == 3) ? 16 : 128;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar0, 256 * c.row_bytes + s_rows * c.s_row_bytes);
    tma_load_2d(g, &map_g, bar0, 0, 0);
    tma_load_2d(s, &map_s, bar0, 0, 0);
    mbar_wait(bar0, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint64_t da, db;
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);
    if (c.mode == 0) {
      da = desc(g + c.off_rows * c.row_bytes + c.kstep * 32, c.row_bytes, c.sbo_rows * c.row_bytes, c.base_mode);
      db = desc(s + c.kstep * 32, c.s_row_bytes, 8 * c.s_row_bytes, 0);
      idesc |= (uint32_t)(16 >> 3) << 17;
    } else if (c.mode == 1) {
      da = desc(s, c.s_row_bytes, 8 * c.s_row_bytes, 0);
      db = desc(g + c.off_rows * c.row_bytes, c.row_bytes, 8 * c.row_bytes, c.base_mode);
      idesc |= (1u << 16) | ((uint32_t)((c.row_bytes / 2) >> 3) << 17);
    } else if (c.mode == 2) {
      // B MN-major, N = 3 atoms that OVERLAP: atom i starts one row (pixel) after atom i-1 (LBO = row_bytes)
      da = desc(s, c.s_row_bytes, 8 * c.s_row_bytes, 0);
      db = desc(g + c.off_rows * c.row_bytes, c.row_bytes, 8 * c.row_bytes, 0);
      db = (db & ~((uint64_t)0x3fff << 16)) | ((uint64_t)((c.row_bytes >> 4) & 0x3fff) << 16);
      idesc |= (1u << 16) | ((uint32_t)((3 * c.row_bytes / 2) >> 3) << 17);
    } else {
      // A MN-major, M = 128 made of 128 / atom aliases of ONE atom (LBO = 0): rows m and m % atom are equal
      da = desc(g + c.off_rows * c.row_bytes, c.row_bytes, 8 * c.row_bytes, 0);
      da = da & ~((uint64_t)0x3fff << 16);
      db = desc(s, c.s_row_bytes, 8 * c.s_row_bytes, 0);  // B K-major: 16 rows (N) x 16 k
      idesc |= (1u << 15) | ((uint32_t)(16 >> 3) << 17);
    }
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                 "l"(da), "l"(db), "r"(idesc), "r"(0)
                 : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar1) : "memory");
  }
  __syncwarp();
  mbar_wait(bar1, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int ncols = (c.mode == 0 || c.mode == 3) ? 16 : (c.mode == 1 ? c.row_bytes / 2 : 3 * c.row_bytes / 2);
  for (int c0 = 0; c0 < ncols; c0 += 16) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 256 + c0 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_tiled;

static CUtensorMap make_map(void* ptr, int cols, int rows, int row_bytes) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)cols, (cuuint32_t)rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = g_tiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                       CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode failed %d\n", (int)r);
    exit(1);
  }
  return m;
}

int main() {
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return 2;
  g_tiled = (EncodeTiledFn)fn;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  float* d_out;
  cudaMalloc(&d_out, 128 * 256 * 4);
  std::vector<float> h_out(128 * 256);
  srand(1);
  int n_fail = 0, n_run = 0;
  for (int mode = 0; mode < 4; ++mode)
    for (int row_bytes : {128, 64, 32}) {
      const int cols = row_bytes / 2;
      const int s_row_bytes = mode == 0 ? row_bytes : 64, s_cols = s_row_bytes / 2, s_rows = (mode == 0 || mode == 3) ? 16 : 128;
      std::vector<float> G(256 * cols), S(s_rows * s_cols);
      std::vector<__nv_bfloat16> Gb(G.size()), Sb(S.size());
      for (size_t i = 0; i < G.size(); ++i) { G[i] = (float)(rand() % 7 - 3); Gb[i] = __float2bfloat16(G[i]); }
      for (size_t i = 0; i < S.size(); ++i) { S[i] = (float)(rand() % 5 - 2); Sb[i] = __float2bfloat16(S[i]); }
      __nv_bfloat16 *dG, *dS;
      cudaMalloc(&dG, Gb.size() * 2);
      cudaMalloc(&dS, Sb.size() * 2);
      cudaMemcpy(dG, Gb.data(), Gb.size() * 2, cudaMemcpyHostToDevice);
      cudaMemcpy(dS, Sb.data(), Sb.size() * 2, cudaMemcpyHostToDevice);
      CUtensorMap mg = make_map(dG, cols, 256, row_bytes), ms = make_map(dS, s_cols, s_rows, s_row_bytes);
      const int offs[] = {0, 1, 2, 3, 5, 7, 8, 9, 10, 11, 12, 20, 21, 22};
      for (int sbo_rows : {8, 10, 18}) {
        if (mode >= 1 && sbo_rows != 8) continue;
        for (int off : offs)
          for (int kstep = 0; kstep < (mode == 0 ? cols / 16 : 1); kstep += (cols / 16 > 1 ? cols / 16 - 1 : 1))
            for (int base_mode = 0; base_mode < 2; ++base_mode) {
              if (mode == 0 && off + 15 * sbo_rows + 8 > 256) continue;
              Cfg c{mode, row_bytes, off, sbo_rows, kstep, base_mode, s_row_bytes};
              cudaMemset(d_out, 0xff, 128 * 256 * 4);
              probe<<<1, 128, 64 * 1024>>>(mg, ms, c, d_out);
              cudaError_t e = cudaDeviceSynchronize();
              if (e != cudaSuccess) {
                printf("CUDA error %s at mode=%d rb=%d off=%d sbo=%d k=%d base=%d\n", cudaGetErrorString(e), mode, row_bytes, off, sbo_rows, kstep, base_mode);
                return 3;
              }
              cudaMemcpy(h_out.data(), d_out, 128 * 256 * 4, cudaMemcpyDeviceToHost);
              int bad = 0;
              const int ncols = (mode == 0 || mode == 3) ? 16 : (mode == 1 ? cols : 3 * cols);
              for (int i = 0; i < 128; ++i)
                for (int n = 0; n < ncols; ++n) {
                  float ref = 0;
                  if (mode == 0) {
                    const int row = off + (i / 8) * sbo_rows + i % 8;
                    for (int k = 0; k < 16; ++k) ref += G[row * cols + kstep * 16 + k] * S[n * s_cols + kstep * 16 + k];
                  } else if (mode == 1) {
                    for (int k = 0; k < 16; ++k) ref += S[i * s_cols + k] * G[(off + k) * cols + n];
                  } else if (mode == 2) {
                    for (int k = 0; k < 16; ++k) ref += S[i * s_cols + k] * G[(off + k + n / cols) * cols + n % cols];
                  } else {
                    for (int k = 0; k < 16; ++k) ref += G[(off + k) * cols + i % cols] * S[n * s_cols + k];
                  }
                  if (ref != h_out[i * 256 + n]) ++bad;
                }
              ++n_run;
              if (bad) ++n_fail;
              printf("mode=%d row_bytes=%3d off=%2d sbo_rows=%2d kstep=%d base_mode=%d : %s (%d wrong)\n", mode, row_bytes, off, sbo_rows, kstep,
                     base_mode, bad ? "FAIL" : "ok", bad);
            }
      }
      cudaFree(dG);
      cudaFree(dS);
    }
  printf("%d configs, %d failed\n", n_run, n_fail);
  return 0;
}
