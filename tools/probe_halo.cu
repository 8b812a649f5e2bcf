// Hardware probe (GPU box only): can a tcgen05 shared-memory descriptor address a ROW-SHIFTED window of a
// 128B-swizzled TMA tile?  The 3x3 convolutions would like to load one haloed activation tile per
// 32-channel chunk and feed the nine taps from it through descriptors whose start address is offset by a
// multiple of 128 bytes (one pixel row), instead of nine tap-shifted TMA loads.
//   mode 0: K-major A operand (forward / dgrad style).  Tile = (16+2) x (8+2) haloed pixels x 32 ch;
//           M = 128 output pixels as 16 groups of 8, SBO = 10 pixels * 128 B = 1280 B (not a multiple of
//           1024), start = ((dy+1)*10 + dx+1) * 128.
//   mode 1: MN-major operands (wgrad style, "128B swizzle, 32B atom"): K = 8 pixels per MMA, start shifted
//           by delta pixel rows.
// Prints the max abs error against a host reference for every shift, with the descriptor base-offset field
// either zero or (start >> 7) & 7.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_halo tools/probe_halo.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra.uni WD;\n\tbra.uni WL;\n\tWD:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)layout << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int HALO_ROWS = 192;     // rows of 128 B loaded for the A operand (>= 18*10 = 180)

// mode 0: out[128][32] = A_window(128 x 32) * B(32 x 32)^T with the window described above
// mode 1: out[128][32] (rows 0..31 real) = sum over 8 pixels of X[p+delta][m] * DY[p][n]
__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                    float* out, int mode, int shift_rows, int sbo_bytes, int use_base_off) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_tile = smem;                              // HALO_ROWS x 128 B
  uint8_t* b_tile = smem + HALO_ROWS * 128;            // 64 x 128 B (1024-aligned: 192*128 = 24576)
  __shared__ __align__(8) uint64_t full_bar, mma_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&full_bar, 1);
    mbar_init(&mma_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&full_bar, HALO_ROWS * 128 + 64 * 128);
    tma_load_2d(a_tile, &tmA, &full_bar, 0, 0);
    tma_load_2d(b_tile, &tmB, &full_bar, 0, 0);
    mbar_wait(&full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t a0 = smem_u32(a_tile) + (uint32_t)shift_rows * 128u, b0 = smem_u32(b_tile);
    const uint32_t boff = use_base_off ? ((a0 >> 7) & 7) : 0;
    if (mode == 0) {
      const uint32_t idesc = make_idesc_tf32(128, 32, 0, 0);
      for (int k = 0; k < 4; ++k)
        umma_tf32(tmem, make_desc(a0 + k * 32, 16, (uint32_t)sbo_bytes, 2, boff), make_desc(b0 + k * 32, 16, 1024, 2, 0), idesc, k != 0);
    } else {
      // MN-major, 128B swizzle with 32B atoms (layout type 1): LBO = stride between 32-element M groups
      // (0: all four groups alias the one box), SBO = 512 (next 4 K-rows)
      const uint32_t idesc = make_idesc_tf32(128, 32, 1, 1);
      umma_tf32(tmem, make_desc(a0, 0, 512, 1, boff), make_desc(b0, 0, 512, 1, 0), idesc, 0);
    }
    umma_commit(&mma_bar);
  }
  mbar_wait(&mma_bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  float v[32];
  tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16), v);
  for (int j = 0; j < 32; ++j) out[(threadIdx.x) * 32 + j] = v[j];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(32) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; float r; memcpy(&r, &u, 4); return r; }

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  std::vector<float> hA(HALO_ROWS * 32), hB(64 * 32);
  srand(1);
  for (auto& x : hA) x = tf32((float)(rand() % 2001 - 1000) / 1000.f);
  for (auto& x : hB) x = tf32((float)(rand() % 2001 - 1000) / 1000.f);
  float *dA, *dB, *dO;
  CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dB, hB.size() * 4)); CK(cudaMalloc(&dO, 128 * 32 * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  std::vector<float> hO(128 * 32);
  const int smem = HALO_ROWS * 128 + 64 * 128 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int mode = 0; mode < 2; ++mode) {
    CUtensorMap tmA, tmB;
    cuuint64_t dimsA[2] = {32, HALO_ROWS}, dimsB[2] = {32, 64};
    cuuint64_t str[1] = {128};
    cuuint32_t boxA[2] = {32, HALO_ROWS}, boxB[2] = {32, 64}, ones[2] = {1, 1};
    CUtensorMapSwizzle swz = mode == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    CUresult r1 = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dimsA, str, boxA, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, dimsB, str, boxB, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { printf("encode failed %d %d\n", (int)r1, (int)r2); return 1; }
    for (int ubo = 0; ubo < 2; ++ubo) {
      if (mode == 0) {
        const int sbos[2] = {1024, 1280};
        for (int si = 0; si < 2; ++si)
          for (int shift = 0; shift <= 22; shift += (shift < 12 ? 1 : 10)) {
            probe_kernel<<<1, 128, smem>>>(tmA, tmB, dO, 0, shift, sbos[si], ubo);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
            double worst = 0;
            for (int m = 0; m < 128; ++m) {
              const int row = shift + (m / 8) * (sbos[si] / 128) + (m % 8);
              for (int n = 0; n < 32; ++n) {
                double ref = 0;
                for (int k = 0; k < 32; ++k) ref += (double)hA[row * 32 + k] * hB[n * 32 + k];
                worst = fmax(worst, fabs(ref - hO[m * 32 + n]));
              }
            }
            printf("mode 0 (K-major)  base_off=%d SBO=%4d shift=%2d rows: max err %.3e %s\n", ubo, sbos[si], shift, worst, worst < 1e-3 ? "OK" : "WRONG");
          }
      } else {
        for (int shift = 0; shift <= 10; ++shift) {
          probe_kernel<<<1, 128, smem>>>(tmA, tmB, dO, 1, shift, 512, ubo);
          CK(cudaDeviceSynchronize());
          CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
          double worst = 0;
          for (int m = 0; m < 32; ++m)
            for (int n = 0; n < 32; ++n) {
              double ref = 0;
              for (int p = 0; p < 8; ++p) ref += (double)hA[(p + shift) * 32 + m] * hB[p * 32 + n];
              worst = fmax(worst, fabs(ref - hO[m * 32 + n]));
            }
          printf("mode 1 (MN-major) base_off=%d shift=%2d rows: max err %.3e %s\n", ubo, shift, worst, worst < 1e-3 ? "OK" : "WRONG");
        }
      }
    }
  }
  return 0;
}
