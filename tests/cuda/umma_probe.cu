// Descriptor probe for the tcgen05 building blocks the mLSTM kernels rely on (run on a B200):
// TMA 128B-swizzled loads, K-/MN-major UMMA operands, M=64 and M=128 TMEM layouts, manual
// swizzled operand writes and a TMA store.  Prints one PASS/FAIL line per variant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o umma_probe umma_probe.cu
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#include "../../xlstm_yolo_clean_b200/csrc/sm100.cuh"

using namespace sm100;
typedef __nv_bfloat16 bf16;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

// A: K-major -> global [M][K]; MN-major -> global [K][M].  Same for B with N.
template <int M, int N, int K, bool A_MN, bool B_MN, bool MANUAL_A, bool TMA_OUT, bool A_F16 = false, bool B_F16 = false>
__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap mapA,
                                             const __grid_constant__ CUtensorMap mapB,
                                             const __grid_constant__ CUtensorMap mapO, const bf16* __restrict__ gA,
                                             float* __restrict__ out, bf16* __restrict__ out16) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                      // M*K*2 bytes
  uint8_t* sB = sA + M * K * 2;            // N*K*2 bytes
  uint8_t* sO = sB + N * K * 2;            // 128 x 64 bf16 staging for the TMA store
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<128>(&tmem_base_s);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (tid == 0) {
    uint32_t bytes = N * K * 2 + (MANUAL_A ? 0 : M * K * 2);
    mbar_expect_tx(&bar_full, bytes);
    if (!MANUAL_A) {
      if (A_MN) { for (int mb = 0; mb < M / 64; ++mb) tma_load_4d(sA + mb * K * 128, &mapA, &bar_full, mb * 64, 0, 0, 0); }
      else      { for (int kb = 0; kb < K / 64; ++kb) tma_load_4d(sA + kb * M * 128, &mapA, &bar_full, kb * 64, 0, 0, 0); }
    }
    if (B_MN) { for (int nb = 0; nb < N / 64; ++nb) tma_load_4d(sB + nb * K * 128, &mapB, &bar_full, nb * 64, 0, 0, 0); }
    else      { for (int kb = 0; kb < K / 64; ++kb) tma_load_4d(sB + kb * N * 128, &mapB, &bar_full, kb * 64, 0, 0, 0); }
  }
  if (MANUAL_A) {  // K-major A written by threads with the swizzle formula (row = thread)
    static_assert(!MANUAL_A || (!A_MN && M == 128), "manual A: K-major, M=128");
    for (int c = 0; c < K; ++c) {
      *(bf16*)(sA + (c / 64) * M * 128 + swz128(tid, c % 64)) = gA[tid * K + c];
    }
    fence_proxy_async_smem();
  }
  __syncthreads();
  mbar_wait(&bar_full, 0, 101);
  tc_fence_after_sync();

  if (warp == 0 && elect_one()) {
    constexpr uint32_t idesc = umma_idesc(M, N, A_MN, B_MN, !A_F16, !B_F16);
    for (int kk = 0; kk < K / 16; ++kk) {
      uint64_t ad, bd;
      if (A_MN) ad = umma_smem_desc(smem_u32(sA) + kk * 2048, K * 128, 1024);
      else      ad = umma_smem_desc(smem_u32(sA) + (kk / 4) * M * 128 + (kk % 4) * 32, 0, 1024);
      if (B_MN) bd = umma_smem_desc(smem_u32(sB) + kk * 2048, K * 128, 1024);
      else      bd = umma_smem_desc(smem_u32(sB) + (kk / 4) * N * 128 + (kk % 4) * 32, 0, 1024);
      umma_f16(tmem, ad, bd, idesc, kk > 0);
    }
    umma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0, 102);
  tc_fence_after_sync();

  // read back: M=128 -> lane = row; M=64 -> row r lives in lane (r%16) + 32*(r/16)
  int row = (M == 128) ? tid : (lane < 16 ? warp * 16 + lane : -1);
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    if (row >= 0) {
      for (int j = 0; j < 32; ++j) {
        out[row * N + c0 + j] = v[j];
        if (TMA_OUT && c0 + j < 64) *(bf16*)(sO + swz128(row, c0 + j)) = __float2bfloat16_rn(v[j]);
      }
    }
  }
  if (TMA_OUT) {
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tma_store_4d(&mapO, sO, 0, 0, 0, 0);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tmem);
}

static float frand() { return (float)(rand() % 2001 - 1000) / 1000.f; }

// 16-bit pattern of x in the operand's format, and the value that pattern stands for
template <bool F16>
static bf16 enc16(float x, float* back) {
  if (F16) {
    __half h = __float2half(x);
    *back = __half2float(h);
    bf16 r;
    memcpy(&r, &h, 2);
    return r;
  }
  bf16 r = __float2bfloat16(x);
  *back = __bfloat162float(r);
  return r;
}

template <int M, int N, int K, bool A_MN, bool B_MN, bool MANUAL_A, bool TMA_OUT, bool A_F16 = false, bool B_F16 = false>
int run(const char* name) {
  std::vector<bf16> hA(M * K), hB(N * K);
  std::vector<float> fA(M * K), fB(N * K);  // logical A[m][k], B[n][k]
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) { bf16 x = enc16<A_F16>(frand(), &fA[m * K + k]); hA[A_MN ? k * M + m : m * K + k] = x; }
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { bf16 x = enc16<B_F16>(frand(), &fB[n * K + k]); hB[B_MN ? k * N + n : n * K + k] = x; }
  bf16 *dA, *dB, *dO16; float* dO;
  CK(cudaMalloc(&dA, M * K * 2)); CK(cudaMalloc(&dB, N * K * 2)); CK(cudaMalloc(&dO, M * N * 4)); CK(cudaMalloc(&dO16, 128 * 64 * 2));
  CK(cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dO, 0xff, M * N * 4)); CK(cudaMemset(dO16, 0, 128 * 64 * 2));
  CUtensorMap mA, mB, mO;
  int rA = A_MN ? sm100_host::make_map_bhsd(&mA, dA, true, 1, 1, K, M, (int64_t)K * M, (int64_t)K * M, M, K)
                : sm100_host::make_map_bhsd(&mA, dA, true, 1, 1, M, K, (int64_t)K * M, (int64_t)K * M, K, M);
  int rB = B_MN ? sm100_host::make_map_bhsd(&mB, dB, true, 1, 1, K, N, (int64_t)K * N, (int64_t)K * N, N, K)
                : sm100_host::make_map_bhsd(&mB, dB, true, 1, 1, N, K, (int64_t)K * N, (int64_t)K * N, K, N);
  int rO = sm100_host::make_map_bhsd(&mO, dO16, true, 1, 1, 128, 64, 128 * 64, 128 * 64, 64, 128);
  if (rA || rB || rO) { printf("%s: tensor map encode failed %d %d %d\n", name, rA, rB, rO); return 1; }
  size_t smem = (size_t)M * K * 2 + (size_t)N * K * 2 + 128 * 128 + 2048;
  auto kern = probe<M, N, K, A_MN, B_MN, MANUAL_A, TMA_OUT, A_F16, B_F16>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int* hdbg = nullptr; int* ddbg = nullptr;
  CK(cudaHostAlloc(&hdbg, 64, cudaHostAllocMapped)); hdbg[0] = 0; hdbg[1] = 0;
  CK(cudaHostGetDevicePointer(&ddbg, hdbg, 0));
  CK(cudaMemcpyToSymbol(g_dbg, &ddbg, sizeof(ddbg)));
  kern<<<1, 128, smem>>>(mA, mB, mO, dA, dO, dO16);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: FAIL kernel error %s (wait tag %d, thread %d)\n", name, cudaGetErrorString(e), hdbg[0], hdbg[1]); return 2; }
  std::vector<float> hO(M * N); std::vector<bf16> hO16(128 * 64);
  CK(cudaMemcpy(hO.data(), dO, M * N * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hO16.data(), dO16, 128 * 64 * 2, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxerr16 = 0;
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    double ref = 0; for (int k = 0; k < K; ++k) ref += (double)fA[m * K + k] * fB[n * K + k];
    double d = fabs(ref - hO[m * N + n]); if (!(d <= maxerr)) maxerr = d;  // NaN-propagating
    if (TMA_OUT && n < 64) { double d16 = fabs(ref - __bfloat162float(hO16[m * 64 + n])); if (!(d16 <= maxerr16)) maxerr16 = d16; }
  }
  bool ok = maxerr < 1e-2 && (!TMA_OUT || maxerr16 < 0.3);
  printf("%s: %s  max|err|=%.3e%s\n", name, ok ? "PASS" : "FAIL", maxerr, TMA_OUT ? (maxerr16 < 0.3 ? "  tma_store PASS" : "  tma_store FAIL") : "");
  cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dO16);
  return ok ? 0 : 3;
}

int time_main();
int sw64_main();
int main(int argc, char** argv) {
  setvbuf(stdout, NULL, _IONBF, 0);
  if (argc > 1 && argv[1][0] == 't') return time_main();
  if (argc > 1 && argv[1][0] == 's') return sw64_main();
  int bad = 0;
  if (argc > 1 && argv[1][0] == 'm') {  // `umma_probe mixed`: one operand fp16, the other bf16, in one kind::f16 instruction
    bad += run<128, 128, 64, false, false, false, false, true, false>("mixed  A:f16 (K)  B:bf16 (K)   M128 N128 K64 ") != 0;
    bad += run<128, 128, 64, false, false, false, false, false, true>("mixed  A:bf16 (K) B:f16 (K)    M128 N128 K64 ") != 0;
    bad += run<128, 64, 128, false, true, false, false, false, true>("mixed  A:bf16 (K) B:f16 (MN)   M128 N64  K128") != 0;
    bad += run<64, 64, 128, true, true, false, false, false, true>("mixed  A:bf16 (MN) B:f16 (MN)  M64  N64  K128") != 0;
    bad += run<128, 64, 128, true, true, false, false, true, false>("mixed  A:f16 (MN) B:bf16 (MN)  M128 N64  K128") != 0;
    bad += run<128, 128, 64, false, false, false, false, true, true>("plain  A:f16 (K)  B:f16 (K)    M128 N128 K64 ") != 0;
    printf("%d variant(s) failed\n", bad);
    return bad ? 1 : 0;
  }
  bad += run<128, 128, 64, false, false, false, false>("S=QK^T      M128 N128 K64  A:K  B:K ") != 0;
  bad += run<128, 64, 128, false, true, false, true>("PV          M128 N64  K128 A:K  B:MN +tma_store") != 0;
  bad += run<128, 64, 128, false, true, true, false>("PV manualA  M128 N64  K128 A:K* B:MN") != 0;
  bad += run<128, 64, 64, false, true, false, false>("QC          M128 N64  K64  A:K  B:MN") != 0;
  bad += run<64, 64, 128, true, true, false, false>("dC=K^TV     M64  N64  K128 A:MN B:MN") != 0;
  bad += run<128, 64, 128, true, true, false, false>("dS^T Q      M128 N64  K128 A:MN B:MN") != 0;
  bad += run<128, 64, 64, false, false, false, false>("dH C^T      M128 N64  K64  A:K  B:K ") != 0;
  bad += run<128, 128, 128, false, false, false, false>("S d128      M128 N128 K128 A:K  B:K ") != 0;
  bad += run<128, 128, 128, true, true, false, false>("d128 MN     M128 N128 K128 A:MN B:MN") != 0;
  bad += run<64, 128, 64, true, true, false, false>("M64 N128    M64  N128 K64  A:MN B:MN") != 0;
  printf("%d variant(s) failed\n", bad);
  return bad ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// timing mode: `umma_probe time` -- cycles per tcgen05.mma for the operand layouts the kernels use
// (a batch of REP x (K/16) MMAs issued back to back, clock64 from first issue to commit arrival).
// ------------------------------------------------------------------------------------------
template <int M, int N, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(128) time_mma(long long* out, int rep) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                 // 128 x 128 bf16 (two [128][64] subtiles)
  uint8_t* sB = smem + 2 * 128 * 128;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 2 * 2 * 128 * 128 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<256>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  constexpr int K = 128;
  if (warp == 0 && elect_one()) {
    constexpr uint32_t idesc = umma_idesc(M, N, A_MN, B_MN, true);
    const uint64_t dA = umma_smem_desc(smem_u32(sA), A_MN ? 128 * 128 : 0, 1024);
    const uint64_t dB = umma_smem_desc(smem_u32(sB), B_MN ? 128 * 128 : 0, 1024);
    long long t0 = clock64();
    for (int r = 0; r < rep; ++r) {
#pragma unroll
      for (int kk = 0; kk < K / 16; ++kk) {
        const uint64_t a = umma_desc_advance(dA, A_MN ? kk * 2048 : (kk / 4) * 128 * 128 + (kk % 4) * 32);
        const uint64_t b = umma_desc_advance(dB, B_MN ? kk * 2048 : (kk / 4) * 128 * 128 + (kk % 4) * 32);
        umma_f16(tmem, a, b, idesc, true);
      }
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0, 7);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

template <int M, int N, bool A_MN, bool B_MN>
void run_time(const char* name) {
  long long *d, h[2];
  cudaMalloc(&d, 16);
  auto kern = time_mma<M, N, A_MN, B_MN>;
  size_t smem = 4 * 128 * 128 + 2048;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int rep = 8;
  for (int it = 0; it < 2; ++it) {
    kern<<<1, 128, smem>>>(d, rep);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: error %s\n", name, cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%s: %d MMAs  issue %.1f cyc/MMA   complete %.1f cyc/MMA\n", name, rep * 8, h[0] / (rep * 8.0), h[1] / (rep * 8.0));
  cudaFree(d);
}

int time_main() {
  run_time<128, 128, false, false>("M128 N128 A:K  B:K ");
  run_time<128, 64, false, false>("M128 N64  A:K  B:K ");
  run_time<128, 64, false, true>("M128 N64  A:K  B:MN");
  run_time<128, 64, true, true>("M128 N64  A:MN B:MN");
  run_time<128, 64, true, false>("M128 N64  A:MN B:K ");
  run_time<64, 64, true, true>("M64  N64  A:MN B:MN");
  run_time<64, 64, false, false>("M64  N64  A:K  B:K ");
  run_time<128, 128, true, true>("M128 N128 A:MN B:MN");
  run_time<128, 256, false, false>("M128 N256 A:K  B:K ");
  return 0;
}

// ------------------------------------------------------------------------------------------
// `umma_probe sw64`: operand layouts for head dim 32 (64-byte swizzle) mixed with 128-byte ones.
// Operands are written by threads with the swizzle formulas (so the formulas themselves are tested);
// one variant loads A through TMA SWIZZLE_64B.
// ------------------------------------------------------------------------------------------
struct OpLayout {
  int mn, k;        // logical extent
  int mn_major;     // 0: K-major, 1: MN-major
  int rowb;         // swizzle / row width in bytes: 64 or 128
};
__host__ __device__ inline uint32_t op_offset(const OpLayout& L, int mn, int k) {
  const int epr = L.rowb / 2;  // elements per row
  if (!L.mn_major) {
    const int sub = k / epr, kc = k % epr;
    const int sw = L.rowb == 128 ? (mn & 7) : ((mn >> 1) & 3);
    return sub * L.mn * L.rowb + mn * L.rowb + ((((kc >> 3) ^ sw)) << 4) + ((kc & 7) << 1);
  } else {
    const int blk = mn / epr, mc = mn % epr;
    const int sw = L.rowb == 128 ? (k & 7) : ((k >> 1) & 3);
    return blk * L.k * L.rowb + k * L.rowb + ((((mc >> 3) ^ sw)) << 4) + ((mc & 7) << 1);
  }
}
__device__ inline uint64_t op_desc(const OpLayout& L, uint32_t base, int kk /*k-step of 16*/, int dup_blocks) {
  const uint32_t lt = L.rowb == 128 ? 2u : 4u;
  uint32_t start, lbo, sbo = 8 * L.rowb;
  if (!L.mn_major) {
    const int epr = L.rowb / 2, k0 = kk * 16;
    start = base + (k0 / epr) * L.mn * L.rowb + (k0 % epr) * 2;
    lbo = 0;
  } else {
    start = base + kk * 16 * L.rowb;
    lbo = dup_blocks ? 0 : L.k * L.rowb;
  }
  uint64_t d = 0;
  d |= (uint64_t)((start & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)lt << 61;
  return d;
}

__global__ void __launch_bounds__(128) probe_sw(OpLayout LA, OpLayout LB, int M, int N, int dupA, const bf16* gA,
                                                const bf16* gB, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + 32768;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < LA.mn * LA.k; e += 128) *(bf16*)(sA + op_offset(LA, e / LA.k, e % LA.k)) = gA[e];
  for (int e = tid; e < LB.mn * LB.k; e += 128) *(bf16*)(sB + op_offset(LB, e / LB.k, e % LB.k)) = gB[e];
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<128>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = umma_idesc(M, N, LA.mn_major, LB.mn_major, true);
    for (int kk = 0; kk < LA.k / 16; ++kk)
      umma_f16(tmem, op_desc(LA, smem_u32(sA), kk, dupA), op_desc(LB, smem_u32(sB), kk, 0), idesc, kk > 0);
    umma_commit(&bar);
  }
  __syncwarp();
  mbar_wait(&bar, 0, 103);
  tc_fence_after_sync();
  int row = (M == 128) ? tid : (lane < 16 ? warp * 16 + lane : -1);
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    if (row >= 0)
      for (int j = 0; j < 32 && c0 + j < N; ++j) out[row * N + c0 + j] = v[j];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tmem);
}

int run_sw(const char* name, int M, int N, int K, int a_mn, int a_rowb, int b_mn, int b_rowb, int validM) {
  // logical A[M'][K] with M' = validM rows of real data (the MMA's M may be larger: duplicated block)
  OpLayout LA{validM, K, a_mn, a_rowb}, LB{N, K, b_mn, b_rowb};
  std::vector<bf16> hA(validM * K), hB(N * K);
  std::vector<float> fA(validM * K), fB(N * K);
  for (int i = 0; i < validM * K; ++i) { bf16 x = __float2bfloat16(frand()); hA[i] = x; fA[i] = __bfloat162float(x); }
  for (int i = 0; i < N * K; ++i) { bf16 x = __float2bfloat16(frand()); hB[i] = x; fB[i] = __bfloat162float(x); }
  bf16 *dA, *dB; float* dO;
  cudaMalloc(&dA, validM * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dO, M * N * 4);
  cudaMemcpy(dA, hA.data(), validM * K * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
  cudaMemset(dO, 0xff, M * N * 4);
  size_t smem = 65536 + 2048;
  cudaFuncSetAttribute(probe_sw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_sw<<<1, 128, smem>>>(LA, LB, M, N, validM < M, dA, dB, dO);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: FAIL kernel error %s\n", name, cudaGetErrorString(e)); return 2; }
  std::vector<float> hO(M * N);
  cudaMemcpy(hO.data(), dO, M * N * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int m = 0; m < validM; ++m) for (int n = 0; n < N; ++n) {
    double ref = 0; for (int k = 0; k < K; ++k) ref += (double)fA[m * K + k] * fB[n * K + k];
    double d = fabs(ref - hO[m * N + n]); if (!(d <= maxerr)) maxerr = d;
  }
  printf("%s: %s  max|err|=%.3e\n", name, maxerr < 1e-2 ? "PASS" : "FAIL", maxerr);
  cudaFree(dA); cudaFree(dB); cudaFree(dO);
  return maxerr < 1e-2 ? 0 : 1;
}

// A operand in TMEM (TS mode): every thread writes its row of A with tcgen05.st (two bf16 per column),
// B from shared memory (MN-major, 128- or 64-byte rows).  M = 128, K = 128.
__global__ void __launch_bounds__(128) probe_ts(OpLayout LB, int N, const bf16* gA, const bf16* gB, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sB = smem;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr int K = 128;
  for (int e = tid; e < LB.mn * LB.k; e += 128) *(bf16*)(sB + op_offset(LB, e / LB.k, e % LB.k)) = gB[e];
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<256>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tA = tmem + 128;  // 64 columns of packed A
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  for (int c0 = 0; c0 < K / 2; c0 += 16) {
    uint32_t r[16];
    for (int j = 0; j < 16; ++j) {
      const uint16_t lo = *(const uint16_t*)&gA[tid * K + 2 * (c0 + j)];
      const uint16_t hi = *(const uint16_t*)&gA[tid * K + 2 * (c0 + j) + 1];
      r[j] = (uint32_t)lo | ((uint32_t)hi << 16);
    }
    tmem_st16(tA + lane_base + c0, r);
  }
  tmem_st_wait();
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0 && elect_one()) {
    tc_fence_after_sync();
    const uint32_t idesc = umma_idesc(128, N, false, LB.mn_major, true);
    for (int kk = 0; kk < K / 16; ++kk) umma_f16_ts(tmem, tA + kk * 8, op_desc(LB, smem_u32(sB), kk, 0), idesc, kk > 0);
    umma_commit(&bar);
  }
  __syncwarp();
  mbar_wait(&bar, 0, 104);
  tc_fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + lane_base + c0, v);
    for (int j = 0; j < 32 && c0 + j < N; ++j) out[tid * N + c0 + j] = v[j];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

int run_ts(const char* name, int N, int b_mn, int b_rowb) {
  const int M = 128, K = 128;
  OpLayout LB{N, K, b_mn, b_rowb};
  std::vector<bf16> hA(M * K), hB(N * K);
  std::vector<float> fA(M * K), fB(N * K);
  for (int i = 0; i < M * K; ++i) { bf16 x = __float2bfloat16(frand()); hA[i] = x; fA[i] = __bfloat162float(x); }
  for (int i = 0; i < N * K; ++i) { bf16 x = __float2bfloat16(frand()); hB[i] = x; fB[i] = __bfloat162float(x); }
  bf16 *dA, *dB; float* dO;
  cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dO, M * N * 4);
  cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
  cudaMemset(dO, 0xff, M * N * 4);
  size_t smem = 32768 + 2048;
  cudaFuncSetAttribute(probe_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_ts<<<1, 128, smem>>>(LB, N, dA, dB, dO);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: FAIL kernel error %s\n", name, cudaGetErrorString(e)); return 2; }
  std::vector<float> hO(M * N);
  cudaMemcpy(hO.data(), dO, M * N * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    double ref = 0; for (int k = 0; k < K; ++k) ref += (double)fA[m * K + k] * fB[n * K + k];
    double d = fabs(ref - hO[m * N + n]); if (!(d <= maxerr)) maxerr = d;
  }
  printf("%s: %s  max|err|=%.3e\n", name, maxerr < 1e-2 ? "PASS" : "FAIL", maxerr);
  cudaFree(dA); cudaFree(dB); cudaFree(dO);
  return maxerr < 1e-2 ? 0 : 1;
}

int sw64_main() {
  int bad = 0;
  bad += run_sw("S d32      M128 N128 K32  A:K/64  B:K/64 ", 128, 128, 32, 0, 64, 0, 64, 128);
  bad += run_sw("PV d32     M128 N32  K128 A:K/128 B:MN/64", 128, 32, 128, 0, 128, 1, 64, 128);
  bad += run_sw("QC d32     M128 N32  K32  A:K/64  B:MN/64", 128, 32, 32, 0, 64, 1, 64, 128);
  bad += run_sw("dC d32     M64  N32  K128 A:MN/64(dup) B:MN/64", 64, 32, 128, 1, 64, 1, 64, 32);
  bad += run_sw("dSK d32    M128 N32  K128 A:K/128 B:MN/64", 128, 32, 128, 0, 128, 1, 64, 128);
  bad += run_sw("SbT dH d32 M128 N32  K128 A:MN/128 B:MN/64", 128, 32, 128, 1, 128, 1, 64, 128);
  bad += run_sw("dH CT d32  M128 N32  K32  A:K/64  B:K/64 ", 128, 32, 32, 0, 64, 0, 64, 128);
  bad += run_sw("ref d64    M128 N64  K128 A:MN/128 B:MN/128", 128, 64, 128, 1, 128, 1, 128, 128);
  bad += run_ts("TS PV d64  M128 N64  K128 A:TMEM   B:MN/128", 64, 1, 128);
  bad += run_ts("TS PV d32  M128 N32  K128 A:TMEM   B:MN/64", 32, 1, 64);
  bad += run_ts("TS    d64  M128 N64  K128 A:TMEM   B:K/128", 64, 0, 128);
  printf("%d sw64 variant(s) failed\n", bad);
  return bad ? 1 : 0;
}
