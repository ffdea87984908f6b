// Issue-rate probe for tcgen05.mma (run on a B200): cycles per MMA for the operand layouts the mLSTM
// kernels use (K-major / MN-major, 128-byte and 64-byte swizzle, A from shared memory or from TMEM).
// Values are garbage (operands are whatever shared memory holds); only the timing matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o umma_rate umma_rate.cu
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../xlstm_yolo_clean_b200/csrc/sm100.cuh"

using namespace sm100;

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, bool accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"((uint32_t)accum)
      : "memory");
}

struct Cfg {
  int M, N;
  int a_mn, b_mn;    // 1: MN-major
  int a_rowb, b_rowb;  // 128 / 64 byte rows
  int ts;            // A from TMEM
  int reps;
  int hammer;        // other warps stream shared memory meanwhile
};

template <int REPS>
__global__ void __launch_bounds__(256) rate(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 4; i += 256) ((uint32_t*)smem)[i] = 0;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); stop = 0; }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  uint8_t* sA = smem;           // up to 32 KB
  uint8_t* sB = smem + 65536;   // up to 32 KB
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = umma_idesc(c.M, c.N, c.a_mn, c.b_mn, true);
    const uint32_t lta = c.a_rowb == 128 ? 2u : 4u, ltb = c.b_rowb == 128 ? 2u : 4u;
    const uint64_t da = umma_smem_desc_lt(smem_u32(sA), c.a_mn ? 128 * c.a_rowb : 0, 8 * c.a_rowb, lta);
    const uint64_t db = umma_smem_desc_lt(smem_u32(sB), c.b_mn ? 128 * c.b_rowb : 0, 8 * c.b_rowb, ltb);
    const uint32_t adv_a = c.a_mn ? 16 * c.a_rowb : 32, adv_b = c.b_mn ? 16 * c.b_rowb : 32;
    uint64_t da4[4], db4[4];
    uint32_t ta4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      da4[i] = umma_desc_advance(da, (c.a_mn ? i : (i % (c.a_rowb / 32))) * adv_a);
      db4[i] = umma_desc_advance(db, (c.b_mn ? i : (i % (c.b_rowb / 32))) * adv_b);
      ta4[i] = tmem + 256 + i * 8;
    }
    const long long t0 = clock64();
    if (c.ts) {
#pragma unroll
      for (int r = 0; r < REPS; ++r) umma_f16_ts(tmem, ta4[r & 3], db4[r & 3], idesc, r > 0);
    } else {
#pragma unroll
      for (int r = 0; r < REPS; ++r) umma_f16(tmem, da4[r & 3], db4[r & 3], idesc, r > 0);
    }
    umma_commit(&bar);
    const long long t1 = clock64();
    mbar_wait(&bar, 0, 1);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
    stop = 1;
  } else if (c.hammer && warp >= 4) {
    // stream 16-byte loads / stores over a private 32 KB region
    uint4* base = (uint4*)(smem + 131072) + (tid - 128);
    uint4 acc = make_uint4(0, 0, 0, 0);
    while (!stop) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint4 v = base[i * 128];
        acc.x ^= v.x; acc.y += v.y;
        base[i * 128 + 1024] = acc;
      }
    }
    if (acc.x == 0x12345) out[2] = acc.y;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int run(const char* name, Cfg c) {
  long long* d;
  cudaMalloc(&d, 64);
  cudaMemset(d, 0, 64);
  cudaFuncSetAttribute(rate<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
  cudaFuncSetAttribute(rate<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
  long long h[2][2];
  for (int pass = 0; pass < 2; ++pass) {
    Cfg cc = c;
    if (pass) rate<192><<<1, 256, 170 * 1024>>>(cc, d); else rate<64><<<1, 256, 170 * 1024>>>(cc, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h[pass], d, 16, cudaMemcpyDeviceToHost);
  }
  printf("%-58s issue %6.1f clk/mma   complete %6.1f clk/mma   (64 reps: %lld clk total)\n", name,
         (double)(h[1][0] - h[0][0]) / 128.0, (double)(h[1][1] - h[0][1]) / 128.0, h[0][1]);
  cudaFree(d);
  return 0;
}

int main() {
  int bad = 0;
  for (int hammer = 0; hammer < 2; ++hammer) {
    printf(hammer ? "--- with 4 warps streaming shared memory ---\n" : "--- tensor pipe alone ---\n");
    bad += run("S    M128 N128 A:K/128  B:K/128", Cfg{128, 128, 0, 0, 128, 128, 0, 0, hammer});
    bad += run("     M128 N256 A:K/128  B:K/128", Cfg{128, 256, 0, 0, 128, 128, 0, 0, hammer});
    bad += run("PV   M128 N64  A:K/128  B:MN/128", Cfg{128, 64, 0, 1, 128, 128, 0, 0, hammer});
    bad += run("dV1  M128 N64  A:MN/128 B:MN/128", Cfg{128, 64, 1, 1, 128, 128, 0, 0, hammer});
    bad += run("dQb  M128 N64  A:K/128  B:K/128", Cfg{128, 64, 0, 0, 128, 128, 0, 0, hammer});
    bad += run("     M128 N64  A:MN/128 B:K/128", Cfg{128, 64, 1, 0, 128, 128, 0, 0, hammer});
    bad += run("dC   M64  N64  A:MN/128 B:MN/128", Cfg{64, 64, 1, 1, 128, 128, 0, 0, hammer});
    bad += run("     M64  N64  A:K/128  B:MN/128", Cfg{64, 64, 0, 1, 128, 128, 0, 0, hammer});
    bad += run("     M128 N128 A:MN/128 B:MN/128", Cfg{128, 128, 1, 1, 128, 128, 0, 0, hammer});
    bad += run("TS   M128 N64  A:TMEM   B:MN/128", Cfg{128, 64, 0, 1, 128, 128, 1, 0, hammer});
    bad += run("TS   M128 N64  A:TMEM   B:K/128", Cfg{128, 64, 0, 0, 128, 128, 1, 0, hammer});
    bad += run("TS   M128 N128 A:TMEM   B:K/128", Cfg{128, 128, 0, 0, 128, 128, 1, 0, hammer});
    bad += run("d32 S   M128 N128 A:K/64   B:K/64", Cfg{128, 128, 0, 0, 64, 64, 0, 0, hammer});
    bad += run("d32 PV  M128 N32  A:K/128  B:MN/64", Cfg{128, 32, 0, 1, 128, 64, 0, 0, hammer});
    bad += run("d32 dV1 M128 N32  A:MN/128 B:MN/64", Cfg{128, 32, 1, 1, 128, 64, 0, 0, hammer});
    bad += run("d32 dQb M128 N32  A:K/64   B:K/64", Cfg{128, 32, 0, 0, 64, 64, 0, 0, hammer});
    bad += run("d32 dC  M64  N32  A:MN/64  B:MN/64", Cfg{64, 32, 1, 1, 64, 64, 0, 0, hammer});
    bad += run("d32 TS  M128 N32  A:TMEM   B:MN/64", Cfg{128, 32, 0, 1, 128, 64, 1, 0, hammer});
  }
  return bad ? 1 : 0;
}
