// TMEM -> register (tcgen05.ld) and register -> TMEM (tcgen05.st) bandwidth probe (run on a B200).
// The mLSTM kernels read every accumulator back through tcgen05.ld; whether that path or instruction issue bounds
// a tile decides what is worth fusing.  Prints cycles per 4 KB (32 lanes x 32 columns x 32 bit) transfer for
// 1 / 4 / 8 / 16 concurrently loading warps (1 = one quadrant, 4 = one warp per quadrant, 8 = two per quadrant).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tmem_rate tmem_rate.cu
#include <stdio.h>
#include <stdlib.h>

#include "../../xlstm_yolo_clean_b200/csrc/sm100.cuh"

using namespace sm100;

template <int MODE>  // 0: ld x32, 1: ld x16 (two per 4 KB), 2: st x16 (two per 4 KB... 2 KB each), 3: ld x32 + 32 FMAs on the result
__global__ void __launch_bounds__(512) probe(int active_warps, int reps, long long* out, float* sink) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  float acc = 0.f;
  long long t0 = 0, t1 = 0;
  __syncthreads();
  if (warp < active_warps) {
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t col = (uint32_t)((r * 32 + (warp >> 2) * 128) & 511);
      if (MODE == 0 || MODE == 3) {
        uint32_t v[32];
        tmem_ld32_nowait(tmem + lane_base + (col & 480), v);
        tmem_ld_wait();
        if (MODE == 3) {
#pragma unroll
          for (int j = 0; j < 32; ++j) acc = fmaf(__uint_as_float(v[j]), 1.0001f, acc);
        } else {
          acc += __uint_as_float(v[0]) + __uint_as_float(v[31]);
        }
      } else if (MODE == 1) {
        uint32_t a[16], b[16];
        tmem_ld16_nowait(tmem + lane_base + (col & 480), a);
        tmem_ld16_nowait(tmem + lane_base + (col & 480) + 16, b);
        tmem_ld_wait();
        acc += __uint_as_float(a[0]) + __uint_as_float(b[15]);
      } else {
        uint32_t a[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = r + j;
        tmem_st16(tmem + lane_base + (col & 480), a);
        tmem_st16(tmem + lane_base + (col & 480) + 16, a);
        tmem_st_wait();
      }
    }
    t1 = clock64();
  }
  if ((tid & 31) == 0 && warp < active_warps) out[warp] = t1 - t0;
  if (acc == 123.456f) sink[tid] = acc;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// pipelined variant: 4 loads in flight before the wait (what an epilogue that loads Hi and Hx together does)
__global__ void __launch_bounds__(512) probe_pipelined(int active_warps, int reps, long long* out, float* sink) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  float acc = 0.f;
  long long t0 = 0, t1 = 0;
  __syncthreads();
  if (warp < active_warps) {
    t0 = clock64();
    for (int r = 0; r < reps; r += 4) {
      uint32_t a[32], b[32], c[32], d[32];
      tmem_ld32_nowait(tmem + lane_base + 0, a);
      tmem_ld32_nowait(tmem + lane_base + 32, b);
      tmem_ld32_nowait(tmem + lane_base + 64, c);
      tmem_ld32_nowait(tmem + lane_base + 96, d);
      tmem_ld_wait();
      acc += __uint_as_float(a[0]) + __uint_as_float(b[31]) + __uint_as_float(c[7]) + __uint_as_float(d[9]);
    }
    t1 = clock64();
  }
  if ((tid & 31) == 0 && warp < active_warps) out[warp] = t1 - t0;
  if (acc == 123.456f) sink[tid] = acc;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  long long* out;
  float* sink;
  cudaMalloc(&out, 16 * sizeof(long long));
  cudaMalloc(&sink, 512 * sizeof(float));
  const int reps = 1024;
  const char* names[] = {"ld 32x32b.x32", "ld 2 x (32x32b.x16)", "st 2 x (32x32b.x16)", "ld x32 + 32 FFMA", "ld x32, 4 in flight"};
  for (int mode = 0; mode < 5; ++mode) {
    for (int aw : {1, 4, 8, 16}) {
      cudaMemset(out, 0, 16 * sizeof(long long));
      for (int it = 0; it < 2; ++it) {
        if (mode == 0) probe<0><<<1, 512>>>(aw, reps, out, sink);
        if (mode == 1) probe<1><<<1, 512>>>(aw, reps, out, sink);
        if (mode == 2) probe<2><<<1, 512>>>(aw, reps, out, sink);
        if (mode == 3) probe<3><<<1, 512>>>(aw, reps, out, sink);
        if (mode == 4) probe_pipelined<<<1, 512>>>(aw, reps, out, sink);
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
      long long h[16];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int w = 0; w < aw; ++w) mx = h[w] > mx ? h[w] : mx;
      const double per = (double)mx / reps;  // cycles per 4 KB transfer of one warp
      printf("%-22s warps=%2d  %7.1f cycles per 4 KB per warp   SM-wide %6.1f B/clk   per quadrant %6.1f B/clk\n", names[mode], aw,
             per, 4096.0 * aw / per, 4096.0 * aw / per / (aw >= 4 ? 4 : 1));
    }
  }
  return 0;
}
