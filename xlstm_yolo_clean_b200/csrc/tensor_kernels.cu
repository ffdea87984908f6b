// Tensor-core kernel family of the sm_100a mLSTM chunkwise path: tcgen05.mma with TMEM
// accumulators, operands staged by TMA (128B swizzle), one persistent CTA per (batch, head)
// that walks the sequence in 128-token tiles and keeps the C / n / m state on chip
// (C: fp32 master copy in registers + bf16 MMA operand copy in shared memory).
//
// Forward, per 128-token tile k (math: SURVEY.md Appendix A; reference native/fw.py:29-221):
//   S      = Q K^T                      tcgen05  M128 N128 K64   (A, B K-major from TMA)
//   dC     = (abar.K)^T V               tcgen05  M64  N64  K128  (A, B MN-major)
//   P      = S . scale . exp(b_t - b_s + i_s - m_t), s <= t      (registers, from TMEM)
//   Hintra = P V                        tcgen05  M128 N64  K128  (A = P written swizzled)
//   Hinter = Q C_{k-1}                  tcgen05  M128 N64  K64   (B = bf16 copy of C)
//   h      = (Hintra + bbar.scale.Hinter) / (max(|den|, exp(-m_t)) + eps)   -> TMA store
//   C_k    = gbar C_{k-1} + dC          fp32 registers (the only sequential dependency)
// h and the final states do not depend on the tile length (m_t equals the step-recurrent
// stabiliser), so a 128-token tile is used although the API chunk size is 64.
#include <type_traits>

#include "common.cuh"
#include "sm100.cuh"

namespace mlstm {
namespace {

using namespace sm100;

constexpr int LT = 128;          // tokens per tile
constexpr int kTcThreads = 256;  // 8 warps: warp w owns rows 32*(w%4).., column half w/4
constexpr float kLog2e = 1.4426950408889634f;

struct TcFwParams {
  int B, NH, S, NT;  // NT = number of 128-token tiles
  float eps, scale;
  const void *ig, *fg;
  int64_t ig_sb, ig_sh, ig_ss, fg_sb, fg_sh, fg_ss;
  const float *c0, *n0, *m0;
  float *n_out, *m_out;
  float *c_last, *n_last, *m_last;
};

template <int D, int NSTAGE>
struct FwSmem {
  static constexpr int kTile = LT * 128;                 // one [128][64] 16-bit tile
  static constexpr int oQ = 0;                           // [NSTAGE] Q tiles
  static constexpr int oK = oQ + NSTAGE * kTile;
  static constexpr int oV = oK + NSTAGE * kTile;
  static constexpr int oKb = oV + NSTAGE * kTile;        // abar . K
  static constexpr int oP = oKb + kTile;                 // P: two K-halves; h staging aliases half 0
  static constexpr int oC = oP + 2 * kTile;              // bf16 copy of C (64 x 64), MMA B operand
  static constexpr int oSmall = oC + D * 128;
  // small region (floats): sb, sy, spm, sabar [LT each]; srs[2][LT]; sqn[2][LT]; sN[2][D]; scalars
  static constexpr int kSmallFloats = 4 * LT + 2 * LT + 2 * LT + 2 * D + 8;
  static constexpr int kBytes = oSmall + kSmallFloats * 4 + 1024 /*alignment slack*/;
};

template <typename T>
__device__ __forceinline__ uint32_t pack2(float a, float b);
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <typename T>
__device__ __forceinline__ float2 unpack2(uint32_t u);
template <>
__device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}
template <>
__device__ __forceinline__ float2 unpack2<__half>(uint32_t u) {
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}

// store 32 consecutive columns (col0 multiple of 32) of row `row` of a [128][64]-subtiled,
// 128B-swizzled 16-bit matrix; `base` points at the first subtile, subtiles are kTile apart.
template <typename T>
__device__ __forceinline__ void store_row32(uint8_t* base, int row, int col0, const float (&v)[32]) {
  uint8_t* tile = base + (col0 >> 6) * (LT * 128);
  const int c = col0 & 63;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 u;
    u.x = pack2<T>(v[8 * j + 0], v[8 * j + 1]);
    u.y = pack2<T>(v[8 * j + 2], v[8 * j + 3]);
    u.z = pack2<T>(v[8 * j + 4], v[8 * j + 5]);
    u.w = pack2<T>(v[8 * j + 6], v[8 * j + 7]);
    *reinterpret_cast<uint4*>(tile + swz128(row, c + 8 * j)) = u;
  }
}

template <typename T, int NSTAGE>
__global__ void __launch_bounds__(kTcThreads, NSTAGE == 1 ? 2 : 1)
tc_fw_d64(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
          const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapH, TcFwParams p) {
  constexpr int D = 64;
  constexpr bool kBf16 = std::is_same<T, __nv_bfloat16>::value;
  using SM = FwSmem<D, NSTAGE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* sb = (float*)(smem + SM::oSmall);  // chunk-local cumsum of logsigmoid(f)
  float* sy = sb + LT;                      // (i_s - b_s) * log2e
  float* spm = sy + LT;                     // prefix max of (i_s - b_s)
  float* sabar = spm + LT;                  // exp(a_t - m_next)
  float* srs = sabar + LT;                  // [2][LT] partial row sums of P
  float* sqn = srs + 2 * LT;                // [2][LT] partial q . n
  float* sN = sqn + 2 * LT;                 // [2][D]
  float* sscal = sN + 2 * D;                // g, amax
  __shared__ uint64_t bar_full[NSTAGE], bar_s, bar_dc, bar_h;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rb = warp & 3, ch = warp >> 2;
  const int row = rb * 32 + lane;  // tile row == TMEM lane of this thread
  const int bh = blockIdx.x, b = bh / p.NH, hh = bh % p.NH;
  const uint32_t lane_base = (uint32_t)(rb * 32) << 16;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) mbar_init(&bar_full[s], 1);
    mbar_init(&bar_s, 1);
    mbar_init(&bar_dc, 1);
    mbar_init(&bar_h, 1);
    fence_mbar_init();
    prefetch_tmap(&mapQ);
    prefetch_tmap(&mapK);
    prefetch_tmap(&mapV);
    prefetch_tmap(&mapH);
  }
  if (warp == 1) tmem_alloc<256>(&tmem_base_s);

  // state: fp32 master copy of C in registers of the threads with lane < 16:
  // row d = 16*rb + lane (M=64 TMEM layout), columns 32*ch .. 32*ch+31
  float Creg[32];
  const int drow = rb * 16 + (lane & 15);
  const bool owns_c = lane < 16;
#pragma unroll
  for (int j = 0; j < 32; ++j) Creg[j] = 0.f;
  if (p.c0 && owns_c) {
    const float* src = p.c0 + ((int64_t)bh * D + drow) * D + ch * 32;
#pragma unroll
    for (int j = 0; j < 32; ++j) Creg[j] = src[j];
  }
  if (owns_c) store_row32<T>(smem + SM::oC, drow, ch * 32, Creg);  // [64][64] tile: rows < 64 of subtile 0
  if (tid < D) sN[tid] = p.n0 ? p.n0[(int64_t)bh * D + tid] : 0.f;
  float m_run = p.m0 ? p.m0[bh] : 0.f;
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tS = tmem, tHi = tmem, tHx = tmem + 64, tDC = tmem + 128;

  constexpr uint32_t kStageBytes = 3 * SM::kTile;
  if (tid == 0) {
    for (int s = 0; s < NSTAGE && s < p.NT; ++s) {
      mbar_expect_tx(&bar_full[s], kStageBytes);
      tma_load_4d(smem + SM::oQ + s * SM::kTile, &mapQ, &bar_full[s], 0, s * LT, hh, b);
      tma_load_4d(smem + SM::oK + s * SM::kTile, &mapK, &bar_full[s], 0, s * LT, hh, b);
      tma_load_4d(smem + SM::oV + s * SM::kTile, &mapV, &bar_full[s], 0, s * LT, hh, b);
    }
  }

  const T* ip = (const T*)p.ig + b * p.ig_sb + hh * p.ig_sh;
  const T* fp = (const T*)p.fg + b * p.fg_sb + hh * p.fg_sh;
  int cur = 0;

  for (int c = 0; c < p.NT; ++c) {
    const int s = c % NSTAGE;
    const uint32_t par_full = (c / NSTAGE) & 1, par = c & 1;
    uint8_t* sQ = smem + SM::oQ + s * SM::kTile;
    uint8_t* sK = smem + SM::oK + s * SM::kTile;
    uint8_t* sV = smem + SM::oV + s * SM::kTile;
    uint8_t* sKb = smem + SM::oKb;
    uint8_t* sP = smem + SM::oP;
    uint8_t* sCc = smem + SM::oC;
    const float* sNc = sN + cur * D;
    float* sNn = sN + (cur ^ 1) * D;
    const int t0 = c * LT;
    const int n_valid = min(LT, p.S - t0);

    // ---- A. gates of this tile (one warp, warp-shuffle scans) -------------------------------
    if (warp == 2) {
      float amax;
      float g = chunk_gate_scan<T>(ip + (int64_t)t0 * p.ig_ss, fp + (int64_t)t0 * p.fg_ss, p.ig_ss, LT, n_valid, sb, sy,
                                   spm, &amax);
      if (lane == 0) {
        sscal[0] = g;
        sscal[1] = amax;
      }
    }
    // ---- B. S = Q K^T -----------------------------------------------------------------------
    if (warp == 0) {
      mbar_wait(&bar_full[s], par_full, 1);
      tc_fence_after_sync();
      if (elect_one()) {
        constexpr uint32_t idesc = umma_idesc(128, 128, false, false, kBf16);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_f16(tS, umma_smem_desc(smem_u32(sQ) + kk * 32, 0, 1024), umma_smem_desc(smem_u32(sK) + kk * 32, 0, 1024),
                   idesc, kk > 0);
        umma_commit(&bar_s);
      }
      __syncwarp();
    }
    __syncthreads();  // gates visible
    // ---- C. per-token factors; Kbar = abar . K ----------------------------------------------
    const float g = sscal[0];
    const float m_next = fmaxf(g + m_run, g + sscal[1]);  // fw.py:96-98
    const float gbar = __expf(g + m_run - m_next);        // fw.py:106
    const float b_t = sb[row], i_t = sy[row];             // sy holds raw i at this point
    const float m_t = b_t + fmaxf(m_run, spm[row]);       // fw.py:178-184
    __syncthreads();                                      // everyone has read raw i from sy
    if (ch == 0) sy[row] = (i_t - b_t) * kLog2e;
    mbar_wait(&bar_full[s], par_full, 2);  // K tile landed (generic-proxy read below)
    {
      const float ab = __expf(g - b_t + i_t - m_next);  // fw.py:102 (exp(-inf) = 0 for tail tokens)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t off = swz128(row, ch * 32 + 8 * j);
        uint4 u = *reinterpret_cast<const uint4*>(sK + off);
        float2 a0 = unpack2<T>(u.x), a1 = unpack2<T>(u.y), a2 = unpack2<T>(u.z), a3 = unpack2<T>(u.w);
        u.x = pack2<T>(a0.x * ab, a0.y * ab);
        u.y = pack2<T>(a1.x * ab, a1.y * ab);
        u.z = pack2<T>(a2.x * ab, a2.y * ab);
        u.w = pack2<T>(a3.x * ab, a3.y * ab);
        *reinterpret_cast<uint4*>(sKb + off) = u;
      }
    }
    // partial q . n_{k-1} over this thread's 32 columns
    {
      float qn = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 u = *reinterpret_cast<const uint4*>(sQ + swz128(row, ch * 32 + 8 * j));
        float2 a0 = unpack2<T>(u.x), a1 = unpack2<T>(u.y), a2 = unpack2<T>(u.z), a3 = unpack2<T>(u.w);
        const float* nn = sNc + ch * 32 + 8 * j;
        qn += a0.x * nn[0] + a0.y * nn[1] + a1.x * nn[2] + a1.y * nn[3] + a2.x * nn[4] + a2.y * nn[5] + a3.x * nn[6] +
              a3.y * nn[7];
      }
      sqn[ch * LT + row] = qn;
    }
    if (tid == 0) tma_store_wait_read<0>();  // previous tile's h store has left sP
    fence_proxy_async_smem();
    __syncthreads();
    // ---- D. dC = Kbar^T V --------------------------------------------------------------------
    if (warp == 0) {
      if (elect_one()) {
        constexpr uint32_t idesc = umma_idesc(64, 64, true, true, kBf16);
#pragma unroll
        for (int kk = 0; kk < LT / 16; ++kk)
          umma_f16(tDC, umma_smem_desc(smem_u32(sKb) + kk * 2048, LT * 128, 1024),
                   umma_smem_desc(smem_u32(sV) + kk * 2048, LT * 128, 1024), idesc, kk > 0);
        umma_commit(&bar_dc);
      }
      __syncwarp();
    }
    // ---- E. P = S . D (causal), row sums ------------------------------------------------------
    mbar_wait(&bar_s, par, 3);
    tc_fence_after_sync();
    {
      const float x_t = (b_t - m_t) * kLog2e + log2f(p.scale);
      float rs = 0.f;
      for (int u = 0; u < 4; ++u) {
        if ((u & 1) != ch) continue;  // warp-uniform
        float v[32];
        if (u <= rb) {
          tmem_ld32(tS + lane_base + u * 32, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float pv = v[j] * exp2f(x_t + sy[u * 32 + j]);
            pv = (u < rb || j <= lane) ? pv : 0.f;
            rs += pv;
            v[j] = pv;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        store_row32<T>(sP, row, u * 32, v);
      }
      srs[ch * LT + row] = rs;
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    // ---- F. Hintra = P V ; Hinter = Q C_{k-1} -------------------------------------------------
    if (warp == 0) {
      tc_fence_after_sync();
      if (elect_one()) {
        constexpr uint32_t idesc = umma_idesc(128, 64, false, true, kBf16);
#pragma unroll
        for (int kk = 0; kk < LT / 16; ++kk)
          umma_f16(tHi, umma_smem_desc(smem_u32(sP) + (kk / 4) * SM::kTile + (kk % 4) * 32, 0, 1024),
                   umma_smem_desc(smem_u32(sV) + kk * 2048, LT * 128, 1024), idesc, kk > 0);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_f16(tHx, umma_smem_desc(smem_u32(sQ) + kk * 32, 0, 1024),
                   umma_smem_desc(smem_u32(sCc) + kk * 2048, D * 128, 1024), idesc, kk > 0);
        umma_commit(&bar_h);
      }
      __syncwarp();
    }
    // ---- G. state update C_k = gbar C_{k-1} + dC (overlaps the H MMAs) ------------------------
    mbar_wait(&bar_dc, par, 4);
    tc_fence_after_sync();
    {
      float v[32];
      tmem_ld32(tDC + lane_base + ch * 32, v);  // M=64 layout: lanes 0-15 of each quadrant hold rows
      if (owns_c) {
#pragma unroll
        for (int j = 0; j < 32; ++j) Creg[j] = gbar * Creg[j] + v[j];
        if (ch == 0) {  // n_k = gbar n_{k-1} + column sums of Kbar (fw.py:116)
          float acc = 0.f;
          for (int t = 0; t < LT; ++t) acc += to_f32<T>(*reinterpret_cast<const T*>(sKb + swz128(t, drow)));
          sNn[drow] = gbar * sNc[drow] + acc;
        }
      }
    }
    // ---- H. epilogue -----------------------------------------------------------------------
    mbar_wait(&bar_h, par, 5);
    tc_fence_after_sync();
    if (owns_c) store_row32<T>(sCc, drow, ch * 32, Creg);  // the Q C_{k-1} MMA has finished reading the old copy
    {
      float hi[32], hx[32];
      tmem_ld32(tHi + lane_base + ch * 32, hi);
      tmem_ld32(tHx + lane_base + ch * 32, hx);
      const float bq = __expf(b_t + m_run - m_t) * p.scale;                       // fw.py:197-198
      const float den = bq * (sqn[row] + sqn[LT + row]) + srs[row] + srs[LT + row];  // fw.py:204-206
      const float nmax = fmaxf(fabsf(den), __expf(-m_t));                          // fw.py:208-210
      const float inv = 1.f / (nmax + p.eps);
#pragma unroll
      for (int j = 0; j < 32; ++j) hi[j] = (hi[j] + bq * hx[j]) * inv;  // fw.py:200-212
      store_row32<T>(sP, row, ch * 32, hi);  // h staging aliases P half 0 (PV MMA has completed)
      if (ch == 0 && row < n_valid) {
        p.n_out[(int64_t)bh * p.S + t0 + row] = nmax;
        p.m_out[(int64_t)bh * p.S + t0 + row] = m_t;
      }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      tma_store_4d(&mapH, sP, 0, t0, hh, b);
      tma_store_commit();
      // ---- I. refill this stage with tile c + NSTAGE ------------------------------------------
      const int cn = c + NSTAGE;
      if (cn < p.NT) {
        mbar_expect_tx(&bar_full[s], kStageBytes);
        tma_load_4d(sQ, &mapQ, &bar_full[s], 0, cn * LT, hh, b);
        tma_load_4d(sK, &mapK, &bar_full[s], 0, cn * LT, hh, b);
        tma_load_4d(sV, &mapV, &bar_full[s], 0, cn * LT, hh, b);
      }
    }
    m_run = m_next;
    cur ^= 1;
  }

  // final states (fw.py:302-309)
  if (p.c_last) {
    if (owns_c) {
      float* dst = p.c_last + ((int64_t)bh * D + drow) * D + ch * 32;
#pragma unroll
      for (int j = 0; j < 32; ++j) dst[j] = Creg[j];
    }
    if (tid < D) p.n_last[(int64_t)bh * D + tid] = sN[cur * D + tid];
    if (tid == 0) p.m_last[bh] = m_run;
  }
  if (tid == 0) tma_store_wait_all<0>();
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem);
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

template <typename T, int NSTAGE>
int launch_fw_d64(const mlstm_b200_fw_args& a, const TcFwParams& p, const CUtensorMap& mq, const CUtensorMap& mk,
                  const CUtensorMap& mv, const CUtensorMap& mh, cudaStream_t st) {
  using SM = FwSmem<64, NSTAGE>;
  auto kern = tc_fw_d64<T, NSTAGE>;
  MLSTM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kBytes));
  kern<<<p.B * p.NH, kTcThreads, SM::kBytes, st>>>(mq, mk, mv, mh, p);
  count_launch();
  MLSTM_CUDA_CHECK(cudaGetLastError());
  return 0;
}

bool tma_ok(const mlstm_b200_tensor& t) {
  return ((uintptr_t)t.ptr & 15) == 0 && t.stride[3] == 1 && (t.stride[0] % 8) == 0 && (t.stride[1] % 8) == 0 &&
         (t.stride[2] % 8) == 0;
}

}  // namespace

bool tensor_supported(const mlstm_b200_shape& s) {
  if (s.dtype != MLSTM_B200_BF16 && s.dtype != MLSTM_B200_F16) return false;
  if (s.DHQK != 64 || s.DHHV != 64) return false;
  if (s.chunk_size % 64 != 0 || s.S % 64 != 0) return false;
  return true;
}

size_t tensor_workspace_bytes(const mlstm_b200_shape& s, int backward) {
  return backward ? exact_workspace_bytes(s, 1) : 256;
}

int tensor_fw(const mlstm_b200_fw_args& a, cudaStream_t st) {
  const mlstm_b200_shape& s = a.shape;
  if (!tma_ok(a.q) || !tma_ok(a.k) || !tma_ok(a.v) || !tma_ok(a.h)) {
    set_error("tensor path needs 16-byte aligned q/k/v/h with strides that are multiples of 8 elements");
    return MLSTM_B200_EUNSUPPORTED;
  }
  const bool bf = s.dtype == MLSTM_B200_BF16;
  CUtensorMap mq, mk, mv, mh;
  int r = 0;
  r |= sm100_host::make_map_bhsd(&mq, a.q.ptr, bf, s.B, s.NH, s.S, s.DHQK, a.q.stride[0], a.q.stride[1], a.q.stride[2], LT);
  r |= sm100_host::make_map_bhsd(&mk, a.k.ptr, bf, s.B, s.NH, s.S, s.DHQK, a.k.stride[0], a.k.stride[1], a.k.stride[2], LT);
  r |= sm100_host::make_map_bhsd(&mv, a.v.ptr, bf, s.B, s.NH, s.S, s.DHHV, a.v.stride[0], a.v.stride[1], a.v.stride[2], LT);
  r |= sm100_host::make_map_bhsd(&mh, a.h.ptr, bf, s.B, s.NH, s.S, s.DHHV, a.h.stride[0], a.h.stride[1], a.h.stride[2], LT);
  if (r) {
    set_error("cuTensorMapEncodeTiled failed (%d)", r);
    return MLSTM_B200_ENODEVICE;
  }
  TcFwParams p{};
  p.B = s.B; p.NH = s.NH; p.S = s.S; p.NT = (s.S + LT - 1) / LT;
  p.eps = s.eps;
  p.scale = s.qk_scale > 0.f ? s.qk_scale : 1.f / sqrtf((float)s.DHQK);
  p.ig = a.i.ptr; p.ig_sb = a.i.stride[0]; p.ig_sh = a.i.stride[1]; p.ig_ss = a.i.stride[2];
  p.fg = a.f.ptr; p.fg_sb = a.f.stride[0]; p.fg_sh = a.f.stride[1]; p.fg_ss = a.f.stride[2];
  p.c0 = a.c_initial; p.n0 = a.n_initial; p.m0 = a.m_initial;
  p.n_out = a.n_out; p.m_out = a.m_out;
  p.c_last = a.c_last; p.n_last = a.n_last; p.m_last = a.m_last;
  const bool two_per_sm = (long)s.B * s.NH > num_sms();
  if (bf) {
    return two_per_sm ? launch_fw_d64<__nv_bfloat16, 1>(a, p, mq, mk, mv, mh, st)
                      : launch_fw_d64<__nv_bfloat16, 2>(a, p, mq, mk, mv, mh, st);
  }
  return two_per_sm ? launch_fw_d64<__half, 1>(a, p, mq, mk, mv, mh, st)
                    : launch_fw_d64<__half, 2>(a, p, mq, mk, mv, mh, st);
}

int tensor_bw(const mlstm_b200_bw_args& a, cudaStream_t st) { return exact_bw(a, st); }

}  // namespace mlstm
