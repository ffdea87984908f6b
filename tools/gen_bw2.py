import re
p='/root/repo/xlstm_yolo_clean_b200/csrc/tensor_kernels.cu'
s=open(p).read()
# remove a previous generation of bw2, if any
if "// >>> BW2 BEGIN" in s:
    s=s[:s.index("// >>> BW2 BEGIN")]+s[s.index("// <<< BW2 END")+len("// <<< BW2 END\n"):]
a=s.index("template <typename T, int D, bool REV>\n__global__ void __launch_bounds__(kTcThreads, 1)\ntc_bw(")
b=s.index("long long* g_prof = nullptr;")
k=s[a:b]
def rep(old,new,count=1):
    global k
    assert k.count(old)==count, (old[:80],k.count(old))
    k=k.replace(old,new)
def cut(start,end,new):
    global k
    i=k.index(start); j=k.index(end,i)
    k=k[:i]+new+k[j:]

rep("tc_bw(","tc_bw2(")
rep("  using SM = BwSmem<D>;","  using SM = Bw2Smem<D>;")
rep("  uint8_t* sSb = smem + SM::oSb;\n","")
# TMEM addresses
cut("  const uint32_t tS = tmem + SM::cS, tdSb = tmem + SM::cdSb;","  const T* ip = (const T*)p.ig",
'''  const uint32_t tST = tmem + SM::cST, tdST = tmem + SM::cdST;
  const uint32_t tdV = tmem + SM::cdV, tdK = tmem + SM::cdK, tdQ = tmem + SM::cdQ, tddC = tmem + SM::cddC;
  // packed 16-bit A operands written by the workers: Sb'^T / dS^T in the first 16 columns of every 32-column
  // unit of tST / tdST; Kbar, Vbar, dHbar in the second 16 columns ("slots") of units ch (K, V) and ch + 2 (dH)
  auto slot = [](uint32_t base, int u) { return base + 32u * (uint32_t)u + 16u; };
  // A address of k-chunk kk (16 elements = 8 columns) of an operand whose rows are split in CW-element halves
  auto half_op = [&](uint32_t base, int u0, int kk) { return slot(base, u0 + kk / (CW / 16)) + 8u * (uint32_t)(kk % (CW / 16)); };

''')
# control: descriptors
rep("    const uint64_t mSb = umma_smem_desc(smem_u32(sSb), SM::kPTile, 1024);\n","")
rep("    constexpr uint32_t id_k_mn = umma_idesc(128, D, false, true, kBf16);   // A K-major, B MN-major","    constexpr uint32_t id_ts_mn = umma_idesc(128, D, false, true, kBf16);  // A from TMEM, B MN-major")
rep("    constexpr uint32_t id_k_k = umma_idesc(128, D, false, false, kBf16);   // A K-major, B K-major","    constexpr uint32_t id_ts_k = umma_idesc(128, D, false, false, kBf16);  // A from TMEM, B K-major")
# issue_s: transposed
cut("    auto issue_s = [&](int it) {","    if (elect_one()) issue_s(0);",
'''    auto issue_s = [&](int it) {  // S^T = K Q^T, dSb^T = V dH^T of processing step `it` (its loads are in flight)
      const int s = it % SM::kNST;
      const uint32_t so = (uint32_t)s * SM::kStage;
      const uint64_t kQ = umma_desc_advance(kQ0, so), kK = umma_desc_advance(kK0, so);
      const uint64_t kH = umma_desc_advance(kH0, so), kV = umma_desc_advance(kV0, so);
      mbar_wait(&bar_full[s], (it / SM::kNST) & 1, 11);
      tc_fence_after_sync();
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk)
        umma_f16(tST, umma_desc_advance(kK, kk * 32), umma_desc_advance(kQ, kk * 32), id_s, kk > 0);
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk)
        umma_f16(tdST, umma_desc_advance(kV, kk * 32), umma_desc_advance(kH, kk * 32), id_s, kk > 0);
      umma_commit(&bar_s);
    };

''')
# batch
cut("      // MMA batch, ordered (a) so that the first epilogue (dk)","      __syncwarp();\n      TC_PROF(it, 13);",
'''      // MMA batch.  Every 128 x 128 A operand that has the tile row on its M axis comes from TMEM (TS mode: no
      // shared-memory read for A), the inter-chunk terms are accumulated into the same TMEM columns through
      // row-scaled operand copies, so each output has ONE accumulator and needs no scaling in its epilogue.
      if (elect_one()) {
        tc_fence_after_sync();
#pragma unroll
        for (int kk = 0; kk < LT / 16; ++kk)  // dk  = dS^T Q            (A = packed dS^T in tdST)
          umma_f16_ts(tdK, tdST + 32 * (kk / 2) + 8 * (kk % 2), umma_desc_advance(mQ, kk * L::kAdvMN), id_ts_mn, kk > 0);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)   //     + (abar V) dC_k^T   (A = Vbar slots of tdST)
          umma_f16_ts(tdK, half_op(tdST, 0, kk), umma_desc_advance(kdC, kk * 32), id_ts_k, true);
        umma_commit(&bar_k);  // Q consumed; dk complete
#pragma unroll
        for (int kk = 0; kk < LT / 16; ++kk)  // dq  = dS K              (A = dS^T rows in shared memory, MN-major)
          umma_f16(tdQ, umma_desc_advance(mdS, kk * 2048), umma_desc_advance(mK, kk * L::kAdvMN), id_mn_mn, kk > 0);
      }
      __syncwarp();
      named_sync(NB_A, kNbAB);  // Kbar / dHbar copies written
      if (elect_one()) {
        tc_fence_after_sync();
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)   //     + (wb dH) C_{k-1}^T (A = dHbar slots of tST)
          umma_f16_ts(tdQ, half_op(tST, 2, kk), umma_desc_advance(kCs, kk * 32), id_ts_k, true);
        umma_commit(&bar_q);  // K, C_{k-1} consumed; dq complete
#pragma unroll
        for (int kk = 0; kk < LT / 16; ++kk)  // dv  = Sb'^T dH          (A = packed Sb'^T in tST)
          umma_f16_ts(tdV, tST + 32 * (kk / 2) + 8 * (kk % 2), umma_desc_advance(mH, kk * L::kAdvMN), id_ts_mn, kk > 0);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)   //     + (abar K) dC_k     (A = Kbar slots of tST)
          umma_f16_ts(tdV, half_op(tST, 0, kk), umma_desc_advance(mdC, kk * L::kAdvMN), id_ts_mn, true);
        umma_commit(&bar_v);  // dv complete (dC_k consumed)
#pragma unroll
        for (int kk = 0; kk < LT / 16; ++kk)  // ddC = Qt^T dH
          umma_f16(tddC, umma_desc_advance(mQt, kk * L::kAdvMN), umma_desc_advance(mH, kk * L::kAdvMN), id_c, kk > 0);
        umma_commit(&bar_d);  // dH consumed; last group of the batch
      }
      __syncwarp();
      if (lane == 0) {  // (off the MMA issue path: 128 CTAs store in lockstep, the reads take a while to drain)
        tma_store_wait_read<0>();  // the previous tile's dq / dk / dv stores have left their staging buffers
        mbar_arrive(&bar_st);
      }
      __syncwarp();
      TC_PROF(it, 11);
      if (SM::kNST == 1) {
        if (c > 0 && elect_one()) {  // re-fill the single stage tile by tile, each as soon as its last reader has completed
          const int r = mt(c - 1) * LT;
          mbar_expect_tx(&bar_full[0], SM::kLoadBytes);
          tma_load_4d(sV, &mapV, &bar_full[0], 0, r, hh, b);  // V: only dSb^T (complete) and the workers (NB_B) read it
          mbar_wait(&bar_k, par, 21);
          tma_load_4d(sQ, &mapQ, &bar_full[0], 0, r, hh, b);
          mbar_wait(&bar_q, par, 24);
          tma_load_4d(sK, &mapK, &bar_full[0], 0, r, hh, b);
          tma_load_4d(sCs, &mapCs, &bar_full[0], 0, mt(c - 1) * D, hh, b);
          mbar_wait(&bar_v, par, 25);
          tma_load_4d(sdH, &mapdH, &bar_full[0], 0, r, hh, b);
        }
      } else if (c >= SM::kNST && elect_one()) {  // this stage is free once the whole batch has completed
        mbar_wait(&bar_v, par, 25);
        load_stage(it % SM::kNST, c - SM::kNST);
      }
''')
rep("      if (SM::kAlias && c > 0 && elect_one()) issue_s(it + 1);  // S / dSb of the next tile (their TMEM columns were read by this epilogue)",
    "      if (c > 0 && elect_one()) issue_s(it + 1);  // S^T / dSb^T of the next tile")
rep("bar_q, bar_v, bar_k, bar_d, bar_b, bar_st, bar_g[2];","bar_q, bar_v, bar_k, bar_d, bar_st, bar_g[2];")
rep("    mbar_init(&bar_b, 1);\n","")
# scan warp: publish extras
rep('''      gate_scan_regs(gb, r.g, REV, p.sig != 0);
      reinterpret_cast<float4*>(gb + GateBuf::oMt)[lane] = r.mt;
      reinterpret_cast<float4*>(gb + GateBuf::oNt)[lane] = r.nt;''','''      gate_scan_regs(gb, r.g, REV, p.sig != 0);
      reinterpret_cast<float4*>(gb + GateBuf::oMt)[lane] = r.mt;
      reinterpret_cast<float4*>(gb + GateBuf::oNt)[lane] = r.nt;
      {  // column factors of W^T: X_t = (b_t - m_t) log2e + log2(scale / (n_t + eps)), -inf for tail tokens;
         // per 32-column unit: XM_u = max X, cx_t = exp2(X_t - XM_u) <= 1 (rank-1 form of the blocks off the diagonal)
        __syncwarp();
        const float4 bb = reinterpret_cast<const float4*>(gb + GateBuf::oB)[lane];
        const float bv[4] = {bb.x, bb.y, bb.z, bb.w}, mv[4] = {r.mt.x, r.mt.y, r.mt.z, r.mt.w},
                    nv[4] = {r.nt.x, r.nt.y, r.nt.z, r.nt.w};
        float X[4], xm = -INFINITY;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          X[e] = lane * 4 + e < r.g.n_valid ? (bv[e] - mv[e]) * kLog2e + log2f(p.scale / (nv[e] + p.eps)) : -INFINITY;
          xm = fmaxf(xm, X[e]);
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) xm = fmaxf(xm, __shfl_xor_sync(0xffffffffu, xm, o));
        xm = fmaxf(xm, -1e30f);
        reinterpret_cast<float4*>(gb + GateBuf::oPm)[lane] = make_float4(X[0], X[1], X[2], X[3]);
        reinterpret_cast<float4*>(gb + GateBuf::oCf)[lane] =
            make_float4(ex2_approx(X[0] - xm), ex2_approx(X[1] - xm), ex2_approx(X[2] - xm), ex2_approx(X[3] - xm));
        if ((lane & 7) == 0) gb[GateBuf::oScal + 4 + (lane >> 3)] = xm;
      }''')
# workers: W phase
cut("      // ---- W = D / (n + eps); Sb' = S.W, dS = dSb.W","      fence_proxy_async_smem();\n      tc_fence_before_sync();\n      named_arrive(NB_B, kNbAB);\n      TC_PROF(it, 3);",
'''      // ---- W^T (this thread: key / value row s = row, its two 32-column units of query columns t) ------------
      //   W_ts = exp2(X_t + Y_s) for t >= s (mirrored in the anti-causal direction), Y_s = (i_s - b_s) log2e;
      //   Sb'^T = scale S^T W^T and dS^T = scale dSb^T W^T are packed in place (TMEM A operands of dv / dk),
      //   dS^T rows also go to shared memory (MN-major A operand of dq = dS K)
      mbar_wait(&bar_s, par, 14);
      tc_fence_after_sync();
      TC_PROF(it, 2);
      {
        const float y_s = gb[GateBuf::oY + row];
        const float* sx = gb + GateBuf::oPm;
        const float* scx = gb + GateBuf::oCf;
#pragma unroll 1
        for (int u = ch; u < 4; u += 2) {  // warp-uniform branches
          float v[32], w[32];
          const bool nonzero = REV ? u <= rb : u >= rb;
          if (nonzero) {
            uint32_t rv[32], rw[32];
            tmem_ld32_nowait(tST + lane_base + u * 32, rv);
            tmem_ld32_nowait(tdST + lane_base + u * 32, rw);
            tmem_ld_wait();
            if (u != rb) {  // fully unmasked 32x32 block: rank-1 weights, one exp per row
              const float r_s = ex2_approx(y_s + gb[GateBuf::oScal + 4 + u]);
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 cf = *reinterpret_cast<const float4*>(scx + u * 32 + 4 * j4);
                const float cc[4] = {cf.x, cf.y, cf.z, cf.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int j = 4 * j4 + e;
                  const float wg = cc[e] * r_s;
                  v[j] = __uint_as_float(rv[j]) * wg;
                  w[j] = __uint_as_float(rw[j]) * wg;
                }
              }
            } else {  // diagonal block: mask, one exp per entry
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 x = *reinterpret_cast<const float4*>(sx + u * 32 + 4 * j4);
                const float xx[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int j = 4 * j4 + e;
                  float wg = ex2_approx(xx[e] + y_s);
                  wg = (REV ? j <= lane : j >= lane) ? wg : 0.f;
                  v[j] = __uint_as_float(rv[j]) * wg;
                  w[j] = __uint_as_float(rw[j]) * wg;
                }
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) { v[j] = 0.f; w[j] = 0.f; }
          }
          uint32_t pv[16], pw[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            pv[j] = pack2<T>(v[2 * j], v[2 * j + 1]);
            pw[j] = pack2<T>(w[2 * j], w[2 * j + 1]);
          }
          tmem_st16(tST + lane_base + u * 32, pv);
          tmem_st16(tdST + lane_base + u * 32, pw);
          if (nonzero || it == 0) {  // the zero blocks of the shared-memory copy are written once
            uint8_t* tile = sdS + (u >> 1) * SM::kPTile;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(tile + swz128(row, (u & 1) * 32 + 8 * j)) =
                  make_uint4(pw[4 * j], pw[4 * j + 1], pw[4 * j + 2], pw[4 * j + 3]);
          }
        }
      }
      // ---- row-scaled operand copies into the free TMEM slots; k / v row slices kept for the gate gradients ----
      uint32_t ks[CW / 2], vs[CW / 2];
      {
        uint32_t ob[CW / 2];
        const T ab = from_f32<T>(abar);
        const T wbs = from_f32<T>(p.scale * bbar * rinv);
#pragma unroll
        for (int j = 0; j < CW / 8; ++j) {
          const uint32_t off = L::swz(row, ch * CW + 8 * j);
          const uint4 uk = *reinterpret_cast<const uint4*>(sK + so + off);
          ks[4 * j] = uk.x; ks[4 * j + 1] = uk.y; ks[4 * j + 2] = uk.z; ks[4 * j + 3] = uk.w;
          const uint4 uv = *reinterpret_cast<const uint4*>(sV + so + off);
          vs[4 * j] = uv.x; vs[4 * j + 1] = uv.y; vs[4 * j + 2] = uv.z; vs[4 * j + 3] = uv.w;
        }
#pragma unroll
        for (int j = 0; j < CW / 2; ++j) ob[j] = mul2<T>(vs[j], ab);   // Vbar = abar v  (bw.py:192)
        tmem_st(slot(tdST, ch) + lane_base, ob);
        // first hand-off: everything the dk group and the shared-memory half of the dq group need
        tmem_st_wait();
        fence_proxy_async_smem();
        tc_fence_before_sync();
        named_arrive(NB_B, kNbAB);
        TC_PROF(it, 3);
#pragma unroll
        for (int j = 0; j < CW / 2; ++j) ob[j] = mul2<T>(ks[j], ab);   // Kbar = abar k  (bw.py:190)
        tmem_st(slot(tST, ch) + lane_base, ob);
#pragma unroll
        for (int j = 0; j < CW / 8; ++j) {
          const uint4 uh = *reinterpret_cast<const uint4*>(sdH + so + L::swz(row, ch * CW + 8 * j));
          ob[4 * j] = mul2<T>(uh.x, wbs); ob[4 * j + 1] = mul2<T>(uh.y, wbs);
          ob[4 * j + 2] = mul2<T>(uh.z, wbs); ob[4 * j + 3] = mul2<T>(uh.w, wbs);  // dHbar = scale bbar/(n+eps) dh (bw.py:193)
        }
        tmem_st(slot(tST, ch + 2) + lane_base, ob);
      }
      // second hand-off: the row-scaled copies the remaining TS instructions read (written while the dk group runs)
      tmem_st_wait();
      tc_fence_before_sync();
      named_arrive(NB_A, kNbAB);
''')
# remove old k/v row block + NB_A
cut("      // ---- this thread's k / v row slices for the gate gradients (the inputs may be re-filled afterwards) --","      TC_PROF(it, 4);","")
# epilogues
cut("      // ---- epilogues, pipelined with the MMA batch through three commits","      fence_proxy_async_smem();\n      tc_fence_before_sync();\n      named_arrive(NB_C, kNbC);\n      TC_PROF(it, 8);",
'''      // ---- epilogues (one accumulator per output, no scaling), pipelined with the MMA batch ------------------
      {
        uint32_t ra[CW];
        float o[CW];
        float dot;
        // dk
        mbar_wait(&bar_k, par, 18);
        tc_fence_after_sync();
        TC_PROF(it, 5);
        tmem_ld_nowait(tdK + lane_base + ch * CW, ra);
        tmem_ld_wait();
        dot = 0.f;
#pragma unroll
        for (int j = 0; j < CW / 2; ++j) {
          o[2 * j] = __uint_as_float(ra[2 * j]);  // bw.py:170,192
          o[2 * j + 1] = __uint_as_float(ra[2 * j + 1]);
          float2 kv = unpack2<T>(ks[j]);
          dot += kv.x * o[2 * j] + kv.y * o[2 * j + 1];
        }
        mbar_wait(&bar_st, par, 19);  // staging buffers free (the previous tile's stores have read them)
        store_cols<T, D>(sdK, row, ch * CW, o);
        spart[(1 * 2 + ch) * LT + row] = dot;
        // dq
        mbar_wait(&bar_q, par, 16);
        tc_fence_after_sync();
        TC_PROF(it, 6);
        tmem_ld_nowait(tdQ + lane_base + ch * CW, ra);
        tmem_ld_wait();
        dot = 0.f;
#pragma unroll
        for (int j = 0; j < CW / 2; ++j) {
          o[2 * j] = __uint_as_float(ra[2 * j]);  // bw.py:169,193
          o[2 * j + 1] = __uint_as_float(ra[2 * j + 1]);
          float2 qv = unpack2<T>(qs[j]);
          dot += qv.x * o[2 * j] + qv.y * o[2 * j + 1];
        }
        store_cols<T, D>(sdQ, row, ch * CW, o);
        spart[(0 * 2 + ch) * LT + row] = dot;
      }
      // ---- dv ---------------------------------------------------------------------------------------
      {
        uint32_t ra[CW];
        float o[CW];
        float dot = 0.f;
        mbar_wait(&bar_v, par, 17);
        tc_fence_after_sync();
        tmem_ld_nowait(tdV + lane_base + ch * CW, ra);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < CW / 2; ++j) {
          o[2 * j] = __uint_as_float(ra[2 * j]);  // bw.py:164,190
          o[2 * j + 1] = __uint_as_float(ra[2 * j + 1]);
          float2 vv = unpack2<T>(vs[j]);
          dot += vv.x * o[2 * j] + vv.y * o[2 * j + 1];
        }
        store_cols<T, D>(sdV, row, ch * CW, o);
        spart[(2 * 2 + ch) * LT + row] = dot;
      }
      // ---- dC_{k-1} = gbar dC_k + ddC ----------------------------------------------------------------
      mbar_wait(&bar_d, par, 15);
      tc_fence_after_sync();
      TC_PROF(it, 7);
      {
        float v[CW];
        tmem_ld(tddC + lane_base + ch * CW, v);
        if (owns_c) {
#pragma unroll
          for (int j = 0; j < CW; ++j) dCreg[j] = gbar * dCreg[j] + v[j];  // bw.py:93-95
          store_cols<T, D>(sdC, drow, ch * CW, dCreg);                     // its readers (dk, dv groups) have completed
        }
      }
''')

# ---------------------------------------------------------------- stage 2: per-tensor buffers, early S^T issue
rep("  uint8_t* sQ = smem + SM::oQ;  // stage 0; stage s is SM::kStage bytes further","  uint8_t* sQ = smem + SM::oQ;  // buffer 0 of each input; processing step `it` uses buffer it % n")
rep("  __shared__ uint64_t bar_full[SM::kNST], bar_s,","  __shared__ uint64_t bar_fa[2], bar_fb[2], bar_fc[2], bar_s,")
cut("  auto load_stage = [&](int s, int c) {","  if (tid == 0) {\n    mbar_init(&bar_s, 1);",
"""  // loads of processing step `it` (memory tile mt(NT-1-it)): Q, K complete on bar_fa[it & 1] (operands of S^T),
  // V, dH on bar_fb (operands of dSb^T), C_{k-1} on bar_fc
  auto expect_step = [&](int it) {
    mbar_expect_tx(&bar_fa[it & 1], 2 * SM::kTile);
    mbar_expect_tx(&bar_fb[it & 1], 2 * SM::kTile);
    mbar_expect_tx(&bar_fc[it & 1], SM::kState);
  };
  auto load_Q = [&](int it) { tma_load_4d(sQ + (it % SM::nQ) * SM::kTile, &mapQ, &bar_fa[it & 1], 0, mt(p.NT - 1 - it) * LT, hh, b); };
  auto load_K = [&](int it) { tma_load_4d(sK + (it % SM::nK) * SM::kTile, &mapK, &bar_fa[it & 1], 0, mt(p.NT - 1 - it) * LT, hh, b); };
  auto load_V = [&](int it) { tma_load_4d(sV + (it % SM::nV) * SM::kTile, &mapV, &bar_fb[it & 1], 0, mt(p.NT - 1 - it) * LT, hh, b); };
  auto load_H = [&](int it) { tma_load_4d(sdH + (it % SM::nH) * SM::kTile, &mapdH, &bar_fb[it & 1], 0, mt(p.NT - 1 - it) * LT, hh, b); };
  auto load_C = [&](int it) { tma_load_4d(sCs + (it % SM::nCs) * SM::kState, &mapCs, &bar_fc[it & 1], 0, mt(p.NT - 1 - it) * D, hh, b); };
  // cold start: the first input tiles are requested before anything else happens in the CTA
  if (tid == kCtlWarp * 32) {
    for (int s = 0; s < 2; ++s) { mbar_init(&bar_fa[s], 1); mbar_init(&bar_fb[s], 1); mbar_init(&bar_fc[s], 1); }
    fence_mbar_init();
    expect_step(0);
    load_Q(0); load_K(0); load_V(0); load_H(0); load_C(0);
  }
""")
cut("    auto issue_s = [&](int it) {","    if (elect_one()) issue_s(0);",
"""    auto issue_s = [&](int it) {  // S^T = K Q^T, dSb^T = V dH^T of processing step `it` (its loads are in flight)
      const uint64_t kQ = umma_desc_advance(kQ0, (it % SM::nQ) * SM::kTile), kK = umma_desc_advance(kK0, (it % SM::nK) * SM::kTile);
      const uint64_t kH = umma_desc_advance(kH0, (it % SM::nH) * SM::kTile), kV = umma_desc_advance(kV0, (it % SM::nV) * SM::kTile);
      mbar_wait(&bar_fa[it & 1], (it >> 1) & 1, 11);
      tc_fence_after_sync();
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk)
        umma_f16(tST, umma_desc_advance(kK, kk * 32), umma_desc_advance(kQ, kk * 32), id_s, kk > 0);
      mbar_wait(&bar_fb[it & 1], (it >> 1) & 1, 11);
      tc_fence_after_sync();
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk)
        umma_f16(tdST, umma_desc_advance(kV, kk * 32), umma_desc_advance(kH, kk * 32), id_s, kk > 0);
      umma_commit(&bar_s);
    };

""")
cut("      const uint32_t so = (uint32_t)(it % SM::kNST) * SM::kStage;\n      const uint64_t kQ = umma_desc_advance(kQ0, so), mQ","      named_sync(NB_B, kNbAB);  // Sb', dS written",
"""      const uint64_t mQ = umma_desc_advance(mQ0, (it % SM::nQ) * SM::kTile), mK = umma_desc_advance(mK0, (it % SM::nK) * SM::kTile);
      const uint64_t mH = umma_desc_advance(mH0, (it % SM::nH) * SM::kTile), kCs = umma_desc_advance(kCs0, (it % SM::nCs) * SM::kState);
      TC_PROF(it, 9);
      if (c > 0 && lane == 0) {  // inputs of the next tile
        expect_step(it + 1);
        if (SM::nK == 2) load_K(it + 1);  // double-buffered: the other buffer was released by the previous tile's batch
        load_H(it + 1);
        if (SM::nQ == 2) { load_Q(it + 1); load_V(it + 1); load_C(it + 1); }
        else {  // single buffers: pull the rows into L2 now, so that the re-fills behind the MMA batch are L2 hits
          const int r = mt(c - 1) * LT;
          tma_prefetch_4d(&mapQ, 0, r, hh, b);
          tma_prefetch_4d(&mapV, 0, r, hh, b);
          tma_prefetch_4d(&mapCs, 0, mt(c - 1) * D, hh, b);
          if (SM::nK == 1) tma_prefetch_4d(&mapK, 0, r, hh, b);
        }
      }
""")
rep("      named_sync(NB_B, kNbAB);  // Sb', dS written\n      TC_PROF(it, 10);\n","""      named_sync(NB_B, kNbAB);  // operands written, input rows read
      TC_PROF(it, 10);
      if (lane == 0) {
        tma_store_wait_read<0>();  // the previous tile's dq / dk / dv stores (issued a W phase ago) have left their
        mbar_arrive(&bar_st);      // staging buffers
        mbar_wait(&bar_fc[it & 1], (it >> 1) & 1, 26);  // C_{k-1} has landed
      }
      __syncwarp();
""")
cut("      __syncwarp();\n      if (lane == 0) {  // (off the MMA issue path: 128 CTAs store in lockstep","      __syncwarp();\n      TC_PROF(it, 13);",
"""      __syncwarp();
      TC_PROF(it, 11);
      if (c > 0 && elect_one()) {
        if (SM::nQ == 1) {  // single-buffered inputs: re-fill each as soon as its last reader has completed
          load_V(it + 1);   // V: only dSb^T (complete) and the workers (before NB_B) read it
          mbar_wait(&bar_k, par, 21);
          load_Q(it + 1);
        }
        mbar_wait(&bar_v, par, 25);  // the TS groups have read their packed TMEM operands (ddC, still queued, has none) (measured: an MMA that writes
        issue_s(it + 1);             // TMEM columns may overtake the A-operand reads of the instruction before it)
        if (SM::nQ == 1) {
          mbar_wait(&bar_q, par, 24);
          if (SM::nK == 1) load_K(it + 1);
          load_C(it + 1);
        }
      }
""")
rep("      if (c > 0 && elect_one()) issue_s(it + 1);  // S^T / dSb^T of the next tile\n      __syncwarp();\n","")
# workers
rep("      const uint32_t so = (uint32_t)(it % SM::kNST) * SM::kStage;  // input stage of this tile\n","")
rep("      mbar_wait(&bar_full[it % SM::kNST], (it / SM::kNST) & 1, 13);","      mbar_wait(&bar_fa[it & 1], (it >> 1) & 1, 13);")
rep("          uint4 u = *reinterpret_cast<const uint4*>(sQ + so + off);","          uint4 u = *reinterpret_cast<const uint4*>(sQ + (it % SM::nQ) * SM::kTile + off);")
rep("          const uint4 uk = *reinterpret_cast<const uint4*>(sK + so + off);","          const uint4 uk = *reinterpret_cast<const uint4*>(sK + (it % SM::nK) * SM::kTile + off);")
rep("          const uint4 uv = *reinterpret_cast<const uint4*>(sV + so + off);","          const uint4 uv = *reinterpret_cast<const uint4*>(sV + (it % SM::nV) * SM::kTile + off);")
rep("          const uint4 uh = *reinterpret_cast<const uint4*>(sdH + so + L::swz(row, ch * CW + 8 * j));","          const uint4 uh = *reinterpret_cast<const uint4*>(sdH + (it % SM::nH) * SM::kTile + L::swz(row, ch * CW + 8 * j));")
rep("      named_arrive(NB_A, kNbAB);\n      fence_proxy_async_smem();\n      tc_fence_before_sync();\n      named_arrive(NB_B, kNbAB);\n      TC_PROF(it, 3);\n","      named_arrive(NB_A, kNbAB);\n")
rep("      uint32_t ks[CW / 2], vs[CW / 2];\n      {\n        uint32_t ob[CW / 2];","      mbar_wait(&bar_fb[it & 1], (it >> 1) & 1, 13);\n      uint32_t ks[CW / 2], vs[CW / 2];\n      {\n        uint32_t ob[CW / 2];")
k=k.replace("// >>> BW2","")
SMEM = '''template <int D_>
struct Bw2Smem {
  static constexpr int D = D_;
  static constexpr int kTile = Lay<D>::kTile;   // one [128][D] tile
  static constexpr int kPTile = LT * 128;       // one [128][64] half of dS^T
  static constexpr int kState = Lay<D>::kState;
  // Input buffers.  D = 32: two of each (a whole tile prefetched ahead).  D = 64: two K and two dH tiles (loaded a
  // tile ahead), one Q, V, C_{k-1}, re-filled behind the MMA batch as soon as their last reader has completed.
  static constexpr int nQ = D == 64 ? 1 : 2, nK = 2, nV = nQ, nH = 2, nCs = nQ;
  static constexpr int oQ = 0, oK = oQ + nQ * kTile, oV = oK + nK * kTile, odH = oV + nV * kTile, oCs = odH + nH * kTile;
  static constexpr int oQt = oCs + nCs * kState;  // wq . Q
  static constexpr int odS = oQt + kTile;       // dS^T rows, two halves (query columns 0-63 / 64-127)
  static constexpr int odQ = odS + 2 * kPTile;  // dq / dv / dk staging
  static constexpr int odV = odQ + kTile;
  static constexpr int odK = odV + kTile;
  static constexpr int odC = odK + kTile;       // dC_k 16-bit operand copy
  static constexpr int oSmall = odC + kState;
  static constexpr int fGates = 0, fPart = 2 * GateBuf::kFloats, kSmallFloats = fPart + 12 * LT;
  static constexpr int kBytes = oSmall + kSmallFloats * 4 + 1024;
  static constexpr bool kAlias = false;
  // TMEM: S^T and dSb^T (128 columns each; later the packed operands), one accumulator per output
  static constexpr uint32_t cST = 0, cdST = 128, cdV = 256, cdK = cdV + D, cdQ = cdK + D, cddC = cdQ + D;
};

'''
out = "// >>> BW2 BEGIN\n// =============================================================================================\n// Backward, transposed formulation (tc_bw2): S^T = K Q^T and dSb^T = V dH^T put the key / value index on the\n// TMEM lanes, so the weighted tiles Sb'^T and dS^T are packed IN PLACE by the thread that owns the row and feed\n// dv = Sb'^T dH and dk = dS^T Q as TMEM A operands (TS mode, no shared-memory read for A, no Sb' buffer); the\n// inter-chunk terms ride in the same accumulators through row-scaled operand copies (abar k, abar v,\n// scale bbar/(n+eps) dh) kept in the unused TMEM columns.  Same warp roles, barriers and scan warp as tc_bw.\n// =============================================================================================\n" + SMEM + k + "// <<< BW2 END\n"
s=s[:b]+out+s[b:]
open(p,'w').write(s)
