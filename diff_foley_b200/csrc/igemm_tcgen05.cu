// igemm_tcgen05.cu -- the one tensor-core kernel of the UNet hot path.
//
// Implicit GEMM on the 5th-gen tensor cores:  out[m, n] = epilogue( sum_k A[m, k] * Wt[n, k] )
//   * A is an fp16 channels-last activation tensor [B,T,H,W,C]; an M tile is a 5-D TMA box of 128
//     output positions x 64 channels, one box per (filter tap, 64-channel slab).  Shifting the box
//     origin by the tap offset and letting TMA zero-fill out-of-bounds coordinates *is* the pad=1
//     halo of the 3x3 convolutions -- no im2col buffer, no boundary branches.  A plain Linear /
//     1x1 conv is the single-tap case with W := M.
//   * Wt is fp16 [N, taps*C] (K-major), 2-D TMA boxes of BN x 64.
//   * both operand tiles land in shared memory in the 128-byte-swizzled K-major layout tcgen05.mma
//     consumes directly; accumulators live in TMEM (128 lanes x BN fp32 columns).
//   * warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
//     warps 2..5 = epilogue (each owns the 32 TMEM lanes of its sub-partition = warp_idx % 4).
//   * epilogue (fused): + bias[n], + per-sample vector (timestep embedding), + fp32 residual,
//     SiLU / ReLU / GEGLU gate, fp32 and/or fp16 stores.
//   * split-K for the weight-streaming-bound deep layers (M <= 128 rows against 30-60 MB of
//     weights): up to 8 CTAs along grid.z each stream a K range; they form a thread-block cluster,
//     park their fp32 partial tiles in their own shared memory and reduce them through distributed
//     shared memory in rank order (deterministic, no atomics, no global workspace), each CTA running
//     the fused epilogue on its share of the tile's columns.
//
// Replaces (reference, all library calls): nn.Conv2d 3x3/1x1 (openai_unetmodel.py:204,230,241,
// 107; attention_openai.py:233,244) and nn.Linear (attention_openai.py:40,60,161-168;
// openai_unetmodel.py:218-224,507-511).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "dfb_internal.h"
#include "dfb_ptx.cuh"

namespace dfb {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 fp16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int IGEMM_THREADS = 320;  // 1 TMA + 1 MMA warp; warps 2..5 read TMEM, warps 2..9 finish the tile

struct IGemmKParams {
  int M, N;
  int B, T, H, W;
  int bb, bt, bh, bw;
  int tw, th, tt;  // tiles along W, H, T (tiles along B implied)
  int ntaps, kpt;  // taps, 64-channel k-blocks per tap
  int kpt2;        // 64-channel k-blocks of the second A source (after the taps), 0 = none
  int8_t dt[9], dh[9], dw[9];
  float* out_f32;
  __half* out_f16;
  int ldo;
  const float* bias;
  const float* rowvec;
  int ld_rowvec;
  int rows_per_sample;
  const int* rowvec_row;
  const float* residual;
  const __half* residual_f16;
  int ld_res;
  int act;
  int splits;  // == cluster size along z (<= 8): the CTAs of one output tile reduce through DSMEM
  const uint8_t* next_w;  // optional: weights of the next GEMM, prefetched into L2 slice-wise
  unsigned long long next_w_bytes;
  unsigned long long* trace;  // optional timeline record (diagnostics, see trace_mark)
  int debug_skip;             // diagnostics: bit 0 = skip phase 1 (TMEM -> smem), bit 1 = skip phase 2
};

template <int BN, int STAGES>
struct IGemmSmem {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int W_BYTES = BN * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + W_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  // full[STAGES], empty[STAGES], tmem_full, red (split-K partials landed), then tmem slot + flag, then
  // the row table (m, sample)
  static constexpr int ROW_OFFSET = BAR_OFFSET + (2 * STAGES + 2) * 8 + 16;
  static constexpr int TOTAL = ROW_OFFSET + BLOCK_M * 8 + 16;  // + 128 valid-row bits
  static constexpr int DYN_BYTES = TOTAL + 1024;  // slack to align the base to 1024 B
};

// ---- fused epilogue on one float4 (4 consecutive output columns of one row).  The accumulator tile
// is first staged in shared memory (see the kernel), so here a warp covers whole rows: every global
// access (bias / timestep-embedding vector / residual loads, fp32 / fp16 stores) is a contiguous
// 256-512 B run per row instead of 32 rows x 16 B.
__device__ __forceinline__ float4 f4_add(float4 a, const float4 b) {
  a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  return a;
}
__device__ __forceinline__ uint2 f4_to_h4(const float4 v) {
  const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<const uint32_t*>(&h0);
  u.y = *reinterpret_cast<const uint32_t*>(&h1);
  return u;
}

// STAGES-deep operand ring; the shallow-ring instantiations (<=3-4 stages, ~97 KB) let two CTAs share
// an SM so one CTA's epilogue overlaps the other's main loop -- used for short-K problems.
template <int BN, int STAGES>
__global__ void __launch_bounds__(IGEMM_THREADS, (STAGES <= 4) ? 2 : 1)
igemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                     const __grid_constant__ CUtensorMap tmW,
                     const __grid_constant__ CUtensorMap tmA2,
                     const __grid_constant__ IGemmKParams p) {
  using L = IGemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* red_bar = tmem_full_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(red_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) trace_mark(p.trace, 0);

  // ---- tile coordinates
  const int n0 = blockIdx.x * BN;
  int tm = blockIdx.y;
  const int tw_i = tm % p.tw; tm /= p.tw;
  const int th_i = tm % p.th; tm /= p.th;
  const int tt_i = tm % p.tt; tm /= p.tt;
  const int tb_i = tm;
  const int x0 = tw_i * p.bw, y0 = th_i * p.bh, t0 = tt_i * p.bt, b0 = tb_i * p.bb;

  const int kb_taps = p.ntaps * p.kpt;
  const int kb_total = kb_taps + p.kpt2;
  const int kb0 = (int)(((long)blockIdx.z * kb_total) / p.splits);
  const int kb1 = (int)(((long)(blockIdx.z + 1) * kb_total) / p.splits);
  const int nkb = kb1 - kb0;

  // ---- one-time setup.  Weights never depend on the preceding kernel, so the producer thread
  // initialises the barriers itself and issues the weight (W) loads of the first ring of stages
  // BEFORE the programmatic-dependent-launch wait: they stream from HBM while the predecessor is still
  // finishing.  Only the activation (A) loads and the epilogue's reads wait for it.
  const int npre = min(nkb, STAGES);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    if (p.kpt2) tma_prefetch_desc(&tmA2);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(red_bar, 1);
    fence_mbar_init();
    for (int i = 0; i < npre; ++i) {
      mbar_expect_tx(&full_bar[i], L::STAGE_BYTES);
      tma_load_2d(smem + i * L::STAGE_BYTES + L::A_BYTES, &tmW, &full_bar[i], (kb0 + i) * BLOCK_K, n0);
    }
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BN);  // BN fp32 columns x 128 lanes (power of two >= 32)
    tmem_relinquish();
  }
  // row table: accumulator row r of this tile -> output row m (-1 = outside the tensor) and the
  // sample it belongs to (for the per-sample vector); filled by the epilogue warps while the ring fills
  const uint32_t row_tab = smem_u32(smem + L::ROW_OFFSET);
  if (warp >= 2 && warp < 6) {
    const int r = (warp & 3) * 32 + lane;
    const int w_i = r % p.bw;
    const int h_i = (r / p.bw) % p.bh;
    const int t_i = (r / (p.bw * p.bh)) % p.bt;
    const int b_i = r / (p.bw * p.bh * p.bt);
    const int x = x0 + w_i, y = y0 + h_i, t = t0 + t_i, b = b0 + b_i;
    const bool row_ok = (x < p.W) && (y < p.H) && (t < p.T) && (b < p.B);
    const int m = ((b * p.T + t) * p.H + y) * p.W + x;
    sts_i2(row_tab + 8 * r, row_ok ? m : -1, (p.rows_per_sample > 0) ? m / p.rows_per_sample : b);
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // everything above overlapped the predecessor's tail; operands are valid from here
  if (threadIdx.x == 0) trace_mark(p.trace, 1);
  pdl_launch_dependents();  // only after our own wait: at most two grids of the chain overlap
  const uint32_t tmem_base = *tmem_slot;
  const bool geglu = (p.act == ACT_GEGLU);
  const bool split = (p.splits > 1);

  // ---- epilogue geometry (phase 2 below).  Warps 2..9 finish the tile: a warp takes whole rows, lane =
  // float4 group (two rows per pass when BN = 64).  Their operands do not depend on the accumulator, so
  // the first batch (bias, per-sample vector, residual) is fetched NOW, under the main loop.
  constexpr int G = BN / 4;            // float4 groups per accumulator row
  constexpr int NW = IGEMM_THREADS / 32 - 2;
  constexpr int RPW = 32 / G;          // rows per warp pass (1 or 2)
  constexpr int UNR = 4;
  constexpr int STEP = NW * RPW * UNR;
  const int S = p.splits, R = (BLOCK_M + S - 1) / S;
  const int Rz = min(R, BLOCK_M - (int)blockIdx.z * R);  // rows this CTA finishes (<= 0: none)
  const uint32_t tab = row_tab + 8u * (uint32_t)((int)blockIdx.z * R);
  const int g = lane % G, sr = lane / G;
  const bool lane_on = (n0 + 4 * g < p.N);
  const int col = n0 + 4 * g;
  const int colc = lane_on ? col : n0;  // clamped: the operand loads are unconditional
  const bool has_rv = (p.rowvec != nullptr), has_res = (p.residual != nullptr),
             has_res16 = (p.residual_f16 != nullptr);
  const int base0 = (warp - 2) * RPW + sr;
  struct EpiOps {
    int m[UNR];
    float4 rv[UNR], res[UNR];
  };
  // every load of a batch is issued (at clamped, always-valid addresses) before any of them is consumed
  const int rv_row = (has_rv && p.rowvec_row != nullptr) ? __ldg(p.rowvec_row) : -1;
  auto epi_fetch = [&](int base, EpiOps& o) {
    int bsm[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int lr = base + u * NW * RPW;
      const int2 ri = lds_i2(tab + 8u * (uint32_t)max(min(lr, Rz - 1), 0));
      o.m[u] = (lane_on && lr < Rz) ? ri.x : -1;
      bsm[u] = rv_row >= 0 ? rv_row : ri.y;
      o.rv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      o.res[u] = o.rv[u];
    }
    if (has_rv) {
#pragma unroll
      for (int u = 0; u < UNR; ++u)
        o.rv[u] = __ldg(reinterpret_cast<const float4*>(p.rowvec + (long)bsm[u] * p.ld_rowvec + colc));
    }
    if (has_res) {
#pragma unroll
      for (int u = 0; u < UNR; ++u)
        o.res[u] = *reinterpret_cast<const float4*>(p.residual + (long)max(o.m[u], 0) * p.ld_res + colc);
    } else if (has_res16) {
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const uint2 h = *reinterpret_cast<const uint2*>(p.residual_f16 + (long)max(o.m[u], 0) * p.ld_res + colc);
        const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&h.x));
        const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
        o.res[u] = make_float4(lo.x, lo.y, hi.x, hi.y);
      }
    }
  };
  EpiOps ops;
  float4 bias = make_float4(0.f, 0.f, 0.f, 0.f), bias_g = bias;
  if (warp >= 2) {
    if (!geglu) {
      if (lane_on && p.bias != nullptr) bias = __ldg(reinterpret_cast<const float4*>(p.bias + col));
      if (base0 < Rz) epi_fetch(base0, ops);
    } else {
      bias = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + 4 * (lane % (G / 2))));
      bias_g = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + BN / 2 + 4 * (lane % (G / 2))));
    }
  }
  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const int kb = kb0 + i;
        uint8_t* sa = smem + s * L::STAGE_BYTES;
        if (i >= npre) {  // ring slot reuse: wait for the MMAs that read it, then arm + load W as well
          mbar_wait(&empty_bar[s], ((i / STAGES) & 1) ^ 1);
          mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
          tma_load_2d(sa + L::A_BYTES, &tmW, &full_bar[s], kb * BLOCK_K, n0);
        }
        if (kb < kb_taps) {
          const int tap = kb / p.kpt;
          const int c0 = (kb - tap * p.kpt) * BLOCK_K;
          tma_load_5d(sa, &tmA, &full_bar[s], c0, x0 + p.dw[tap], y0 + p.dh[tap], t0 + p.dt[tap], b0);
        } else {  // second source: same positions, no tap shift
          tma_load_5d(sa, &tmA2, &full_bar[s], (kb - kb_taps) * BLOCK_K, x0, y0, t0, b0);
        }
      }
    }
  } else if (warp == 1) {
    // ====================================================================== MMA issuer
    constexpr uint32_t idesc = umma_idesc_f16(BLOCK_M, BN);
    for (int i = 0; i < nkb; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (lane == 0) {
        if (i == 0) trace_mark(p.trace, 2);
        const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
        const uint32_t sw = sa + L::A_BYTES;
        const uint64_t da = umma_desc_k_sw128(sa);
        const uint64_t db = umma_desc_k_sw128(sw);
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          // advancing 16 fp16 (32 B) along K inside the 128-B swizzle row = +2 in the addr field
          umma_f16_ss(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc,
                      (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs have read it
        if (i == nkb - 1) {
          umma_commit(tmem_full_bar);
          trace_mark(p.trace, 3);
        }
      }
      __syncwarp();
    }
  } else {
    // ================================================= epilogue, phase 0: wait for the accumulator
    if (warp == 2 && lane == 0 && p.next_w != nullptr) {
      // (an epilogue warp: idle until the accumulator is complete, and not on the TMA/MMA critical path)
      // software pipelining across layers: pull this CTA's slice of the next GEMM's weights into L2
      const unsigned long long nct = (unsigned long long)gridDim.x * gridDim.y * gridDim.z;
      const unsigned long long cta = ((unsigned long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      const unsigned long long per = ((p.next_w_bytes + nct - 1) / nct + 4095ull) & ~4095ull;
      unsigned long long off = cta * per;
      const unsigned long long end = min(off + per, p.next_w_bytes & ~15ull);
      for (; off < end; off += 4096ull) {
        const unsigned int sz = (unsigned int)min(4096ull, end - off);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.next_w + off), "r"(sz) : "memory");
      }
    }
    __syncwarp();
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (warp == 2 && lane == 0) trace_mark(p.trace, 4);
  }

  // ---- epilogue, phase 1: TMEM -> shared-memory staging (the operand ring is dead once the accumulator
  // is complete).  Layout: rows of BN floats, float4 group g of row r at [r][g ^ (lr & 7)] -- the XOR
  // keeps both the row-per-thread writes here and the row-per-warp reads of phase 2 bank-conflict free.
  // Split-K: the `splits` CTAs of one output tile form a thread-block cluster (1,1,splits); CTA z
  // finishes rows [z*R, (z+1)*R) of the tile, R = ceil(128/splits), lr = r - z*R.  After ONE cluster
  // barrier (every CTA's ring is dead; no memory ordering needed) each CTA sends every peer that peer's rows of
  // its partial tile as one contiguous bulk copy (cp.async.bulk shared::cta -> shared::cluster, bytes
  // counted on the receiver's mbarrier), into slot [sender rank] of the receive area behind the local
  // tile.  The receiver then sums the slots in rank order from its own shared memory: deterministic, no
  // atomics, no global workspace, no per-element remote traffic (DSMEM moves ~20 B/clk/SM at best, and a
  // per-float4 st.async costs one mbarrier update each -- both measured).
  constexpr uint32_t ROWB = BN * 4;
  const uint32_t tile = smem_u32(smem);             // this CTA's (partial) accumulator tile, 128 rows
  const uint32_t rcv = tile + BLOCK_M * ROWB;        // split-K: [slot][R rows] received partials
  uint32_t* vmask = reinterpret_cast<uint32_t*>(smem + L::ROW_OFFSET + BLOCK_M * 8);  // valid-row bits
  __syncwarp();  // re-converge the single-lane roles before the .aligned barriers
  // "this CTA's operand ring is dead" (warps >= 2 get here after the accumulator completed): arrive now,
  // wait only before the copies are issued -- the barrier's latency hides behind the staging below
  if (split) cluster_arrive_relaxed();
  if (warp >= 2 && warp < 6) {
    const int sub = warp & 3;          // TMEM sub-partition this warp may read
    const int r = sub * 32 + lane;     // accumulator row (= TMEM lane) owned by this thread
    const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16);
    const int lr = r % R, sw = lr & 7;
    const uint32_t dst = tile + (uint32_t)r * ROWB;
    if (split) {
      // rows outside the tensor (M = 32 tiles: three quarters of them) are neither sent nor summed
      const uint32_t vm = __ballot_sync(0xffffffffu, lds_i2(row_tab + 8u * (uint32_t)r).x >= 0);
      if (lane == 0) vmask[sub] = vm;
    }
    if (warp == 2 && lane == 0) trace_mark(p.trace, 12);
    // NB: tcgen05.ld is warp-collective (.sync.aligned): every lane loads
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      if (n0 + c >= p.N || (p.debug_skip & 1)) break;
      if (c == 32 && warp == 2 && lane == 0) trace_mark(p.trace, 13);
      uint32_t raw[32];
      tmem_ld_32x32(taddr + c, raw);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sts_f4(dst + (uint32_t)((((c >> 2) + j) ^ sw) << 4), __uint_as_float(raw[4 * j]),
               __uint_as_float(raw[4 * j + 1]), __uint_as_float(raw[4 * j + 2]),
               __uint_as_float(raw[4 * j + 3]));
    }
    if (split) fence_proxy_async_smem();  // the staged tile is read by the bulk-copy engine
    if (warp == 2 && lane == 0) trace_mark(p.trace, 5);
  }
  tc_fence_before();
  __syncthreads();  // the staged tile (and the valid-row bits) are complete
  if (split) {
    cluster_wait();
    if (warp == 2 && lane < S) {
      // lane z: the span of valid rows among those CTA z finishes -> one bulk copy to CTA z; the lane
      // whose z is this CTA's own rank also arms the receive barrier with what all S senders deliver
      const int z = lane;
      int lo = BLOCK_M, hi = 0;
      for (int r = z * R; r < min(BLOCK_M, (z + 1) * R); ++r)
        if ((vmask[r >> 5] >> (r & 31)) & 1u) { lo = min(lo, r); hi = r + 1; }
      const uint32_t bytes = hi > lo ? (uint32_t)(hi - lo) * ROWB : 0u;
      if (z == (int)blockIdx.z) mbar_expect_tx(red_bar, (uint32_t)S * bytes);
      if (bytes)
        bulk_copy_s2c(mapa_shared(rcv + (uint32_t)((int)blockIdx.z * R + (lo - z * R)) * ROWB, (uint32_t)z),
                      tile + (uint32_t)lo * ROWB, bytes, mapa_shared(smem_u32(red_bar), (uint32_t)z));
    }
    if (warp == 2 && lane == 0) trace_mark(p.trace, 15);
    mbar_wait_cluster(red_bar, 0);
  }
  const uint32_t stg = split ? rcv : tile;

  // ---- epilogue, phase 2 (warps 2..9): sum the K-split partials from local shared memory and run the
  // fused epilogue with coalesced global accesses.  These warps run alone on their schedulers, so the
  // code is bound by its own instruction latencies: UNR rows are processed as independent straight-line
  // chains.
  if (warp >= 2 && !(p.debug_skip & 2)) {
    if (warp == 2 && lane == 0) trace_mark(p.trace, 8);
    int nbatch = 0;
    const uint32_t slot_stride = (uint32_t)(R * G) << 4;
    if (!geglu) {
#pragma unroll 1
      for (int base = base0; base < Rz; base += STEP) {
        if (base != base0) epi_fetch(base, ops);  // (the first batch was fetched under the main loop)
        float4 acc[UNR];
        uint32_t sa[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int lrc = min(base + u * NW * RPW, Rz - 1);
          sa[u] = stg + ((uint32_t)(lrc * G + (g ^ (lrc & 7))) << 4);
          acc[u] = bias;
        }
#pragma unroll 1
        for (int sidx = 0; sidx < S; ++sidx) {
#pragma unroll
          for (int u = 0; u < UNR; ++u) acc[u] = f4_add(acc[u], lds_f4(sa[u] + (uint32_t)sidx * slot_stride));
        }
        if (warp == 2 && lane == 0 && nbatch == 0) trace_mark(p.trace, 14);
#pragma unroll
        for (int u = 0; u < UNR; ++u) acc[u] = f4_add(f4_add(acc[u], ops.rv[u]), ops.res[u]);
        if (p.act == ACT_SILU) {
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            acc[u].x = silu_f(acc[u].x); acc[u].y = silu_f(acc[u].y);
            acc[u].z = silu_f(acc[u].z); acc[u].w = silu_f(acc[u].w);
          }
        } else if (p.act == ACT_RELU) {
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            acc[u].x = fmaxf(acc[u].x, 0.f); acc[u].y = fmaxf(acc[u].y, 0.f);
            acc[u].z = fmaxf(acc[u].z, 0.f); acc[u].w = fmaxf(acc[u].w, 0.f);
          }
        }
        if (p.debug_skip & 4) {  // diagnostics: everything but the global stores
          float z = 0.f;
#pragma unroll
          for (int u = 0; u < UNR; ++u) z += acc[u].x + acc[u].y + acc[u].z + acc[u].w;
          if (z == 1.2345e-30f) p.out_f32[0] = z;
          continue;
        }
        if (p.out_f32 != nullptr) {
#pragma unroll
          for (int u = 0; u < UNR; ++u)
            if (ops.m[u] >= 0) *reinterpret_cast<float4*>(p.out_f32 + (long)ops.m[u] * p.ldo + col) = acc[u];
        }
        if (p.out_f16 != nullptr) {
#pragma unroll
          for (int u = 0; u < UNR; ++u)
            if (ops.m[u] >= 0) *reinterpret_cast<uint2*>(p.out_f16 + (long)ops.m[u] * p.ldo + col) = f4_to_h4(acc[u]);
        }
        if (warp == 2 && lane == 0 && nbatch < 3) trace_mark(p.trace, 9 + nbatch);
        ++nbatch;
      }
    } else {
      // GEGLU (attention_openai.py:42-44): tile columns [0,BN/2) are the value half, [BN/2,BN) the gate
      // half of the same BN/2 output features (weights were interleaved per tile when packed).  A lane
      // owns one float4 of values and the matching float4 of gates: G/2 lanes per row.
      constexpr int G2 = G / 2, RPW2 = 32 / G2, UG = 4;
      const int g2 = lane % G2, sr2 = lane / G2;
      const int colg = (int)blockIdx.x * (BN / 2) + 4 * g2;
#pragma unroll 1
      for (int base = (warp - 2) * RPW2 + sr2; base < Rz; base += NW * RPW2 * UG) {
        int m[UG];
        uint32_t sa[UG], sg[UG];
        float4 va[UG], vg[UG];
#pragma unroll
        for (int u = 0; u < UG; ++u) {
          const int lr = base + u * NW * RPW2;
          const int lrc = min(lr, Rz - 1);
          const int mi = lds_i2(tab + 8u * (uint32_t)lrc).x;
          m[u] = (lr < Rz) ? mi : -1;
          sa[u] = stg + ((uint32_t)(lrc * G + (g2 ^ (lrc & 7))) << 4);
          sg[u] = stg + ((uint32_t)(lrc * G + ((g2 + G2) ^ (lrc & 7))) << 4);
          va[u] = bias;
          vg[u] = bias_g;
        }
#pragma unroll 1
        for (int sidx = 0; sidx < S; ++sidx) {
#pragma unroll
          for (int u = 0; u < UG; ++u) {
            va[u] = f4_add(va[u], lds_f4(sa[u] + (uint32_t)sidx * slot_stride));
            vg[u] = f4_add(vg[u], lds_f4(sg[u] + (uint32_t)sidx * slot_stride));
          }
        }
#pragma unroll
        for (int u = 0; u < UG; ++u) {
          float4 v = va[u];
          v.x *= gelu_erf_f(vg[u].x); v.y *= gelu_erf_f(vg[u].y);
          v.z *= gelu_erf_f(vg[u].z); v.w *= gelu_erf_f(vg[u].w);
          if (m[u] >= 0) *reinterpret_cast<uint2*>(p.out_f16 + (long)m[u] * p.ldo + colg) = f4_to_h4(v);
        }
      }
    }
    if (warp == 2 && lane == 0) trace_mark(p.trace, 6);
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
  if (threadIdx.x == 0) trace_mark(p.trace, 7);
}

// =========================================================================== host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes /* rank-1 entries */, const uint32_t* box,
                  CUtensorMapSwizzle swz) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available (driver too old?)");
    return -3;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                   gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    std::string s = "cuTensorMapEncodeTiled failed, CUresult=" + std::to_string((int)r) + " rank=" +
                    std::to_string(rank) + " dims=";
    for (int i = 0; i < rank; ++i) s += std::to_string((unsigned long long)dims[i]) + ",";
    s += " box=";
    for (int i = 0; i < rank; ++i) s += std::to_string(box[i]) + ",";
    set_error(s);
    return -3;
  }
  return 0;
}

IGemmGeom gemm_geom(int M, int K) {
  IGemmGeom g;
  memset(&g, 0, sizeof(g));
  g.B = 1; g.T = 1; g.H = 1; g.W = M; g.C = K;
  g.bb = 1; g.bt = 1; g.bh = 1; g.bw = 128;
  g.ntaps = 1;
  return g;
}

IGemmGeom conv_taps_geom(int B, int T, int H, int W, int C, int kt, int kh, int kw) {
  IGemmGeom g;
  memset(&g, 0, sizeof(g));
  g.B = B; g.T = T; g.H = H; g.W = W; g.C = C;
  // 128 output positions per tile, filled W-first; each box extent divides 128
  int rem = 128;
  auto take = [&](int extent) {
    int b = std::min(extent, rem);
    while (rem % b) --b;
    rem /= b;
    return b;
  };
  g.bw = take(W); g.bh = take(H); g.bt = take(T); g.bb = rem;
  g.ntaps = 0;
  for (int a = 0; a < kt; ++a)
    for (int b = 0; b < kh; ++b)
      for (int c = 0; c < kw; ++c) {
        g.dt[g.ntaps] = (int8_t)(a - kt / 2); g.dh[g.ntaps] = (int8_t)(b - kh / 2);
        g.dw[g.ntaps] = (int8_t)(c - kw / 2);
        ++g.ntaps;
      }
  return g;
}

IGemmGeom conv3x3_geom(int B, int H, int W, int C) {
  IGemmGeom g;
  memset(&g, 0, sizeof(g));
  g.B = B; g.T = 1; g.H = H; g.W = W; g.C = C;
  // 128 output positions per tile: as much of a row as fits, then rows, then samples
  int bw = std::min(W, 128);
  while (128 % bw) --bw;  // W is a power of two in this model; keep it general anyway
  int rem = 128 / bw;
  int bh = std::min(H, rem);
  while (rem % bh) --bh;
  rem /= bh;
  g.bw = bw; g.bh = bh; g.bt = 1; g.bb = rem;
  g.ntaps = 9;
  int i = 0;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      g.dt[i] = 0; g.dh[i] = (int8_t)dy; g.dw[i] = (int8_t)dx;
      ++i;
    }
  return g;
}

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}


// Measured plan table: (M, N, K, taps, GEGLU?) -> (tile width, K-splits, ring depth), generated on a
// B200 by tools/autotune_igemm.py from event-timed, cold-weight replays of every distinct GEMM of the
// UNet plans (B_eff = 2 and 16).  Shapes that are not listed fall back to the cost model below.
struct IGemmTuned { int M, N, K, ntaps, geglu, bn, splits, deep; };
static const IGemmTuned kTuned[] = {
#include "igemm_tuned.inc"
    {0, 0, 0, 0, 0, 0, 0, 0}};
static int g_force_bn = 0, g_force_deep = -1;
void igemm_force(int bn, int deep) { g_force_bn = bn; g_force_deep = deep; }

int igemm_plan(IGemmPlan* plan, const __half* A, const __half* Wt, int N, const IGemmGeom& g,
               const IGemmEpilogue& e, int splits) {
  if (g.C % BLOCK_K != 0) {
    set_error("igemm: channel count must be a multiple of 64, got " + std::to_string(g.C));
    return -1;
  }
  if (g.bb * g.bt * g.bh * g.bw != BLOCK_M) {
    set_error("igemm: tile box must cover exactly 128 output positions");
    return -1;
  }
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(Wt) & 15)) {
    set_error("igemm: operand pointers must be 16-byte aligned");
    return -1;
  }
  plan->g = g;
  plan->e = e;
  plan->M = g.B * g.T * g.H * g.W;
  plan->N = N;
  plan->K = g.ntaps * g.C + g.C2;
  if (splits > 8) splits = 8;  // cluster size limit (portable)
  const bool geglu = (e.act == ACT_GEGLU);
  const int tw = (g.W + g.bw - 1) / g.bw, th = (g.H + g.bh - 1) / g.bh,
            tt = (g.T + g.bt - 1) / g.bt, tb = (g.B + g.bb - 1) / g.bb;
  plan->tiles_m = tw * th * tt * tb;
  if (g.C2 % BLOCK_K != 0 || (g.C2 > 0 && (g.A2 == nullptr || (reinterpret_cast<uintptr_t>(g.A2) & 15)))) {
    set_error("igemm: second A source needs a 16-byte aligned pointer and a channel count that is a multiple of 64");
    return -1;
  }
  const int kb_total = g.ntaps * (g.C / BLOCK_K) + g.C2 / BLOCK_K;
  plan->deep = g_force_deep;
  static const bool no_table = getenv("DFB_NO_TUNED") != nullptr;
  const IGemmTuned* tuned = nullptr;
  if (splits == 0 && g_force_bn == 0 && !no_table)
    for (const IGemmTuned* t = kTuned; t->M; ++t)
      if (t->M == plan->M && t->N == N && t->K == plan->K && t->ntaps == g.ntaps && t->geglu == (int)geglu) { tuned = t; break; }
  // ---- tile width BN and split-K factor (= cluster size, <= 8).  Tiny cost model: a CTA moves one
  // (A rows that exist + BN weight rows) x 128 B stage per k-block at ~55 GB/s (its share of L2
  // bandwidth), pays ~1 us for a DSMEM reduction, and the grid runs in ceil(ctas / #SMs) waves.
  {
    const int nsm = num_sms();
    const int rows_per_tile = std::min(BLOCK_M, plan->M);  // rows TMA really fetches (rest is zero fill)
    double best = 1e30;
    int best_bn = 128, best_s = 1;
    int bn_lo = (geglu ? 128 : 64), bn_hi = (N <= 64 ? 64 : 128);
    if (g_force_bn == 64 || g_force_bn == 128) bn_lo = bn_hi = (geglu ? 128 : g_force_bn);
    for (int bn = bn_hi; bn >= bn_lo; bn /= 2) {
      const int tiles = plan->tiles_m * ((N + bn - 1) / bn);
      const int smax = (splits > 0) ? splits : 8;
      for (int sp = (splits > 0 ? splits : 1); sp <= smax; ++sp) {
        if (sp > kb_total) break;
        const int kb_cta = (kb_total + sp - 1) / sp;
        const double stage_kb = (rows_per_tile + bn) * 128.0 / 1024.0;
        const double waves = std::ceil((double)tiles * sp / nsm);
        // per-CTA fill at ~55 KB/us, but the whole grid cannot pull more than ~3.3-4 GB/ms through
        // L2 (measured: tools/microbench_deep.py) -- the N-tile CTAs of a k-block all re-read the
        // same A lines, so narrower tiles / more CTAs do not help once that limit is reached
        const double t_cta = waves * kb_cta * stage_kb / 55.0;
        const double t_grid = (double)tiles * sp * kb_cta * stage_kb / 4000.0;
        const double t = std::max(t_cta, t_grid) + (sp > 1 ? 1.0 : 0.0) + 0.05 * (bn / 16) * waves + 3.0;
        if (t < best - 1e-9) { best = t; best_bn = bn; best_s = sp; }
      }
    }
    plan->BN = best_bn;
    splits = best_s;
    if (tuned != nullptr && tuned->splits <= kb_total) {
      plan->BN = tuned->bn;
      splits = tuned->splits;
      if (plan->deep < 0) plan->deep = tuned->deep;
    }
  }
  const int BN = plan->BN;
  plan->tiles_n = (N + BN - 1) / BN;
  plan->splits = splits;
  if ((e.out_f16 && (e.ldo % 8)) || (e.out_f32 && (e.ldo % 4)) || (e.residual && (e.ld_res % 4)) || (e.residual_f16 && (e.ld_res % 8)) ||
      (N % 16) != 0) {
    set_error("igemm: N must be a multiple of 16 and output/residual row strides 16-byte aligned");
    return -1;
  }
  if (e.out_f16 == nullptr && e.out_f32 == nullptr) {
    set_error("igemm: no output pointer");
    return -1;
  }
  if (geglu && (e.out_f16 == nullptr || e.bias == nullptr || (N % 128) != 0)) {
    set_error("igemm: GEGLU epilogue needs fp16 output, bias and N % 128 == 0");
    return -1;
  }
  // ---- tensor maps
  {
    uint64_t dims[5] = {(uint64_t)g.C, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.T, (uint64_t)g.B};
    uint64_t str[4];
    str[0] = (uint64_t)g.C * 2;
    str[1] = str[0] * g.W;
    str[2] = str[1] * g.H;
    str[3] = str[2] * g.T;
    uint32_t box[5] = {BLOCK_K, (uint32_t)g.bw, (uint32_t)g.bh, (uint32_t)g.bt, (uint32_t)g.bb};
    int rc = make_tmap_f16(&plan->tmA, A, 5, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    plan->tmA2 = plan->tmA;
    if (g.C2 > 0) {
      dims[0] = (uint64_t)g.C2;
      str[0] = (uint64_t)g.C2 * 2;
      str[1] = str[0] * g.W;
      str[2] = str[1] * g.H;
      str[3] = str[2] * g.T;
      rc = make_tmap_f16(&plan->tmA2, g.A2, 5, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
  }
  {
    uint64_t dims[2] = {(uint64_t)plan->K, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)plan->K * 2};
    uint32_t box[2] = {BLOCK_K, (uint32_t)BN};
    int rc = make_tmap_f16(&plan->tmW, Wt, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  return 0;
}

int igemm_init() {
  DFB_CUDA_OK(cudaFuncSetAttribute(igemm_tcgen05_kernel<64, 4>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   IGemmSmem<64, 4>::DYN_BYTES));
  DFB_CUDA_OK(cudaFuncSetAttribute(igemm_tcgen05_kernel<128, 3>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   IGemmSmem<128, 3>::DYN_BYTES));
  DFB_CUDA_OK(cudaFuncSetAttribute(igemm_tcgen05_kernel<64, 8>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   IGemmSmem<64, 8>::DYN_BYTES));
  DFB_CUDA_OK(cudaFuncSetAttribute(igemm_tcgen05_kernel<128, 6>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   IGemmSmem<128, 6>::DYN_BYTES));
  return 0;
}

template <int BN, int STAGES>
static int launch_t(const IGemmPlan& plan, const IGemmKParams& kp, cudaStream_t stream) {
  using L = IGemmSmem<BN, STAGES>;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(plan.tiles_n, plan.tiles_m, plan.splits);
  cfg.blockDim = dim3(IGEMM_THREADS);
  cfg.dynamicSmemBytes = L::DYN_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;  // the K-splits of a tile share a cluster
  attr[1].val.clusterDim.x = 1;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = plan.splits;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  DFB_CUDA_OK(cudaLaunchKernelEx(&cfg, igemm_tcgen05_kernel<BN, STAGES>, plan.tmA, plan.tmW, plan.tmA2, kp));
  return 0;
}

int igemm_launch(const IGemmPlan& plan, cudaStream_t stream) {
  IGemmKParams kp;
  memset(&kp, 0, sizeof(kp));
  const IGemmGeom& g = plan.g;
  kp.M = plan.M; kp.N = plan.N;
  kp.B = g.B; kp.T = g.T; kp.H = g.H; kp.W = g.W;
  kp.bb = g.bb; kp.bt = g.bt; kp.bh = g.bh; kp.bw = g.bw;
  kp.tw = (g.W + g.bw - 1) / g.bw;
  kp.th = (g.H + g.bh - 1) / g.bh;
  kp.tt = (g.T + g.bt - 1) / g.bt;
  kp.ntaps = g.ntaps;
  kp.kpt = g.C / BLOCK_K;
  kp.kpt2 = g.C2 / BLOCK_K;
  for (int i = 0; i < 9; ++i) { kp.dt[i] = g.dt[i]; kp.dh[i] = g.dh[i]; kp.dw[i] = g.dw[i]; }
  kp.out_f32 = plan.e.out_f32; kp.out_f16 = plan.e.out_f16; kp.ldo = plan.e.ldo;
  kp.bias = plan.e.bias; kp.rowvec = plan.e.rowvec; kp.ld_rowvec = plan.e.ld_rowvec;
  kp.rows_per_sample = plan.e.rows_per_sample;
  kp.rowvec_row = plan.e.rowvec_row;
  kp.residual = plan.e.residual; kp.residual_f16 = plan.e.residual_f16; kp.ld_res = plan.e.ld_res;
  kp.act = plan.e.act;
  kp.splits = plan.splits;
  kp.next_w = reinterpret_cast<const uint8_t*>(plan.next_w);
  kp.next_w_bytes = plan.next_w_bytes;
  kp.trace = trace_record();
  static const int dbg_skip = getenv("DFB_DEBUG_SKIP") ? atoi(getenv("DFB_DEBUG_SKIP")) : 0;
  kp.debug_skip = dbg_skip;
  {
    const double out_b = (plan.e.out_f32 ? 4.0 : 0.0) + (plan.e.out_f16 ? 2.0 : 0.0);
    note(g.ntaps == 9 ? "igemm_conv3x3" : "igemm_linear", 2.0 * plan.M * plan.N * plan.K,
         2.0 * plan.N * plan.K + 2.0 * plan.M * (g.C + g.C2) + out_b * plan.M * plan.e.ldo +
             (plan.e.residual ? 4.0 * plan.M * plan.e.ldo : 0.0),
         plan.M, plan.N, plan.K, plan.splits, plan.tiles_m * plan.tiles_n * plan.splits);
  }
  // Shallow operand ring (~97 KB) by default: two CTAs fit on an SM, so (a) one CTA's epilogue overlaps
  // the other's main loop and (b) under programmatic dependent launch the NEXT kernel's CTAs become
  // resident -- and stream their first weight tiles -- while this kernel is still running.  Measured
  // 3.37 ms/step vs 3.46 with the deep ring everywhere (DFB_SHALLOW=0 selects the deep ring).
  static const int force_shallow = getenv("DFB_SHALLOW") ? atoi(getenv("DFB_SHALLOW")) : -1;
  // split-K with 128-wide tiles stages its own tile (64 KB) plus the received partials (<= 68 KB): that
  // needs the deep ring's shared memory
  const bool shallow = (plan.deep >= 0 ? plan.deep == 0 : force_shallow != 0) && !(plan.BN == 128 && plan.splits > 1);
  if (plan.BN == 64) return shallow ? launch_t<64, 4>(plan, kp, stream) : launch_t<64, 8>(plan, kp, stream);
  return shallow ? launch_t<128, 3>(plan, kp, stream) : launch_t<128, 6>(plan, kp, stream);
}

}  // namespace dfb
