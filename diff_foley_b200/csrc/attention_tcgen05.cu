// attention_tcgen05.cu -- fused softmax(q k^T * scale) v for the UNet's self- and cross-attention on
// the 5th-gen tensor cores (tcgen05) with both accumulators in TMEM.
//
// One CTA owns a 128-query tile of one (sample, head).  Per block of KB keys:
//   S = Q K^T        tcgen05.mma, fp16 operands from shared memory, fp32 S[128 x KB] in TMEM (double
//                    buffered: S_{j+1} is issued while the softmax warps work on S_j), KB <= 64
//   P = softmax-blk  4 warps, one query row per thread (row = TMEM lane, no shuffles): ONE tcgen05.ld
//                    pass pulls the thread's whole S row of the block into registers,
//                    running max / sum in registers, exp2 with the d^-0.5 scale folded in, P written
//                    as fp16 into shared memory in the MMA's canonical K-major layout
//   O += P V         tcgen05.mma with V as an MN-major B operand (no transposed copy of V); O[128 x d]
//                    stays in TMEM for the whole key loop and is rescaled in place (tcgen05.ld/st)
//                    only by warps whose running max actually moved
// so the [B*8, Nq, Nk] score tensor the reference materialises in HBM (attention_openai.py:178-190:
// einsum -> softmax -> einsum, 32 MB per sample at the 16x64 level) never exists.
//
// Operands are the fp16 projections written by the QKV GEMM epilogue (head h at columns h*dpad of
// each row, dpad = head dim padded to a multiple of 16 with zero columns).  TMA loads them as one
// (8 halves x rows) box per 16-byte chunk, which lands them in shared memory as
// [chunk][row][8 halves] -- the un-swizzled canonical core-matrix layout (8 rows x 16 B contiguous)
// that serves as K-major A/B operand (Q, K) *and* as MN-major B operand (V) without any shuffle.
// The output is fp16 [B*Lq, heads*d], the A operand of the to_out GEMM -- the reference's
// '(b h) n d -> b n (h d)' rearrange is free.
//
// warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..9 = softmax / correction /
// epilogue.  The softmax warps are bound by their own instruction stream (one query row per thread,
// ~6 instructions per score), so TWO warps share each TMEM lane quarter (warp_idx % 4): warps 2..5 take
// the even 16-key chunks of a block, warps 6..9 the odd ones; the two threads of a row exchange their
// partial row maxima through shared memory (one 64-thread named barrier per block), keep separate
// partial row sums, and split the O columns for the rescale and the final normalisation.
#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

#include "dfb_internal.h"
#include "dfb_ptx.cuh"

namespace dfb {

int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz);

constexpr int ATT_THREADS = 320;  // 1 TMA + 1 MMA warp + 2 x 4 softmax warps
constexpr int ATT_BM = 128;  // queries per CTA
constexpr int ATT_POLY_DEFAULT = 0;

struct AttParams {
  __half* out;
  int ldo;
  int heads, Lq, Lk, d, dpad, KB, nblocks, tmem_cols;
  int kvs;  // depth of the K/V tile ring (2..4)
  float scale_log2;  // d^-0.5 * log2(e)
  unsigned long long* trace;  // optional timeline record (diagnostics)
};

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// smem carve-up (bytes), all tiles in the [chunk][row][8 halves] layout.  K/V tiles live in a ring of `kvs`
// stages (K then V per stage): with only two, a stage is refilled after the P V of its previous block,
// and the ~1 us L2 round trip of that load -- not the softmax -- set the pace of the key loop (ncu source
// view: 31 % of all warp samples on the softmax warps' wait for S).
struct AttSmem {
  int q_bytes, kv_tile_bytes, p_bytes;
  int off_q, off_kv, off_p, off_bar, off_xch, total;
};
constexpr int ATT_MAX_KVS = 4;
__host__ __device__ inline AttSmem att_smem_layout(int dpad, int KB, int kvs) {
  AttSmem s;
  s.q_bytes = ATT_BM * dpad * 2;
  s.kv_tile_bytes = KB * dpad * 2;
  s.p_bytes = ATT_BM * KB * 2;
  int o = 0;
  s.off_q = o; o += s.q_bytes;
  s.off_kv = o; o += kvs * 2 * s.kv_tile_bytes;
  s.off_p = o; o += 2 * s.p_bytes;
  s.off_bar = (o + 15) & ~15;
  s.off_xch = s.off_bar + 16 * 8 + 16;            // float [2 parities][2 halves][128 rows] maxima + [2][128] sums
  s.total = s.off_xch + (4 + 2) * ATT_BM * 4;
  return s;
}

// Resources are sized so that two CTAs share an SM for head dims <= 128 (<= 97 KB of shared memory,
// 256 TMEM columns): one CTA's softmax overlaps the other's tensor-core work and loads.
// NQ = 16-column chunks of S per key block: 4 (KB <= 64; 2 CTAs/SM -- grids larger than the machine) or
// 8 (KB <= 128; half as many softmax/MMA round trips per CTA -- small grids, latency-bound).
// POLY = how many of every four exponentials are evaluated by ex2_poly on the FMA pipe instead of MUFU.EX2
template <int NQ, int POLY>
__global__ void __launch_bounds__(ATT_THREADS, (NQ <= 4) ? 2 : 1)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttParams p) {
  extern __shared__ uint8_t att_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(att_raw) + 127) &
                                             ~static_cast<uintptr_t>(127));
  const AttSmem L = att_smem_layout(p.dpad, p.KB, p.kvs);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bar);
  uint64_t* q_full = bars + 0;
  uint64_t* s_full = bars + 1;     // [2]
  uint64_t* p_full = bars + 3;     // [2], 256 arrivals each
  uint64_t* pv_done = bars + 5;
  uint64_t* kv_full = bars + 6;                 // [kvs]
  uint64_t* kv_empty = bars + 6 + ATT_MAX_KVS;  // [kvs]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6 + 2 * ATT_MAX_KVS);
  const int KVS = p.kvs;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) trace_mark(p.trace, 0);
  const int bh = blockIdx.y;
  const int b = bh / p.heads, h = bh % p.heads;
  const int q0 = blockIdx.x * ATT_BM;
  const int nb = p.nblocks;
  const int KB = p.KB;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
  }
  if (warp == 1) {
    if (elect_one()) {
      mbar_init(q_full, 1);
      for (int i = 0; i < KVS; ++i) {
        mbar_init(&kv_full[i], 1);
        mbar_init(&kv_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&p_full[i], 256);
      }
      mbar_init(pv_done, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // q/k/v come from the preceding GEMM
  if (threadIdx.x == 0) trace_mark(p.trace, 1);
  pdl_launch_dependents();  // only after our own wait: at most two grids of the chain overlap
  const uint32_t tmem_base = *tmem_slot;
  // S double buffer: 2 x (NQ*16 fp32 columns) at tmem_base + s*NQ*16 (arithmetic, not a runtime-indexed
  // array: that would live in local memory)
  const uint32_t tm_o = tmem_base + 2 * NQ * 16;               // O: dpad fp32 columns

  if (warp == 0) {
    // ======================================================================== TMA producer
    // (single-thread roles are entered through elect_one(), see dfb_ptx.cuh: the loop stays on the uniform datapath)
    if (elect_one()) {
      int st = 0;
      uint32_t eph = 1;
      // one 2-D box (8 halves x rows) per 16-byte chunk: chunk c lands at tile + c*rows*16
      const int nch = p.dpad >> 3, col0 = h * p.dpad;
      mbar_expect_tx(q_full, L.q_bytes);
      for (int c = 0; c < nch; ++c)
        tma_load_2d(smem + L.off_q + c * (ATT_BM * 16), &tmQ, q_full, col0 + 8 * c, b * p.Lq + q0);
      for (int j = 0; j < nb; ++j) {
        mbar_wait(&kv_empty[st], eph);
        mbar_expect_tx(&kv_full[st], 2 * L.kv_tile_bytes);
        uint8_t* kt = smem + L.off_kv + st * 2 * L.kv_tile_bytes;
        for (int c = 0; c < nch; ++c) {
          tma_load_2d(kt + c * (KB * 16), &tmK, &kv_full[st], col0 + 8 * c, b * p.Lk + j * KB);
          tma_load_2d(kt + L.kv_tile_bytes + c * (KB * 16), &tmV, &kv_full[st], col0 + 8 * c, b * p.Lk + j * KB);
        }
        if (++st == KVS) { st = 0; eph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ========================================================================= MMA issuer
    const uint32_t idesc_s = umma_idesc_f16(ATT_BM, KB, 0, 0);
    const uint32_t idesc_o = umma_idesc_f16(ATT_BM, p.dpad, 0, 1);  // B = V is MN-major
    const uint32_t q_lbo = ATT_BM * 16, k_lbo = KB * 16;
    auto desc = [&](uint32_t addr, uint32_t lbo, uint32_t sbo) {
      return umma_desc_nosw(addr, lbo, sbo);
    };
    const uint32_t sq = smem_u32(smem + L.off_q);
    auto issue_s = [&](int j) {
      const int s = j & 1;
      const uint32_t sk = smem_u32(smem + L.off_kv + (j % KVS) * 2 * L.kv_tile_bytes);
#pragma unroll 1
      for (int k = 0; k < p.dpad / 16; ++k)
        umma_f16_ss(tmem_base + (uint32_t)(s * NQ * 16), desc(sq + k * 2 * q_lbo, q_lbo, 128), desc(sk + k * 2 * k_lbo, k_lbo, 128),
                    idesc_s, k > 0 ? 1u : 0u);
      umma_commit(&s_full[s]);
    };
    if (elect_one()) {
    mbar_wait(q_full, 0);
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    issue_s(0);
    for (int j = 0; j < nb; ++j) {
      const int s = j & 1;
      if (j + 1 < nb) {
        const int s1 = (j + 1) & 1;
        mbar_wait(&kv_full[(j + 1) % KVS], ((j + 1) / KVS) & 1);
        // S buffer s1 was last read by the softmax of block j-1 (done once P_{j-1} is full)
        if (j >= 1) mbar_wait(&p_full[s1], ((j - 1) >> 1) & 1);
        tc_fence_after();
        issue_s(j + 1);
      }
      mbar_wait(&p_full[s], (j >> 1) & 1);
      tc_fence_after();
      {
        const uint32_t sp = smem_u32(smem + L.off_p + s * L.p_bytes);
        const uint32_t sv = smem_u32(smem + L.off_kv + (j % KVS) * 2 * L.kv_tile_bytes + L.kv_tile_bytes);
#pragma unroll 1
        for (int k = 0; k < KB / 16; ++k)
          umma_f16_ss(tm_o, desc(sp + k * 2 * q_lbo, q_lbo, 128),
                      // V: MN-major; 8-key groups are 128 B apart (LBO), 8-column groups KB*16 B (SBO)
                      desc(sv + k * 256, 128, k_lbo), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(pv_done);
        umma_commit(&kv_empty[j % KVS]);
      }
    }
    }
    __syncwarp();
  } else {
    // ============================================================ softmax / correction / epilogue
    const int sub = warp & 3;
    const int half = (warp - 2) >> 2;   // 0: even 16-key chunks of a block, 1: odd ones
    const int r = sub * 32 + lane;      // query row within the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(sub * 32) << 16;
    float* xmax = reinterpret_cast<float*>(smem + L.off_xch);   // [parity][half][row]
    float* xsum = xmax + 4 * ATT_BM;                             // [half][row]
    constexpr int NQH = NQ / 2;         // chunks per thread per block
    float m_run = -INFINITY, l_run = 0.f;  // l_run: this thread's share of the row sum
    const float sl = p.scale_log2;
    const uint32_t p_base = smem_u32(smem + L.off_p) + (uint32_t)r * 16u;
    for (int j = 0; j < nb; ++j) {
      const int s = j & 1;
      mbar_wait(&s_full[s], (j >> 1) & 1);
      tc_fence_after();
      const int kvalid = min(KB, p.Lk - j * KB);  // keys of this block that exist
      const bool full = (kvalid == KB);           // CTA-uniform: only a ragged last block is masked
      const uint32_t prow = p_base + (uint32_t)s * (uint32_t)L.p_bytes;
      const uint32_t ts = tmem_base + (uint32_t)(s * NQ * 16) + lane_off;
      // ---- one TMEM pass: this thread's chunks of the S row into registers
      uint32_t raw[NQH][16];
#pragma unroll
      for (int q = 0; q < NQH; ++q)
        if ((2 * q + half) * 16 < KB) tmem_ld_32x16(ts + (2 * q + half) * 16, raw[q]);
      tmem_ld_wait();
      // the softmax warps are bound by their own instruction stream: the common (full-block) path is
      // FMNMX / FFMA / MUFU.EX2 / FADD / half a F2FP per score, nothing else; four independent chains
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      if (full) {
#pragma unroll
        for (int q = 0; q < NQH; ++q)
          if ((2 * q + half) * 16 < KB) {
#pragma unroll
            for (int i = 0; i < 16; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(raw[q][i]));
          }
      } else {
#pragma unroll
        for (int q = 0; q < NQH; ++q)
          if ((2 * q + half) * 16 < KB) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if ((2 * q + half) * 16 + i < kvalid) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(raw[q][i]));
          }
      }
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      // ---- row maximum across the two threads of this row
      xmax[(s * 2 + half) * ATT_BM + r] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + sub) : "memory");
      mx = fmaxf(mx, xmax[(s * 2 + (half ^ 1)) * ATT_BM + r]);
      const float m_new = fmaxf(m_run, mx * sl);
      const float alpha = ex2_approx(m_run - m_new);  // 0 on the first block (m_run = -inf)
      const float neg_m = -m_new;
      // ---- p = exp2(s*scale - m), partial row sum, fp16 P tile in the canonical K-major layout
      float ps4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int q = 0; q < NQH; ++q)
        if ((2 * q + half) * 16 < KB) {
          const int c0 = (2 * q + half) * 16;
          float e[16];
          if (full) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float xs = fmaf(__uint_as_float(raw[q][i]), sl, neg_m);
              e[i] = ((i & 3) < POLY) ? ex2_poly(xs) : ex2_approx(xs);
              ps4[i & 3] += e[i];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float v = ex2_approx(fmaf(__uint_as_float(raw[q][i]), sl, neg_m));
              e[i] = (c0 + i < kvalid) ? v : 0.f;
              ps4[i & 3] += e[i];
            }
          }
#pragma unroll
          for (int g = 0; g < 2; ++g)
            sts_u4(prow + (uint32_t)(((c0 >> 3) + g) * (ATT_BM * 16)), pack_h2(e[8 * g + 0], e[8 * g + 1]),
                   pack_h2(e[8 * g + 2], e[8 * g + 3]), pack_h2(e[8 * g + 4], e[8 * g + 5]),
                   pack_h2(e[8 * g + 6], e[8 * g + 7]));
        }
      l_run = fmaf(l_run, alpha, (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]));
      m_run = m_new;
      fence_proxy_async_smem();  // generic-proxy P stores -> visible to the tensor core (async proxy)
      // ---- correction: O *= alpha, in TMEM, once the previous P V has landed (columns split by half)
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll 1
          for (int c = half * 16; c < p.dpad; c += 32) {
            uint32_t rw[16];
            tmem_ld_32x16(tm_o + lane_off + c, rw);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) rw[i] = __float_as_uint(__uint_as_float(rw[i]) * alpha);
            tmem_st_32x16(tm_o + lane_off + c, rw);
          }
          tmem_st_wait();
        }
      }
      tc_fence_before();
      mbar_arrive(&p_full[s]);
    }
    // ---- epilogue: O / l -> fp16 [row, h*d + c]; the row sum is the two threads' shares
    xsum[half * ATT_BM + r] = l_run;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + sub) : "memory");
    const float inv = 1.f / (l_run + xsum[(half ^ 1) * ATT_BM + r]);
    mbar_wait(pv_done, (nb - 1) & 1);
    tc_fence_after();
    const bool row_ok = (q0 + r) < p.Lq;
    __half* orow = p.out + ((size_t)b * p.Lq + q0 + r) * p.ldo + h * p.d;
#pragma unroll 1
    for (int c = half * 16; c < p.dpad; c += 32) {
      uint32_t rw[16];
      tmem_ld_32x16(tm_o + lane_off + c, rw);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          if (c + i < p.d)
            *reinterpret_cast<__half2*>(orow + c + i) =
                __floats2half2_rn(__uint_as_float(rw[i]) * inv, __uint_as_float(rw[i + 1]) * inv);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
  if (threadIdx.x == 0) trace_mark(p.trace, 7);
}

int attention_init() {
#define DFB_ATT_ATTR(NQ, PL)                                                                                    \
  DFB_CUDA_OK(cudaFuncSetAttribute(attention_tcgen05_kernel<NQ, PL>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                   200 * 1024))
  DFB_ATT_ATTR(4, 0); DFB_ATT_ATTR(4, 1); DFB_ATT_ATTR(4, 2);
  DFB_ATT_ATTR(8, 0); DFB_ATT_ATTR(8, 1); DFB_ATT_ATTR(8, 2);
#undef DFB_ATT_ATTR
  return 0;
}

// tensor maps are cached: the engine's buffers are stable, so each (pointer, geometry) is encoded once
struct TmapKey {
  const void* p; int ld, rows, dpad, box_rows;
  bool operator<(const TmapKey& o) const {
    return std::tie(p, ld, rows, dpad, box_rows) < std::tie(o.p, o.ld, o.rows, o.dpad, o.box_rows);
  }
};
static std::map<TmapKey, CUtensorMap> g_tmaps;
static std::mutex g_tmap_mu;

static int get_tmap(CUtensorMap* out, const __half* base, int ld, long rows, int dpad, int box_rows) {
  TmapKey key{base, ld, (int)rows, dpad, box_rows};
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  auto it = g_tmaps.find(key);
  if (it != g_tmaps.end()) { *out = it->second; return 0; }
  // plain 2-D view [rows, ld] fp16; box = 8 halves (one 16-byte chunk) x box_rows
  uint64_t dims[2] = {(uint64_t)ld, (uint64_t)rows};
  uint64_t strides[1] = {(uint64_t)ld * 2};
  uint32_t box[2] = {8, (uint32_t)box_rows};
  int rc = make_tmap_f16(out, base, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (rc) return rc;
  if (g_tmaps.size() > 4096) g_tmaps.clear();
  g_tmaps[key] = *out;
  return 0;
}

int attention_launch(const __half* q, int ldq, const __half* k, int ldk, const __half* v, int ldv,
                     __half* out, int ldo, int B, int heads, int Lq, int Lk, int d, int dpad,
                     float scale, cudaStream_t stream) {
  if (d > 160 || (d & 1) || dpad < d || (dpad % 16) || (ldq % 8) || (ldk % 8) || (ldv % 8) || Lk < 1) {
    set_error("attention: need even head dim <= 160, dpad a multiple of 16 and row strides multiple of 8");
    return -1;
  }
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) {
    set_error("attention: q/k/v must be 16-byte aligned");
    return -1;
  }
  AttParams p;
  p.out = out; p.ldo = ldo; p.heads = heads; p.Lq = Lq; p.Lk = Lk; p.d = d; p.dpad = dpad;
  // key-block size: 128 when the grid does not even fill the machine once (fewer, longer round trips
  // per CTA) and the tiles fit; otherwise 64, which lets two CTAs share an SM
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
  const long ctas = (long)((Lq + ATT_BM - 1) / ATT_BM) * B * heads;
  const bool wide = (ctas <= n_sm) && dpad <= 96 && Lk > 64;
  const int nq = wide ? 8 : 4;
  p.KB = std::min(nq * 16, (Lk + 15) / 16 * 16);
  {
    const int need = 2 * nq * 16 + dpad;  // S double buffer + O
    p.tmem_cols = need <= 256 ? 256 : 512;
  }
  p.nblocks = (Lk + p.KB - 1) / p.KB;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.trace = trace_record();
  CUtensorMap tq, tk, tv;
  int rc = get_tmap(&tq, q, ldq, (long)B * Lq, dpad, ATT_BM);
  if (!rc) rc = get_tmap(&tk, k, ldk, (long)B * Lk, dpad, p.KB);
  if (!rc) rc = get_tmap(&tv, v, ldv, (long)B * Lk, dpad, p.KB);
  if (rc) return rc;
  // K/V ring depth: as deep as the key loop needs (<= 4) while two CTAs still share an SM when the head
  // dim allows it (~110 KB per CTA), else whatever fits in one SM's shared memory
  p.kvs = 2;
  for (int k = 3; k <= ATT_MAX_KVS && k <= std::max(2, p.nblocks); ++k) {
    const int tot = att_smem_layout(dpad, p.KB, k).total + 128;
    const bool two_ctas = (nq == 4) && att_smem_layout(dpad, p.KB, 2).total + 128 <= 110 * 1024;
    if (tot <= (two_ctas ? 110 : 200) * 1024) p.kvs = k;
  }
  const AttSmem L = att_smem_layout(dpad, p.KB, p.kvs);
  dim3 grid((Lq + ATT_BM - 1) / ATT_BM, B * heads);
  note("attention", 4.0 * B * heads * (double)Lq * Lk * d,
       2.0 * B * heads * ((double)Lq * d * 2 + (double)Lk * d * 2), Lq, Lk, d, 1, grid.x * grid.y);
  // share of the exponentials moved from MUFU to the FMA pipe (quarters); DFB_ATT_POLY=0|1|2 overrides
  static const int poly_env = getenv("DFB_ATT_POLY") ? atoi(getenv("DFB_ATT_POLY")) : -1;
  const int poly = poly_env >= 0 ? std::min(poly_env, 2) : ATT_POLY_DEFAULT;
#define DFB_ATT_LAUNCH(NQ, PL) \
  DFB_CUDA_OK(launch_pdl(attention_tcgen05_kernel<NQ, PL>, dim3(grid), dim3(ATT_THREADS), L.total + 128, stream, tq, tk, tv, p))
  if (nq == 8) {
    if (poly == 0) DFB_ATT_LAUNCH(8, 0); else if (poly == 1) DFB_ATT_LAUNCH(8, 1); else DFB_ATT_LAUNCH(8, 2);
  } else {
    if (poly == 0) DFB_ATT_LAUNCH(4, 0); else if (poly == 1) DFB_ATT_LAUNCH(4, 1); else DFB_ATT_LAUNCH(4, 2);
  }
#undef DFB_ATT_LAUNCH
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dfb
