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
constexpr int IGEMM_THREADS = 192;

struct IGemmKParams {
  int M, N;
  int B, T, H, W;
  int bb, bt, bh, bw;
  int tw, th, tt;  // tiles along W, H, T (tiles along B implied)
  int ntaps, kpt;  // taps, 64-channel k-blocks per tap
  int8_t dt[9], dh[9], dw[9];
  float* out_f32;
  __half* out_f16;
  int ldo;
  const float* bias;
  const float* rowvec;
  int ld_rowvec;
  int rows_per_sample;
  const float* residual;
  const __half* residual_f16;
  int ld_res;
  int act;
  int splits;  // == cluster size along z (<= 8): the CTAs of one output tile reduce through DSMEM
  const uint8_t* next_w;  // optional: weights of the next GEMM, prefetched into L2 slice-wise
  unsigned long long next_w_bytes;
};

template <int BN, int STAGES>
struct IGemmSmem {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int W_BYTES = BN * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + W_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  // full[STAGES], empty[STAGES], tmem_full, then tmem slot + flag
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;  // slack to align the base to 1024 B
};

// Epilogue on 16 consecutive columns of one accumulator row.  Kept deliberately compact (rolled
// column-chunk loops, one code path): the kernel is launched ~200x per UNet forward between other
// kernels, so it starts with a cold instruction cache every time -- a 100 KB fully unrolled
// epilogue cost ~10 us per launch in instruction fetch alone.
__device__ __forceinline__ void epilogue16(const IGemmKParams& p, float (&v)[16], long m, int bs,
                                           int n_base) {
  if (p.bias != nullptr) {
    const float4* bp = reinterpret_cast<const float4*>(p.bias + n_base);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 t = __ldg(bp + j);
      v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
    }
  }
  if (p.rowvec != nullptr) {
    const float4* rp = reinterpret_cast<const float4*>(p.rowvec + (long)bs * p.ld_rowvec + n_base);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 t = __ldg(rp + j);
      v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
    }
  }
  if (p.residual != nullptr) {
    const float4* rs = reinterpret_cast<const float4*>(p.residual + m * p.ld_res + n_base);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 t = rs[j];
      v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
    }
  }
  if (p.residual_f16 != nullptr) {
    const uint4* rs = reinterpret_cast<const uint4*>(p.residual_f16 + m * p.ld_res + n_base);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint4 u = rs[j];
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
      const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
      const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&u.z));
      const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&u.w));
      v[8 * j] += a.x; v[8 * j + 1] += a.y; v[8 * j + 2] += b.x; v[8 * j + 3] += b.y;
      v[8 * j + 4] += c.x; v[8 * j + 5] += c.y; v[8 * j + 6] += d.x; v[8 * j + 7] += d.y;
    }
  }
  if (p.act == ACT_SILU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = silu_f(v[j]);
  } else if (p.act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (p.out_f32 != nullptr) {
    float4* o = reinterpret_cast<float4*>(p.out_f32 + m * p.ldo + n_base);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
  if (p.out_f16 != nullptr) {
    uint4* o = reinterpret_cast<uint4*>(p.out_f16 + m * p.ldo + n_base);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      __half2 h0 = __floats2half2_rn(v[8 * j], v[8 * j + 1]);
      __half2 h1 = __floats2half2_rn(v[8 * j + 2], v[8 * j + 3]);
      __half2 h2 = __floats2half2_rn(v[8 * j + 4], v[8 * j + 5]);
      __half2 h3 = __floats2half2_rn(v[8 * j + 6], v[8 * j + 7]);
      uint4 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
      u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
      o[j] = u;
    }
  }
}

// GEGLU on 16 value columns + their 16 gate columns -> 16 fp16 outputs (attention_openai.py:42-44)
__device__ __forceinline__ void epilogue16_geglu(const IGemmKParams& p, const float (&a)[16],
                                                 const float (&g)[16], long m, int nbv, int nbg,
                                                 int out_col) {
  float o[16];
#pragma unroll
  for (int j = 0; j < 16; ++j)
    o[j] = (a[j] + __ldg(p.bias + nbv + j)) * gelu_erf_f(g[j] + __ldg(p.bias + nbg + j));
  uint4* dst = reinterpret_cast<uint4*>(p.out_f16 + m * p.ldo + out_col);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    __half2 h0 = __floats2half2_rn(o[8 * j], o[8 * j + 1]);
    __half2 h1 = __floats2half2_rn(o[8 * j + 2], o[8 * j + 3]);
    __half2 h2 = __floats2half2_rn(o[8 * j + 4], o[8 * j + 5]);
    __half2 h3 = __floats2half2_rn(o[8 * j + 6], o[8 * j + 7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
    u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
    dst[j] = u;
  }
}

// STAGES-deep operand ring; the shallow-ring instantiations (<=3-4 stages, ~97 KB) let two CTAs share
// an SM so one CTA's epilogue overlaps the other's main loop -- used for short-K problems.
template <int BN, int STAGES>
__global__ void __launch_bounds__(IGEMM_THREADS, (STAGES <= 4) ? 2 : 1)
igemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                     const __grid_constant__ CUtensorMap tmW,
                     const __grid_constant__ IGemmKParams p) {
  using L = IGemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates
  const int n0 = blockIdx.x * BN;
  int tm = blockIdx.y;
  const int tw_i = tm % p.tw; tm /= p.tw;
  const int th_i = tm % p.th; tm /= p.th;
  const int tt_i = tm % p.tt; tm /= p.tt;
  const int tb_i = tm;
  const int x0 = tw_i * p.bw, y0 = th_i * p.bh, t0 = tt_i * p.bt, b0 = tb_i * p.bb;

  const int kb_total = p.ntaps * p.kpt;
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
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // everything above overlapped the predecessor's tail; operands are valid from here
  pdl_launch_dependents();  // only after our own wait: at most two grids of the chain overlap
  const uint32_t tmem_base = *tmem_slot;
  const bool geglu = (p.act == ACT_GEGLU);
  const bool split = (p.splits > 1);
  // output row owned by this thread when it acts as an epilogue thread (warps 2..5)
  bool row_ok = false;
  long m = 0;
  int bs = 0;
  if (warp >= 2) {
    const int r = (warp & 3) * 32 + lane;
    const int w_i = r % p.bw;
    const int h_i = (r / p.bw) % p.bh;
    const int t_i = (r / (p.bw * p.bh)) % p.bt;
    const int b_i = r / (p.bw * p.bh * p.bt);
    const int x = x0 + w_i, y = y0 + h_i, t = t0 + t_i, b = b0 + b_i;
    row_ok = (x < p.W) && (y < p.H) && (t < p.T) && (b < p.B);
    m = (((long)b * p.T + t) * p.H + y) * p.W + x;
    bs = (p.rows_per_sample > 0) ? (int)(m / p.rows_per_sample) : b;
  }

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const int kb = kb0 + i;
        const int tap = kb / p.kpt;
        const int c0 = (kb - tap * p.kpt) * BLOCK_K;
        uint8_t* sa = smem + s * L::STAGE_BYTES;
        if (i >= npre) {  // ring slot reuse: wait for the MMAs that read it, then arm + load W as well
          mbar_wait(&empty_bar[s], ((i / STAGES) & 1) ^ 1);
          mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
          tma_load_2d(sa + L::A_BYTES, &tmW, &full_bar[s], kb * BLOCK_K, n0);
        }
        tma_load_5d(sa, &tmA, &full_bar[s], c0, x0 + p.dw[tap], y0 + p.dh[tap], t0 + p.dt[tap], b0);
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
        if (i == nkb - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else {
    // ========================================================================= epilogue
    const int sub = warp & 3;          // TMEM sub-partition this warp may read
    const int r = sub * 32 + lane;     // accumulator row (= TMEM lane) owned by this thread
    const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16);
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
    if (!split) {
      // NB: tcgen05.ld is warp-collective (.sync.aligned): every lane loads, stores are predicated
      if (!geglu) {
#pragma unroll 1
        for (int c = 0; c < BN; c += 16) {
          if (n0 + c >= p.N) break;
          uint32_t raw[16];
          tmem_ld_32x16(taddr + c, raw);
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
          if (row_ok) epilogue16(p, v, m, bs, n0 + c);
        }
      } else {
        // tile columns [0,BN/2) are the value half, [BN/2,BN) the gate half of the same BN/2
        // output features (weights were interleaved per tile when packed)
#pragma unroll 1
        for (int c = 0; c < BN / 2; c += 16) {
          uint32_t ra[16], rg[16];
          tmem_ld_32x16(taddr + c, ra);
          tmem_ld_32x16(taddr + BN / 2 + c, rg);
          tmem_ld_wait();
          float a[16], g[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { a[j] = __uint_as_float(ra[j]); g[j] = __uint_as_float(rg[j]); }
          if (row_ok) epilogue16_geglu(p, a, g, m, n0 + c, n0 + BN / 2 + c, blockIdx.x * (BN / 2) + c);
        }
      }
    } else {
      // split-K: park this CTA's partial tile in its own shared memory (the operand stages are dead
      // once tmem_full fired) as [4-column group][row] float4, conflict-free for row-per-thread access
      float4* stg = reinterpret_cast<float4*>(smem);
#pragma unroll 1
      for (int c = 0; c < BN; c += 16) {
        uint32_t raw[16];
        tmem_ld_32x16(taddr + c, raw);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          stg[((c >> 2) + j) * BLOCK_M + r] =
              make_float4(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]),
                          __uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3]));
      }
    }
  }

  if (split) {
    // ---- split-K reduction through distributed shared memory.  The `splits` CTAs of one output
    // tile form a thread-block cluster (cluster dims (1,1,splits)); after a cluster barrier each CTA
    // reduces every splits-th 16-column chunk by reading all peers' parked tiles with
    // ld.shared::cluster in rank order -- deterministic, no atomics, no HBM/L2 workspace -- and runs
    // the fused epilogue on it.
    cluster_sync_all();
    if (warp >= 2) {
      const int sub = warp & 3;
      const int r = sub * 32 + lane;
      const uint32_t stg_base = smem_u32(smem);
      const int S = p.splits, rank = blockIdx.z;
      auto load16 = [&](int c, float (&v)[16]) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
#pragma unroll 1
        for (int sidx = 0; sidx < S; ++sidx) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t local = stg_base + (uint32_t)((((c >> 2) + j) * BLOCK_M + r) * 16);
            const float4 f = ld_dsmem_f4(mapa_shared(local, (uint32_t)sidx));
            v[4 * j] += f.x; v[4 * j + 1] += f.y; v[4 * j + 2] += f.z; v[4 * j + 3] += f.w;
          }
        }
      };
      if (row_ok) {
        if (!geglu) {
#pragma unroll 1
          for (int q = rank; q < BN / 16; q += S) {
            const int c = q * 16;
            if (n0 + c >= p.N) break;
            float v[16];
            load16(c, v);
            epilogue16(p, v, m, bs, n0 + c);
          }
        } else {
#pragma unroll 1
          for (int q = rank; q < BN / 32; q += S) {
            const int c = q * 16;
            float a[16], g[16];
            load16(c, a);
            load16(BN / 2 + c, g);
            epilogue16_geglu(p, a, g, m, n0 + c, n0 + BN / 2 + c, blockIdx.x * (BN / 2) + c);
          }
        }
      }
    }
    cluster_sync_all();  // nobody's shared memory may go away while a peer is still reading it
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
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
  plan->K = g.ntaps * g.C;
  if (splits > 8) splits = 8;  // cluster size limit (portable)
  const bool geglu = (e.act == ACT_GEGLU);
  const int tw = (g.W + g.bw - 1) / g.bw, th = (g.H + g.bh - 1) / g.bh,
            tt = (g.T + g.bt - 1) / g.bt, tb = (g.B + g.bb - 1) / g.bb;
  plan->tiles_m = tw * th * tt * tb;
  const int kb_total = g.ntaps * (g.C / BLOCK_K);
  // ---- tile width BN and split-K factor (= cluster size, <= 8).  Tiny cost model: a CTA moves one
  // (A rows that exist + BN weight rows) x 128 B stage per k-block at ~55 GB/s (its share of L2
  // bandwidth), pays ~1 us for a DSMEM reduction, and the grid runs in ceil(ctas / #SMs) waves.
  {
    const int nsm = num_sms();
    const int rows_per_tile = std::min(BLOCK_M, plan->M);  // rows TMA really fetches (rest is zero fill)
    double best = 1e30;
    int best_bn = 128, best_s = 1;
    const int bn_lo = (geglu ? 128 : 64), bn_hi = (N <= 64 ? 64 : 128);
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
  DFB_CUDA_OK(cudaLaunchKernelEx(&cfg, igemm_tcgen05_kernel<BN, STAGES>, plan.tmA, plan.tmW, kp));
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
  for (int i = 0; i < 9; ++i) { kp.dt[i] = g.dt[i]; kp.dh[i] = g.dh[i]; kp.dw[i] = g.dw[i]; }
  kp.out_f32 = plan.e.out_f32; kp.out_f16 = plan.e.out_f16; kp.ldo = plan.e.ldo;
  kp.bias = plan.e.bias; kp.rowvec = plan.e.rowvec; kp.ld_rowvec = plan.e.ld_rowvec;
  kp.rows_per_sample = plan.e.rows_per_sample;
  kp.residual = plan.e.residual; kp.residual_f16 = plan.e.residual_f16; kp.ld_res = plan.e.ld_res;
  kp.act = plan.e.act;
  kp.splits = plan.splits;
  kp.next_w = reinterpret_cast<const uint8_t*>(plan.next_w);
  kp.next_w_bytes = plan.next_w_bytes;
  {
    const double out_b = (plan.e.out_f32 ? 4.0 : 0.0) + (plan.e.out_f16 ? 2.0 : 0.0);
    note(g.ntaps == 9 ? "igemm_conv3x3" : "igemm_linear", 2.0 * plan.M * plan.N * plan.K,
         2.0 * plan.N * plan.K + 2.0 * plan.M * g.C + out_b * plan.M * plan.e.ldo +
             (plan.e.residual ? 4.0 * plan.M * plan.e.ldo : 0.0),
         plan.M, plan.N, plan.K, plan.splits, plan.tiles_m * plan.tiles_n * plan.splits);
  }
  // Shallow operand ring (~97 KB) by default: two CTAs fit on an SM, so (a) one CTA's epilogue overlaps
  // the other's main loop and (b) under programmatic dependent launch the NEXT kernel's CTAs become
  // resident -- and stream their first weight tiles -- while this kernel is still running.  Measured
  // 3.37 ms/step vs 3.46 with the deep ring everywhere (DFB_SHALLOW=0 selects the deep ring).
  static const int force_shallow = getenv("DFB_SHALLOW") ? atoi(getenv("DFB_SHALLOW")) : -1;
  const bool shallow = (force_shallow != 0);
  if (plan.BN == 64) return shallow ? launch_t<64, 4>(plan, kp, stream) : launch_t<64, 8>(plan, kp, stream);
  return shallow ? launch_t<128, 3>(plan, kp, stream) : launch_t<128, 6>(plan, kp, stream);
}

}  // namespace dfb
