// attn_tcgen05.cu — fused attention for head_dim 64 on tcgen05 tensor cores (sm_100a).
//
// Kernels in this file:
//   flash_attn_v5_kernel     self-attention (AttnProcessor2_0, src/models/attention_processor.py:69-81) and the
//                            Resampler's perceiver attention (src/models/resampler.py:64-74): softmax(Q K^T * scale) V,
//                            no mask.  Two independent online-softmax streams per CTA, S / O / P all in TMEM (section 1).
//                            Q/K/V are addressed by 3-D tensor maps {columns, tokens, batch} over the fused projection
//                            output, so the head split / transposes of the reference are never materialised.
//   cross_ip_attn_v2_kernel  out = softmax(Q Kt^T/8) Vt + scale * softmax(Q Kip^T/8 + M(bbox)) Vip
//                            (MaskedIPAttnProcessor2_0, :231-258) in ONE pass: persistent, two softmax warp groups,
//                            P in TMEM; the additive bbox mask M in {0,-10000} (:115-169) is evaluated in registers
//                            with the reference's closed-interval / derived-(H',W') semantics (ip_mask.cuh) (section 2).
// The round-1 kernels these replaced (P through shared memory; one tile per CTA) are in git history (commit 8e22d8c).
#include <cstdlib>

#include "ds_common.cuh"
#include "ds_host.h"
#include "ip_mask.cuh"

namespace ds {

constexpr int kTile = 128;            // query rows per CTA == kv rows per tile
constexpr int kHd = 64;               // head dim
constexpr int kTileBytes = kTile * kHd * 2;  // 16 KiB
constexpr float kLog2e = 1.4426950408889634f;

// swizzled (SWIZZLE_128B) byte offset of 16-byte chunk `q16` (0..7) of row `r` inside a [rows][64 bf16] atom
__device__ __forceinline__ uint32_t sw128_off(int r, int q16) { return r * 128 + ((q16 ^ (r & 7)) << 4); }

// 2^x on the MUFU pipe (ex2.approx.ftz): inputs here are <= ~8, results feed a bf16 rounding
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void st_shared_16(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

struct FlashParams {
  __nv_bfloat16* out;  // [B][Nq][ldo]
  int Nq, Nkv, ldo;
  int q_col0, k_col0, v_col0;  // column of head 0 inside the respective tensor map
  float scale_log2;            // softmax scale * log2(e)
};

// =================================================================================================
// Shared by the TMEM-P flash kernels.  History of the self-attention kernel at B8 N4096 h10 (all measured on B200,
// sources in git history, analysis in profiles/r01_ncu_flash_v3.md):
//   v2 (git history: P through shared memory, 2 threads per row)                              634 us
//   v3 (P in TMEM via tcgen05.st + ts-form PV MMA, 1 thread per row, sum-bounded lazy rescale) 562 us
//   v4 (v3 + the kv tile pipelined through the MMA warp in two 64-key halves)                  559 us
//   v5 (two independent online-softmax streams per CTA)                                        436 us
// =================================================================================================
constexpr int kFlash3Ring = 4;
constexpr float kSumOverflow = 1024.0f;  // 2^10

// Degree-3 minimax 2^f on f in [-0.5, 0.5] after Cody-Waite range reduction, all on the FMA/ALU pipes (FA4-style
// MUFU offload).  Max rel. error 7.5e-5 — far below the bf16 rounding (2^-9) applied to P right after.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float r = x + 12582912.0f;            // 1.5 * 2^23: integer part lands in the low mantissa bits
  const float fi = r - 12582912.0f;
  const float f = x - fi;                     // [-0.5, 0.5]
  float p = fmaf(f, 0.05517132f, 0.24261054f);  // minimax (Lawson) fit, max rel. err 7.5e-5
  p = fmaf(p, f, 0.69326099f);
  p = fmaf(p, f, 0.99992811f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(r) << 23));
}

// =================================================================================================
// (1) flash attention v5 — two INDEPENDENT online-softmax streams per CTA.
//      ncu on v3/v4 (profiles/r01_ncu_flash_v3.md): with one softmax warp per SM sub-partition per CTA the exp
//      loop is a single dependent instruction stream (issue: selected 33 % / fixed-latency wait 35 %), i.e. bound
//      by per-warp issue latency, not by the MUFU or the tensor pipe — which is also why moving exponentials to
//      the FMA pipe made it slower.  v5 doubles the warps that are in the exp loop at any time:
//        stream h (h = 0, 1) owns keys [64h, 64h+64) of EVERY 128-key tile, with its own warps (4 per stream, one
//        thread per query row), its own reference max m_h, row sum l_h and TMEM accumulator O_h.  The streams
//        never synchronise until the epilogue, where   out = (w_0 O_0 + w_1 O_1) / (w_0 l_0 + w_1 l_1),
//        w_h = 2^(m_h - max(m_0, m_1))   (the split-KV identity).
//      P_h is written over the head of S_h's own columns (thread-local rows; in-order MMA execution makes
//      S_h(j+1) land after PV_h(j) has consumed it), which frees the TMEM columns for the second accumulator:
//        TMEM (256 columns, 2 CTAs/SM): S_0|P_0 [0,64) | S_1|P_1 [64,128) | O_0 [128,192) | O_1 [192,256).
//      Both chunks of P are built in registers before the overflow check, so the (rare) redo still finds S intact.
// =================================================================================================
// packed fp32x2 arithmetic (FFMA2 / FADD2 on sm_100): one instruction for two lanes-worth of scale-and-subtract /
// row-sum work in the exp loops
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

constexpr int kFlash5Threads = 320;  // warp 0: TMA + TMEM alloc, warp 1: MMA issue, warps 2-9: two softmax streams
constexpr int kFlash5SmemBytes = kTileBytes * (1 + kFlash3Ring) + 1024 + 256 + 2 * kTile * 2 * 4;

// DUAL: one MMA-issue thread PER STREAM (warp 1 lane 0 -> stream 0, warp 10 lane 0 -> stream 1; 352 threads).  The
// -DDS_ATTN_TRACE build shows why: the single issuer spends ~1200 clk per 128-key tile inside its 16 tcgen05.mma
// (M128 N64 K16: 32 clk of tensor time each, ~75 clk each to dispatch from one thread — the converged `elect.sync`
// variant is no faster, so it is not ptxas's waterfall), and a stream whose P is ready queues behind the other stream's
// 8 dispatches.  With one issuer per stream the two dispatch sequences run side by side; K/V ring slots are released by
// BOTH issuers' commits (`empty` barriers count 2).
template <int POLY, bool F2, bool ELECT, bool DUAL = false>
__global__ void __launch_bounds__(DUAL ? kFlash5Threads + 32 : kFlash5Threads, 2)
flash_attn_v5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const FlashParams p) {
  constexpr int RING = kFlash3Ring;
  constexpr int kHalf = 64;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sRing = sQ + kTileBytes;  // RING x 16 KiB: K_0 V_0 K_1 V_1 ...
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRing + RING * kTileBytes);
  uint64_t* q_full = bars;
  uint64_t* full = bars + 1;        // [RING]
  uint64_t* empty = full + RING;    // [RING]
  uint64_t* s_full = empty + RING;  // [2]  S_h readable (and PV_h of the previous tile retired)
  uint64_t* p_full = s_full + 2;    // [2]  P_h written
  uint64_t* o_full = p_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
  float2* s_ml = reinterpret_cast<float2*>(bars + 32);  // [2][128] (m_h, l_h) exchange for the epilogue

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kTile;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int num_kv_tiles = (p.Nkv + kTile - 1) / kTile;

  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < RING; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], DUAL ? 2 : 1);
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(&s_full[h], 1);
      mbar_init(&p_full[h], 4);  // one arrival per softmax warp of the stream
    }
    mbar_init(o_full, DUAL ? 2 : 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // the prologue above overlapped the producer GEMM's tail
  const uint32_t tS = tmem_base;        // S_h (and P_h over its first 32 columns) at +64h
  const uint32_t tO = tmem_base + 128;  // O_h at +64h

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kTileBytes);
      tma_load_3d(sQ, &tmQ, q_full, p.q_col0 + head * kHd, q0, batch);
      for (int idx = 0; idx < 2 * num_kv_tiles; ++idx) {  // K_0 V_0 K_1 V_1 ...
        const int slot = idx % RING;
        const uint32_t ph = (idx / RING) & 1;
        const int j = idx >> 1, which = idx & 1;
        mbar_wait(&empty[slot], ph ^ 1);
        mbar_arrive_expect_tx(&full[slot], kTileBytes);
        tma_load_3d(sRing + slot * kTileBytes, which == 0 ? &tmK : &tmV, &full[slot],
                    (which == 0 ? p.k_col0 : p.v_col0) + head * kHd, j * kTile, batch);
      }
    }
  } else if (DUAL && (warp == 1 || warp == 10)) {
    if (lane == 0) {
      // one issuer per stream: S_h(0); then per tile PV_h(j), S_h(j+1); every ring slot it has finished with gets one
      // of the two commits its `empty` barrier waits for
      const int h = warp == 1 ? 0 : 1;
      constexpr uint32_t idesc_qk = make_idesc_bf16(kTile, kHalf, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(kTile, kHd, 0, 1);
      const uint32_t q_addr = smem_u32(sQ);
      auto issue_s = [&](uint32_t k_addr) {
#pragma unroll
        for (int k = 0; k < kHd / 16; ++k)
          umma_ss(tS + h * kHalf, make_sw128_desc(q_addr + k * 32, 1024, 16),
                  make_sw128_desc(k_addr + h * (kHalf * 128) + k * 32, 1024, 16), idesc_qk, k != 0 ? 1u : 0u);
        umma_commit(&s_full[h]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&full[0], 0);
      tc_fence_after();
      issue_s(smem_u32(sRing));
      umma_commit(&empty[0]);
      for (int j = 0; j < num_kv_tiles; ++j) {
        const int vi = 2 * j + 1, vslot = vi % RING;
        const int ki = 2 * j + 2, kslot = ki % RING;
        const bool more = j + 1 < num_kv_tiles;
        mbar_wait(&full[vslot], (vi / RING) & 1);
        if (more) mbar_wait(&full[kslot], (ki / RING) & 1);
        const uint32_t v_addr = smem_u32(sRing + vslot * kTileBytes);
        mbar_wait(&p_full[h], j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kHalf / 16; ++k)
          umma_ts(tO + h * kHd, tS + h * kHalf + k * 8, make_sw128_desc(v_addr + h * (kHalf * 128) + k * 2048, 1024, 1024),
                  idesc_pv, (j == 0 && k == 0) ? 0u : 1u);
        umma_commit(&empty[vslot]);
        if (more) {
          issue_s(smem_u32(sRing + kslot * kTileBytes));
          umma_commit(&empty[kslot]);
        }
      }
      umma_commit(o_full);
    }
  } else if (warp == 1) {
    if (ELECT) {
      // MMA issue warp, CONVERGED: lane 0 alone polls the mbarriers (32 polling lanes made this variant slower:
      // 433 -> 474 us), the warp re-converges with __syncwarp and one elected lane issues every tcgen05 instruction
      // with uniform operands — no ELECT / BRA.U.ANY waterfall around each MMA (~50 issue cycles each)
      auto wait1 = [&](uint64_t* bar, uint32_t ph) {
        if (lane == 0) mbar_wait(bar, ph);
        __syncwarp();
      };
      constexpr uint32_t idesc_qk = make_idesc_bf16(kTile, kHalf, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(kTile, kHd, 0, 1);
      const uint32_t q_addr = smem_u32(sQ);
      auto issue_s = [&](uint32_t k_addr, int h) {
#pragma unroll
        for (int k = 0; k < kHd / 16; ++k)
          umma_ss_e(tS + h * kHalf, make_sw128_desc(q_addr + k * 32, 1024, 16),
                    make_sw128_desc(k_addr + h * (kHalf * 128) + k * 32, 1024, 16), idesc_qk, k != 0 ? 1u : 0u);
        umma_commit_e(&s_full[h]);
      };
      auto issue_pv = [&](uint32_t v_addr, int h, bool first) {
#pragma unroll
        for (int k = 0; k < kHalf / 16; ++k)
          umma_ts_e(tO + h * kHd, tS + h * kHalf + k * 8,
                    make_sw128_desc(v_addr + h * (kHalf * 128) + k * 2048, 1024, 1024), idesc_pv,
                    (first && k == 0) ? 0u : 1u);
      };
#ifdef DS_ATTN_TRACE
      long long fe_poll = 0, fe_issue = 0;
#endif
      wait1(q_full, 0);
      wait1(&full[0], 0);
      tc_fence_after();
      issue_s(smem_u32(sRing), 0);
      issue_s(smem_u32(sRing), 1);
      umma_commit_e(&empty[0]);
      for (int j = 0; j < num_kv_tiles; ++j) {
        const int vi = 2 * j + 1, vslot = vi % RING;
        const int ki = 2 * j + 2, kslot = ki % RING;
        const bool more = j + 1 < num_kv_tiles;
        wait1(&full[vslot], (vi / RING) & 1);
        if (more) wait1(&full[kslot], (ki / RING) & 1);
        const uint32_t v_addr = smem_u32(sRing + vslot * kTileBytes);
        const uint32_t k_addr = smem_u32(sRing + kslot * kTileBytes);
        int first = 0;
#ifdef DS_ATTN_TRACE
        long long fe_a = clock64();
#endif
        if (lane == 0) {
          for (;;) {
            if (mbar_try_wait(&p_full[0], j & 1)) break;
            if (mbar_try_wait(&p_full[1], j & 1)) {
              first = 1;
              break;
            }
          }
        }
        first = __shfl_sync(0xffffffffu, first, 0);
#ifdef DS_ATTN_TRACE
        fe_poll += clock64() - fe_a;
#endif
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const int h = o == 0 ? first : 1 - first;
#ifdef DS_ATTN_TRACE
          fe_a = clock64();
#endif
          if (o == 1) wait1(&p_full[h], j & 1);
#ifdef DS_ATTN_TRACE
          fe_poll += clock64() - fe_a;
          fe_a = clock64();
#endif
          tc_fence_after();
          issue_pv(v_addr, h, j == 0);
          if (o == 1) umma_commit_e(&empty[vslot]);
          if (more) issue_s(k_addr, h);
          if (more && o == 1) umma_commit_e(&empty[kslot]);
#ifdef DS_ATTN_TRACE
          fe_issue += clock64() - fe_a;
#endif
        }
      }
      umma_commit_e(o_full);
#ifdef DS_ATTN_TRACE
      if ((blockIdx.x % 601) == 0 && lane == 0)
        printf("[trace] flash(elect) blk %d MMA warp: per kv tile: poll p_full %.0f clk, issue %.0f clk\n", blockIdx.x,
               double(fe_poll) / num_kv_tiles, double(fe_issue) / num_kv_tiles);
#endif
    } else if (lane == 0) {
      // MMA issue: ONE lane in a divergent region (ptxas wraps every tcgen05.mma in an ELECT / BRA.U.ANY waterfall)
      constexpr uint32_t idesc_qk = make_idesc_bf16(kTile, kHalf, 0, 0);  // M128 N64, both K-major
      constexpr uint32_t idesc_pv = make_idesc_bf16(kTile, kHd, 0, 1);    // M128 N64, A from TMEM, B (=V) MN-major
      const uint32_t q_addr = smem_u32(sQ);
      auto issue_s = [&](uint32_t k_addr, int h) {  // S_h = Q K[64h..64h+64)^T
#pragma unroll
        for (int k = 0; k < kHd / 16; ++k)
          umma_ss(tS + h * kHalf, make_sw128_desc(q_addr + k * 32, 1024, 16),
                  make_sw128_desc(k_addr + h * (kHalf * 128) + k * 32, 1024, 16), idesc_qk, k != 0 ? 1u : 0u);
        umma_commit(&s_full[h]);
      };
      auto issue_pv = [&](uint32_t v_addr, int h, bool first) {  // O_h (+)= P_h V[64h..64h+64)
#pragma unroll
        for (int k = 0; k < kHalf / 16; ++k)
          umma_ts(tO + h * kHd, tS + h * kHalf + k * 8, make_sw128_desc(v_addr + h * (kHalf * 128) + k * 2048, 1024, 1024),
                  idesc_pv, (first && k == 0) ? 0u : 1u);
      };
#ifdef DS_ATTN_TRACE
      long long fm_poll = 0, fm_issue = 0;
#endif
      mbar_wait(q_full, 0);
      mbar_wait(&full[0], 0);  // K_0
      tc_fence_after();
      issue_s(smem_u32(sRing), 0);
      issue_s(smem_u32(sRing), 1);
      umma_commit(&empty[0]);
      for (int j = 0; j < num_kv_tiles; ++j) {
        const int vi = 2 * j + 1, vslot = vi % RING;
        const int ki = 2 * j + 2, kslot = ki % RING;
        const bool more = j + 1 < num_kv_tiles;
        mbar_wait(&full[vslot], (vi / RING) & 1);            // V_j
        if (more) mbar_wait(&full[kslot], (ki / RING) & 1);  // K_{j+1}
        const uint32_t v_addr = smem_u32(sRing + vslot * kTileBytes);
        const uint32_t k_addr = smem_u32(sRing + kslot * kTileBytes);
        // serve whichever stream has its P ready first (no head-of-line blocking behind the slower stream); both
        // streams of tile j use the same V_j / K_{j+1} slots, which are released after the second one is served
        int first = 0;
#ifdef DS_ATTN_TRACE
        long long fm_a = clock64();
#endif
        for (;;) {
          if (mbar_try_wait(&p_full[0], j & 1)) break;
          if (mbar_try_wait(&p_full[1], j & 1)) {
            first = 1;
            break;
          }
        }
#ifdef DS_ATTN_TRACE
        fm_poll += clock64() - fm_a;
#endif
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const int h = o == 0 ? first : 1 - first;
#ifdef DS_ATTN_TRACE
          fm_a = clock64();
#endif
          if (o == 1) mbar_wait(&p_full[h], j & 1);
#ifdef DS_ATTN_TRACE
          fm_poll += clock64() - fm_a;
          fm_a = clock64();
#endif
          tc_fence_after();
          issue_pv(v_addr, h, j == 0);
          if (o == 1) umma_commit(&empty[vslot]);
          if (more) issue_s(k_addr, h);  // executes after PV_h(j): P_h(j) is consumed before S_h(j+1) overwrites it
          if (more && o == 1) umma_commit(&empty[kslot]);
#ifdef DS_ATTN_TRACE
          fm_issue += clock64() - fm_a;
#endif
        }
      }
      umma_commit(o_full);
#ifdef DS_ATTN_TRACE
      if ((blockIdx.x % 601) == 0)
        printf("[trace] flash blk %d MMA thread: per kv tile: poll p_full %.0f clk, issue (2 x (4 PV + 4 S MMAs + commits)) %.0f clk\n",
               blockIdx.x, double(fm_poll) / num_kv_tiles, double(fm_issue) / num_kv_tiles);
#endif
    }
  } else {
    // ---------------------------------------------------------------- softmax: 2 streams x 4 warps, thread <-> (row, h)
    const int h = (warp - 2) >> 2;
    const int wq = warp & 3;  // TMEM lane quadrant this warp may touch
    const int row = wq * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t tSh = tS + lane_base + h * kHalf;  // this stream's scores; P_h goes over [tSh, tSh + 32)
    float m_ref = -INFINITY, l = 0.f;

    // 32 scores -> 16 packed bf16x2 probabilities; partial sums into s0/s1
    auto chunk = [&](const uint32_t(&raw)[32], uint32_t(&pk)[16], float neg_m, int valid, float& s0, float& s1) {
      if (valid >= 32) {
        if (F2) {
          const uint64_t sc2 = f2_pack(p.scale_log2, p.scale_log2), nm2 = f2_pack(neg_m, neg_m);
          uint64_t acc = f2_pack(s0, s1);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float t0, t1;
            f2_unpack(f2_fma(f2_pack(__uint_as_float(raw[2 * i]), __uint_as_float(raw[2 * i + 1])), sc2, nm2), t0, t1);
            const float e0 = ((2 * i) % 8 < POLY) ? ex2_poly(t0) : ex2(t0);
            const float e1 = ((2 * i + 1) % 8 < POLY) ? ex2_poly(t1) : ex2(t1);
            acc = f2_add(acc, f2_pack(e0, e1));
            pk[i] = pack_bf16_alu(e0, e1);
          }
          f2_unpack(acc, s0, s1);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float t0 = fmaf(__uint_as_float(raw[2 * i]), p.scale_log2, neg_m);
            const float t1 = fmaf(__uint_as_float(raw[2 * i + 1]), p.scale_log2, neg_m);
            const float e0 = ((2 * i) % 8 < POLY) ? ex2_poly(t0) : ex2(t0);
            const float e1 = ((2 * i + 1) % 8 < POLY) ? ex2_poly(t1) : ex2(t1);
            s0 += e0;
            s1 += e1;
            pk[i] = pack_bf16_alu(e0, e1);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float t0 = fmaf(__uint_as_float(raw[2 * i]), p.scale_log2, neg_m);
          const float t1 = fmaf(__uint_as_float(raw[2 * i + 1]), p.scale_log2, neg_m);
          const float e0 = (2 * i < valid) ? ex2(t0) : 0.f;
          const float e1 = (2 * i + 1 < valid) ? ex2(t1) : 0.f;
          s0 += e0;
          s1 += e1;
          pk[i] = pack_bf16_alu(e0, e1);
        }
      }
    };

#ifdef DS_ATTN_TRACE
    long long fr_wait = 0, fr_t0 = clock64(), fr_a;
#endif
    for (int j = 0; j < num_kv_tiles; ++j) {
      const int valid = p.Nkv - j * kTile - h * kHalf;  // >= 64: whole half valid; <= 0: nothing valid
#ifdef DS_ATTN_TRACE
      fr_a = clock64();
#endif
      mbar_wait(&s_full[h], j & 1);
#ifdef DS_ATTN_TRACE
      fr_wait += clock64() - fr_a;
#endif
      tc_fence_after();
      bool need_max = (j == 0);
      float lsum;
#pragma unroll 1
      for (;;) {
        uint32_t ra[32];
        if (need_max) {  // exact maximum of this stream's valid scores -> move the reference, rescale O_h and l_h
          float mx = -INFINITY;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            tmem_ld32(tSh + c * 32, ra);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(ra[i]));
          }
          const float t_new = mx * p.scale_log2;
          float alpha = 1.0f;  // rows whose reference does not move are scaled by 1
          if (t_new > m_ref) {
            alpha = ex2(m_ref - t_new);  // first tile: m_ref = -inf -> alpha = 0 (l = 0, O_h unset)
            m_ref = t_new;
            l *= alpha;
          }
          if (j > 0) {  // warp-uniform (tcgen05.ld/st are .sync.aligned); s_full[h](j) was committed after
                        // PV_h(j-1), so O_h is quiescent
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              tmem_ld32(tO + lane_base + h * kHd + c * 32, ra);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) ra[i] = __float_as_uint(__uint_as_float(ra[i]) * alpha);
              tmem_st32(tO + lane_base + h * kHd + c * 32, ra);
            }
          }
        }
        float s0 = 0.f, s1 = 0.f;
        uint32_t pk0[16], pk1[16];
        tmem_ld32(tSh, ra);
        tmem_ld_wait();
        chunk(ra, pk0, -m_ref, valid, s0, s1);
        tmem_ld32(tSh + 32, ra);
        tmem_ld_wait();
        chunk(ra, pk1, -m_ref, valid - 32, s0, s1);
        lsum = s0 + s1;
        if (!need_max) need_max = __any_sync(0xffffffffu, !(lsum < kSumOverflow));  // warp-uniform, rare
        else need_max = false;  // exact reference: every term <= 1
        if (!need_max) {  // S_h is only overwritten once the row sum is known to be safe
          tmem_st16(tSh, pk0);
          tmem_st16(tSh + 16, pk1);
          break;
        }
      }
      l += lsum;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[h]);
    }
#ifdef DS_ATTN_TRACE
    if ((blockIdx.x % 601) == 0 && lane == 0 && wq == 0)
      printf("[trace] flash blk %d stream %d: %d kv tiles, loop %lld clk, wait s_full %lld (%.0f / tile), softmax %.0f / tile\n",
             blockIdx.x, h, num_kv_tiles, clock64() - fr_t0, fr_wait, double(fr_wait) / num_kv_tiles,
             double(clock64() - fr_t0 - fr_wait) / num_kv_tiles);
#endif
    // ---- epilogue: merge the two streams, normalise, store (thread (row, h) writes output columns [32h, 32h+32))
    s_ml[h * kTile + row] = make_float2(m_ref, l);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float2 mine = make_float2(m_ref, l), other = s_ml[(h ^ 1) * kTile + row];
    const float m = fmaxf(mine.x, other.x);
    const float w_me = ex2(mine.x - m), w_ot = ex2(other.x - m);  // -inf - m = -inf -> 0 for an empty stream
    const float inv = 1.0f / (mine.y * w_me + other.y * w_ot);
    const float w0 = (h == 0 ? w_me : w_ot) * inv, w1 = (h == 0 ? w_ot : w_me) * inv;
    mbar_wait(o_full, 0);
    tc_fence_after();
    const int q_row = q0 + row;
    __nv_bfloat16* orow = p.out + (static_cast<size_t>(batch) * p.Nq + q_row) * p.ldo + head * kHd + h * 32;
    uint32_t a[32], b[32];
    tmem_ld32(tO + lane_base + h * 32, a);           // O_0[:, 32h .. 32h+32)
    tmem_ld32(tO + lane_base + kHd + h * 32, b);     // O_1[:, 32h .. 32h+32)
    tmem_ld_wait();
    if (q_row < p.Nq) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaf(__uint_as_float(a[q * 8 + e]), w0, __uint_as_float(b[q * 8 + e]) * w1);
        uint4 u;
        u.x = pack_bf16(o[0], o[1]);
        u.y = pack_bf16(o[2], o[3]);
        u.z = pack_bf16(o[4], o[5]);
        u.w = pack_bf16(o[6], o[7]);
        reinterpret_cast<uint4*>(orow)[q] = u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// =================================================================================================
// (2) fused text + masked-IP cross-attention
// =================================================================================================
struct CrossParams {
  __nv_bfloat16* out;   // [B][N][C]
  const float* bbox;    // [B][num_ips][4]
  int N, C;
  int n_text, n_ip, nt_pad, nip_pad;  // pads are multiples of 16; nt_pad + nip_pad <= 192
  int num_ips, tokens_per_ip, num_dummy;
  int Hd, Wd;           // derived (H', W')
  float ip_scale;
};

// =================================================================================================
// (2) cross_ip_attn_v2_kernel — restructured in round 1 because the one-tile-per-CTA
//      version was latency-bound (one serial load -> S -> softmax -> PV -> store chain per CTA lifetime; 4-7x off its
//      HBM floor, profiles/r01_ncu_summary.md):
//        * persistent: 2 CTAs per SM, CTA c owns a contiguous range of (batch, head, q-tile) items, so K|V of a
//          (batch, head) are loaded once per ~8 tiles and the next tile's Q is prefetched through a 2-slot ring
//          while the current tile is in its softmax;
//        * P goes registers -> TMEM (bf16x2, aliasing the dead head of the S columns: the thread that read row r's
//          scores is the only writer of row r's probabilities) and feeds the PV MMAs as a TMEM A operand — no
//          st.shared / fence.proxy.async / 48 KB P staging buffer;
//        * the text softmax and the masked IP softmax of a row are independent: each gets its own 4 warps (one
//          thread per row and softmax), so 8 exp-loop warps per CTA hide each other's issue latency; TMEM loads are
//          software-pipelined one 16-column chunk ahead; one compact predicated code path per pass (the round-1 code
//          was 75 KB of SASS per kernel and instruction-fetch bound).
// =================================================================================================
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

constexpr int kCross2Threads = 320;  // warp 0: TMA + TMEM alloc, warp 1: MMA issue, warps 2-5: text keys, 6-9: IP keys

template <bool UNIFORM>  // UNIFORM: tokens_per_ip and num_dummy are multiples of 16 -> one mask bit per 16-key chunk
__global__ void __launch_bounds__(kCross2Threads, 2)
cross_ip_attn_v2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKVt,
                        const __grid_constant__ CUtensorMap tmKVip, const CrossParams p, int heads, int q_tiles,
                        int total_items) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int n_keys = p.nt_pad + p.nip_pad;  // multiple of 16, <= 192
  const int kv_bytes = n_keys * 128;        // [n_keys][64] bf16
  const int kv_stride = (kv_bytes + 1023) & ~1023;
  uint8_t* sQ = smem;                       // 2 x 16 KiB ring
  uint8_t* sK = sQ + 2 * kTileBytes;
  uint8_t* sV = sK + kv_stride;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kv_stride);
  uint64_t* q_full = bars;         // [2]
  uint64_t* q_empty = bars + 2;    // [2]
  uint64_t* kv_full = bars + 4;
  uint64_t* kv_empty = bars + 5;
  uint64_t* s_full = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* o_full = bars + 8;
  uint64_t* o_empty = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  float* s_l = reinterpret_cast<float*>(bars + 16);  // [2][128]: row sums of the two softmaxes (epilogue exchange)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = static_cast<int>(static_cast<long long>(blockIdx.x) * total_items / gridDim.x);
  const int i1 = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * total_items / gridDim.x);
  const int n_items = i1 - i0;

  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKVt);
    tma_prefetch_desc(&tmKVip);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    mbar_init(kv_full, 1);
    mbar_init(kv_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 8);   // one arrival per softmax warp (4 text + 4 IP)
    mbar_init(o_full, 1);
    mbar_init(o_empty, 8);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  // TMEM (256 columns): S text [0, nt_pad) | S ip [nt_pad, n_keys);  P_text over [0, nt_pad/2), P_ip over
  // [nt_pad, nt_pad + nip_pad/2) (each stream overwrites the head of its OWN score columns, thread-local rows);
  // O_text [128,192) (dead IP score columns + free ones; needs nt_pad + nip_pad/2 <= 128), O_ip [192,256)
  const uint32_t tS = tmem_base;
  const uint32_t tOt = tmem_base + 128;
  const uint32_t tOi = tmem_base + 192;

  if (warp == 0) {
    if (lane == 0) {
      int cur_bh = -1, kv_loads = 0;
      for (int n = 0; n < n_items; ++n) {
        const int item = i0 + n;
        const int qt = item % q_tiles, bh = item / q_tiles;
        const int head = bh % heads, batch = bh / heads;
        if (bh != cur_bh) {
          if (kv_loads > 0) mbar_wait(kv_empty, (kv_loads - 1) & 1);  // every MMA on the old K|V has completed
          mbar_arrive_expect_tx(kv_full, 2 * kv_bytes);
          tma_load_3d(sK, &tmKVt, kv_full, head * kHd, 0, batch);
          tma_load_3d(sK + p.nt_pad * 128, &tmKVip, kv_full, head * kHd, 0, batch);
          tma_load_3d(sV, &tmKVt, kv_full, p.C + head * kHd, 0, batch);
          tma_load_3d(sV + p.nt_pad * 128, &tmKVip, kv_full, p.C + head * kHd, 0, batch);
          ++kv_loads;
          cur_bh = bh;
        }
        const int slot = n & 1;
        mbar_wait(&q_empty[slot], ((n >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[slot], kTileBytes);
        tma_load_3d(sQ + slot * kTileBytes, &tmQ, &q_full[slot], head * kHd, qt * kTile, batch);
      }
    }
  } else if (warp == 1) {
    {  // converged MMA-issue warp (elected lane issues; see umma_ss_e)
      const uint32_t idesc_qk = make_idesc_bf16(kTile, n_keys, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(kTile, kHd, 0, 1);
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      const int kt = p.nt_pad / 16;
      int cur_bh = -1, kv_uses = 0;
      for (int n = 0; n < n_items; ++n) {
        const int item = i0 + n;
        const int bh = item / q_tiles;
        if (bh != cur_bh) {
          mbar_wait(kv_full, kv_uses & 1);
          ++kv_uses;
          cur_bh = bh;
        }
        const int slot = n & 1;
        mbar_wait(&q_full[slot], (n >> 1) & 1);
        if (n > 0) mbar_wait(o_empty, (n - 1) & 1);  // the previous tile's outputs (aliasing S) have been read
        tc_fence_after();
        const uint32_t q_addr = smem_u32(sQ + slot * kTileBytes);
#pragma unroll
        for (int k = 0; k < kHd / 16; ++k)
          umma_ss_e(tS, make_sw128_desc(q_addr + k * 32, 1024, 16), make_sw128_desc(k_addr + k * 32, 1024, 16), idesc_qk,
                  k != 0 ? 1u : 0u);
        umma_commit_e(&q_empty[slot]);
        umma_commit_e(s_full);
        mbar_wait(p_full, n & 1);
        tc_fence_after();
        for (int k = 0; k < n_keys / 16; ++k) {
          const uint64_t bdesc = make_sw128_desc(v_addr + k * 2048, 1024, 1024);
          if (k < kt)
            umma_ts_e(tOt, tS + k * 8, bdesc, idesc_pv, k != 0 ? 1u : 0u);
          else
            umma_ts_e(tOi, tS + p.nt_pad + (k - kt) * 8, bdesc, idesc_pv, k != kt ? 1u : 0u);
        }
        umma_commit_e(o_full);
        if (n + 1 < n_items && (item + 1) / q_tiles != bh) umma_commit_e(kv_empty);
      }
    }
  } else {
    // ------------------------------------------------------------ softmax: group 0 = text keys, group 1 = IP keys
    const int g = (warp - 2) >> 2;
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(wq * 32) << 16;
    constexpr float kS2 = 0.125f * kLog2e;
    constexpr float kMask2 = -10000.0f * kLog2e;
    const int chunks = (g == 0 ? p.nt_pad : p.nip_pad) / 16;        // 16-key chunks of this group's softmax
    const int n_real = g == 0 ? p.n_text : p.n_ip;                  // keys that are not padding
    const uint32_t tSg = tS + lane_base + (g == 0 ? 0 : p.nt_pad);  // this group's scores; its P goes over their head
#ifdef DS_ATTN_TRACE
    long long tr_s = 0, tr_p1 = 0, tr_p2 = 0, tr_o = 0, tr_ep = 0, tr_t0 = clock64(), tr_a;
#endif

    for (int n = 0; n < n_items; ++n) {
      const int item = i0 + n;
      const int qt = item % q_tiles, bh = item / q_tiles;
      const int head = bh % heads, batch = bh / heads;
      const int q_row = qt * kTile + row;
      // bit k of `open16` (UNIFORM): IP chunk k (16 keys of one character / the dummies) is visible from this row
      uint32_t bits = 0, open16 = 0xffffffffu;
      if (g == 1) {
        bits = ip_inside_bits(p.bbox + static_cast<size_t>(batch) * p.num_ips * 4, p.num_ips, min(q_row, p.N - 1), p.Hd,
                              p.Wd);
        if (UNIFORM) {
          open16 = 0;
          for (int k = 0; k < chunks; ++k)
            open16 |= static_cast<uint32_t>(ip_key_open(bits, k * 16, p.tokens_per_ip, p.num_dummy)) << k;
        }
      }
      // additive log2-domain term of key i of chunk c (reference :142,162-163: M in {0, -10000}; text keys: 0)
      auto add_of = [&](int c, int i) -> float {
        if (g == 0) return 0.0f;
        const bool open = UNIFORM ? ((open16 >> c) & 1u) != 0
                                  : ip_key_open(bits, c * 16 + i, p.tokens_per_ip, p.num_dummy);
        return open ? 0.0f : kMask2;
      };

#ifdef DS_ATTN_TRACE
      tr_a = clock64();
#endif
      mbar_wait(s_full, n & 1);
#ifdef DS_ATTN_TRACE
      tr_s += clock64() - tr_a;
      tr_a = clock64();
#endif
      tc_fence_after();
      // ---- pass 1: row maximum (log2 domain)
      float m = -INFINITY;
      auto max_chunk = [&](int c, const uint32_t(&raw)[16]) {
        const int nv = n_real - c * 16;
        float mx = -INFINITY;
        if (UNIFORM) {
#pragma unroll
          for (int i = 0; i < 16; ++i) mx = fmaxf(mx, i < nv ? __uint_as_float(raw[i]) : -INFINITY);
          mx = fmaf(mx, kS2, add_of(c, 0));
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            mx = fmaxf(mx, i < nv ? fmaf(__uint_as_float(raw[i]), kS2, add_of(c, i)) : -INFINITY);
        }
        m = fmaxf(m, mx);
      };
      // ---- pass 2: unnormalised P = 2^(t - m) -> bf16x2 -> TMEM; row sum
      float l = 0.f;
      auto exp_chunk = [&](int c, const uint32_t(&raw)[16]) {
        const int nv = n_real - c * 16;
        const float off_u = add_of(c, 0) - m;  // UNIFORM: one offset per chunk
        uint32_t pk[8];
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float o0 = UNIFORM ? off_u : add_of(c, 2 * i) - m;
          const float o1 = UNIFORM ? off_u : add_of(c, 2 * i + 1) - m;
          const float e0 = 2 * i < nv ? ex2(fmaf(__uint_as_float(raw[2 * i]), kS2, o0)) : 0.f;
          const float e1 = 2 * i + 1 < nv ? ex2(fmaf(__uint_as_float(raw[2 * i + 1]), kS2, o1)) : 0.f;
          s0 += e0;
          s1 += e1;
          pk[i] = pack_bf16_alu(e0, e1);
        }
        l += s0 + s1;
        tmem_st8(tSg + c * 8, pk);  // columns [8c, 8c+8) <= the chunk just read: never ahead of an unread score
      };
      {
        // single register buffer: the 4 softmax warps per SM sub-partition (2 groups x 2 CTAs) hide the TMEM load
        // latency.  MEASURED (round 2): software-pipelining the loads one chunk ahead through a second 16-register
        // buffer costs 3 spilled registers at the kernel's 96-register budget and LOSES: 57.0 -> 65.5 us at B8 N4096
        // h10, 33.4 -> 36.9 us at B8 N1024 h20 (spill reloads are L2 round trips: the L1 carve-out is all shared memory)
        uint32_t ra[16];
#pragma unroll 1
        for (int c = 0; c < chunks; ++c) {
          tmem_ld16(tSg + c * 16, ra);
          tmem_ld_wait();
          max_chunk(c, ra);
        }
#ifdef DS_ATTN_TRACE
        tr_p1 += clock64() - tr_a;
        tr_a = clock64();
#endif
#pragma unroll 1
        for (int c = 0; c < chunks; ++c) {
          tmem_ld16(tSg + c * 16, ra);
          tmem_ld_wait();
          exp_chunk(c, ra);
        }
      }
      // the other group's row sum travels with the barriers: st.shared -> arrive(p_full) [release] -> MMA thread
      // [acquire] -> commit(o_full) -> our wait below [acquire]; and s_l is only rewritten after o_empty(n)
      s_l[g * kTile + row] = l;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
#ifdef DS_ATTN_TRACE
      tr_p2 += clock64() - tr_a;
      tr_a = clock64();
#endif

      // ---- epilogue: out = O_text / l_t + scale * O_ip / l_i   (blend BEFORE to_out, reference :258);
      //      group g writes output columns [32g, 32g+32)
      mbar_wait(o_full, n & 1);
#ifdef DS_ATTN_TRACE
      tr_o += clock64() - tr_a;
      tr_a = clock64();
#endif
      tc_fence_after();
      const float w_t = 1.0f / s_l[row], w_i = p.ip_scale / s_l[kTile + row];
      __nv_bfloat16* orow = p.out + (static_cast<size_t>(batch) * p.N + q_row) * p.C + head * kHd + g * 32;
      {
        uint32_t rt[32], ri[32];
        tmem_ld32(tOt + lane_base + g * 32, rt);
        tmem_ld32(tOi + lane_base + g * 32, ri);
        tmem_ld_wait();
        if (q_row < p.N) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              o[e] = fmaf(__uint_as_float(rt[q * 8 + e]), w_t, __uint_as_float(ri[q * 8 + e]) * w_i);
            uint4 u;
            u.x = pack_bf16(o[0], o[1]);
            u.y = pack_bf16(o[2], o[3]);
            u.z = pack_bf16(o[4], o[5]);
            u.w = pack_bf16(o[6], o[7]);
            reinterpret_cast<uint4*>(orow)[q] = u;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
#ifdef DS_ATTN_TRACE
      tr_ep += clock64() - tr_a;
#endif
    }
#ifdef DS_ATTN_TRACE
    if ((blockIdx.x % 97) == 0 && lane == 0 && (warp == 2 || warp == 6))
      printf("[trace] cross blk %d warp %d items %d: total %lld | wait S %lld, pass1 %lld, pass2 %lld, wait O %lld, epilogue %lld\n",
             blockIdx.x, warp, n_items, clock64() - tr_t0, tr_s, tr_p1, tr_p2, tr_o, tr_ep);
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static bool make_tok_map(CUtensorMap* m, const void* base, int cols, int ld, int tokens, int batch, int box_rows) {
  const uint64_t dims[3] = {static_cast<uint64_t>(cols), static_cast<uint64_t>(tokens), static_cast<uint64_t>(batch)};
  const uint64_t strides[2] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(tokens) * ld * 2};
  const uint32_t box[3] = {kHd, static_cast<uint32_t>(box_rows), 1};
  return encode_tmap_bf16(m, base, 3, dims, strides, box, nullptr);
}

static int launch_flash(const void* q, int ldq, int q_cols, const void* k, const void* v, int ldkv, int kv_cols,
                        int k_col0, int v_col0, void* out, int ldo, int B, int Nq, int Nkv, int heads, float scale,
                        cudaStream_t st) {
  DeviceInfo dev;
  if (!get_device(&dev)) return DS_ERR_CUDA;
  CUtensorMap tmQ, tmK, tmV;
  if (!make_tok_map(&tmQ, q, q_cols, ldq, Nq, B, kTile)) return DS_ERR_CUDA;
  if (!make_tok_map(&tmK, k, kv_cols, ldkv, Nkv, B, kTile)) return DS_ERR_CUDA;
  if (!make_tok_map(&tmV, v, kv_cols, ldkv, Nkv, B, kTile)) return DS_ERR_CUDA;
  FlashParams p;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.Nq = Nq;
  p.Nkv = Nkv;
  p.ldo = ldo;
  p.q_col0 = 0;
  p.k_col0 = k_col0;
  p.v_col0 = v_col0;
  p.scale_log2 = scale * kLog2e;
  dim3 grid((Nq + kTile - 1) / kTile, heads, B);
  // DS_FLASH_POLY=n sends n of every 8 exponentials to the FMA pipe (measured neutral: profiles/r02_attn_sweep.log)
  static const int flash_poly = [] {
    const char* e = getenv("DS_FLASH_POLY");
    return e ? atoi(e) : 0;
  }();
  {
    static const int flash_f2 = [] {  // DS_FLASH_F2=0: scalar FFMA/FADD instead of the packed fp32x2 forms
      const char* e = getenv("DS_FLASH_F2");
      return e ? atoi(e) : 1;
    }();
    static const int flash_elect = [] {  // DS_FLASH_ELECT=1: converged MMA-issue warp (A/B)
      const char* e = getenv("DS_FLASH_ELECT");
      return e ? atoi(e) : 0;
    }();
    static const int flash_dual = [] {  // DS_FLASH_DUAL=1: one MMA-issue thread per stream
      const char* e = getenv("DS_FLASH_DUAL");
      return e ? atoi(e) : 0;
    }();
    static bool attr5_set_dev[kMaxDevices] = {};
    bool& attr5_set = attr5_set_dev[device_slot()];
#define DS_F5_ATTR(P, F, E) \
  DS_CUDA_OK(cudaFuncSetAttribute(flash_attn_v5_kernel<P, F, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFlash5SmemBytes))
    if (!attr5_set) {
      DS_F5_ATTR(0, true, false);
      DS_F5_ATTR(1, true, false);
      DS_F5_ATTR(2, true, false);
      DS_F5_ATTR(0, false, false);
      DS_F5_ATTR(0, true, true);
      DS_F5_ATTR(2, true, true);
      DS_CUDA_OK(cudaFuncSetAttribute(flash_attn_v5_kernel<0, true, false, true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, kFlash5SmemBytes));
      attr5_set = true;
    }
#undef DS_F5_ATTR
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kFlash5Threads);
    cfg.dynamicSmemBytes = kFlash5SmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    pdl_attr(&attr[0]);
    cfg.attrs = attr;
    cfg.numAttrs = 1;
#define DS_F5_LAUNCH(P, F, E) DS_CUDA_OK(cudaLaunchKernelEx(&cfg, flash_attn_v5_kernel<P, F, E>, tmQ, tmK, tmV, p))
    if (flash_dual && flash_f2 && !flash_elect && flash_poly == 0) {
      cfg.blockDim = dim3(kFlash5Threads + 32);
      DS_CUDA_OK(cudaLaunchKernelEx(&cfg, flash_attn_v5_kernel<0, true, false, true>, tmQ, tmK, tmV, p));
    } else if (!flash_f2)
      DS_F5_LAUNCH(0, false, false);
    else if (flash_elect && flash_poly >= 2)
      DS_F5_LAUNCH(2, true, true);
    else if (flash_elect)
      DS_F5_LAUNCH(0, true, true);
    else if (flash_poly == 1)
      DS_F5_LAUNCH(1, true, false);
    else if (flash_poly == 2)
      DS_F5_LAUNCH(2, true, false);
    else
      DS_F5_LAUNCH(0, true, false);
#undef DS_F5_LAUNCH
    DS_LAUNCH_OK("flash_attn_v5_kernel");
    return DS_OK;
  }
}

}  // namespace ds

using namespace ds;

extern "C" int ds_attention_self(const void* qkv, void* out, int B, int N, int heads, void* stream) {
  DS_REQUIRE(qkv && out, "ds_attention_self: NULL pointer");
  DS_REQUIRE(B > 0 && N > 0 && heads > 0, "ds_attention_self: bad shape");
  DS_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             "ds_attention_self: pointers must be 16-byte aligned");
  const int C = heads * kHd;
  // one tensor map over the fused [B][N][3C] projection; K and V are column offsets C and 2C
  return launch_flash(qkv, 3 * C, 3 * C, qkv, qkv, 3 * C, 3 * C, C, 2 * C, out, C, B, N, N, heads, 0.125f,
                      static_cast<cudaStream_t>(stream));
}

extern "C" int ds_resampler_attn(const void* q, const void* kv, void* out, int Bc, int nq, int n_kv, int heads,
                                 void* stream) {
  DS_REQUIRE(q && kv && out, "ds_resampler_attn: NULL pointer");
  DS_REQUIRE(Bc > 0 && nq > 0 && n_kv > 0 && heads > 0, "ds_resampler_attn: bad shape");
  DS_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(kv) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             "ds_resampler_attn: pointers must be 16-byte aligned");
  const int C = heads * kHd;
  // (q * d^-1/4)(k * d^-1/4)^T == q k^T / sqrt(d)   (src/models/resampler.py:69-70)
  return launch_flash(q, C, C, kv, kv, 2 * C, 2 * C, 0, C, out, C, Bc, nq, n_kv, heads, 0.125f,
                      static_cast<cudaStream_t>(stream));
}

extern "C" int ds_attention_cross_ip(const ds_cross_ip_args* a, void* stream) {
  DS_REQUIRE(a != nullptr, "ds_attention_cross_ip: args is NULL");
  DS_REQUIRE(a->q && a->kv_text && a->kv_ip && a->bbox && a->out, "ds_attention_cross_ip: NULL pointer");
  DS_REQUIRE(a->B > 0 && a->N > 0 && a->heads > 0, "ds_attention_cross_ip: bad shape");
  DS_REQUIRE(a->n_text > 0 && a->n_ip > 0, "ds_attention_cross_ip: n_text and n_ip must be positive");
  DS_REQUIRE(a->num_ips > 0 && a->num_ips <= kMaxIps && a->tokens_per_ip > 0 && a->num_dummy >= 0 &&
                 a->num_dummy + a->num_ips * a->tokens_per_ip == a->n_ip,
             "ds_attention_cross_ip: n_ip (%d) != num_dummy (%d) + num_ips (%d) * tokens_per_ip (%d)", a->n_ip,
             a->num_dummy, a->num_ips, a->tokens_per_ip);
  const int nt_pad = (a->n_text + 15) / 16 * 16, nip_pad = (a->n_ip + 15) / 16 * 16;
  DS_REQUIRE(nt_pad + nip_pad <= 192, "ds_attention_cross_ip: padded key count %d exceeds 192", nt_pad + nip_pad);
  int Hd, Wd;
  if (!derive_hw(a->N, a->aspect_ratio, &Hd, &Wd)) {
    set_error("ds_attention_cross_ip: cannot factor N=%d for aspect_ratio=%f", a->N, a->aspect_ratio);
    return DS_ERR_INVALID;
  }
  DeviceInfo dev;
  if (!get_device(&dev)) return DS_ERR_CUDA;
  const int C = a->heads * kHd;
  CUtensorMap tmQ, tmT, tmI;
  if (!make_tok_map(&tmQ, a->q, C, C, a->N, a->B, kTile)) return DS_ERR_CUDA;
  if (!make_tok_map(&tmT, a->kv_text, 2 * C, 2 * C, a->n_text, a->B, nt_pad)) return DS_ERR_CUDA;
  if (!make_tok_map(&tmI, a->kv_ip, 2 * C, 2 * C, a->n_ip, a->B, nip_pad)) return DS_ERR_CUDA;
  const int n_keys = nt_pad + nip_pad;
  const int kv_bytes = ((n_keys * 128) + 1023) & ~1023;
  // TMEM layout of the kernel: P_ip must end below column 128 (77 text + 80 IP keys: 80 + 40 = 120)
  DS_REQUIRE(nt_pad + nip_pad / 2 <= 128,
             "ds_attention_cross_ip: padded key counts (%d text, %d ip) do not fit the kernel's TMEM layout "
             "(text + ip/2 <= 128)", nt_pad, nip_pad);
  CrossParams p;
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.bbox = a->bbox;
  p.N = a->N;
  p.C = C;
  p.n_text = a->n_text;
  p.n_ip = a->n_ip;
  p.nt_pad = nt_pad;
  p.nip_pad = nip_pad;
  p.num_ips = a->num_ips;
  p.tokens_per_ip = a->tokens_per_ip;
  p.num_dummy = a->num_dummy;
  p.Hd = Hd;
  p.Wd = Wd;
  p.ip_scale = a->ip_scale;
  {
    const int q_tiles = (a->N + kTile - 1) / kTile;
    const long long total_ll = static_cast<long long>(a->B) * a->heads * q_tiles;
    DS_REQUIRE(total_ll < (1ll << 30), "ds_attention_cross_ip: too many tiles");
    const int total = static_cast<int>(total_ll);
    const int smem2 = 2 * kTileBytes + 2 * kv_bytes + 1024 + 128 + 2 * kTile * 4;
    static int attr_smem2_dev[kMaxDevices] = {};
    int& attr_smem2 = attr_smem2_dev[device_slot()];
    if (smem2 > attr_smem2) {
      DS_CUDA_OK(cudaFuncSetAttribute(cross_ip_attn_v2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
      DS_CUDA_OK(cudaFuncSetAttribute(cross_ip_attn_v2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
      attr_smem2 = smem2;
    }
    const int grid2 = total < 2 * dev.num_sms ? total : 2 * dev.num_sms;
    const bool uniform = (a->tokens_per_ip % 16 == 0) && (a->num_dummy % 16 == 0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid2);
    cfg.blockDim = dim3(kCross2Threads);
    cfg.dynamicSmemBytes = smem2;
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    pdl_attr(&attr[0]);
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int heads_i = a->heads;
    if (uniform)
      DS_CUDA_OK(cudaLaunchKernelEx(&cfg, cross_ip_attn_v2_kernel<true>, tmQ, tmT, tmI, p, heads_i, q_tiles, total));
    else
      DS_CUDA_OK(cudaLaunchKernelEx(&cfg, cross_ip_attn_v2_kernel<false>, tmQ, tmT, tmI, p, heads_i, q_tiles, total));
    DS_LAUNCH_OK("cross_ip_attn_v2_kernel");
    return DS_OK;
  }
}
