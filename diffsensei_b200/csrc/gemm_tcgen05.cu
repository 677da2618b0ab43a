// gemm_tcgen05.cu — persistent, warp-specialised bf16 GEMM / implicit-GEMM conv3x3 for sm_100a.
//
//   out[M][Nout] = epilogue( A[M][K] * W[N][K]^T ),  fp32 accumulation in TMEM.
//
// One CTA per SM (persistent, static round-robin over 128 x BN output tiles), 384 threads:
//   warp 0    TMA producer   : one lane streams A (128x64) and W (BNx64) k-slices into a STAGES-deep
//                              shared-memory ring (128-byte swizzle), arming a "full" mbarrier per stage
//   warp 1    MMA issuer     : one lane issues 4 x tcgen05.mma (M=128, N=BN, K=16) per stage into one of
//                              two TMEM accumulators, tcgen05.commit frees the stage / publishes the tile
//   warp 2    TMEM allocator : 2*BN columns (double-buffered accumulator)
//   warps 4-11 epilogue      : tcgen05.ld the finished accumulator (thread <-> output row; 4 lane quadrants
//                              x 2 column halves), apply
//                              bias / row-bias / GEGLU / activation / residual in fp32, round once, store —
//                              overlapping the next tile's MMAs thanks to the second accumulator
//
// conv3x3 mode (ds_conv3x3_nhwc): identical MMA pipeline; only the producer and the row->address map
// change.  An M tile is an 8x16 patch of output pixels of one image; for filter tap (r,s) and channel
// chunk c the A slice is the 4-D TMA box {64 ch, 16 px, 8 px, 1 img} of the NHWC input at pixel offset
// (r-1, s-1) — halo and zero padding come from TMA out-of-bounds fill, stride-2 from the tensor map's
// element strides.  K runs over (tap, channel): weights are packed [Cout][3][3][Cin].
//
// Reference arithmetic replaced: every nn.Linear / nn.Conv2d on the UNet sampling path
// (src/models/attention_processor.py:56-84,207-261; diffusers blocks reached from src/models/unet.py:190-338).
#include "ds_common.cuh"
#include "ds_host.h"

namespace ds {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kGemmThreads = 384;  // 4 control warps + 8 epilogue warps
constexpr int kABytes = kBM * kBK * 2;  // 16 KiB per stage
constexpr int kConvTileH = 8;
constexpr int kConvTileW = 16;

struct GemmParams {
  const float* bias;
  const float* rowbias;
  const __nv_bfloat16* residual;
  void* out;
  int M, N, K;
  int n_out;  // output columns (N, or N/2 for GEGLU)
  int ldo, ldres, rows_per_batch, ldrb;
  int epilogue, out_fp32;
  float out_scale;
  int num_m_tiles, num_n_tiles, num_k_iters;
  // conv geometry
  int conv, stride, Ho, Wo, tiles_x, tiles_y, cin_chunks;
};

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * Cfg::kBBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;  // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;      // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);  // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.num_m_tiles * p.num_n_tiles;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_blk = tile % p.num_n_tiles;
        const int m_blk = tile / p.num_n_tiles;
        int img = 0, x0 = 0, y0 = 0;
        if (p.conv) {
          const int per_img = p.tiles_x * p.tiles_y;
          img = m_blk / per_img;
          const int rem = m_blk - img * per_img;
          y0 = (rem / p.tiles_x) * kConvTileH;
          x0 = (rem % p.tiles_x) * kConvTileW;
        }
        for (int kb = 0; kb < p.num_k_iters; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          if (p.conv) {
            const int tap = kb / p.cin_chunks;
            const int cc = kb - tap * p.cin_chunks;
            const int r = tap / 3, s = tap - r * 3;
            tma_load_4d(sA + stage * kABytes, &tmA, &full_bar[stage], cc * kBK, x0 * p.stride + s - 1,
                        y0 * p.stride + r - 1, img);
          } else {
            tma_load_2d(sA + stage * kABytes, &tmA, &full_bar[stage], kb * kBK, m_blk * kBM);
          }
          tma_load_2d(sB + stage * Cfg::kBBytes, &tmB, &full_bar[stage], kb * kBK, n_blk * BN);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kBM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < p.num_k_iters; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * kABytes);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t adesc = make_sw128_desc(a_addr + k * 32, 1024, 16);
            const uint64_t bdesc = make_sw128_desc(b_addr + k * 32, 1024, 16);
            umma_ss(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // stage reusable once these MMAs have read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: 8 warps = 4 TMEM lane quadrants
    // (rows) x 2 column halves; thread <-> one output row, 32-column chunks
    const int wq = warp & 3;             // TMEM lane quadrant this warp may access
    const int half = (warp - 4) >> 2;    // which half of the tile's output columns
    const bool geglu = p.epilogue == DS_EPI_GEGLU;
    const int bn_out = geglu ? BN / 2 : BN;
    const int c_begin = half * (bn_out / 2), c_end = c_begin + bn_out / 2;
    const bool vec_ok = (p.n_out % 8 == 0) && (p.ldo % 8 == 0) && (!p.residual || p.ldres % 8 == 0);

    // v[j] += src[j] (src already offset to the chunk's first column n0), guarded by n0 + j < N
    auto add32 = [&](float(&v)[32], const float* __restrict__ src, int n0) {
      if (n0 + 32 <= p.N && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 f = __ldg(reinterpret_cast<const float4*>(src) + q);
          v[q * 4 + 0] += f.x;
          v[q * 4 + 1] += f.y;
          v[q * 4 + 2] += f.z;
          v[q * 4 + 3] += f.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + j < p.N) v[j] += __ldg(src + j);
      }
    };

    int iter = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      const int n_blk = tile % p.num_n_tiles;
      const int m_blk = tile / p.num_n_tiles;
      const int r_local = wq * 32 + lane;

      // row -> (valid, output row index, batch index)
      bool row_ok;
      long long orow;
      int batch;
      if (p.conv) {
        const int per_img = p.tiles_x * p.tiles_y;
        const int img = m_blk / per_img;
        const int rem = m_blk - img * per_img;
        const int y = (rem / p.tiles_x) * kConvTileH + r_local / kConvTileW;
        const int x = (rem % p.tiles_x) * kConvTileW + r_local % kConvTileW;
        row_ok = (y < p.Ho) && (x < p.Wo);
        orow = (static_cast<long long>(img) * p.Ho + y) * p.Wo + x;
        batch = img;
      } else {
        const int row = m_blk * kBM + r_local;
        row_ok = row < p.M;
        orow = row;
        batch = p.rowbias ? row / p.rows_per_batch : 0;
      }

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * BN;

      for (int c0 = c_begin; c0 < c_end; c0 += 32) {
        const int nw0 = n_blk * BN + c0;      // weight-row index of column 0 of this chunk
        const int no0 = n_blk * bn_out + c0;  // output column of column 0 of this chunk
        if (no0 >= p.n_out) break;            // warp-uniform
        const bool full_chunk = vec_ok && (no0 + 32 <= p.n_out);
        uint32_t raw[32];
        tmem_ld32(t_row + c0, raw);
        // issue the residual loads before waiting on TMEM so their latency overlaps
        uint4 rres[4];
        const bool vec_res = p.residual != nullptr && row_ok && full_chunk;
        if (vec_res) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.residual + orow * p.ldres + no0);
#pragma unroll
          for (int q = 0; q < 4; ++q) rres[q] = __ldg(rp + q);
        }
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
        if (p.bias) add32(v, p.bias + nw0, nw0);
        if (p.rowbias && row_ok) add32(v, p.rowbias + static_cast<long long>(batch) * p.ldrb + nw0, nw0);
        if (geglu) {
          uint32_t graw[32];
          tmem_ld32(t_row + BN / 2 + c0, graw);
          tmem_ld_wait();
          float g[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) g[j] = __uint_as_float(graw[j]);
          if (p.bias) add32(g, p.bias + nw0 + BN / 2, nw0 + BN / 2);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= gelu_erf_fast(g[j]);
        } else if (p.epilogue == DS_EPI_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf_fast(v[j]);
        } else if (p.epilogue == DS_EPI_SILU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
        }

        if (row_ok) {
          if (vec_res) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              v[q * 8 + 0] += bf16_lo(rres[q].x);
              v[q * 8 + 1] += bf16_hi(rres[q].x);
              v[q * 8 + 2] += bf16_lo(rres[q].y);
              v[q * 8 + 3] += bf16_hi(rres[q].y);
              v[q * 8 + 4] += bf16_lo(rres[q].z);
              v[q * 8 + 5] += bf16_hi(rres[q].z);
              v[q * 8 + 6] += bf16_lo(rres[q].w);
              v[q * 8 + 7] += bf16_hi(rres[q].w);
            }
          } else if (p.residual) {
            const __nv_bfloat16* rp = p.residual + orow * p.ldres + no0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (no0 + j < p.n_out) v[j] += __bfloat162float(rp[j]);
          }
          if (p.out_scale != 0.0f) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= p.out_scale;
          }
          if (p.out_fp32) {
            float* op = reinterpret_cast<float*>(p.out) + orow * p.ldo + no0;
            if (full_chunk) {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                reinterpret_cast<float4*>(op)[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (no0 + j < p.n_out) op[j] = v[j];
            }
          } else {
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldo + no0;
            if (full_chunk) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint4 u;
                u.x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
                u.y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
                u.z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
                u.w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
                reinterpret_cast<uint4*>(op)[q] = u;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (no0 + j < p.n_out) op[j] = __float2bfloat16(v[j]);
            }
          }
        }
      }
      // release the accumulator back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }

  // ---------------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <int BN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int num_sms,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;  // benign race: idempotent
  if (!attr_set) {
    DS_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tcgen05<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg::kSmemBytes));
    attr_set = true;
  }
  const int total = p.num_m_tiles * p.num_n_tiles;
  const int grid = total < num_sms ? total : num_sms;
  gemm_bf16_tcgen05<BN><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, p);
  DS_LAUNCH_OK("gemm_bf16_tcgen05");
  return DS_OK;
}

static int pick_bn(int N, int epilogue) {
  if (epilogue == DS_EPI_GEGLU) return 256;
  if (N <= 128) return 128;
  const double e256 = static_cast<double>(N) / (((N + 255) / 256) * 256);
  const double e128 = static_cast<double>(N) / (((N + 127) / 128) * 128);
  return (e256 + 0.04 >= e128) ? 256 : 128;
}

static int run_gemm(const CUtensorMap& tmA, const void* w, int ldw, GemmParams& p, cudaStream_t stream) {
  DeviceInfo dev;
  if (!get_device(&dev)) return DS_ERR_CUDA;
  const int bn = pick_bn(p.N, p.epilogue);
  CUtensorMap tmB;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(p.K), static_cast<uint64_t>(p.N)};
    const uint64_t strides[1] = {static_cast<uint64_t>(ldw) * 2};
    const uint32_t box[2] = {kBK, static_cast<uint32_t>(bn)};
    if (!encode_tmap_bf16(&tmB, w, 2, dims, strides, box, nullptr)) return DS_ERR_CUDA;
  }
  p.num_n_tiles = (p.N + bn - 1) / bn;
  if (bn == 256) return launch_gemm<256>(tmA, tmB, p, dev.num_sms, stream);
  return launch_gemm<128>(tmA, tmB, p, dev.num_sms, stream);
}

}  // namespace ds

extern "C" int ds_gemm_bf16(const ds_gemm_args* a, void* stream) {
  using namespace ds;
  DS_REQUIRE(a != nullptr, "ds_gemm_bf16: args is NULL");
  DS_REQUIRE(a->a && a->w && a->out, "ds_gemm_bf16: a/w/out must be non-NULL");
  DS_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "ds_gemm_bf16: M,N,K must be positive (got %d,%d,%d)", a->M, a->N,
             a->K);
  DS_REQUIRE(a->K % 8 == 0 && a->lda % 8 == 0 && a->ldw % 8 == 0,
             "ds_gemm_bf16: K, lda, ldw must be multiples of 8 (got %d,%d,%d)", a->K, a->lda, a->ldw);
  DS_REQUIRE(a->lda >= a->K && a->ldw >= a->K, "ds_gemm_bf16: lda/ldw smaller than K");
  DS_REQUIRE((reinterpret_cast<uintptr_t>(a->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->w) & 15) == 0,
             "ds_gemm_bf16: a and w must be 16-byte aligned");
  DS_REQUIRE(a->epilogue >= DS_EPI_NONE && a->epilogue <= DS_EPI_SILU, "ds_gemm_bf16: bad epilogue %d", a->epilogue);
  if (a->epilogue == DS_EPI_GEGLU)
    DS_REQUIRE(a->N % 256 == 0, "ds_gemm_bf16: GEGLU needs N %% 256 == 0 (128 value + 128 gate rows per block)");
  if (a->rowbias) DS_REQUIRE(a->rows_per_batch > 0, "ds_gemm_bf16: rowbias needs rows_per_batch > 0");
  const int n_out = a->epilogue == DS_EPI_GEGLU ? a->N / 2 : a->N;
  DS_REQUIRE(a->ldo >= n_out, "ds_gemm_bf16: ldo (%d) smaller than output width (%d)", a->ldo, n_out);
  if (a->residual) DS_REQUIRE(a->ldres >= n_out, "ds_gemm_bf16: ldres smaller than output width");

  CUtensorMap tmA;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(a->K), static_cast<uint64_t>(a->M)};
    const uint64_t strides[1] = {static_cast<uint64_t>(a->lda) * 2};
    const uint32_t box[2] = {kBK, kBM};
    if (!encode_tmap_bf16(&tmA, a->a, 2, dims, strides, box, nullptr)) return DS_ERR_CUDA;
  }
  GemmParams p{};
  p.bias = a->bias;
  p.rowbias = a->rowbias;
  p.residual = static_cast<const __nv_bfloat16*>(a->residual);
  p.out = a->out;
  p.M = a->M;
  p.N = a->N;
  p.K = a->K;
  p.n_out = n_out;
  p.ldo = a->ldo;
  p.ldres = a->ldres;
  p.rows_per_batch = a->rows_per_batch > 0 ? a->rows_per_batch : 1;
  p.ldrb = a->rowbias_ld > 0 ? a->rowbias_ld : a->N;
  p.epilogue = a->epilogue;
  p.out_fp32 = a->out_fp32;
  p.out_scale = a->out_scale == 1.0f ? 0.0f : a->out_scale;
  p.num_m_tiles = (a->M + kBM - 1) / kBM;
  p.num_k_iters = (a->K + kBK - 1) / kBK;
  p.conv = 0;
  return run_gemm(tmA, a->w, a->ldw, p, static_cast<cudaStream_t>(stream));
}

extern "C" int ds_conv3x3_nhwc(const ds_conv3x3_args* a, void* stream) {
  using namespace ds;
  DS_REQUIRE(a != nullptr, "ds_conv3x3_nhwc: args is NULL");
  DS_REQUIRE(a->x && a->w && a->out, "ds_conv3x3_nhwc: x/w/out must be non-NULL");
  DS_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0 && a->Cin > 0 && a->Cout > 0, "ds_conv3x3_nhwc: bad geometry");
  DS_REQUIRE(a->Cin % 64 == 0, "ds_conv3x3_nhwc: Cin must be a multiple of 64 (got %d)", a->Cin);
  DS_REQUIRE(a->stride == 1 || a->stride == 2, "ds_conv3x3_nhwc: stride must be 1 or 2 (got %d)", a->stride);
  DS_REQUIRE((reinterpret_cast<uintptr_t>(a->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->w) & 15) == 0,
             "ds_conv3x3_nhwc: x and w must be 16-byte aligned");
  const int Ho = (a->H - 1) / a->stride + 1;
  const int Wo = (a->W - 1) / a->stride + 1;

  CUtensorMap tmA;
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(a->Cin), static_cast<uint64_t>(a->W),
                              static_cast<uint64_t>(a->H), static_cast<uint64_t>(a->B)};
    const uint64_t strides[3] = {static_cast<uint64_t>(a->Cin) * 2, static_cast<uint64_t>(a->W) * a->Cin * 2,
                                 static_cast<uint64_t>(a->H) * a->W * a->Cin * 2};
    const uint32_t box[4] = {kBK, static_cast<uint32_t>(kConvTileW * a->stride),
                             static_cast<uint32_t>(kConvTileH * a->stride), 1};
    const uint32_t es[4] = {1, static_cast<uint32_t>(a->stride), static_cast<uint32_t>(a->stride), 1};
    if (!encode_tmap_bf16(&tmA, a->x, 4, dims, strides, box, es)) return DS_ERR_CUDA;
  }
  GemmParams p{};
  p.bias = a->bias;
  p.rowbias = a->rowbias;
  p.residual = static_cast<const __nv_bfloat16*>(a->residual);
  p.out = a->out;
  p.tiles_x = (Wo + kConvTileW - 1) / kConvTileW;
  p.tiles_y = (Ho + kConvTileH - 1) / kConvTileH;
  p.M = a->B * Ho * Wo;
  p.N = a->Cout;
  p.K = 9 * a->Cin;
  p.n_out = a->Cout;
  p.ldo = a->Cout;
  p.ldres = a->Cout;
  p.rows_per_batch = 1;
  p.ldrb = a->rowbias_ld > 0 ? a->rowbias_ld : a->Cout;
  p.epilogue = DS_EPI_NONE;
  p.out_fp32 = a->out_fp32;
  p.out_scale = a->out_scale == 1.0f ? 0.0f : a->out_scale;
  p.num_m_tiles = a->B * p.tiles_x * p.tiles_y;
  p.num_k_iters = 9 * (a->Cin / kBK);
  p.conv = 1;
  p.stride = a->stride;
  p.Ho = Ho;
  p.Wo = Wo;
  p.cin_chunks = a->Cin / kBK;
  return run_gemm(tmA, a->w, 9 * a->Cin, p, static_cast<cudaStream_t>(stream));
}
