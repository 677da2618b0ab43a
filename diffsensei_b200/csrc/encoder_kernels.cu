// encoder_kernels.cu — what the conditioning encoders need beyond the UNet's kernels (SURVEY.md §8f ranks 2-3:
// the CLIP ViT-H and Magi ViT-MAE image encoders of prepare_ip_image_embeds, src/pipelines/pipeline_diffsensei.py:125-128,
// and the two SDXL CLIP text encoders of encode_prompt, :232-245).  Their linears / MLPs run on the tcgen05 GEMM
// (ds_gemm_bf16, bias / GELU / quick-GELU / residual epilogues) and their LayerNorms on ds_layernorm; this file adds
//   ds_attention_small : softmax(scale * Q K^T [+ causal mask]) V for SHORT sequences (<= 320 keys) and any head
//                        width that is a multiple of 8 up to 128 — 77 text tokens (causal), 197 / 257 image tokens, heads
//                        of 64 (text, ViT-MAE) and 80 (ViT-H, outside the flash kernel's head_dim 64).  These encoders
//                        run once per panel (1.3 TFLOP of GEMMs for four 224x224 character crops, 43 GFLOP of
//                        attention), so the attention is a plain CUDA-core kernel: K / V of one (batch, head) staged
//                        in shared memory, one warp per query row, fp32 softmax.                   [latency-bound]
//   ds_embed_tokens    : out[b][t][:] = token_embedding[ids[b][t]][:] + position_embedding[t][:]   (CLIPTextEmbeddings)
#include "ds_common.cuh"
#include "ds_host.h"

namespace ds {

constexpr int kSmallAttnThreads = 128;   // 4 warps, one query row each per sweep
constexpr int kSmallAttnRows = 32;       // query rows per CTA
constexpr int kSmallAttnMaxKeys = 320;   // 10 keys per lane

__global__ void __launch_bounds__(kSmallAttnThreads)
attention_small_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                       const __nv_bfloat16* __restrict__ v, __nv_bfloat16* __restrict__ out, int Nq, int Nk, int hd,
                       long long ldq, long long ldk, long long ldv, long long ldo, float scale_log2, int causal) {
  extern __shared__ uint32_t sm[];
  const int hw = hd >> 1;        // bf16x2 words per row
  const int row_words = (hw & 1) ? hw : hw + 1;           // odd word stride: lane j reads row j without bank conflicts
  uint32_t* sK = sm;                                      // [Nk][row_words]
  uint32_t* sV = sK + static_cast<size_t>(Nk) * row_words;
  float* sQ = reinterpret_cast<float*>(sV + static_cast<size_t>(Nk) * row_words);   // [4 warps][hd]
  float* sP = sQ + 4 * hd;                                                         // [4 warps][Nk]
  const int b = blockIdx.z, head = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16* kb = k + static_cast<long long>(b) * Nk * ldk + head * hd;
  const __nv_bfloat16* vb = v + static_cast<long long>(b) * Nk * ldv + head * hd;
  // stage K and V of this (batch, head): 16-byte global loads, 4-byte shared stores into the padded rows
  const int vec_per_row = hd >> 3;
  for (int i = threadIdx.x; i < Nk * vec_per_row; i += blockDim.x) {
    const int j = i / vec_per_row, c = i - j * vec_per_row;
    const uint4 uk = __ldg(reinterpret_cast<const uint4*>(kb + static_cast<long long>(j) * ldk) + c);
    const uint4 uv = __ldg(reinterpret_cast<const uint4*>(vb + static_cast<long long>(j) * ldv) + c);
    uint32_t* dk = sK + static_cast<size_t>(j) * row_words + c * 4;
    uint32_t* dv = sV + static_cast<size_t>(j) * row_words + c * 4;
    dk[0] = uk.x; dk[1] = uk.y; dk[2] = uk.z; dk[3] = uk.w;
    dv[0] = uv.x; dv[1] = uv.y; dv[2] = uv.z; dv[3] = uv.w;
  }
  __syncthreads();
  const int r_end = min(Nq, (static_cast<int>(blockIdx.x) + 1) * kSmallAttnRows);
  float* myQ = sQ + warp * hd;
  float* myP = sP + static_cast<size_t>(warp) * Nk;
  for (int r = blockIdx.x * kSmallAttnRows + warp; r < r_end; r += 4) {
    const __nv_bfloat16* qr = q + (static_cast<long long>(b) * Nq + r) * ldq + head * hd;
    for (int d = lane; d < hd; d += 32) myQ[d] = __bfloat162float(qr[d]);
    __syncwarp();
    const int kmax = causal ? min(Nk, r + 1) : Nk;       // keys a causal row may see
    float s[kSmallAttnMaxKeys / 32];
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < kSmallAttnMaxKeys / 32; ++t) {
      const int j = lane + 32 * t;
      float acc = -INFINITY;
      if (j < kmax) {
        const uint32_t* kr = sK + static_cast<size_t>(j) * row_words;
        float a0 = 0.f, a1 = 0.f;
        for (int w = 0; w < hw; ++w) {
          const uint32_t kw = kr[w];
          const float2 qq = *reinterpret_cast<const float2*>(myQ + 2 * w);
          a0 = fmaf(qq.x, bf16_lo(kw), a0);
          a1 = fmaf(qq.y, bf16_hi(kw), a1);
        }
        acc = (a0 + a1) * scale_log2;
      }
      s[t] = acc;
      m = fmaxf(m, acc);
    }
    m = warp_max(m);
    float l = 0.f;
#pragma unroll
    for (int t = 0; t < kSmallAttnMaxKeys / 32; ++t) {
      const int j = lane + 32 * t;
      const float pj = j < kmax ? exp2f(s[t] - m) : 0.f;
      l += pj;
      if (j < Nk) myP[j] = pj;
    }
    l = warp_sum(l);
    __syncwarp();
    const float inv = 1.0f / l;
    __nv_bfloat16* orow = out + (static_cast<long long>(b) * Nq + r) * ldo + head * hd;
    for (int w = lane; w < hw; w += 32) {
      float o0 = 0.f, o1 = 0.f;
      for (int j = 0; j < kmax; ++j) {
        const float pj = myP[j];
        const uint32_t vw = sV[static_cast<size_t>(j) * row_words + w];
        o0 = fmaf(pj, bf16_lo(vw), o0);
        o1 = fmaf(pj, bf16_hi(vw), o1);
      }
      *reinterpret_cast<uint32_t*>(orow + 2 * w) = pack_bf16(o0 * inv, o1 * inv);
    }
    __syncwarp();
  }
}

__global__ void embed_tokens_kernel(const int* __restrict__ ids, const uint4* __restrict__ tok,
                                    const uint4* __restrict__ pos, uint4* __restrict__ out, int L, int cv, int vocab,
                                    long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % cv);
  const long long row = i / cv;
  const int t = static_cast<int>(row % L);
  int id = ids[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const uint4 a = __ldg(tok + static_cast<long long>(id) * cv + c), p = __ldg(pos + static_cast<long long>(t) * cv + c);
  uint4 o;
  o.x = pack_bf16(bf16_lo(a.x) + bf16_lo(p.x), bf16_hi(a.x) + bf16_hi(p.x));
  o.y = pack_bf16(bf16_lo(a.y) + bf16_lo(p.y), bf16_hi(a.y) + bf16_hi(p.y));
  o.z = pack_bf16(bf16_lo(a.z) + bf16_lo(p.z), bf16_hi(a.z) + bf16_hi(p.z));
  o.w = pack_bf16(bf16_lo(a.w) + bf16_lo(p.w), bf16_hi(a.w) + bf16_hi(p.w));
  out[i] = o;
}

}  // namespace ds

extern "C" int ds_attention_small(const void* q, const void* k, const void* v, void* out, int B, int Nq, int Nk,
                                  int heads, int head_dim, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                                  float scale, int causal, void* stream) {
  using namespace ds;
  DS_REQUIRE(q && k && v && out, "ds_attention_small: NULL pointer");
  DS_REQUIRE(B > 0 && Nq > 0 && Nk > 0 && heads > 0, "ds_attention_small: bad shape");
  DS_REQUIRE(Nk <= kSmallAttnMaxKeys, "ds_attention_small: at most %d keys (got %d); long sequences use ds_attention_self",
             kSmallAttnMaxKeys, Nk);
  DS_REQUIRE(head_dim >= 8 && head_dim <= 256 && head_dim % 8 == 0, "ds_attention_small: head_dim %% 8 == 0, <= 256");
  DS_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0 && ldq >= heads * head_dim &&
                 ldk >= heads * head_dim && ldv >= heads * head_dim && ldo >= heads * head_dim,
             "ds_attention_small: row strides must cover heads*head_dim and be multiples of 8 elements");
  DS_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(k) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(v) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0,
             "ds_attention_small: q / k / v must be 16-byte aligned");
  DS_REQUIRE(!causal || Nq == Nk, "ds_attention_small: the causal mask needs Nq == Nk");
  DeviceInfo dev;
  if (!get_device(&dev)) return DS_ERR_CUDA;
  const int hw = head_dim / 2;
  const int row_words = (hw & 1) ? hw : hw + 1;
  const size_t smem = static_cast<size_t>(2) * Nk * row_words * 4 + 4 * head_dim * 4 + static_cast<size_t>(4) * Nk * 4;
  static size_t attr[kMaxDevices] = {};
  if (smem > 48 * 1024 && smem > attr[device_slot()]) {
    DS_CUDA_OK(cudaFuncSetAttribute(attention_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr[device_slot()] = smem;
  }
  dim3 grid((Nq + kSmallAttnRows - 1) / kSmallAttnRows, heads, B);
  attention_small_kernel<<<grid, kSmallAttnThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k), static_cast<const __nv_bfloat16*>(v),
      static_cast<__nv_bfloat16*>(out), Nq, Nk, head_dim, ldq, ldk, ldv, ldo, scale * 1.4426950408889634f, causal);
  DS_LAUNCH_OK("attention_small_kernel");
  return DS_OK;
}

extern "C" int ds_embed_tokens(const int* ids, const void* tok_emb, const void* pos_emb, void* out, int B, int L, int C,
                               int vocab, void* stream) {
  using namespace ds;
  DS_REQUIRE(ids && tok_emb && pos_emb && out && B > 0 && L > 0 && C > 0 && C % 8 == 0 && vocab > 0,
             "ds_embed_tokens: bad arguments (C %% 8 == 0)");
  DeviceInfo dev;
  if (!get_device(&dev)) return DS_ERR_CUDA;
  const long long total = static_cast<long long>(B) * L * (C / 8);
  embed_tokens_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      ids, static_cast<const uint4*>(tok_emb), static_cast<const uint4*>(pos_emb), static_cast<uint4*>(out), L, C / 8,
      vocab, total);
  DS_LAUNCH_OK("embed_tokens_kernel");
  return DS_OK;
}
