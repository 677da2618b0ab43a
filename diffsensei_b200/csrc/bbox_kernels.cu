// bbox_kernels.cu — the two bbox-driven HBM kernels of the manga UNet.
//
//   ds_dialog_embed_add : UNetMangaModel.encode_dialog_bbox            (src/models/unet.py:88-114)
//   ds_ip_mask          : MaskedIPAttnProcessor2_0.prepare_attention_mask_ip
//                                                                      (src/models/attention_processor.py:115-169)
// The reference drives both from Python loops with one device->host sync per scalar (256 resp. 4480 syncs per
// step at cfg2); here the boxes are kernel arguments read from device memory.
// ip_mask.cuh holds the predicate shared with the fused cross-attention kernel.
#include "ds_common.cuh"
#include "ds_host.h"
#include "ip_mask.cuh"

namespace ds {

constexpr int kMaxDialogs = 32;

// One thread per (pixel, 8-channel vector); only pixels inside a box are touched (read-modify-write).
__global__ void dialog_embed_add_kernel(uint4* __restrict__ sample, const float* __restrict__ emb,
                                        const float* __restrict__ dialog_bbox, int H, int W, int C, int nd,
                                        int round_bf16) {
  __shared__ int box[kMaxDialogs][4];
  const int b = blockIdx.y;
  if (threadIdx.x < nd) {
    const float* bb = dialog_bbox + (static_cast<size_t>(b) * nd + threadIdx.x) * 4;
    // int(bbox * size): product in the unet dtype (bf16 rounding when round_bf16), truncation toward zero,
    // then clamp to the image (unet.py:102-108).  bf16 value x small int is exact in fp32, so the fp32 product
    // followed by one bf16 rounding equals torch's bf16 multiply.
    float px[4] = {bb[0] * static_cast<float>(W), bb[1] * static_cast<float>(H), bb[2] * static_cast<float>(W),
                   bb[3] * static_cast<float>(H)};
    int v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float f = px[i];
      if (round_bf16) f = __bfloat162float(__float2bfloat16_rn(f));
      v[i] = static_cast<int>(f);  // truncates toward zero like Python int()
    }
    box[threadIdx.x][0] = max(0, v[0]);
    box[threadIdx.x][1] = max(0, v[1]);
    // x2 / y2 are only clamped from above (unet.py:107-108) and then used as Python slice ends: a NEGATIVE end
    // counts from the far edge (sample[..., y1:y2, x1:x2] with y2 = -3 stops 3 rows before the bottom).
    box[threadIdx.x][2] = v[2] < 0 ? max(0, W + v[2]) : min(W, v[2]);
    box[threadIdx.x][3] = v[3] < 0 ? max(0, H + v[3]) : min(H, v[3]);
  }
  __syncthreads();
  const int cv = C >> 3;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(H) * W * cv;
  if (idx >= total) return;
  const int cvec = static_cast<int>(idx % cv);
  const int pix = static_cast<int>(idx / cv);
  const int y = pix / W, x = pix - y * W;
  bool inside = false;
  for (int j = 0; j < nd; ++j) inside |= (x >= box[j][0]) && (x < box[j][2]) && (y >= box[j][1]) && (y < box[j][3]);
  if (!inside) return;
  uint4* ptr = sample + (static_cast<size_t>(b) * H * W + pix) * cv + cvec;
  uint4 u = *ptr;
  const float4 e0 = __ldg(reinterpret_cast<const float4*>(emb) + 2 * cvec);
  const float4 e1 = __ldg(reinterpret_cast<const float4*>(emb) + 2 * cvec + 1);
  u.x = pack_bf16(bf16_lo(u.x) + e0.x, bf16_hi(u.x) + e0.y);
  u.y = pack_bf16(bf16_lo(u.y) + e0.z, bf16_hi(u.y) + e0.w);
  u.z = pack_bf16(bf16_lo(u.z) + e1.x, bf16_hi(u.z) + e1.y);
  u.w = pack_bf16(bf16_lo(u.w) + e1.z, bf16_hi(u.w) + e1.w);
  *ptr = u;
}

__global__ void ip_mask_kernel(const float* __restrict__ bbox, float* __restrict__ mask, int N, int Hd, int Wd,
                               int num_ips, int tokens_per_ip, int num_dummy) {
  const int b = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const uint32_t bits = ip_inside_bits(bbox + static_cast<size_t>(b) * num_ips * 4, num_ips, n, Hd, Wd);
  const int nk = num_dummy + num_ips * tokens_per_ip;
  float* out = mask + (static_cast<size_t>(b) * N + n) * nk;
  for (int j = 0; j < nk; ++j) out[j] = ip_key_open(bits, j, tokens_per_ip, num_dummy) ? 0.0f : -10000.0f;
}

}  // namespace ds

extern "C" int ds_dialog_embed_add(void* sample, const float* emb, const float* dialog_bbox, int B, int H, int W,
                                   int C, int num_dialogs, int round_bf16, void* stream) {
  using namespace ds;
  DS_REQUIRE(sample && emb && dialog_bbox, "ds_dialog_embed_add: NULL pointer");
  DS_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "ds_dialog_embed_add: bad shape (C %% 8 == 0 required)");
  DS_REQUIRE(num_dialogs >= 0 && num_dialogs <= kMaxDialogs, "ds_dialog_embed_add: num_dialogs must be in [0, %d]",
             kMaxDialogs);
  DS_REQUIRE((reinterpret_cast<uintptr_t>(sample) & 15) == 0 && (reinterpret_cast<uintptr_t>(emb) & 15) == 0,
             "ds_dialog_embed_add: sample and emb must be 16-byte aligned");
  DeviceInfo dev;
  if (!get_device(&dev)) return DS_ERR_CUDA;
  if (num_dialogs == 0) return DS_OK;
  const long long total = static_cast<long long>(H) * W * (C / 8);
  dim3 grid(static_cast<unsigned>((total + 255) / 256), B);
  dialog_embed_add_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<uint4*>(sample), emb,
                                                                               dialog_bbox, H, W, C, num_dialogs,
                                                                               round_bf16);
  DS_LAUNCH_OK("dialog_embed_add_kernel");
  return DS_OK;
}

extern "C" int ds_ip_mask(const float* bbox, float* mask, int B, int N, double aspect_ratio, int num_ips,
                          int tokens_per_ip, int num_dummy, void* stream) {
  using namespace ds;
  DS_REQUIRE(bbox && mask, "ds_ip_mask: NULL pointer");
  DS_REQUIRE(B > 0 && N > 0 && aspect_ratio > 0.0, "ds_ip_mask: bad shape");
  DS_REQUIRE(num_ips > 0 && num_ips <= kMaxIps && tokens_per_ip > 0 && num_dummy >= 0,
             "ds_ip_mask: num_ips must be in [1, %d]", kMaxIps);
  int Hd, Wd;
  if (!derive_hw(N, aspect_ratio, &Hd, &Wd)) {
    set_error("ds_ip_mask: cannot factor N=%d for aspect_ratio=%f", N, aspect_ratio);
    return DS_ERR_INVALID;
  }
  DeviceInfo dev;
  if (!get_device(&dev)) return DS_ERR_CUDA;
  dim3 grid((N + 127) / 128, B);
  ip_mask_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(bbox, mask, N, Hd, Wd, num_ips, tokens_per_ip,
                                                                      num_dummy);
  DS_LAUNCH_OK("ip_mask_kernel");
  return DS_OK;
}
