// Native (torch-free) check of ds_attention_self / ds_resampler_attn / ds_attention_cross_ip through the C ABI
// against a double-precision CPU restatement.  Usage: test_attn <case-id>.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "dsengine.h"

static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  uint32_t r = u + 0x7FFF + ((u >> 16) & 1);
  return (uint16_t)(r >> 16);
}
static float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static uint32_t rng_state = 777;
static uint32_t urand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return rng_state >> 4;
}
static float frand() { return (urand() & 0xFFFFFF) / 8388608.0f - 1.0f; }

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)

enum Kind { SELF, RESAMPLER, CROSS };
struct Case {
  const char* name;
  Kind kind;
  int B, N, heads, Nkv;  // self: Nkv ignored; resampler: N = nq
  float amp;             // input amplitude (larger -> peakier softmax, exercises rescaling)
  int rows_checked;      // 0 = all rows
  int time_it;
  int Hf, Wf;            // cross: feature map (N = Hf*Wf), aspect = Hf/Wf
};
static const Case cases[] = {
    {"self_1x128_h1", SELF, 1, 128, 1, 0, 1.0f, 0, 0, 0, 0},
    {"self_1x256_h1_twotiles", SELF, 1, 256, 1, 0, 1.0f, 0, 0, 0, 0},
    {"self_2x200_h2_tail", SELF, 2, 200, 2, 0, 1.0f, 0, 0, 0, 0},
    {"self_1x1024_h2_peaky", SELF, 1, 1024, 2, 0, 4.0f, 64, 0, 0, 0},
    {"self_2x77_h1_small", SELF, 2, 77, 1, 0, 2.0f, 0, 0, 0, 0},
    {"self_1x2048_h1_verypeaky_rescale", SELF, 1, 2048, 1, 0, 8.0f, 96, 0, 0, 0},
    {"self_1x300_h1_peaky_tail", SELF, 1, 300, 1, 0, 6.0f, 0, 0, 0, 0},
    {"resampler_4x16_kv274_h2", RESAMPLER, 4, 16, 2, 274, 1.5f, 0, 0, 0, 0},
    {"cross_2x(12x20)_h2", CROSS, 2, 240, 2, 0, 1.0f, 0, 0, 12, 20},
    {"cross_2x(32x32)_h1", CROSS, 2, 1024, 1, 0, 2.0f, 0, 0, 32, 32},
    {"cross_2x264_derived24x11_quirk", CROSS, 2, 264, 1, 0, 1.0f, 0, 0, 44, 23},
    {"T_self_8x4096_h10", SELF, 8, 4096, 10, 0, 1.0f, 48, 1, 0, 0},
    {"T_self_8x1024_h20", SELF, 8, 1024, 20, 0, 1.0f, 48, 1, 0, 0},
    {"T_cross_8x(64x64)_h10", CROSS, 8, 4096, 10, 0, 1.0f, 48, 1, 64, 64},
    {"T_cross_8x(32x32)_h20", CROSS, 8, 1024, 20, 0, 1.0f, 48, 1, 32, 32},
};
static const int num_cases = sizeof(cases) / sizeof(cases[0]);

// ---- CPU restatement of the reference mask (attention_processor.py:131-167)
static void derive_hw(int N, double ar, int* H, int* W) {
  long long w = (long long)pow((double)N / ar, 0.5), h = N / w;
  while (w * h != N) {
    if (w * h < N) w += 1; else w -= 1;
    h = N / w;
  }
  *H = (int)h;
  *W = (int)w;
}
static float lin01(int i, int steps) {
  if (steps <= 1) return 0.f;
  volatile float step = 1.0f / (float)(steps - 1);
  if (i < steps / 2) {
    volatile float r = step * (float)i;
    return r;
  }
  return fmaf(-step, (float)(steps - 1 - i), 1.0f);
}

int main(int argc, char** argv) {
  if (argc < 2) {
    printf("%d\n", num_cases);
    return 0;
  }
  const int id = atoi(argv[1]);
  if (id < 0 || id >= num_cases) return 1;
  const Case c = cases[id];
  rng_state = 4242 + id;
  const int D = 64, C = c.heads * D;
  const int n_text = 77, n_ip = 80, num_ips = 4, tpi = 16, ndummy = 16;

  std::vector<uint16_t> hQ, hKV, hKVip;  // SELF: hQ is the fused qkv
  int Nq = c.N, Nkv = c.kind == RESAMPLER ? c.Nkv : c.N;
  if (c.kind == SELF) {
    hQ.resize((size_t)c.B * c.N * 3 * C);
    for (auto& v : hQ) v = f2bf(frand() * c.amp);
  } else if (c.kind == RESAMPLER) {
    hQ.resize((size_t)c.B * Nq * C);
    hKV.resize((size_t)c.B * Nkv * 2 * C);
    for (auto& v : hQ) v = f2bf(frand() * c.amp);
    for (auto& v : hKV) v = f2bf(frand() * c.amp);
  } else {
    hQ.resize((size_t)c.B * c.N * C);
    hKV.resize((size_t)c.B * n_text * 2 * C);
    hKVip.resize((size_t)c.B * n_ip * 2 * C);
    for (auto& v : hQ) v = f2bf(frand() * c.amp);
    for (auto& v : hKV) v = f2bf(frand() * c.amp);
    for (auto& v : hKVip) v = f2bf(frand() * c.amp);
  }
  // boxes: batch 0 = all zero (CFG-negative branch), others = 3 characters + one padded box
  std::vector<float> hBox((size_t)c.B * num_ips * 4, 0.f);
  const float boxes[4][4] = {{.05f, .10f, .50f, .95f}, {.50f, .15f, .95f, .90f}, {.30f, .55f, .70f, 1.0f}, {0, 0, 0, 0}};
  for (int b = 1; b < c.B; ++b)
    for (int i = 0; i < num_ips; ++i)
      for (int k = 0; k < 4; ++k) hBox[((size_t)b * num_ips + i) * 4 + k] = boxes[i][k];
  const float ip_scale = 0.6f;
  const double aspect = c.kind == CROSS ? (double)c.Hf / (double)c.Wf : 1.0;

  void *dQ, *dKV = nullptr, *dKVip = nullptr, *dOut;
  float* dBox = nullptr;
  const size_t out_elems = (size_t)c.B * Nq * C;
  CK(cudaMalloc(&dQ, hQ.size() * 2));
  CK(cudaMemcpy(dQ, hQ.data(), hQ.size() * 2, cudaMemcpyHostToDevice));
  if (!hKV.empty()) {
    CK(cudaMalloc(&dKV, hKV.size() * 2));
    CK(cudaMemcpy(dKV, hKV.data(), hKV.size() * 2, cudaMemcpyHostToDevice));
  }
  if (!hKVip.empty()) {
    CK(cudaMalloc(&dKVip, hKVip.size() * 2));
    CK(cudaMemcpy(dKVip, hKVip.data(), hKVip.size() * 2, cudaMemcpyHostToDevice));
  }
  CK(cudaMalloc(&dBox, hBox.size() * 4));
  CK(cudaMemcpy(dBox, hBox.data(), hBox.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dOut, out_elems * 2));
  CK(cudaMemset(dOut, 0xFF, out_elems * 2));

  auto run = [&]() -> int {
    if (c.kind == SELF) return ds_attention_self(dQ, dOut, c.B, c.N, c.heads, nullptr);
    if (c.kind == RESAMPLER) return ds_resampler_attn(dQ, dKV, dOut, c.B, Nq, Nkv, c.heads, nullptr);
    ds_cross_ip_args a;
    memset(&a, 0, sizeof(a));
    a.q = dQ;
    a.kv_text = dKV;
    a.kv_ip = dKVip;
    a.bbox = dBox;
    a.out = dOut;
    a.B = c.B;
    a.N = c.N;
    a.heads = c.heads;
    a.n_text = n_text;
    a.n_ip = n_ip;
    a.num_ips = num_ips;
    a.tokens_per_ip = tpi;
    a.num_dummy = ndummy;
    a.aspect_ratio = aspect;
    a.ip_scale = ip_scale;
    return ds_attention_cross_ip(&a, nullptr);
  };
  int rc = run();
  if (rc != DS_OK) {
    printf("CASE %d %s FAIL rc=%d err=%s\n", id, c.name, rc, ds_last_error());
    return 1;
  }
  cudaError_t se = cudaDeviceSynchronize();
  if (se != cudaSuccess) {
    printf("CASE %d %s FAIL sync: %s\n", id, c.name, cudaGetErrorString(se));
    return 1;
  }
  std::vector<uint16_t> hOut(out_elems);
  CK(cudaMemcpy(hOut.data(), dOut, out_elems * 2, cudaMemcpyDeviceToHost));

  int Hd = 1, Wd = 1;
  if (c.kind == CROSS) derive_hw(c.N, aspect, &Hd, &Wd);

  // softmax(q k^T / 8 + mask) v for one (b, h, row) over a key set; returns 64 outputs
  auto attend = [&](const uint16_t* qrow, const uint16_t* kbase, const uint16_t* vbase, int ld, int nk,
                    const double* addmask, double* out64) {
    std::vector<double> s(nk);
    double mx = -1e300;
    for (int j = 0; j < nk; ++j) {
      double acc = 0;
      for (int d = 0; d < D; ++d) acc += (double)bf2f(qrow[d]) * (double)bf2f(kbase[(size_t)j * ld + d]);
      s[j] = acc * 0.125 + (addmask ? addmask[j] : 0.0);
      if (s[j] > mx) mx = s[j];
    }
    double l = 0;
    for (int j = 0; j < nk; ++j) {
      s[j] = exp(s[j] - mx);
      l += s[j];
    }
    for (int d = 0; d < D; ++d) {
      double acc = 0;
      for (int j = 0; j < nk; ++j) acc += s[j] * (double)bf2f(vbase[(size_t)j * ld + d]);
      out64[d] = acc / l;
    }
  };

  double max_err = 0, max_ref = 0;
  long long bad = 0, checked = 0;
  int shown = 0;
  const long long total_rows = (long long)c.B * c.heads * Nq;
  const long long nrows = c.rows_checked ? c.rows_checked * (long long)c.B : total_rows;
  for (long long t = 0; t < nrows; ++t) {
    long long ridx = c.rows_checked ? (long long)(urand() % total_rows) : t;
    if (c.rows_checked && t < 4) ridx = (t & 1) ? total_rows - 1 - t : t;  // include first/last rows
    const int row = (int)(ridx % Nq);
    const int h = (int)((ridx / Nq) % c.heads);
    const int b = (int)(ridx / ((long long)Nq * c.heads));
    double ref[64];
    if (c.kind == SELF) {
      const uint16_t* base = &hQ[(size_t)b * c.N * 3 * C];
      attend(base + (size_t)row * 3 * C + h * D, base + C + h * D, base + 2 * C + h * D, 3 * C, c.N, nullptr, ref);
    } else if (c.kind == RESAMPLER) {
      const uint16_t* kb = &hKV[(size_t)b * Nkv * 2 * C];
      attend(&hQ[((size_t)b * Nq + row) * C + h * D], kb + h * D, kb + C + h * D, 2 * C, Nkv, nullptr, ref);
    } else {
      const uint16_t* q = &hQ[((size_t)b * c.N + row) * C + h * D];
      const uint16_t* kt = &hKV[(size_t)b * n_text * 2 * C];
      const uint16_t* ki = &hKVip[(size_t)b * n_ip * 2 * C];
      double o_t[64], o_i[64], mask[80];
      const int yi = row / Wd, xi = row % Wd;
      const float x = lin01(xi, Wd), y = lin01(yi, Hd);
      bool any = false, in[4];
      for (int i = 0; i < num_ips; ++i) {
        const float* bb = &hBox[((size_t)b * num_ips + i) * 4];
        in[i] = x >= bb[0] && x <= bb[2] && y >= bb[1] && y <= bb[3];
        any |= in[i];
      }
      for (int j = 0; j < n_ip; ++j) mask[j] = (j < ndummy ? !any : in[(j - ndummy) / tpi]) ? 0.0 : -10000.0;
      attend(q, kt + h * D, kt + C + h * D, 2 * C, n_text, nullptr, o_t);
      attend(q, ki + h * D, ki + C + h * D, 2 * C, n_ip, mask, o_i);
      for (int d = 0; d < D; ++d) ref[d] = o_t[d] + (double)ip_scale * o_i[d];
    }
    for (int d = 0; d < D; ++d) {
      const double g = bf2f(hOut[((size_t)b * Nq + row) * C + h * D + d]);
      const double err = fabs(g - ref[d]);
      const double tol = 0.02 * c.amp + 0.02 * fabs(ref[d]);  // P and O are rounded to bf16 (2^-8 relative)
      if (!(err <= tol)) {
        ++bad;
        if (shown < 8) {
          printf("  mismatch b=%d h=%d row=%d d=%d ref=%.5f got=%.5f\n", b, h, row, d, ref[d], g);
          ++shown;
        }
      }
      if (err > max_err) max_err = err;
      if (fabs(ref[d]) > max_ref) max_ref = fabs(ref[d]);
      ++checked;
    }
  }

  double ms = 0, tflops = 0;
  if (c.time_it && bad == 0) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) run();
    CK(cudaDeviceSynchronize());
    const int iters = 10;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) run();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float t;
    CK(cudaEventElapsedTime(&t, e0, e1));
    ms = t / iters;
    const double nk = c.kind == CROSS ? (n_text + n_ip) : Nkv;
    tflops = 4.0 * c.B * c.heads * (double)Nq * nk * D / (ms * 1e-3) / 1e12;
  }
  printf("CASE %d %s %s checked=%lld bad=%lld max_err=%.4g max_ref=%.4g", id, c.name, bad == 0 ? "PASS" : "FAIL",
         checked, bad, max_err, max_ref);
  if (ms > 0) printf(" ms=%.4f TFLOPs=%.1f", ms, tflops);
  printf("\n");
  return bad == 0 ? 0 : 1;
}
