// Fused multi-tensor Adam with coupled L2 weight decay over the flat gradient buffer:
// what `torch.optim.Adam(model.parameters(), lr, weight_decay)` (train.py:134, step at train.py:200) does
// for the 1960 trainable tensors of the stack, as ONE launch that also clears the gradients.
//
// Parameters stay the caller's separate tensors; gradients and both moment buffers share one flat layout
// (the gradient views installed by the engine), so a segment is (param pointer, flat offset, length).
#include <vector>
#include "kernels.h"

namespace {

struct AdamItem { float* p; int64_t off; int32_t n; int32_t pad; };   // <= kChunk elements of one tensor
constexpr int kChunk = 8192;
constexpr int kThreadsAdam = 256;

struct AdamHyper {
  float beta1, beta2, eps, weight_decay;
  float step_size;            // lr / (1 - beta1^t)
  float inv_bc2_sqrt;         // 1 / sqrt(1 - beta2^t)
  int zero_grad;
};

__global__ void __launch_bounds__(kThreadsAdam) adam_kernel(const AdamItem* __restrict__ items, float* __restrict__ grad,
                                                           float* __restrict__ m, float* __restrict__ v, AdamHyper h) {
  const AdamItem it = items[blockIdx.x];
  float* __restrict__ p = it.p;
  float* __restrict__ g = grad + it.off;
  float* __restrict__ mm = m + it.off;
  float* __restrict__ vv = v + it.off;
  constexpr int U = 4;
  for (int i0 = threadIdx.x; i0 < it.n; i0 += U * kThreadsAdam) {
    float pv[U], gv[U], mv[U], sv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * kThreadsAdam;
      if (i < it.n) { pv[u] = p[i]; gv[u] = g[i]; mv[u] = mm[i]; sv[u] = vv[i]; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * kThreadsAdam;
      if (i >= it.n) break;
      // torch/optim/adam.py (_single_tensor_adam / fused kernel): coupled L2, lerp form of the first moment
      const float gr = fmaf(h.weight_decay, pv[u], gv[u]);
      const float m1 = mv[u] + (1.f - h.beta1) * (gr - mv[u]);
      const float v1 = h.beta2 * sv[u] + (1.f - h.beta2) * gr * gr;
      const float denom = sqrtf(v1) * h.inv_bc2_sqrt + h.eps;
      p[i] = pv[u] - h.step_size * (m1 / denom);
      mm[i] = m1;
      vv[i] = v1;
      if (h.zero_grad) g[i] = 0.f;
    }
  }
}

}  // namespace

struct rnvp_adam {
  AdamItem* d_items = nullptr;
  int n_items = 0;
  int64_t total = 0;
};

extern "C" {

int rnvp_adam_create(void* const* params_host, const int64_t* offsets_host, const int64_t* sizes_host, int n,
                     rnvp_adam** out) {
  RNVP_REQUIRE(params_host && offsets_host && sizes_host && out && n > 0, "rnvp_adam_create: bad arguments");
  std::vector<AdamItem> items;
  int64_t total = 0;
  for (int i = 0; i < n; ++i) {
    RNVP_REQUIRE(params_host[i] != nullptr && sizes_host[i] >= 0 && offsets_host[i] >= 0, "rnvp_adam_create: bad segment %d", i);
    for (int64_t s = 0; s < sizes_host[i]; s += kChunk) {
      int64_t c = sizes_host[i] - s < kChunk ? sizes_host[i] - s : kChunk;
      items.push_back(AdamItem{reinterpret_cast<float*>(params_host[i]) + s, offsets_host[i] + s, (int32_t)c, 0});
    }
    total = offsets_host[i] + sizes_host[i] > total ? offsets_host[i] + sizes_host[i] : total;
  }
  rnvp_adam* a = new rnvp_adam();
  a->n_items = (int)items.size();
  a->total = total;
  if (cudaMalloc(&a->d_items, items.size() * sizeof(AdamItem)) != cudaSuccess ||
      cudaMemcpy(a->d_items, items.data(), items.size() * sizeof(AdamItem), cudaMemcpyHostToDevice) != cudaSuccess) {
    rnvp::set_error("rnvp_adam_create: %s", cudaGetErrorString(cudaGetLastError()));
    if (a->d_items) cudaFree(a->d_items);
    delete a;
    return RNVP_ERR_CUDA;
  }
  *out = a;
  return RNVP_OK;
}

int rnvp_adam_destroy(rnvp_adam* a) {
  if (!a) return RNVP_OK;
  if (a->d_items) cudaFree(a->d_items);
  delete a;
  return RNVP_OK;
}

int rnvp_adam_step(rnvp_adam* a, float* flat_grad, float* exp_avg, float* exp_avg_sq, int64_t flat_len, double lr,
                   double beta1, double beta2, double eps, double weight_decay, int64_t step, int zero_grad,
                   void* stream) {
  RNVP_REQUIRE(a && flat_grad && exp_avg && exp_avg_sq, "rnvp_adam_step: null argument");
  RNVP_REQUIRE(flat_len >= a->total, "rnvp_adam_step: flat buffers hold %lld elements, the plan needs %lld",
               (long long)flat_len, (long long)a->total);
  RNVP_REQUIRE(step >= 1, "rnvp_adam_step: step counts from 1");
  AdamHyper h;
  h.beta1 = (float)beta1; h.beta2 = (float)beta2; h.eps = (float)eps; h.weight_decay = (float)weight_decay;
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  h.step_size = (float)(lr / bc1);
  h.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  h.zero_grad = zero_grad;
  adam_kernel<<<a->n_items, kThreadsAdam, 0, (cudaStream_t)stream>>>(a->d_items, flat_grad, exp_avg, exp_avg_sq, h);
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

}  // extern "C"
