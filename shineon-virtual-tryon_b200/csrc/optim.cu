// Fused Adam over flat fp32 buffers (reference: torch.optim.Adam built at models/base_model.py:165-168, lr 1e-4,
// default betas/eps, no weight decay, no amsgrad).  One launch updates every parameter of the model; `grad_scale`
// folds the 1/world_size of the data-parallel gradient mean (row U7) and any loss-scale into the same pass.
#include "common.cuh"

namespace shineon {

__global__ void __launch_bounds__(256)
    adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                     long n, float lr, float beta1, float beta2, float eps, float weight_decay, float bc1, float bc2_sqrt,
                     float grad_scale) {
  pdl_grid_sync();
  const float step_size = lr / bc1;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float grad = g[i] * grad_scale;
    const float param = p[i];
    if (weight_decay != 0.f) grad = fmaf(weight_decay, param, grad);
    const float mi = fmaf(beta1, m[i], (1.f - beta1) * grad);           // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = fmaf(beta2, v[i], (1.f - beta2) * grad * grad);    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = param - step_size * (mi / denom);
  }
}

}  // namespace shineon

using namespace shineon;

extern "C" int shineon_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long n, float lr,
                                 float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                                 shineon_stream_t stream) {
  SHINEON_REQUIRE(param && grad && exp_avg && exp_avg_sq, "adam_step: null pointer");
  SHINEON_REQUIRE(n >= 0 && step >= 1 && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f, "adam_step: bad hyper-parameters");
  if (n == 0) return SHINEON_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  klaunch(adam_step_kernel, (int)blocks, 256, 0, (cudaStream_t)stream, param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                 weight_decay, (float)bc1, (float)sqrt(bc2), grad_scale);
  return after_launch("adam_step_kernel");
}
