// AdamW on one element, in torch's single-tensor op order (shared by the flat AdamW kernels and the fused exchange kernel).
#pragma once
#include "common.cuh"
#include <math.h>
namespace addk {
// torch.optim.AdamW, amsgrad=False, maximize=False (torch/optim/adamw.py -> adam.py _single_tensor_adam)
struct AdamK { float lr_wd_factor, one_minus_b1, b2, one_minus_b2, step_size, bc2_sqrt, eps, grad_scale; };
__device__ __forceinline__ void adam1(const AdamK& k, float& p, float g, float& m, float& v) {
  const float grad = mul_rn(g, k.grad_scale);
  const float w = mul_rn(p, k.lr_wd_factor);                                     // param.mul_(1 - lr*wd)
  const float mi = add_rn(m, mul_rn(k.one_minus_b1, sub_rn(grad, m)));           // exp_avg.lerp_(grad, 1-b1)
  const float vi = add_rn(mul_rn(v, k.b2), mul_rn(mul_rn(k.one_minus_b2, grad), grad));  // mul_(b2).addcmul_(g, g, 1-b2)
  const float denom = add_rn(sqrtf(vi) / k.bc2_sqrt, k.eps);
  p = add_rn(w, mul_rn(-k.step_size, mi / denom));                               // addcdiv_(exp_avg, denom, -step_size)
  m = mi; v = vi;
}
// scalar prep in double exactly as torch does it on the host, then rounded once to fp32
inline AdamK adamk_host(int step, double lr, double beta1, double beta2, double eps, double weight_decay, double grad_scale) {
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  return AdamK{(float)(1.0 - lr * weight_decay), (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)(lr / bc1),
               (float)sqrt(bc2), (float)eps, (float)grad_scale};
}
}  // namespace addk
