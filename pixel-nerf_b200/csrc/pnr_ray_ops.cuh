// Per-ray stages of NeRFRenderer.forward as device functions shared by the stage kernels (pnr_stages.cu) and by the
// fused render kernel's ray-completion epilogue (pnr_field_tc.cu): stratified samples, alpha compositing,
// importance + depth-centred resampling with the sorted merge (src/render/nerf.py:98-161, 178-249, 285-295).
//
// Every multiply-add is written with explicit round-to-nearest intrinsics, so the bits do not depend on the
// translation unit's -fmad setting: the reference evaluates these stages as separate elementwise torch ops (one
// rounding per op), and reproducing that rounding keeps the importance sampler's bin edges aligned with it.
#pragma once
#include <math.h>

#include "pnr_common.cuh"

namespace pnr {

// torch.linspace(0, 1 - 1/Kc, Kc) in fp32 (symmetric two-sided formula of ATen)
__device__ __forceinline__ float lin_step_value(int k, int Kc) {
  const float step_sz = 1.0f / (float)Kc;
  const float end = 1.0f - step_sz;
  const float inc = (Kc > 1) ? end / (float)(Kc - 1) : 0.f;
  return (k < Kc / 2) ? __fmul_rn(inc, (float)k) : __fsub_rn(end, __fmul_rn(inc, (float)(Kc - 1 - k)));
}

// sample_coarse (nerf.py:98-113): z_k = near (1 - s) + far s, s = linspace_k + u / Kc
__device__ __forceinline__ float coarse_sample(float near, float far, float lin_k, float u, int Kc) {
  const float step = 1.0f / (float)Kc;
  const float s = __fadd_rn(lin_k, __fmul_rn(u, step));
  return __fadd_rn(__fmul_rn(near, __fsub_rn(1.0f, s)), __fmul_rn(far, s));
}

// Alpha compositing of one ray (nerf.py:178-182, 222-249), sequential transmittance product like torch.cumprod on the
// CPU.  zr [K], fr [K] = (sigmoid rgb, relu sigma); w_out may be NULL.  ldf / ldf4 abstract the load flavour (the
// fused kernel reads values other CTAs wrote during the same launch and must bypass L1).
struct LdPlain {
  __device__ __forceinline__ float operator()(const float* q) const { return *q; }
};
struct LdPlain4 {
  __device__ __forceinline__ float4 operator()(const float4* q) const { return *q; }
};
template <typename LoadF, typename LoadF4>
__device__ __forceinline__ void composite_ray(const float* zr, const float4* fr, float far, int K, int white,
                                              float* w_out, float* rgb_out, float* depth_out, LoadF ldf, LoadF4 ldf4) {
  float T = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f, cd = 0.f, wsum = 0.f;
  float zk = ldf(zr);
  for (int k = 0; k < K; ++k) {
    const float znext = (k + 1 < K) ? ldf(zr + k + 1) : far;
    const float delta = __fsub_rn(znext, zk);
    const float4 f = ldf4(fr + k);
    const float sigma = fmaxf(f.w, 0.f);
    const float alpha = __fsub_rn(1.0f, expf(__fmul_rn(-delta, sigma)));
    const float w = __fmul_rn(alpha, T);
    T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f));
    cr = __fadd_rn(cr, __fmul_rn(w, f.x));
    cg = __fadd_rn(cg, __fmul_rn(w, f.y));
    cb = __fadd_rn(cb, __fmul_rn(w, f.z));
    cd = __fadd_rn(cd, __fmul_rn(w, zk));
    wsum = __fadd_rn(wsum, w);
    if (w_out) w_out[k] = w;
    zk = znext;
  }
  if (white) {
    cr = __fsub_rn(__fadd_rn(cr, 1.0f), wsum);
    cg = __fsub_rn(__fadd_rn(cg, 1.0f), wsum);
    cb = __fsub_rn(__fadd_rn(cb, 1.0f), wsum);
  }
  rgb_out[0] = cr;
  rgb_out[1] = cg;
  rgb_out[2] = cb;
  depth_out[0] = cd;
}

// sample_fine + sample_fine_depth + cat + sort (nerf.py:120-161, 285-295) for one ray by one warp.
//   zc, wc [Kc]: the coarse samples and weights; dc: coarse depth; u, uj [Kf-Kfd]; nd [Kfd]; zout [Kc+Kf] ascending.
//   scratch: Kc + 1 + Kc + Kf floats of shared memory private to the warp.
// cdf is accumulated sequentially (torch.cumsum order on CPU).
template <typename LoadF>
__device__ __forceinline__ void sample_fine_ray(float near, float far, const float* zc, const float* wc, float dc,
                                                const float* u, const float* uj, const float* nd, float depth_std,
                                                float* zout, int Kc, int Kf, int Kfd, float* scratch, int lane,
                                                LoadF ldf) {
  const int K = Kc + Kf, Ku = Kf - Kfd;
  float* cdf = scratch;            // [Kc+1]
  float* zs = scratch + (Kc + 1);  // [K]
  for (int k = lane; k < Kc; k += 32) zs[k] = ldf(zc + k);
  if (Ku > 0) {
    // pdf = (w + 1e-5) / sum ; cdf = [0, cumsum(pdf)]
    float part = 0.f;
    for (int k = lane; k < Kc; k += 32) part = __fadd_rn(part, __fadd_rn(ldf(wc + k), 1e-5f));
    // torch.sum order is not sequential either; use a fixed tree so results are deterministic
    for (int o = 16; o > 0; o >>= 1) part = __fadd_rn(part, __shfl_xor_sync(0xffffffffu, part, o));
    const float total = part;
    if (lane == 0) {
      float acc = 0.f;
      cdf[0] = 0.f;
      for (int k = 0; k < Kc; ++k) {
        acc = __fadd_rn(acc, __fdiv_rn(__fadd_rn(ldf(wc + k), 1e-5f), total));
        cdf[k + 1] = acc;
      }
    }
    __syncwarp();
    for (int j = lane; j < Ku; j += 32) {
      const float uu = u[j];
      // searchsorted(cdf, u, right=True): number of entries <= u  (cdf is non-decreasing)
      int lo = 0, hi = Kc + 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf[mid] <= uu) lo = mid + 1; else hi = mid;
      }
      const float ind = fmaxf((float)lo - 1.0f, 0.f);
      const float s = __fdiv_rn(__fadd_rn(ind, uj[j]), (float)Kc);
      zs[Kc + j] = __fadd_rn(__fmul_rn(near, __fsub_rn(1.0f, s)), __fmul_rn(far, s));
    }
  }
  if (Kfd > 0) {
    for (int j = lane; j < Kfd; j += 32) {
      const float zz = __fadd_rn(dc, __fmul_rn(nd[j], depth_std));
      zs[Kc + Ku + j] = fmaxf(fminf(zz, far), near);
    }
  }
  __syncwarp();
  // rank sort (values only matter; ties broken by index)
  for (int i = lane; i < K; i += 32) {
    const float v = zs[i];
    int rank = 0;
    for (int j = 0; j < K; ++j) {
      const float o = zs[j];
      rank += (o < v) || (o == v && j < i);
    }
    zout[rank] = v;
  }
  __syncwarp();
}

}  // namespace pnr
