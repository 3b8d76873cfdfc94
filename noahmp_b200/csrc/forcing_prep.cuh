// noahmp_b200 -- forcing preparation: the caller-side step just before the column physics.
//
// The reference's namelist carries `input_frequency` and `interval_seconds` (offline/noahmp_config.py:81-82,98)
// but nothing interpolates forcing records to the model step or computes the cosine of the solar zenith angle
// that noahmp_sflx takes as an input (func:67, COSZ).  This kernel does both in one pass over HBM:
//   out[0..7] = rec0 + w * (rec1 - rec0)                      (SFCTMP SFCPRS UU VV Q2 SOLDN LWDN PRCP)
//   out[8]    = sin(lat) sin(decl) + cos(lat) cos(decl) cos(H0 + lon)       with cos(H0+lon) expanded
// Every transcendental is evaluated on the host (per step: decl, H0; per column once: sin/cos of lat, lon), so
// the device arithmetic is +,-,* with explicit round-to-nearest intrinsics and is bit-identical to the numpy
// statement in noahmp_b200/forcing.py regardless of compiler flags.
//
// Pure streaming: 29 words per column (16 read + 4 geo + 9 written), 128-bit accesses, no reuse -> HBM bound.
#pragma once
#include "common.cuh"

namespace nmp {

struct PrepScalars {
  float w;                    // interpolation weight of rec1
  float sd, cd, ch0, sh0;     // sin/cos of the declination, cos/sin of the hour angle at longitude 0
};

__device__ __forceinline__ float lerp_rn(float a, float b, float w) { return __fadd_rn(a, __fmul_rn(w, __fsub_rn(b, a))); }
__device__ __forceinline__ float cosz_rn(float sinlat, float coslat, float sinlon, float coslon, const PrepScalars& p) {
  const float ch = __fsub_rn(__fmul_rn(p.ch0, coslon), __fmul_rn(p.sh0, sinlon));
  // clamped: fp32 rounding can overshoot 1 by an ulp and twostream takes acos(COSZ) (func:2315)
  return fminf(fmaxf(__fadd_rn(__fmul_rn(sinlat, p.sd), __fmul_rn(__fmul_rn(coslat, p.cd), ch)), -1.0f), 1.0f);
}

// VEC = 4: ncol % 4 == 0 and 16-byte aligned planes; VEC = 1: anything
template <int VEC>
__global__ void __launch_bounds__(256) prepare_forcing_kernel(int64_t ncol, const float* __restrict__ rec0,
                                                              const float* __restrict__ rec1,
                                                              const float* __restrict__ geo, PrepScalars p,
                                                              float* __restrict__ out) {
  const int64_t nvec = ncol / VEC;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    if (VEC == 4) {
      const float4* g = reinterpret_cast<const float4*>(geo);
      const float4 sl = __ldg(g + i), cl = __ldg(g + nvec + i), sn = __ldg(g + 2 * nvec + i), cn = __ldg(g + 3 * nvec + i);
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const float4 a = __ldcs(reinterpret_cast<const float4*>(rec0) + v * nvec + i);
        const float4 b = __ldcs(reinterpret_cast<const float4*>(rec1) + v * nvec + i);
        float4 o;
        o.x = lerp_rn(a.x, b.x, p.w); o.y = lerp_rn(a.y, b.y, p.w); o.z = lerp_rn(a.z, b.z, p.w); o.w = lerp_rn(a.w, b.w, p.w);
        reinterpret_cast<float4*>(out)[v * nvec + i] = o;
      }
      float4 z;
      z.x = cosz_rn(sl.x, cl.x, sn.x, cn.x, p); z.y = cosz_rn(sl.y, cl.y, sn.y, cn.y, p);
      z.z = cosz_rn(sl.z, cl.z, sn.z, cn.z, p); z.w = cosz_rn(sl.w, cl.w, sn.w, cn.w, p);
      reinterpret_cast<float4*>(out)[8 * nvec + i] = z;
    } else {
#pragma unroll
      for (int v = 0; v < 8; ++v) out[v * ncol + i] = lerp_rn(rec0[v * ncol + i], rec1[v * ncol + i], p.w);
      out[8 * ncol + i] = cosz_rn(geo[i], geo[ncol + i], geo[2 * ncol + i], geo[3 * ncol + i], p);
    }
  }
}

}  // namespace nmp
