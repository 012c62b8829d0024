// Shared device/host helpers for librbslam (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define RB_PI 3.14159265358979323846
#define RB_LOG2PI 1.8378770664093454835606594728112

#define RB_CUDA(call)                                                          \
  do {                                                                         \
    cudaError_t err__ = (call);                                                \
    if (err__ != cudaSuccess) {                                                \
      this->fail_cuda(err__, #call, __FILE__, __LINE__);                       \
      return RBSLAM_ECUDA;                                                     \
    }                                                                          \
  } while (0)

namespace rb {

// device-side error flags written by kernels (never silently ignored)
struct DevStatus {
  int not_pd;        // !=0: innovation covariance not PD even after jitter
  int not_pd_step;   // time step of the first failure
  int not_pd_particle;
  int used_jitter;   // count of jitter retries (informational)
  int clamp_sample;  // count of u > wc(end) clamps (reference would index-error)
  int peer_timeout;  // sharded filter: a peer never reached the barrier (it failed or died)
  int scan_ambig;    // fast resampling path: a draw could depend on the rounding order of the scan -> exact path
  int clamp_fast;    // clamps counted by the fast path (added to clamp_sample when it stands)
  int scan_fallbacks;   // how often the exact path had to run (informational)
};

// ---------------------------------------------------------------------------
// Philox4x32-10, counter = (particle, step, block, sweep), key = seed.
// Restated in NumPy in oracle/streams.py (philox4x32_10) for parity.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
  return (double)((((uint64_t)(a >> 5)) << 26) + (uint64_t)(b >> 6));
}
__device__ __forceinline__ double philox_uniform(uint64_t seed, uint32_t sweep, uint32_t t,
                                                 uint32_t i) {
  uint4 x = philox4x32_10(make_uint4(i, t, 0u, sweep),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  return u53(x.x, x.y) * 0x1p-53;
}
// Box-Muller pair p (normals 2p, 2p+1) of particle i at (sweep, t)
__device__ __forceinline__ void philox_normal_pair(uint64_t seed, uint32_t sweep, uint32_t t,
                                                   uint32_t i, uint32_t p, double &z0,
                                                   double &z1) {
  uint4 x = philox4x32_10(make_uint4(i, t, 1u + p, sweep),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const double ua = (u53(x.x, x.y) + 1.0) * 0x1p-53;
  const double ub = u53(x.z, x.w) * 0x1p-53;
  const double r = sqrt(-2.0 * log(ua));
  double s, c;
  sincos(2.0 * RB_PI * ub, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

// ---------------------------------------------------------------------------
// quaternion algebra (tools/expq.m, qLeft.m, quat2rmat.m, qInv.m, logq.m)
// ---------------------------------------------------------------------------
// qLeft(p)*q: Hamilton product p (x) q, scalar first (tools/qLeft.m:30-34)
__device__ __forceinline__ void qmul(const double p[4], const double q[4], double o[4]) {
  o[0] = p[0] * q[0] - p[1] * q[1] - p[2] * q[2] - p[3] * q[3];
  o[1] = p[1] * q[0] + p[0] * q[1] - p[3] * q[2] + p[2] * q[3];
  o[2] = p[2] * q[0] + p[3] * q[1] + p[0] * q[2] - p[1] * q[3];
  o[3] = p[3] * q[0] - p[2] * q[1] + p[1] * q[2] + p[0] * q[3];
}
// tools/expq.m:22-32 (single-vector branch; argument is phi, not phi/2)
__device__ __forceinline__ void expq(const double phi[3], double eq[4]) {
  const double mag = sqrt(phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2]);
  const double den = mag + (mag == 0.0 ? 1.0 : 0.0);
  double s, c;
  sincos(mag, &s, &c);
  eq[0] = c;
  eq[1] = phi[0] / den * s;
  eq[2] = phi[1] / den * s;
  eq[3] = phi[2] / den * s;
  if (eq[0] < 0.0) {
    eq[0] = -eq[0]; eq[1] = -eq[1]; eq[2] = -eq[2]; eq[3] = -eq[3];
  }
}
// tools/logq.m:25-30; q0 is clamped to 1 (hazard Q6: MATLAB's acos would go complex)
__device__ __forceinline__ void logq(const double qin[4], double lq[3]) {
  double q[4] = {qin[0], qin[1], qin[2], qin[3]};
  if (q[0] < 0.0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
  const double q0 = fmin(q[0], 1.0);
  const double na = acos(q0);
  const double den = sin(na) + (na == 0.0 ? 1.0 : 0.0);
  lq[0] = na * q[1] / den;
  lq[1] = na * q[2] / den;
  lq[2] = na * q[3] / den;
}
// tools/quat2rmat.m:27-32, row-major R[3][3], no normalisation
__device__ __forceinline__ void quat2rmat(const double q[4], double R[3][3]) {
  const double q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
  R[0][0] = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3;
  R[0][1] = 2 * q1 * q2 - 2 * q0 * q3;
  R[0][2] = 2 * q1 * q3 + 2 * q0 * q2;
  R[1][0] = 2 * q1 * q2 + 2 * q0 * q3;
  R[1][1] = q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3;
  R[1][2] = 2 * q2 * q3 - 2 * q0 * q1;
  R[2][0] = 2 * q1 * q3 - 2 * q0 * q2;
  R[2][1] = 2 * q2 * q3 + 2 * q0 * q1;
  R[2][2] = q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3;
}

// In-place lower Cholesky of a small column-major matrix A[ld*n] (thread-serial).
// Returns 0 on success, k+1 if pivot k is <= 0 or NaN (LAPACK dpotrf semantics).
__device__ __forceinline__ int chol_small(double *A, int n, int ld) {
  for (int j = 0; j < n; ++j) {
    double s = A[j + j * ld];
    for (int k = 0; k < j; ++k) s -= A[j + k * ld] * A[j + k * ld];
    if (!(s > 0.0)) return j + 1;
    const double ljj = sqrt(s);
    A[j + j * ld] = ljj;
    for (int i = j + 1; i < n; ++i) {
      double v = A[i + j * ld];
      for (int k = 0; k < j; ++k) v -= A[i + k * ld] * A[j + k * ld];
      A[i + j * ld] = v / ljj;
    }
  }
  return 0;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace rb
