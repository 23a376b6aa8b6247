// Counter-based RNG and the completion-time samplers used on the device.
//
// The reference draws from three global generators (Python `random`, numpy global state,
// numba's private state; SURVEY.md section 7 "RNG streams").  None of those is reproducible on a
// GPU, so the device uses Philox4x32-10 keyed by the run seed and counted by
// (person global id, step, stream, k).  The Bernoulli stream is additionally injectable for
// bit-exact tests against the reference (jb_inject_uniforms).
//
// Samplers restate the distributions scipy provides to june/epidemiology/infection/
// trajectory_maker.py:95-132 (beta, lognorm, norm, expon, exponweib): statistical parity only.
#pragma once
#include <stdint.h>

#define JB_STREAM_BERNOULLI 0u
#define JB_STREAM_VARIANT 2u
#define JB_STREAM_LEISURE 3u   // block 0: (does any activity, which activity), block 1: which venue
#define JB_STREAM_BLAME 4u     // block 0: (which infector subgroup, which infector) of a new infection
#define JB_STREAM_COMMUTE 5u   // block 0: (which station of the city, which transport of the station)
#define JB_STREAM_VACCINE 6u   // counter word "step" = day; block 0: (gets vaccinated today, which campaign)
#define JB_STREAM_POLICY0 8u   // + index of the individual policy (two uniforms per block)
#define JB_STREAM_SAMPLE0 16u  // + sample slot (0..15) of a new infection, see k_newinf

struct JbPhilox4 {
  uint32_t x[4];
};

__host__ __device__ __forceinline__ uint32_t jb_mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

// Philox4x32-10 (Salmon et al., SC'11).  Known answer: ctr=0,key=0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8
__host__ __device__ __forceinline__ JbPhilox4 jb_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                              uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = jb_mulhi32(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = jb_mulhi32(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  JbPhilox4 out;
  out.x[0] = c0; out.x[1] = c1; out.x[2] = c2; out.x[3] = c3;
  return out;
}

// 52 random bits -> double strictly inside (0,1):  (b + 0.5) * 2^-52
__host__ __device__ __forceinline__ double jb_u52(uint32_t hi, uint32_t lo) {
  uint64_t b = (((uint64_t)hi << 32) | (uint64_t)lo) >> 12;
#ifdef __CUDA_ARCH__
  // the same value without the 64-bit integer -> double conversion: 1.b - 1 = b * 2^-52 and the sum
  // (2b + 1) * 2^-53 are both exact
  return (__longlong_as_double((long long)(0x3FF0000000000000ull | b)) - 1.0) + 1.1102230246251565e-16;
#else
  return ((double)b + 0.5) * 2.220446049250313e-16;
#endif
}

// The uniform a given person consumes for its Bernoulli draw in a given step.
__host__ __device__ __forceinline__ double jb_bernoulli_uniform(uint64_t seed, uint32_t step, uint32_t gid) {
  JbPhilox4 r = jb_philox4x32_10(gid, step, JB_STREAM_BERNOULLI, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
  return jb_u52(r.x[0], r.x[1]);
}
__host__ __device__ __forceinline__ double jb_variant_uniform(uint64_t seed, uint32_t step, uint32_t gid) {
  JbPhilox4 r = jb_philox4x32_10(gid, step, JB_STREAM_VARIANT, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
  return jb_u52(r.x[0], r.x[1]);
}

// Sequential stream of uniforms for one (person, step): two doubles per Philox block.
struct JbStream {
  uint32_t gid, step, stream, k, k0, k1;
  int have;
  double spare;
};

__host__ __device__ __forceinline__ void jb_stream_init(JbStream& s, uint64_t seed, uint32_t step,
                                                       uint32_t gid, uint32_t stream) {
  s.gid = gid; s.step = step; s.stream = stream; s.k = 0;
  s.k0 = (uint32_t)seed; s.k1 = (uint32_t)(seed >> 32);
  s.have = 0; s.spare = 0.0;
}

// (The samplers below are deliberately NOT inlined: the infection sampler runs a few thousand warps
// once per step through straight-line code, so its time is instruction fetch -- one copy of Philox
// and of each transcendental instead of a dozen.)
__host__ __device__ __noinline__ double jb_stream_next(JbStream& s) {
  if (s.have) { s.have = 0; return s.spare; }
  JbPhilox4 r = jb_philox4x32_10(s.gid, s.step, s.stream, s.k++, s.k0, s.k1);
  s.spare = jb_u52(r.x[2], r.x[3]);
  s.have = 1;
  return jb_u52(r.x[0], r.x[1]);
}

// Box-Muller (cosine branch only): consumes two uniforms.
__host__ __device__ __noinline__ double jb_normal(JbStream& s) {
  double u1 = jb_stream_next(s), u2 = jb_stream_next(s);
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

// Marsaglia & Tsang (2000) gamma(k, 1).
__host__ __device__ __noinline__ double jb_gamma(JbStream& s, double k) {
  double boost = 1.0;
  if (k < 1.0) {
    boost = pow(jb_stream_next(s), 1.0 / k);
    k += 1.0;
  }
  const double d = k - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  double out = d;
  for (int it = 0; it < 64; ++it) {
    double z = jb_normal(s);
    double v = 1.0 + c * z;
    if (v <= 0.0) continue;
    v = v * v * v;
    double u = jb_stream_next(s);
    if (log(u) < 0.5 * z * z + d - d * v + d * log(v)) { out = d * v; break; }
  }
  return out * boost;
}

// One completion time (JbDist; trajectory_maker.py:62-132).
__host__ __device__ __noinline__ double jb_sample_dist(JbStream& s, int type, const double* p) {
  switch (type) {
    case 0: return p[0];                                            // constant
    case 1: return p[0] - p[1] * log(jb_stream_next(s));            // expon(loc, scale)
    case 2: {                                                       // beta(a, b, loc, scale)
      double x = jb_gamma(s, p[0]);
      double y = jb_gamma(s, p[1]);
      return p[2] + p[3] * (x / (x + y));
    }
    case 3: return p[1] + p[2] * exp(p[0] * jb_normal(s));          // lognorm(s, loc, scale)
    case 4: return p[0] + p[1] * jb_normal(s);                      // norm(loc, scale)
    case 5: {                                                       // exponweib(a, c, loc, scale)
      double u = jb_stream_next(s);                                 // F(x) = (1 - exp(-x^c))^a
      double q = exp(log(u) / p[0]);                                // u^(1/a), may be << 1e-16
      return p[2] + p[3] * pow(-log1p(-q), 1.0 / p[1]);
    }
    default: return 0.0;
  }
}
