// Host-only: the base samples of Agent.random_vector_within_bounds (src/agent.py:76-104) drawn from torch's CPU generator
// STREAM-IDENTICALLY, without the Python loop.
//
// The reference draws ONE candidate per iteration -- w = torch.normal(0, 1, size=(1, g_ny, H, T)) in float64 from the default
// CPU generator -- keeps it iff every |w| <= beta, and repeats until it has ns candidates, for every (MPC step, SQP
// iteration): 400 000 Python iterations with a growing torch.cat for the car rollout.  A batched
// torch.normal is NOT the same stream: a candidate of n >= 16 scalars goes through ATen's normal_fill (n uniforms, Box-Muller
// in chunks of 16, and 16 MORE uniforms to recompute the tail when n % 16 != 0), a smaller one through the scalar
// at::normal_distribution<double> (two uniforms per pair, the second sample cached in the generator across calls).  So this
// file restates the generator itself -- mt19937, random64, uniform_real<double>, both normal paths -- operating directly on
// the bytes of torch.get_rng_state(); the caller writes the advanced state back with torch.set_rng_state(), which leaves the
// generator exactly where the reference's loop would have left it.  libm is the same shared object ATen calls
// (std::log / std::log1p / std::sin / std::cos / std::sqrt on doubles), so the values are bit-identical, which
// tests/test_base_samples.py checks against the loop itself.
//
// State layout (ATen/CPUGeneratorImpl.cpp, CPUGeneratorImplState; 5056 bytes):
//   u64 seed | i32 left | i32 seeded | u64 next | u64 state[624] | f64 normal_x | f64 normal_y | f64 normal_rho |
//   i32 normal_is_valid (+4 pad) | f32 next_float_normal_sample | bool is_next_float_normal_sample_valid (+3 pad)
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace gpmpc_rng {

constexpr int MT_N = 624, MT_M = 397;
constexpr size_t STATE_BYTES = 5056;
constexpr size_t OFF_LEFT = 8, OFF_NEXT = 16, OFF_STATE = 24, OFF_NORMAL_Y = OFF_STATE + 8 * MT_N + 8,
                 OFF_NORMAL_VALID = OFF_STATE + 8 * MT_N + 24;

struct Engine {
  uint32_t state[MT_N];
  int left;
  uint64_t next;
  bool has_cached;
  double cached;

  void load(const uint8_t* b) {
    int32_t l;
    std::memcpy(&l, b + OFF_LEFT, 4);
    left = l;
    std::memcpy(&next, b + OFF_NEXT, 8);
    for (int i = 0; i < MT_N; ++i) {
      uint64_t v;
      std::memcpy(&v, b + OFF_STATE + 8 * (size_t)i, 8);
      state[i] = (uint32_t)v;
    }
    int32_t valid;
    std::memcpy(&valid, b + OFF_NORMAL_VALID, 4);
    has_cached = valid != 0;
    std::memcpy(&cached, b + OFF_NORMAL_Y, 8);
  }
  void store(uint8_t* b) const {
    const int32_t l = left;
    std::memcpy(b + OFF_LEFT, &l, 4);
    std::memcpy(b + OFF_NEXT, &next, 8);
    for (int i = 0; i < MT_N; ++i) {
      const uint64_t v = state[i];
      std::memcpy(b + OFF_STATE + 8 * (size_t)i, &v, 8);
    }
    const int32_t valid = has_cached ? 1 : 0;
    const double y = has_cached ? cached : 0.0;
    std::memcpy(b + OFF_NORMAL_VALID, &valid, 4);
    std::memcpy(b + OFF_NORMAL_Y, &y, 8);
  }

  static uint32_t twist(uint32_t u, uint32_t v) {
    return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
  }
  void next_state() {  // at::mt19937::next_state
    uint32_t* p = state;
    left = MT_N;
    next = 0;
    for (int j = MT_N - MT_M + 1; --j; p++) *p = p[MT_M] ^ twist(p[0], p[1]);
    for (int j = MT_M; --j; p++) *p = p[MT_M - MT_N] ^ twist(p[0], p[1]);
    *p = p[MT_M - MT_N] ^ twist(p[0], state[0]);
  }
  uint32_t u32() {  // at::mt19937::operator()
    if (--left == 0) next_state();
    uint32_t y = state[next++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  uint64_t u64() {  // CPUGeneratorImpl::random64: hi word first
    const uint32_t hi = u32(), lo = u32();
    return ((uint64_t)hi << 32) | lo;
  }
  double uniform() {  // at::uniform_real_distribution<double>(0, 1): 53 random bits
    constexpr uint64_t MASK = (1ULL << 53) - 1;
    constexpr double DIVISOR = 1.0 / (double)(1ULL << 53);
    return (double)(u64() & MASK) * DIVISOR * (1.0 - 0.0) + 0.0;
  }
  double normal_scalar() {  // at::normal_distribution<double>(0, 1) with the generator's cached second sample
    if (has_cached) {
      has_cached = false;
      return cached * 1.0 + 0.0;
    }
    const double u1 = uniform(), u2 = uniform();
    const double r = std::sqrt(-2.0 * std::log1p(-u2));
    const double theta = 2.0 * M_PI * u1;
    cached = r * std::sin(theta);
    has_cached = true;
    return r * std::cos(theta) * 1.0 + 0.0;
  }
  static void fill16(double* d) {  // ATen normal_fill_16<double>, mean 0, std 1
    for (int j = 0; j < 8; ++j) {
      const double u1 = 1 - d[j];
      const double u2 = d[j + 8];
      const double radius = std::sqrt(-2 * std::log(u1));
      const double theta = 2.0f * M_PI * u2;
      d[j] = radius * std::cos(theta) * 1.0 + 0.0;
      d[j + 8] = radius * std::sin(theta) * 1.0 + 0.0;
    }
  }
  // one torch.normal(0, 1, size = n doubles) call
  void normal_call(double* out, int64_t n) {
    if (n >= 16) {  // normal_fill
      for (int64_t i = 0; i < n; ++i) out[i] = uniform();
      for (int64_t i = 0; i < n - 15; i += 16) fill16(out + i);
      if (n % 16 != 0) {
        double* tail = out + n - 16;
        for (int i = 0; i < 16; ++i) tail[i] = uniform();
        fill16(tail);
      }
    } else {
      for (int64_t i = 0; i < n; ++i) out[i] = normal_scalar();
    }
  }
};

// slots candidates of n doubles each, every one redrawn until all |w| <= beta; returns the number of torch.normal calls
inline int64_t truncated_candidates(uint8_t* rng_state, int64_t slots, int64_t n, double beta, double* out) {
  Engine e;
  e.load(rng_state);
  int64_t calls = 0;
  for (int64_t s = 0; s < slots; ++s) {
    double* w = out + s * n;
    for (;;) {
      e.normal_call(w, n);
      ++calls;
      bool ok = true;
      for (int64_t i = 0; i < n; ++i) ok = ok && (w[i] >= -beta) && (w[i] <= beta);
      if (ok) break;
    }
  }
  e.store(rng_state);
  return calls;
}

}  // namespace gpmpc_rng
