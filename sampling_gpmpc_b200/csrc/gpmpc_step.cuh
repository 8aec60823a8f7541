// K1: fused rollout step (H = 1): posterior + draw + post-processing + rank-T append in one launch.
//
// One WARP owns one batch element (sample s, output j) at a time and loops over samples (persistent CTAs:
// blockIdx.y = output j, so the warps of a CTA share L_oo, the observed real inputs and beta_o in shared
// memory).  The step is a forward substitution  w = L^{-1} k(X, x*)  against the bordered factor
// [[L_oo, 0], [V, L_hh]], done LEFT-LOOKING over SUB-PANELS of 8 rows (layout: gpmpc_state.cuh) on the FP64
// tensor cores: one mma.sync.m8n8k4.f64 multiplies an 8-row x 4-column tile of L with the 4 x T tile of w
// (T <= 7 right-hand sides = the N dimension; unused columns are don't-care), so the inner loop is two shared
// loads and one DMMA per 4 factor columns and needs no cross-lane reduction.
//
//   A  kernel vector k(X, x*): lanes over training scalars, T right-hand sides each, written to the per-warp
//      shared array wv[storage column][T]  (in place: wv holds k first, w = L^{-1} k afterwards)
//   B  the m shared rows against L_oo (shared memory, sub-panel layout)
//   C  the element's own c rows, streamed from HBM exactly once:  the element's factor is one contiguous
//      stream of sub-panels, cut into fixed chunks of STEP_SEG column groups (4 KB, sub-panel boundaries are
//      ignored) that one elected lane pulls into a per-warp ring of STEP_NST shared-memory slots with TMA bulk
//      copies (cp.async.bulk + mbarrier complete_tx).  The ring runs STEP_NST-1 chunks ahead of the consumer and ACROSS elements: while an
//      element's epilogue runs, the first chunks of the warp's next element are already in flight.
//      Per sub-panel: dot = L[rows][cols < n_off] w  (DMMA chain, two accumulator sets), rhs = k - dot, then the
//      8 x 8 diagonal block is applied as w_blk = inv(D) rhs with two more DMMAs (inv(D) is kept transposed in
//      the block's upper triangle) -- no substitution chain anywhere.
//   D  W^T [W | beta] with the same DMMA loop: Sigma* = K** - W^T W, mean = W^T beta
//   F  rank-T append: w goes to the T new rows' column groups
//   then k_step_finish (one THREAD per element): E  T x T Cholesky with the jitter ladder, draw, post-processing;
//   the diagonal-block part of the append and the touched blocks' inverses
//
// HBM traffic per element-step = its own factor read once (8-row granularity) + T new rows written once + O(c)
// inputs: HBM-bound by design (DESIGN.md "Roofline"); the host counts the algorithmic bytes per launch.
#pragma once
#include "gpmpc_state.cuh"
#include "gpmpc_eig.cuh"

#ifndef STEP_MAX_WARPS
#define STEP_MAX_WARPS 16  // warps per CTA are chosen per launch (shared-memory budget), one CTA per SM
#endif
#ifndef STEP_SEG
#define STEP_SEG 64  // 8-row column groups per TMA chunk / ring slot (64 * 64 B = 4 KB); multiple of 8
#endif
#ifndef STEP_NST
#define STEP_NST 2   // ring slots per warp: one being consumed, the other in flight
#endif
#define STEP_SLOT_BYTES (STEP_SEG * 64)
#ifndef GPMPC_PRED_LOADS
// B-fragment loads of the hot loops under a predicate (1) or inside a divergent branch (0).  Measured on one B200 box
// (profiles/r2_horizon_probe.txt): the branch wins -- step-wise rollout 220.4 ms vs 229.8 ms, fused horizon 254.7 vs 276.2 --
// although it costs a BSSY / BSYNC pair per iteration: the predicated form keeps the dead lanes' loads in the LSU queue
#define GPMPC_PRED_LOADS 0
#endif

// ---- TMA bulk copy + mbarrier (one ring per warp; the warp is its own producer and consumer) ----------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

template <int T>
struct TriT {  // lower-triangular T x T in registers
  double v[T * (T + 1) / 2];
  __device__ __forceinline__ double& at(int r, int s) { return v[r * (r + 1) / 2 + s]; }
};

// in-register Cholesky (every lane does the same arithmetic); returns false on a non-positive / NaN pivot
template <int T>
__device__ __forceinline__ bool chol_T(const TriT<T>& S, double add, TriT<T>& L) {
  bool ok = true;
#pragma unroll
  for (int r = 0; r < T; ++r) {
#pragma unroll
    for (int s = 0; s <= r; ++s) {
      double v = S.v[r * (r + 1) / 2 + s];
      if (s == r) v += add;
#pragma unroll
      for (int k = 0; k < s; ++k) v -= L.v[r * (r + 1) / 2 + k] * L.v[s * (s + 1) / 2 + k];
      if (s == r) {
        if (!(v > 0.0)) ok = false;
        L.v[r * (r + 1) / 2 + s] = sqrt(v);
      } else {
        L.v[r * (r + 1) / 2 + s] = v / L.v[s * (s + 1) / 2 + s];
      }
    }
  }
  return ok;
}

// cov( task ta at xa , tasks 0..T-1 at xs ), r = xa - xs  
template <int D, int T>
__device__ __forceinline__ void kernel_row(const double* __restrict__ xa, int ta, const double (&xs)[D],
                                           const double (&il)[D], double os, double (&out)[T]) {
  double g[D], sq = 0.0, ga = 0.0, il2 = 0.0;
#pragma unroll
  for (int a = 0; a < D; ++a) {
    const double r = xa[a] - xs[a];
    const double t = r * il[a];
    sq = fma(t, t, sq);
    g[a] = t * il[a];  // r_a / l_a^2
    if (a == ta - 1) { ga = g[a]; il2 = __dmul_rn(il[a], il[a]); }
  }
  const double k0 = os * exp(-0.5 * sq);
  if (ta == 0) {
    out[0] = k0;
#pragma unroll
    for (int tb = 1; tb < T; ++tb) out[tb] = k0 * g[tb - 1];
  } else {
    out[0] = -k0 * ga;
#pragma unroll
    for (int tb = 1; tb < T; ++tb) {
      double h = -ga * g[tb - 1];
      if (tb == ta) h = __dadd_rn(h, il2);  // explicit: the product 1/l^2 is never fused into this sum (the same entry must
      out[tb] = k0 * h;                     // come out bit-identical from every kernel that evaluates it)
    }
  }
}

// the T x T block cov( task ta at xa , task tb at xs ), one exp per point
template <int D, int T>
__device__ __forceinline__ void kernel_block(const double (&xa)[D], const double (&xs)[D], const double (&il)[D],
                                             double os, double (&out)[T][T]) {
  double g[D], sq = 0.0;
#pragma unroll
  for (int a = 0; a < D; ++a) {
    const double t = (xa[a] - xs[a]) * il[a];
    sq = fma(t, t, sq);
    g[a] = t * il[a];  // r_a / l_a^2
  }
  const double k0 = os * exp(-0.5 * sq);
  out[0][0] = k0;
  if constexpr (T > 1) {
#pragma unroll
    for (int tb = 1; tb < T; ++tb) {
      out[0][tb] = k0 * g[tb - 1];
      out[tb][0] = -out[0][tb];
    }
#pragma unroll
    for (int ta = 1; ta < T; ++ta)
#pragma unroll
      for (int tb = 1; tb < T; ++tb) {
        double h = -g[ta - 1] * g[tb - 1];
        if (ta == tb) h = __dadd_rn(h, __dmul_rn(il[ta - 1], il[ta - 1]));  // explicit roundings, see kernel_row
        out[ta][tb] = k0 * h;
      }
  }
}

// ---- explicit shared-space accesses (32-bit addresses, immediate offsets) for the hot loops -----------------
template <int OFF = 0>
__device__ __forceinline__ double lds(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF) : "memory");
  return v;
}
// four loads at addr + {0, S, 2S, 3S} under ONE predicate (lanes with live == 0 issue no shared-memory access and keep their
// register values): predication instead of a divergent branch -- an `if (live)` around the loads costs a BSSY / BSYNC pair and
// a branch per iteration of the hot loops
template <int S>
__device__ __forceinline__ void lds4_if(uint32_t live, uint32_t addr, double& b0, double& b1, double& b2, double& b3) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.u32 p, %4, 0;\n\t"
      "@p ld.shared.f64 %0, [%5];\n\t"
      "@p ld.shared.f64 %1, [%5+%6];\n\t"
      "@p ld.shared.f64 %2, [%5+%7];\n\t"
      "@p ld.shared.f64 %3, [%5+%8];\n\t}"
      : "+d"(b0), "+d"(b1), "+d"(b2), "+d"(b3)
      : "r"(live), "r"(addr), "n"(S), "n"(2 * S), "n"(3 * S)
      : "memory");
}
template <int S>
__device__ __forceinline__ void lds2_if(uint32_t live, uint32_t addr, double& b0, double& b1) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.u32 p, %2, 0;\n\t"
      "@p ld.shared.f64 %0, [%3];\n\t"
      "@p ld.shared.f64 %1, [%3+%4];\n\t}"
      : "+d"(b0), "+d"(b1)
      : "r"(live), "r"(addr), "n"(S)
      : "memory");
}
__device__ __forceinline__ void sts(uint32_t addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

// D(8x8) += A(8x4) B(4x8) in fp64 on the tensor cores.  Fragments (lane = 4*gid + tig): a = A[gid][tig],
// b = B[tig][gid], c0/c1 = C[gid][2*tig], C[gid][2*tig+1].
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// c += L[8 rows][4*n4 columns] * w[4*n4 columns][.]: k-blocks (256 B each, gpmpc_state.cuh) contiguous from shared
// address `la` (lane address: + a_lane_off), w rows from `wa` (lane address: + (tig*T + gid)*8).  n4 is even; two
// accumulator sets.
// byte offset of this lane's A-fragment element (row gid, column tig) inside a k-block
__device__ __forceinline__ uint32_t a_lane_off(int gid, int tig) { return (uint32_t)sp_idx(tig, gid) * 8; }

// bf: the B-fragment registers, kept by the caller across calls (the dead lanes never load them: any finite-or-not value does,
// so zeroing them per call -- four instructions, ~16 calls per element-step -- is not needed)
template <int T>
__device__ __forceinline__ void mma_accumulate(double (&c)[4], double (&bf)[4], uint32_t la, uint32_t wa, int n4) {
  constexpr int WSTEP = 4 * T * 8;  // bytes of w per 4 columns
  // B fragments: only T < 8 of the 8 columns are live; the dead ones belong to the lanes with gid >= T -- for T <= 4 the
  // whole upper half-warp -- which issue no shared-memory access at all (their wavefront disappears; stale register
  // values only ever reach accumulator columns nobody reads)
  const uint32_t live = (T >= 8 || ((threadIdx.x & 31) >> 2) < T) ? 1u : 0u;
  double &b0 = bf[0], &b1 = bf[1], &b2 = bf[2], &b3 = bf[3];
  int it = 0;
  for (; it + 4 <= n4; it += 4) {
    const double a0 = lds<0>(la), a1 = lds<256>(la), a2 = lds<512>(la), a3 = lds<768>(la);
#if GPMPC_PRED_LOADS
    lds4_if<WSTEP>(live, wa, b0, b1, b2, b3);
#else
    if (live) { b0 = lds<0>(wa); b1 = lds<WSTEP>(wa); b2 = lds<2 * WSTEP>(wa); b3 = lds<3 * WSTEP>(wa); }
#endif
    dmma(c[0], c[1], a0, b0);
    dmma(c[2], c[3], a1, b1);
    dmma(c[0], c[1], a2, b2);
    dmma(c[2], c[3], a3, b3);
    la += 1024;
    wa += 4 * WSTEP;
  }
  if (it < n4) {
    const double a0 = lds<0>(la), a1 = lds<256>(la);
#if GPMPC_PRED_LOADS
    lds2_if<WSTEP>(live, wa, b0, b1);
#else
    if (live) { b0 = lds<0>(wa); b1 = lds<WSTEP>(wa); }
#endif
    dmma(c[0], c[1], a0, b0);
    dmma(c[2], c[3], a1, b1);
  }
}
template <int T>
__device__ __forceinline__ void mma_accumulate(double (&c)[4], uint32_t la, uint32_t wa, int n4) {
  double bf[4] = {0.0, 0.0, 0.0, 0.0};
  mma_accumulate<T>(c, bf, la, wa, n4);
}

// Finishes one sub-panel: rows n_off .. n_off+7 of wv (shared address wblk) hold the kernel entries, c the
// off-diagonal dot products (C layout).  rhs = k - dot goes back to wv, w_blk = inv(D) rhs comes out of two DMMAs
// and replaces it.  Rows >= nvalid (padding / not yet appended) are forced to 0.  dblk = shared (or generic, for
// the L_oo-in-global variant) address of the 8 x 8 diagonal block.
template <int T, bool DBLK_SHARED>
__device__ __forceinline__ void subpanel_finish(const double (&c)[4], uint32_t dblk_s, const double* dblk_g,
                                                uint32_t wblk, int nvalid, int gid, int tig) {
  const bool v0 = 2 * tig < T, v1 = 2 * tig + 1 < T, live = gid < nvalid;
  const uint32_t mine = wblk + (gid * T + 2 * tig) * 8;
  double r0 = 0.0, r1 = 0.0;
  if (v0) r0 = lds(mine) - (c[0] + c[2]);
  if (v1) r1 = lds<8>(mine) - (c[1] + c[3]);
  if (v0) sts(mine, live ? r0 : 0.0);
  if (v1) sts(mine + 8, live ? r1 : 0.0);
  // inv(D)[gid][k], k = tig and tig + 4: slot (row k, column gid) of the block, zero above the diagonal
  double a0 = 0.0, a1 = 0.0;
  if (DBLK_SHARED) {
    if (tig <= gid) a0 = lds(dblk_s + (uint32_t)sp_idx(gid, tig) * 8);
    if (tig + 4 <= gid) a1 = lds(dblk_s + (uint32_t)sp_idx(gid, tig + 4) * 8);
  } else {
    if (tig <= gid) a0 = dblk_g[sp_idx(gid, tig)];
    if (tig + 4 <= gid) a1 = dblk_g[sp_idx(gid, tig + 4)];
  }
  __syncwarp();
  const double b0 = lds(wblk + (tig * T + gid) * 8), b1 = lds(wblk + ((tig + 4) * T + gid) * 8);
  double d0 = 0.0, d1 = 0.0;
  dmma(d0, d1, a0, b0);
  dmma(d0, d1, a1, b1);  // mma.sync: every lane's operand loads above have completed
  if (v0) sts(mine, live ? d0 : 0.0);
  if (v1) sts(mine + 8, live ? d1 : 0.0);
  __syncwarp();
}

// host + device: column groups (64 B) in the factor stream of an element with c own rows
__host__ __device__ __forceinline__ int step_groups_per_element(int c, int mo) {
  return (int)(subpanel_off((c + 7) / 8, mo) / 8);
}

// WO = true (large m, "Regime A"): the shared rows  w_o = inv(L_oo) k_o  of every element were produced by the batched
// tensor-core GEMM k_shared_rows (below) into st.Wo; this kernel copies them into wv instead of running phases A
// (real points) and B, so that inv(L_oo) is streamed once per TILE of elements instead of once per element.
template <int D, int T, bool LOO_SMEM, bool WO = false>
__global__ void __launch_bounds__(STEP_MAX_WARPS * 32, 1)
k_step(DevState st, const double* __restrict__ x, int grow_factor) {
  static_assert(!(WO && LOO_SMEM), "WO replaces the in-kernel product with inv(L_oo)");
  extern __shared__ __align__(128) double smem[];
  const int j_out = blockIdx.y;
  const int nw = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const int m = st.m, mo = st.mo, c = st.c;
  const int Pm = (m + 7) >> 3, P8 = (c + 7) >> 3;
  const int loop_sz = (int)subpanel_off(Pm, 0);
  const int m_even = (m + 1) & ~1;

  // ---- shared-memory carve-up (sizes mirrored by launch_step in gpmpc_api.cu) -----------------------------
  double* sL = smem;                                         // [loop_sz] L_oo sub-panels (only if LOO_SMEM)
  const int nr_even = WO ? 0 : (st.n_real + 1) & ~1;          // WO: no real inputs / row table in shared memory
  double* sXr = sL + (LOO_SMEM ? loop_sz : 0);               // [nr_even*D] real inputs
  double* sBo = sXr + (size_t)nr_even * D;                   // [m_even]
  int* sRrow = (int*)(sBo + m_even);                         // [n_real*T] factor row of (real point, task), -1 = unobserved
  int* sHrow = sRrow + (WO ? 0 : (st.n_real * T + 1) & ~1);  // [np] first own row of hallucinated point p, -1 = not in the factor
  const int wv_rows = mo + 8 * P8;
  const int wv_sz = (wv_rows * T + 8 + 15) & ~15;            // per warp, doubles (+8: don't-care reads of idle lanes)
  const int wb_sz = (wv_rows + 15) & ~15;
  const int per_warp = wv_sz + wb_sz + STEP_NST * STEP_SEG * 8;
  double* warp_base = (double*)(sHrow + ((st.np + 1) & ~1));
  warp_base = (double*)(((uintptr_t)warp_base + 127) & ~(uintptr_t)127);
  double* wv = warp_base + (size_t)warp * per_warp;          // [wv_rows][T]  k, then w
  double* wb = wv + wv_sz;                                   // [wv_rows]     beta by storage column
  double* ring = wb + wb_sz;                                 // [STEP_NST][STEP_SEG*8]
  uint64_t* bars = (uint64_t*)(warp_base + (size_t)nw * per_warp) + warp * STEP_NST;

  if (lane < STEP_NST) mbar_init(bars + lane, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const double* gL = st.LooP + (size_t)j_out * loop_sz;
  if (LOO_SMEM)
    for (int idx = threadIdx.x; idx < loop_sz; idx += blockDim.x) sL[idx] = gL[idx];
  if (!WO) {
    for (int idx = threadIdx.x; idx < st.n_real * D; idx += blockDim.x) sXr[idx] = st.Xr[idx];
    for (int idx = threadIdx.x; idx < st.n_real * T; idx += blockDim.x) sRrow[idx] = -1;
  }
  for (int idx = threadIdx.x; idx < st.np; idx += blockDim.x) sHrow[idx] = st.hrow0[idx];
  for (int idx = threadIdx.x; idx < m; idx += blockDim.x) sBo[idx] = st.beta_o[(size_t)j_out * m + idx];
  __syncthreads();
  if (!WO)
    for (int idx = threadIdx.x; idx < m; idx += blockDim.x) sRrow[st.obs_pt[idx] * T + st.obs_task[idx]] = idx;
  // padding rows [m, mo) and rows >= c of wv / wb are zero for the whole launch
  for (int idx = lane; idx < wv_sz + wb_sz; idx += 32) wv[idx] = 0.0;
  __syncthreads();
  for (int idx = lane; idx < m; idx += 32) wb[idx] = sBo[idx];
  __syncwarp();
  // no block-level synchronisation below this line: every warp runs its own element loop

  const uint32_t wv_s = smem_u32(wv), wb_s = smem_u32(wb), ring_s = smem_u32(ring), bars_s = smem_u32(bars);
  const uint32_t sL_s = smem_u32(sL);
  const uint32_t a_lane = a_lane_off(gid, tig);      // lane offset into a run of k-blocks (A fragment)
  const uint32_t b_lane = (tig * T + gid) * 8;       // lane offset into wv rows (B fragment)
  const int nwarps_total = gridDim.x * nw;
  const int s_first = blockIdx.x * nw + warp;
  double il[D];
#pragma unroll
  for (int a = 0; a < D; ++a) il[a] = 1.0 / st.ls[j_out * D + a];
  const double os = st.os[j_out];

  // ---- producer (warp-uniform state; lane 0 issues): the factor streams of this warp's elements, one after
  //      the other, in chunks of STEP_SLOT_BYTES (the last chunk of an element is shorter) ----------------------
  const unsigned elem_bytes = (unsigned)step_groups_per_element(c, mo) * 64u;
  const size_t elem_step = (size_t)st.elem_stride * st.g_ny * nwarps_total * 8;  // bytes to this warp's next element
  const char* prod_base = (const char*)(st.Lh + (size_t)(s_first * st.g_ny + j_out) * st.elem_stride);
  int prod_left = (elem_bytes > 0 && s_first < st.ns) ? (st.ns - s_first + nwarps_total - 1) / nwarps_total : 0;  // elements
  unsigned prod_off = 0, prod_slot = 0;
  // The LAST sub-panel of the stream holds only v = c mod 8 rows so far (v = 0: it is full).  A k-block is row-major, so
  // its valid rows are the first 32 v bytes of its 256: those k-blocks are fetched one small bulk copy each (lane i takes
  // k-block i of the chunk) into their usual places in the slot -- the rows never fetched keep stale shared memory, which
  // only ever reaches accumulator rows >= v that subpanel_finish discards by selection.  5.5 % fewer factor bytes from DRAM
  // over the car rollout.
  const unsigned tail_rows = (unsigned)(c & 7);
  const unsigned tail_start = tail_rows ? (unsigned)subpanel_off(P8 - 1, mo) * 8u : elem_bytes;
  auto produce_one = [&]() {
    if (prod_left == 0) return;
    const unsigned bytes = min((unsigned)STEP_SLOT_BYTES, elem_bytes - prod_off);
    const unsigned bulk = prod_off < tail_start ? min(prod_off + bytes, tail_start) - prod_off : 0u;  // one copy
    const unsigned nkb_t = (bytes - bulk) >> 8;                                                          // tail k-blocks
    const unsigned tx = bulk + nkb_t * tail_rows * 32u;
    const uint32_t bar = bars_s + prod_slot * 8;
    const uint32_t dst = ring_s + prod_slot * STEP_SLOT_BYTES;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.eq.u32 p, %4, 0;\n\t"
        "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %5;\n\t"
        "setp.ne.and.u32 q, %1, 0, p;\n\t"
        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%2], [%3], %1, [%0];\n\t}"
        ::"r"(bar), "r"(bulk), "r"(dst), "l"(prod_base + prod_off), "r"(lane), "r"(tx) : "memory");
    if ((unsigned)lane < nkb_t) {
      const unsigned o = bulk + (unsigned)lane * 256u;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst + o), "l"(prod_base + prod_off + o), "r"(tail_rows * 32u), "r"(bar) : "memory");
    }
    prod_slot = prod_slot + 1 == STEP_NST ? 0 : prod_slot + 1;
    prod_off += bytes;
    if (prod_off == elem_bytes) {
      prod_off = 0;
      prod_base += elem_step;
      --prod_left;
    }
  };
#pragma unroll
  for (int i = 0; i < STEP_NST; ++i) produce_one();
  unsigned cons_slot = 0, cons_parity = 0;  // slot / phase parity of the chunk being consumed
  double bfrag[4] = {0.0, 0.0, 0.0, 0.0};   // B-fragment registers of mma_accumulate, kept across its calls

  // this element's test input and base samples are loaded one element ahead
  double xs_n[D];
  if (s_first < st.ns) {
#pragma unroll
    for (int a = 0; a < D; ++a) xs_n[a] = x[(size_t)(s_first * st.g_ny + j_out) * D + a];
  }
  const int np = st.np;

  for (int s_idx = s_first; s_idx < st.ns; s_idx += nwarps_total) {
    const int b = s_idx * st.g_ny + j_out;
    __syncwarp();  // wv is about to be rewritten: every lane is done with the previous element
    double xs[D];
#pragma unroll
    for (int a = 0; a < D; ++a) xs[a] = xs_n[a];
    const double* Xb = st.Xh + (size_t)b * st.cap_points * D;
    const double* bh = st.beta_h + (size_t)b * st.c_cap;
    if (s_idx + nwarps_total < st.ns) {
      const size_t bn_ = (size_t)(s_idx + nwarps_total) * st.g_ny + j_out;
#pragma unroll
      for (int a = 0; a < D; ++a) xs_n[a] = x[bn_ * D + a];
      // pull the next element's hallucinated inputs and beta towards L2 while this one is being processed
      const char* nx = (const char*)(st.Xh + bn_ * st.cap_points * D);
      const char* nb = (const char*)(st.beta_h + bn_ * st.c_cap);
      if (lane * 128 < np * D * 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + lane * 128));
      if (lane * 128 < c * 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(nb + lane * 128));
      if (WO) {
        const char* nw_ = (const char*)(st.Wo + bn_ * (size_t)mo * T);
        for (int o = lane * 128; o < mo * T * 8; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nw_ + o));
      }
    }
    if (WO) {
      // shared rows from the batched GEMM: [mo][T] doubles, the layout of wv
      const double2* wo = (const double2*)(st.Wo + (size_t)b * mo * T);
      double2* wv2 = (double2*)wv;
      const int n2 = mo * T / 2;
      int i2 = lane;
      for (; i2 + 96 < n2; i2 += 128) {
        const double2 v0 = __ldcg(wo + i2), v1 = __ldcg(wo + i2 + 32), v2 = __ldcg(wo + i2 + 64), v3 = __ldcg(wo + i2 + 96);
        wv2[i2] = v0; wv2[i2 + 32] = v1; wv2[i2 + 64] = v2; wv2[i2 + 96] = v3;
      }
      for (; i2 < n2; i2 += 32) wv2[i2] = __ldcg(wo + i2);
    }

    // ---- A: kernel vector, one exp per training POINT (and the element's beta) ------------------------------
    // real and hallucinated points share ONE pass over the lanes (point index qi: the n_real real points first, then the np
    // hallucinated ones): the exp / derivative-block sequence runs once per 32 points of either kind -- separate loops ran it
    // once for the <= 32 real points and once or twice more for the hallucinated ones, ~150 warp-instructions per pass
    {
      const int n_rl = WO ? 0 : st.n_real;
      for (int q0 = 0; q0 < n_rl + np; q0 += 32) {
        const int qi = q0 + lane, ph = qi - n_rl;
        const bool is_real = qi < n_rl;
        int ra = -1;  // first own row of hallucinated point ph (>= 0: in the factor; <= -2: null rows at -2 - ra)
        if (!is_real && ph < np) {
          ra = sHrow[ph];
          // grouped rollout: a masked / dropped point keeps its (null) factor rows but contributes no kernel entries
          if (st.pstate && ra >= 0 && st.pstate[(size_t)b * st.cap_points + ph]) ra = -2 - ra;
        }
        double xa[D], ba[T];
        if (is_real) {
#pragma unroll
          for (int a = 0; a < D; ++a) xa[a] = sXr[qi * D + a];
        } else if (ra >= 0) {
#pragma unroll
          for (int a = 0; a < D; ++a) xa[a] = Xb[(size_t)ph * D + a];
#pragma unroll
          for (int r = 0; r < T; ++r) ba[r] = bh[ra + r];
        }
        if (is_real || ra >= 0) {
          double kb[T][T];
          kernel_block<D, T>(xa, xs, il, os, kb);
#pragma unroll
          for (int ta = 0; ta < T; ++ta) {
            const int row = is_real ? sRrow[qi * T + ta] : mo + ra + ta;
            if (row >= 0) {
#pragma unroll
              for (int tb = 0; tb < T; ++tb) wv[row * T + tb] = kb[ta][tb];
              if (!is_real) wb[row] = ba[ta];
            }
          }
        } else if (ra <= -2) {  // null rows: zero kernel entries, zero beta
          const int r0 = -2 - ra;
#pragma unroll
          for (int ta = 0; ta < T; ++ta) {
#pragma unroll
            for (int tb = 0; tb < T; ++tb) wv[(mo + r0 + ta) * T + tb] = 0.0;
            wb[mo + r0 + ta] = 0.0;
          }
        }
      }
    }
    __syncwarp();

    // ---- B: shared rows: w_o = inv(L_oo) k_o, tile-row by tile-row from the LAST one (row i needs k_j, j <= i only,
    //      so the in-place write of a tile-row never disturbs the rows still to be computed) --------------------------
    for (int p8 = Pm - 1; !WO && p8 >= 0; --p8) {
      const uint32_t boff = (uint32_t)subpanel_off(p8, 0) * 8;
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      if (LOO_SMEM) {
        mma_accumulate<T>(acc, bfrag, sL_s + boff + a_lane, wv_s + b_lane, 2 * p8 + 2);
      } else {
        const double* lp = gL + boff / 8 + a_lane / 8;
        for (int it = 0; it < 2 * p8 + 2; it += 2) {
          const double a0 = lp[it * 32], a1 = lp[it * 32 + 32];
          const double b0 = lds(wv_s + b_lane + it * 4 * T * 8), b1 = lds(wv_s + b_lane + (it + 1) * 4 * T * 8);
          dmma(acc[0], acc[1], a0, b0);
          dmma(acc[2], acc[3], a1, b1);
        }
      }
      // mma.sync: every lane's reads of this tile-row's inputs are complete
      const uint32_t mine = wv_s + ((8 * p8 + gid) * T + 2 * tig) * 8;
      if (2 * tig < T) sts(mine, acc[0] + acc[2]);
      if (2 * tig + 1 < T) sts(mine + 8, acc[1] + acc[3]);
    }
    __syncwarp();

    // ---- C: own rows, streamed through the TMA ring -------------------------------------------------------------
    if (P8 > 0) {
      mbar_wait_s(bars_s + cons_slot * 8, cons_parity);
      unsigned left_in_elem = elem_bytes / 64;  // column groups of this element not yet consumed
      int cpos = 0;                             // column groups consumed in the current chunk
      // the current chunk is exhausted: hand its slot back to the producer and wait for the next one
      auto next_chunk = [&]() {
        __syncwarp();
        produce_one();
        cpos = 0;
        if (++cons_slot == STEP_NST) { cons_slot = 0; cons_parity ^= 1; }
        if (left_in_elem > 0) mbar_wait_s(bars_s + cons_slot * 8, cons_parity);
      };
      for (int p8 = 0; p8 < P8; ++p8) {
        const int n_off = mo + 8 * p8;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        uint32_t wa = wv_s + b_lane;
        int rem = n_off;
        while (rem > 0) {
          const int piece = min(rem, STEP_SEG - cpos);
          mma_accumulate<T>(acc, bfrag, ring_s + cons_slot * STEP_SLOT_BYTES + cpos * 64 + a_lane, wa, piece >> 2);
          wa += piece * T * 8;
          rem -= piece;
          cpos += piece;
          left_in_elem -= piece;
          if (cpos == STEP_SEG) next_chunk();
        }
        const uint32_t dblk = ring_s + cons_slot * STEP_SLOT_BYTES + cpos * 64;
        subpanel_finish<T, true>(acc, dblk, nullptr, wv_s + n_off * T * 8, min(8, c - 8 * p8), gid, tig);
        cpos += 8;
        left_in_elem -= 8;
        if (cpos == STEP_SEG || left_in_elem == 0) next_chunk();
      }
    }

    // ---- D: posterior moments: C[r][s] = sum_t w[t][r] w[t][s],  C[r][7] = sum_t w[t][r] beta[t] ------------------
    {
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      uint32_t wa = wv_s + b_lane, ba = wb_s + tig * 8;
      const bool is_beta = gid == 7;
      for (int t = 0; t < wv_rows; t += 8) {
        const double a0 = lds<0>(wa), a1 = lds<4 * T * 8>(wa);
        const double e0 = lds<0>(ba), e1 = lds<32>(ba);
        dmma(acc[0], acc[1], a0, is_beta ? e0 : a0);
        dmma(acc[2], acc[3], a1, is_beta ? e1 : a1);
        wa += 8 * T * 8;
        ba += 64;
      }
      // hand [sum_t w_t beta_t | sum_t w_t w_t^T (lower)] to the per-element finishing kernel: lane (gid, tig) holds
      // C[gid][2 tig], C[gid][2 tig + 1];  C[r][7] = mean_r
      constexpr int FS = T + T * (T + 1) / 2;
      double* fo = st.fin + (size_t)b * FS;
      const double c0 = acc[0] + acc[2], c1 = acc[1] + acc[3];
      if (gid < T) {
        const int s0 = 2 * tig, s1 = 2 * tig + 1, tri = T + gid * (gid + 1) / 2;
        if (s0 <= gid) fo[tri + s0] = c0;
        if (s1 <= gid) fo[tri + s1] = c1;
        if (s1 == 7) fo[gid] = c1;
      }
    }
    if (!grow_factor) continue;
    // F (first half): the new rows' entries left of their diagonal block are w itself; row c+r lives in sub-panel
    // (c+r)/8, entry of storage column t at + t*8.  The diagonal-block part depends on the draw: k_step_finish.
    double* Le = st.Lh + (size_t)b * st.elem_stride;
    double* rowp[T];
#pragma unroll
    for (int r = 0; r < T; ++r) rowp[r] = Le + subpanel_off((c + r) >> 3, mo) + sp_idx(0, (c + r) & 7);
    // (lanes = consecutive storage columns: every 4 lanes fill one 32-byte sector of the row; the padding columns are
    // written too, as zeros, so that the sector of column m - 1 is complete)
    for (int t = lane; t < mo + c; t += 32) {
      const bool pad = t >= m && t < mo;
      const size_t to = sp_idx(t, 0);
#pragma unroll
      for (int r = 0; r < T; ++r) rowp[r][to] = pad ? 0.0 : wv[t * T + r];
    }
  }
}

// Observed real scalars are ordered point-major, task-minor.  True iff all T tasks of scalar i's point are observed
// (their rows are then the consecutive rows i - ta .. i - ta + T - 1): such a point needs one exp for its T x T block.
template <int T>
__device__ __forceinline__ bool real_point_full(const DevState& st, int i, int pt, int ta, int m) {
  if (T == 1) return false;
  const int i0 = i - ta;
  return i0 >= 0 && i0 + T - 1 < m && st.obs_pt[i0] == pt && st.obs_task[i0] == 0 && st.obs_pt[i0 + T - 1] == pt &&
         st.obs_task[i0 + T - 1] == T - 1;
}

// K1a: the shared rows of EVERY element as one batched tensor-core product ("Regime A"): for large m
// the per-element product with inv(L_oo) is FP64-contraction bound and, done one element at a time, re-streams the
// m x m factor from L2 for every element.  Here a CTA takes a TILE of E = 8 NB / T consecutive samples of output j:
//   A  K[mo][8 NB] in shared memory: row i = observed real scalar i, column e*T + tb = cov(scalar i, task tb at x_e)
//      ((scalar, element) pairs over the threads, one exp per pair; columns rotated per row so that the B-fragment
//      loads of mma.m8n8k4 are bank-conflict free without padding the rows)
//   B  W = inv(L_oo) K: every warp takes 8-row panels of inv(L_oo) (longest first), streams the panel ONCE from
//      L2 straight into A fragments (coalesced 256-byte rows of the sub-panel layout, 8 k-steps prefetched in
//      registers) and multiplies it with all NB column blocks: NB DMMAs per A load, all 8 columns of every MMA live
//   C  the accumulators go to st.Wo[b][row][T], the layout of k_step's wv array (k_step<.., WO = true> picks them up)
// inv(L_oo) traffic from L2 per element drops by E (8 at T = 3) against the one-element-per-pass path.
#define SR_THREADS 512
template <int NC>
__device__ __forceinline__ int sr_rot(int row) { return NC == 16 ? 4 * (row & 3) : 4 * ((row >> 1) & 1); }

// k_lo / k_hi (multiples of 8): only the rows [k_lo, k_hi) of K (= columns of inv(L_oo)) are processed by this launch and the
// result is ADDED to st.Wo when k_lo > 0 -- the k-slab form for training sets whose kernel tile K[mo][8 NB] does not fit in
// shared memory (m = 10^4: 640 KB for NB = 1); the host walks the slabs.  k_lo = 0, k_hi = mo is the one-pass form.
template <int D, int T, int NB>
__global__ void __launch_bounds__(SR_THREADS, 1)
k_shared_rows(DevState st, const double* __restrict__ x, int k_lo, int k_hi) {
  constexpr int NC = 8 * NB, E = NC / T;
  constexpr uint32_t KSTEP = 4 * NC * 8;  // bytes of K per k-step (4 rows)
  extern __shared__ __align__(128) double smem[];
  const int j_out = blockIdx.y;
  const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const int m = st.m, mo = st.mo, Pm = mo >> 3;
  const int ks = k_hi - k_lo;         // rows of K in this pass
  const bool one_pass = k_lo == 0 && k_hi == mo;
  double* sK = smem;                  // [ks][NC]   (row i of K at local row i - k_lo)
  double* sx = sK + (size_t)ks * NC;  // [E][D]
  const double* gL = st.LooP + (size_t)j_out * subpanel_off(Pm, 0);
  for (int idx = threadIdx.x; idx < ks * NC; idx += blockDim.x) sK[idx] = 0.0;  // padding rows / columns stay 0
  double il[D];
#pragma unroll
  for (int a = 0; a < D; ++a) il[a] = 1.0 / st.ls[j_out * D + a];
  const double os = st.os[j_out];
  uint32_t colo[NB];  // byte offset of column nb*8 + gid in this lane's rows (row & 3 == tig)
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) colo[nb] = (uint32_t)(((nb * 8 + gid + sr_rot<NC>(tig)) % NC) * 8);
  const uint32_t brow = smem_u32(sK) + tig * NC * 8;
  const int n_tiles = (st.ns + E - 1) / E;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int s0 = tile * E, e_live = min(E, st.ns - s0);
    __syncthreads();  // the previous tile's reads of sK / sx are complete
    if (threadIdx.x < E * D) {
      const int e = threadIdx.x / D, a = threadIdx.x - e * D;
      sx[threadIdx.x] = e < e_live ? x[((size_t)(s0 + e) * st.g_ny + j_out) * D + a] : 0.0;
    }
    __syncthreads();
    const int i_hi = min(m, k_hi);
    for (int idx = threadIdx.x; idx < (i_hi - k_lo) * E; idx += blockDim.x) {
      const int il_ = idx / E, e = idx - il_ * E, i = k_lo + il_;
      double xs[D];
#pragma unroll
      for (int a = 0; a < D; ++a) xs[a] = sx[e * D + a];
      const int pt = st.obs_pt[i], ta = st.obs_task[i];
      if (one_pass && real_point_full<T>(st, i, pt, ta, m)) {  // (a point's rows may straddle a slab boundary: row by row there)
        // all T tasks of the point are observed (rows i - ta .. i - ta + T - 1): ONE exp for its T x T block, by the
        // thread of its task-0 row
        if (ta != 0) continue;
        double xa[D], kb[T][T];
#pragma unroll
        for (int a = 0; a < D; ++a) xa[a] = st.Xr[(size_t)pt * D + a];
        kernel_block<D, T>(xa, xs, il, os, kb);
#pragma unroll
        for (int t2 = 0; t2 < T; ++t2) {
          const int rot = sr_rot<NC>(i + t2);
#pragma unroll
          for (int tb = 0; tb < T; ++tb) sK[(i + t2) * NC + (e * T + tb + rot) % NC] = kb[t2][tb];
        }
      } else {
        double out[T];
        kernel_row<D, T>(st.Xr + (size_t)pt * D, ta, xs, il, os, out);
        const int rot = sr_rot<NC>(il_);
#pragma unroll
        for (int tb = 0; tb < T; ++tb) sK[il_ * NC + (e * T + tb + rot) % NC] = out[tb];
      }
    }
    __syncthreads();

    const int p_first = k_lo >> 3;  // panels above the slab have no columns in it (lower triangular)
    // panels longest first, dealt to the warps in serpentine order (0 .. nw-1, nw-1 .. 0, ...): panel p costs 2 (p + 1)
    // k-steps, so plain round-robin leaves warp 0 with 90 of them and warp 7 with 48 at m = 180 (8 warps); this way 76 / 62
    for (int rnd = 0;; ++rnd) {
#ifdef GPMPC_SR_ROUND_ROBIN
      const int pi = rnd * nw + warp;
#else
      const int pi = rnd * nw + ((rnd & 1) ? nw - 1 - warp : warp);
#endif
      if (rnd * nw >= Pm - p_first) break;
      if (pi >= Pm - p_first) continue;
      const int p = Pm - 1 - pi;
      const int n4 = (min(k_hi, 8 * p + 8) - k_lo) >> 2;  // k-steps of 4 columns of this panel inside the slab
      const double* ap = gL + subpanel_off(p, 0) + sp_idx(tig, gid) + (size_t)(k_lo >> 2) * 32;
      double acc[NB][4];
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) acc[nb][0] = acc[nb][1] = acc[nb][2] = acc[nb][3] = 0.0;
      double abuf[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) abuf[i] = i < n4 ? __ldcg(ap + i * 32) : 0.0;
      uint32_t bb = brow;
      for (int k0 = 0; k0 < n4; k0 += 8) {
        double acur[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acur[i] = abuf[i];
        ap += 256;
#pragma unroll
        for (int i = 0; i < 8; ++i) abuf[i] = k0 + 8 + i < n4 ? __ldcg(ap + i * 32) : 0.0;
#define SR_KSTEP(i)                                                                    \
        if (k0 + (i) < n4) {                                                           \
          _Pragma("unroll") for (int nb = 0; nb < NB; ++nb) {                          \
            const double bv = lds<(i) * KSTEP>(bb + colo[nb]);                         \
            dmma(acc[nb][2 * ((i) & 1)], acc[nb][2 * ((i) & 1) + 1], acur[i], bv);     \
          }                                                                            \
        }
        SR_KSTEP(0) SR_KSTEP(1) SR_KSTEP(2) SR_KSTEP(3) SR_KSTEP(4) SR_KSTEP(5) SR_KSTEP(6) SR_KSTEP(7)
#undef SR_KSTEP
        bb += 8 * KSTEP;
      }
      const int row = 8 * p + gid;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int col = nb * 8 + 2 * tig + hh;
          const int e = col / T, tb = col - e * T;
          if (e < e_live) {
            double* wo = st.Wo + (((size_t)(s0 + e) * st.g_ny + j_out) * mo + row) * T + tb;
            const double v = acc[nb][hh] + acc[nb][2 + hh];
            *wo = k_lo > 0 ? *wo + v : v;
          }
        }
    }
  }
}

// K1s: the step when NO element has own factor rows and none are appended (c == 0, grow == 0): the value-only model
// of benchmarking/simulate_forward_sampling_car.py as shipped (use_model_without_derivatives: the model ignores
// the hallucinated set, src/agent.py:221-226), and posterior-only queries on real data.  Everything is a product
// with the SHARED inv(L_oo): FP64-pipe bound, not HBM bound (SURVEY.md 8d "Regime A"), so the job is to fill the
// tensor-core tile: one warp takes G = 8 / T elements (consecutive samples of output j) at a time and puts their
// G*T right-hand sides side by side in the N = 8 dimension of mma.sync.m8n8k4.f64 -- with T = 1 that is 8 samples per
// DMMA instead of 1 -- and the kernel vectors of the G elements are spread over all 32 lanes ((real point, element)
// pairs), one exp per pair.
//   A  K[mo][8]: column g*T + tb = cov(train scalar, task tb at x_g)              (wv8, row stride 8 doubles)
//   B  W = inv(L_oo) K, tile-rows last to first, in place                         (as K1 phase B, all 8 columns live)
//   D  C = W^T W (diagonal T x T blocks = the elements' W^T W),  M = W^T beta_o    -> st.fin, then k_step_finish
template <int D, int T>
__global__ void __launch_bounds__(STEP_MAX_WARPS * 32, 1)
k_step_shared(DevState st, const double* __restrict__ x) {
  constexpr int G = 8 / T;       // elements per warp pass
  constexpr int FS = T + T * (T + 1) / 2;
  extern __shared__ __align__(128) double smem[];
  const int j_out = blockIdx.y;
  const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const int m = st.m, mo = st.mo, Pm = (m + 7) >> 3;
  const int loop_sz = (int)subpanel_off(Pm, 0);
  const int nr_even = (st.n_real + 1) & ~1;
  // carve-up (mirrored by launch_step_shared): inv(L_oo) | real inputs | beta_o (zero padded to mo) | row table | per warp
  double* sL = smem;
  double* sXr = sL + loop_sz;
  double* sBo = sXr + (size_t)nr_even * D;
  int* sRrow = (int*)(sBo + mo);
  double* warp_base = (double*)(sRrow + ((st.n_real * T + 1) & ~1));
  warp_base = (double*)(((uintptr_t)warp_base + 127) & ~(uintptr_t)127);
  const int per_warp = mo * 8 + 8 * D;
  double* wv = warp_base + (size_t)warp * per_warp;  // [mo][8]
  double* sx = wv + (size_t)mo * 8;                  // [G][D] test inputs of the group

  const double* gL = st.LooP + (size_t)j_out * loop_sz;
  for (int idx = threadIdx.x; idx < loop_sz; idx += blockDim.x) sL[idx] = gL[idx];
  for (int idx = threadIdx.x; idx < st.n_real * D; idx += blockDim.x) sXr[idx] = st.Xr[idx];
  for (int idx = threadIdx.x; idx < st.n_real * T; idx += blockDim.x) sRrow[idx] = -1;
  for (int idx = threadIdx.x; idx < mo; idx += blockDim.x) sBo[idx] = idx < m ? st.beta_o[(size_t)j_out * m + idx] : 0.0;
  __syncthreads();
  for (int idx = threadIdx.x; idx < m; idx += blockDim.x) sRrow[st.obs_pt[idx] * T + st.obs_task[idx]] = idx;
  for (int idx = lane; idx < mo * 8; idx += 32) wv[idx] = 0.0;  // padding rows [m, mo) and unused columns stay 0
  __syncthreads();

  const uint32_t wv_s = smem_u32(wv), sL_s = smem_u32(sL), bo_s = smem_u32(sBo);
  const uint32_t a_lane = a_lane_off(gid, tig); // A fragment: k-block run of inv(L_oo)
  const uint32_t b_lane = (tig * 8 + gid) * 8;  // B fragment: wv[4k + tig][gid]
  double il[D];
#pragma unroll
  for (int a = 0; a < D; ++a) il[a] = 1.0 / st.ls[j_out * D + a];
  const double os = st.os[j_out];
  const int n_groups = (st.ns + G - 1) / G;
  const int nwarps_total = gridDim.x * nw;

  for (int grp = blockIdx.x * nw + warp; grp < n_groups; grp += nwarps_total) {
    const int s0 = grp * G;
    const int g_live = min(G, st.ns - s0);
    __syncwarp();  // previous group's reads of wv / sx are complete
    if (lane < G * D) {
      const int g = lane / D, a = lane % D;
      sx[lane] = g < g_live ? x[((size_t)(s0 + g) * st.g_ny + j_out) * D + a] : 0.0;
    }
    __syncwarp();
    // ---- A: (real point, element) pairs over the lanes ----------------------------------------------------------
    for (int idx = lane; idx < st.n_real * G; idx += 32) {
      const int p = idx / G, g = idx % G;
      double xa[D], xs[D], kb[T][T];
#pragma unroll
      for (int a = 0; a < D; ++a) { xa[a] = sXr[p * D + a]; xs[a] = sx[g * D + a]; }
      kernel_block<D, T>(xa, xs, il, os, kb);
#pragma unroll
      for (int ta = 0; ta < T; ++ta) {
        const int row = sRrow[p * T + ta];
        if (row >= 0) {
#pragma unroll
          for (int tb = 0; tb < T; ++tb) wv[row * 8 + g * T + tb] = kb[ta][tb];
        }
      }
    }
    __syncwarp();
    // ---- B: W = inv(L_oo) K -------------------------------------------------------------------------------------
    for (int p8 = Pm - 1; p8 >= 0; --p8) {
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      mma_accumulate<8>(acc, sL_s + (uint32_t)subpanel_off(p8, 0) * 8 + a_lane, wv_s + b_lane, 2 * p8 + 2);
      const uint32_t mine = wv_s + ((8 * p8 + gid) * 8 + 2 * tig) * 8;
      sts(mine, acc[0] + acc[2]);
      sts(mine + 8, acc[1] + acc[3]);
    }
    __syncwarp();
    // ---- D: C = W^T W and M = W^T beta_o (every column of M is the same vector) ---------------------------------------
    double cw[4] = {0.0, 0.0, 0.0, 0.0}, cm[4] = {0.0, 0.0, 0.0, 0.0};
    {
      uint32_t wa = wv_s + b_lane, ba = bo_s + tig * 8;
      for (int t = 0; t < mo; t += 8) {
        const double a0 = lds<0>(wa), a1 = lds<256>(wa);
        const double e0 = lds<0>(ba), e1 = lds<32>(ba);
        dmma(cw[0], cw[1], a0, a0);
        dmma(cw[2], cw[3], a1, a1);
        dmma(cm[0], cm[1], a0, e0);
        dmma(cm[2], cm[3], a1, e1);
        wa += 512;
        ba += 64;
      }
    }
    // lane (gid, tig) holds C[gid][2 tig], C[gid][2 tig + 1]; row gid belongs to element g = gid / T, task r = gid % T
    const int g = gid / T, r = gid - g * T;
    if (g < g_live) {
      double* fo = st.fin + ((size_t)(s0 + g) * st.g_ny + j_out) * FS;
      const int c0 = 2 * tig - g * T, c1 = c0 + 1;  // task index of the two columns within element g
      if (c0 >= 0 && c0 <= r) fo[T + r * (r + 1) / 2 + c0] = cw[0] + cw[2];
      if (c1 >= 0 && c1 <= r) fo[T + r * (r + 1) / 2 + c1] = cw[1] + cw[3];
      if (tig == 0) fo[r] = cm[0] + cm[2];
    }
  }
}

// K1b: per-element scalar tail of the step, ONE THREAD per batch element (the warp kernel above would do this
// arithmetic 32-fold redundantly): Sigma* = K** - W^T W, T x T Cholesky with GPyTorch's jitter ladder, y = mean +
// L eps, zero-variance / truncation (src/agent.py:646-708), then the diagonal-block part of the rank-T append:
// chol(Sigma* + noise), 1/L_kk, beta_new and the transposed block inverses (gpmpc_state.cuh).
#ifndef FIN_THREADS
#define FIN_THREADS 128
#endif
// modes of k_step_finish: the plain fused step, GPyTorch's eigen-root redo, and the two halves of a GROUPED step (per-Agent
// filter of update_hallucinated_Dyn_dataset, src/agent.py:164-202): draw only, then -- after the group-wise all / any
// reduction of the "too close" flags -- the append, which writes NULL rows for a masked / dropped point
#define FIN_ALL 0
#define FIN_EIG_REDO 1
#define FIN_DRAW 2
#define FIN_APPEND 3
struct GroupArgs {
  const unsigned char* flag;  // [B]  1 = this element's new point is within min_dist of its data set (label becomes NaN)
  const unsigned char* decision;  // [n_groups]  0 append, 1 mask, 2 drop
  int group_elems;            // batch elements per group = group_size * g_ny
};

// The diagonal-block part of the append for ONE 8 x 8 block with everything known at compile time: IF = block row of the first
// new row, new rows [R0, R1) of the T fall into this block (c is uniform over a launch, so the host dispatches on c mod 8).
// The 8 x 8 mirror then lives in registers -- with run-time bounds it is a local-memory array and its loads form a serial chain
// of DRAM round trips (ncu: 63 % of k_step_finish's stall samples sat on these loops) -- and every load is issued before the
// first use.  Same arithmetic, same order as the run-time form in k_step_finish.
template <int T, int IF, int R0, int R1>
__device__ __forceinline__ void finish_block_const(double* Le, int c, int mo, const TriT<T>& Ln, const double (&rdn)[T]) {
  const int kb = c + R0 - IF;  // first own row of this block (a multiple of 8)
  double* sp = Le + subpanel_off(kb >> 3, mo);
  double* gblk = sp + (size_t)(mo + kb) * 8;
  double blk[8][8];
  // old columns t < IF: the old inverse (slots on and above the diagonal: column t, rows <= t) and the new rows' L entries
#pragma unroll
  for (int t = 0; t < 8; ++t)
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (t < IF && (i <= t || (i >= IF && i < IF + (R1 - R0)))) blk[t][i] = gblk[sp_idx(t, i)];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i >= IF && i < IF + (R1 - R0)) {
      const int r = R0 + (i - IF);
#pragma unroll
      for (int s = 0; s < T; ++s)
        if (s >= R0 && s < r) blk[IF + (s - R0)][i] = Ln.v[r * (r + 1) / 2 + s];
      blk[i][i] = rdn[r];
#pragma unroll
      for (int jc = 0; jc < 8; ++jc) {
        if (jc < i) {
          double acc = 0.0;
#pragma unroll
          for (int t = 0; t < 8; ++t)
            if (t >= jc && t < i) acc = fma(blk[t][i], blk[t][jc], acc);
          blk[i][jc] = -rdn[r] * acc;  // slot (row jc, column i)
        }
      }
#pragma unroll
      for (int s = 0; s < T; ++s)
        if (s < R0) sp[sp_idx(mo + c + s, i)] = Ln.v[r * (r + 1) / 2 + s];  // new columns of an earlier block
#pragma unroll
      for (int t = 0; t < 8; ++t)
        if (t >= IF && t <= i) gblk[sp_idx(t, i)] = blk[t][i];  // new columns of row i (incl. 1 / L_ii)
#pragma unroll
      for (int jc = 0; jc < 8; ++jc)
        if (jc < i) gblk[sp_idx(i, jc)] = blk[i][jc];  // transposed inverse row
    }
  }
}

// IFC = c mod 8 when the host knows it (the plain fused step), -1 = run-time block bounds
template <int T, int IFC = -1>
__global__ void __launch_bounds__(FIN_THREADS)
k_step_finish(DevState st, const double* __restrict__ x, const double* __restrict__ eps, gpmpc_sample_opts opts,
              double* __restrict__ mean, double* __restrict__ var, double* __restrict__ y,
              int* __restrict__ jitter_level, int grow_factor, int mode, GroupArgs ga = GroupArgs{}) {
  constexpr int FS = T + T * (T + 1) / 2;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= st.B) return;
  const int eig_redo = mode == FIN_EIG_REDO;
  // eig_redo: the launch queued behind the regular one; it repeats the draw and the append of EVERY element through the
  // eigen root iff some element's jitter ladder failed in this step (GPyTorch's batch-wide symeig fallback, gpmpc_eig.cuh).
  // Everything this kernel writes depends only on st.fin, eps and the old factor rows, so the repeat simply overwrites.
  if (eig_redo && (T == 1 || *(volatile int*)st.eig_flag != st.eig_epoch)) return;
  const int j_out = b % st.g_ny, d = st.d, c = st.c, mo = st.mo;
  if (IFC > 0 && grow_factor && eps && mode != FIN_DRAW) {
    // the diagonal block the new rows fall into (4 lines of 128 B) is needed at the very end: start fetching it now
    const char* blk0 = (const char*)(st.Lh + (size_t)b * st.elem_stride + subpanel_off(c >> 3, mo) + (size_t)(mo + (c & ~7)) * 8);
#pragma unroll
    for (int l = 0; l < 4; ++l) asm volatile("prefetch.global.L1 [%0];" ::"l"(blk0 + 128 * l));
  }
  const double* fi = st.fin + (size_t)b * FS;
  const double os = st.os[j_out];
  double macc[T];
  TriT<T> S;
#pragma unroll
  for (int r = 0; r < T; ++r) macc[r] = fi[r];
#pragma unroll
  for (int r = 0; r < T; ++r)
#pragma unroll
    for (int s = 0; s <= r; ++s) {
      double kss = 0.0;
      if (r == s) {
        if (r == 0) kss = os;
        else { const double il = 1.0 / st.ls[j_out * d + r - 1]; kss = __dmul_rn(os, __dmul_rn(il, il)); }  // never fused into the subtraction below
      }
      S.at(r, s) = kss - fi[T + r * (r + 1) / 2 + s];
    }
  double vr[T];
#pragma unroll
  for (int r = 0; r < T; ++r) {
    vr[r] = fmax(S.at(r, r), GP_MIN_VARIANCE);
    if (mean) mean[(size_t)b * T + r] = macc[r];
    if (var) var[(size_t)b * T + r] = vr[r];
  }
  if (!eps) return;

  // ---- E: draw ------------------------------------------------------------------------------------
  TriT<T> Lc;
  int level = 0;
  double yv[T];
  if (mode == FIN_APPEND) {
#pragma unroll
    for (int r = 0; r < T; ++r) yv[r] = y[(size_t)b * T + r];  // as drawn by the FIN_DRAW launch
  } else if (eig_redo) {
    double R[T * T];
    eig_root_T<T>(S.v, R);
    level = 4;
    for (int r = 0; r < T; ++r) {
      double acc = macc[r];
      for (int s = 0; s < T; ++s) acc += R[r * T + s] * eps[(size_t)b * T + s];
      yv[r] = acc;
    }
    if (b == 0) atomicOr(st.status, GPMPC_ST_SAMPLE_EIG);
  } else {
    if (T == 1) {
      Lc.v[0] = opts.unclamped_sqrt_1x1 ? sqrt(S.v[0]) : sqrt(fmax(S.v[0], 0.0));
    } else {
      bool ok = chol_T<T>(S, 0.0, Lc);
      double jit = st.jitter;
      while (!ok && level < GP_MAX_TRIES) {
        ++level;
        ok = chol_T<T>(S, jit, Lc);
        jit *= 10.0;
      }
      if (!ok) {
        level = 4;
        bool has_nan = false;  // NaN in the matrix: GPyTorch's NanError, no eigen fallback
#pragma unroll
        for (int i = 0; i < T * (T + 1) / 2; ++i) has_nan = has_nan || isnan(S.v[i]);
        if (has_nan) atomicOr(st.status, GPMPC_ST_NAN_INPUT);
        else if (!(opts.flags & GPMPC_OPT_NO_EIG_FALLBACK)) atomicMax(st.eig_flag, st.eig_epoch);
      }
    }
#pragma unroll
    for (int r = 0; r < T; ++r) {
      double acc = macc[r];
#pragma unroll
      for (int s = 0; s <= r; ++s) acc += Lc.at(r, s) * eps[(size_t)b * T + s];
      yv[r] = level < 4 ? acc : nan("");
    }
  }
  if (mode != FIN_APPEND) {
    bool zero = opts.variance_is_zero >= 0.0;
#pragma unroll
    for (int r = 0; r < T; ++r) zero = zero && (vr[r] <= opts.variance_is_zero);
#pragma unroll
    for (int r = 0; r < T; ++r) {
      if (zero) yv[r] = macc[r];
      if (opts.beta >= 0.0) {
        const double sd = sqrt(vr[r]);
        yv[r] = fmin(fmax(yv[r], macc[r] - opts.beta * sd), macc[r] + opts.beta * sd);
      }
      y[(size_t)b * T + r] = yv[r];
    }
    if (jitter_level) jitter_level[b] = level;
    if (level == 4 && !eig_redo) atomicOr(st.status, GPMPC_ST_SAMPLE_NOT_PD);
    if (mode == FIN_DRAW) return;
  }

  // ---- F (second half): condition on (x*, y) --------------------------------------------------------
  int decision = 0;
  bool flagged = false;
  if (mode == FIN_APPEND) {
    decision = ga.decision[b / ga.group_elems];
    flagged = ga.flag[b] != 0;
    st.pstate[(size_t)b * st.cap_points + st.np] = (unsigned char)decision;
  }
  for (int a = 0; a < d; ++a) st.Xh[((size_t)b * st.cap_points + st.np) * d + a] = x[(size_t)b * d + a];
#pragma unroll
  for (int r = 0; r < T; ++r) st.Yh[((size_t)b * st.cap_points + st.np) * T + r] = flagged ? nan("") : yv[r];
  if (b == 0) st.hrow0[st.np] = grow_factor ? c : -1;
  if (!grow_factor) return;
  if (decision != 0) {
    // NULL rows: the point stays out of this Agent's model (masked: a label of the Agent is NaN; dropped: filtered for all
    // samples).  Rows c .. c+T-1 = unit rows (zero off the diagonal, 1 / L_kk = 1, zero inverse entries, beta 0); phase A of
    // later steps gives them zero kernel entries (st.pstate), so w is 0 there and nothing downstream sees the point.
    double* Le0 = st.Lh + (size_t)b * st.elem_stride;
    for (int r = 0; r < T; ++r) {
      const int k = c + r, i = k & 7, kb = k - i;
      double* sp = Le0 + subpanel_off(k >> 3, mo);
      for (int t = 0; t < st.m; ++t) sp[sp_idx(t, i)] = 0.0;
      for (int t = 0; t < k; ++t) sp[sp_idx(mo + t, i)] = 0.0;
      sp[sp_idx(mo + k, i)] = 1.0;
      for (int jc = 0; jc < i; ++jc) sp[sp_idx(mo + kb + i, jc)] = 0.0;  // transposed-inverse slots (row jc, column i)
      st.beta_h[(size_t)b * st.c_cap + k] = 0.0;
    }
    if (b == 0)
      for (int r = 0; r < T; ++r) { st.hobs_pt[c + r] = st.np; st.hobs_task[c + r] = r; }
    return;
  }
  TriT<T> Sn = S, Ln;
#pragma unroll
  for (int r = 0; r < T; ++r) Sn.at(r, r) += st.noise[j_out * T + r];
  if (!chol_T<T>(Sn, 0.0, Ln)) atomicOr(st.status, GPMPC_ST_APPEND_NOT_PD);
  double bn[T], rdn[T];
#pragma unroll
  for (int r = 0; r < T; ++r) {
    double t = yv[r] - macc[r];
#pragma unroll
    for (int s = 0; s < r; ++s) t -= Ln.at(r, s) * bn[s];
    bn[r] = t / Ln.at(r, r);
    rdn[r] = 1.0 / Ln.at(r, r);
    st.beta_h[(size_t)b * st.c_cap + c + r] = bn[r];
  }
  if (b == 0) {
#pragma unroll
    for (int r = 0; r < T; ++r) {
      st.hobs_pt[c + r] = st.np;
      st.hobs_task[c + r] = r;
    }
  }
  // diagonal blocks touched by rows c .. c+T-1 (at most two): blk[col][row] mirrors the 8 x 8 block in memory
  // (lower: L, diagonal: 1/L_kk, upper: transposed inverse); the old rows' inverse and the new rows' entries under
  // old columns (written by k_step) are read back, the new rows are completed and their slots written.
  double* Le = st.Lh + (size_t)b * st.elem_stride;
  if constexpr (IFC >= 0) {
    // c mod 8 known at compile time (the host dispatches on it: c is uniform over the launch): the block lives in registers
    constexpr int R1 = T < 8 - IFC ? T : 8 - IFC;
    finish_block_const<T, IFC, 0, R1>(Le, c, mo, Ln, rdn);
    if constexpr (R1 < T) finish_block_const<T, 0, R1, T>(Le, c, mo, Ln, rdn);
  } else {
    int r0 = 0;
    while (r0 < T) {
      const int kb = (c + r0) & ~7;                    // first own row of this block
      const int i_first = c + r0 - kb;                 // block row of the first new row
      const int r1 = min(T, r0 + 8 - i_first);         // new rows [r0, r1) fall into this block
      double* gblk = Le + subpanel_off(kb >> 3, mo) + (size_t)(mo + kb) * 8;
      double blk[8][8];
      for (int t = 0; t < i_first; ++t)                // old columns: old rows' slots and the new rows' L entries
        for (int i = 0; i < i_first + (r1 - r0); ++i) blk[t][i] = gblk[sp_idx(t, i)];
      for (int r = r0; r < r1; ++r) {
        const int i = i_first + (r - r0);
        for (int s = r0; s < r; ++s) blk[i_first + (s - r0)][i] = Ln.v[r * (r + 1) / 2 + s];
        blk[i][i] = rdn[r];
        for (int jc = 0; jc < i; ++jc) {
          double acc = 0.0;
          for (int t = jc; t < i; ++t) acc = fma(blk[t][i], blk[t][jc], acc);
          blk[i][jc] = -rdn[r] * acc;                  // slot (row jc, column i)
        }
        double* rowi = Le + subpanel_off(kb >> 3, mo);
        for (int s = 0; s < r0; ++s) rowi[sp_idx(mo + c + s, i)] = Ln.v[r * (r + 1) / 2 + s];  // new columns of an earlier block
        for (int t = i_first; t <= i; ++t) gblk[sp_idx(t, i)] = blk[t][i];   // new columns of row i (incl. 1/L_ii)
        for (int jc = 0; jc < i; ++jc) gblk[sp_idx(i, jc)] = blk[i][jc];     // transposed inverse row
      }
      r0 = r1;
    }
  }
}

// K1b: the fused step for training sets too large for the per-warp w array of k_step (m in the thousands and up; the shared
// rows w_o = inv(L_oo) k_o of every element come from the batched GEMM k_shared_rows, st.Wo).  One CTA per batch element, w_o
// stays in global memory (L2), only the own rows' part of w lives in shared memory; plain FP64 FMAs: for such m the step is the
// GEMM (T m^2 flops per element) and this kernel only streams the element's c x (m + c) own rows once (memory bound).
//   A  kernel entries of the hallucinated points                              -> wh [c8][T] (shared)
//   C  own rows, sub-panel by sub-panel: dot = L[rows][cols < n_off] w (k-blocks over the warps, lane = one (row, column mod 4)
//      of a k-block, 256-byte coalesced loads), rhs = k - dot, 8 x 8 block by substitution
//   D  W^T [W | beta] over the m shared rows (global) and the c own rows       -> st.fin (k_step's layout; k_step_finish follows)
//   F  (first half) the new rows' entries left of their diagonal block = w
#define BIG_THREADS 256
__global__ void __launch_bounds__(BIG_THREADS) k_step_big(DevState st, const double* __restrict__ x, int grow_factor) {
  extern __shared__ __align__(16) double bsm[];
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const int j = b % st.g_ny, d = st.d, T = st.T, m = st.m, mo = st.mo, c = st.c, np = st.np;
  const int P8 = (c + 7) >> 3, c8 = 8 * P8;
  const int FS = T + T * (T + 1) / 2;
  double* wh = bsm;                          // [c8][T]   k, then w of the own rows
  double* red = wh + (size_t)max(c8, 1) * T; // [nw][8][T] partial dots of a sub-panel / [nw][FS] partial moments
  const double* wo = st.Wo + (size_t)b * mo * T;
  const double* xb = x + (size_t)b * d;
  const double* ls = st.ls + j * d;
  const double os = st.os[j];
  double* Le = st.Lh + (size_t)b * st.elem_stride;

  for (int idx = tid; idx < c8 * T; idx += nt) wh[idx] = 0.0;
  __syncthreads();
  for (int idx = tid; idx < np * T * T; idx += nt) {
    const int p = idx / (T * T), ta = (idx / T) % T, tb = idx % T;
    const int r0 = st.hrow0[p];
    if (r0 < 0) continue;
    wh[(r0 + ta) * T + tb] = cov_scalar(st.Xh + ((size_t)b * st.cap_points + p) * d, ta, xb, tb, ls, os, d);
  }
  __syncthreads();

  // lane <-> element of a k-block: [row 0..7][column 0..3]  (sp_idx)
  const int l_col = lane & 3, l_row = lane >> 2;
  for (int p8 = 0; p8 < P8; ++p8) {
    const int n_off = mo + 8 * p8, nkb = n_off >> 2;
    const double* sp = Le + subpanel_off(p8, mo);
    double acc[GPMPC_MAX_T];
#pragma unroll
    for (int t = 0; t < GPMPC_MAX_T; ++t) acc[t] = 0.0;
    for (int kb = warp; kb < nkb; kb += nw) {
      const double lv = __ldcg(sp + (size_t)kb * 32 + lane);
      const int col = 4 * kb + l_col;
      const double* wr = col < mo ? wo + (size_t)col * T : wh + (size_t)(col - mo) * T;
#pragma unroll
      for (int t = 0; t < GPMPC_MAX_T; ++t)
        if (t < T) acc[t] = fma(lv, wr[t], acc[t]);
    }
#pragma unroll
    for (int t = 0; t < GPMPC_MAX_T; ++t) {
      if (t < T) {
        double v = acc[t];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (l_col == 0) red[((size_t)warp * 8 + l_row) * T + t] = v;
      }
    }
    __syncthreads();
    if (tid < T) {
      // rhs = k - dot, then the 8 x 8 diagonal block by substitution (lower part: L, diagonal slot: 1 / L_kk); one thread per
      // right-hand side
      const int t = tid, nvalid = min(8, c - 8 * p8);
      const double* blk = sp + (size_t)n_off * 8;
      double w8[8];
      for (int i = 0; i < 8; ++i) {
        double dot = 0.0;
        for (int w = 0; w < nw; ++w) dot += red[((size_t)w * 8 + i) * T + t];
        double v = wh[(8 * p8 + i) * T + t] - dot;
        for (int k = 0; k < i; ++k) v = fma(-__ldcg(blk + sp_idx(k, i)), w8[k], v);
        w8[i] = i < nvalid ? v * __ldcg(blk + sp_idx(i, i)) : 0.0;
      }
      for (int i = 0; i < 8; ++i) wh[(8 * p8 + i) * T + t] = w8[i];
    }
    __syncthreads();
  }

  // ---- D: moments ----
  {
    double mom[GPMPC_MAX_T + GPMPC_MAX_T * (GPMPC_MAX_T + 1) / 2];
#pragma unroll
    for (int q = 0; q < GPMPC_MAX_T + GPMPC_MAX_T * (GPMPC_MAX_T + 1) / 2; ++q) mom[q] = 0.0;
    const double* beta_o = st.beta_o + (size_t)j * m;
    const double* beta_h = st.beta_h + (size_t)b * st.c_cap;
    for (int row = tid; row < m + c; row += nt) {
      const double* wr = row < m ? wo + (size_t)row * T : wh + (size_t)(row - m) * T;
      const double be = row < m ? beta_o[row] : beta_h[row - m];
      double wv_[GPMPC_MAX_T];
#pragma unroll
      for (int t = 0; t < GPMPC_MAX_T; ++t) wv_[t] = t < T ? wr[t] : 0.0;
#pragma unroll
      for (int r = 0; r < GPMPC_MAX_T; ++r) {
        if (r < T) {
          mom[r] = fma(wv_[r], be, mom[r]);
#pragma unroll
          for (int s2 = 0; s2 <= r; ++s2) mom[GPMPC_MAX_T + r * (r + 1) / 2 + s2] = fma(wv_[r], wv_[s2], mom[GPMPC_MAX_T + r * (r + 1) / 2 + s2]);
        }
      }
    }
    // block reduction of the FS moments
    for (int r = 0; r < T; ++r) {
      double v = warp_sum(mom[r]);
      if (lane == 0) red[(size_t)warp * FS + r] = v;
      for (int s2 = 0; s2 <= r; ++s2) {
        double u = warp_sum(mom[GPMPC_MAX_T + r * (r + 1) / 2 + s2]);
        if (lane == 0) red[(size_t)warp * FS + T + r * (r + 1) / 2 + s2] = u;
      }
    }
    __syncthreads();
    if (tid < FS) {
      double v = 0.0;
      for (int w = 0; w < nw; ++w) v += red[(size_t)w * FS + tid];
      st.fin[(size_t)b * FS + tid] = v;
    }
  }
  if (!grow_factor) return;
  // ---- F (first half): L[c + r][t] = w[t][r] for the shared columns and the own columns left of the new block ----
  for (int t = tid; t < m + c; t += nt) {
    const double* wr = t < m ? wo + (size_t)t * T : wh + (size_t)(t - m) * T;
    const int tt = t < m ? t : mo + (t - m);
    for (int r = 0; r < T; ++r) Le[subpanel_off((c + r) >> 3, mo) + sp_idx(tt, (c + r) & 7)] = wr[r];
  }
}
