// Eigen-root fallback of the draw (GPyTorch: MultivariateNormal.rsample -> lazy_covariance_matrix.root_decomposition()):
// when the joint Cholesky of ANY batch element still fails after psd_safe_cholesky's jitter ladder, linear_operator
// catches the NotPSDError and takes  root = evecs * sqrt(clamp(evals, 0))  of torch.linalg.eigh for the WHOLE batch
// ("Runtime Error when computing Cholesky decomposition ... Using symeig method"); y = mean + root eps with eps paired to
// the eigenvalues in ascending order.  Call site: src/agent.py:641 (.sample(base_samples=...)); the car yamls set
// Dyn_gp_jitter 1e-20 (params_car_residual.yaml:51), which makes the ladder a no-op, so there this IS the draw whenever a
// joint covariance is numerically singular.
//
// Protocol (no host round trip): every draw launch carries an epoch; a failing element does atomicMax(eig_flag, epoch);
// the eigen kernel is launched right behind the draw kernel and returns at once unless eig_flag == epoch.
//
// Eigenvectors are determined up to sign (and up to a rotation inside a cluster of equal eigenvalues); LAPACK applies no
// convention, so two eigh implementations agree on root root^T, not on root.  Here every eigenvector is normalised to
// have its largest-magnitude component positive (deterministic; the tests compare draws modulo that sign).
#pragma once
#include "gpmpc_block.cuh"

#define EIG_THREADS 256
#define EIG_MAX_SWEEPS 40
#define EIG_MAX_PAIRS 1024  // q <= 2048

__device__ __forceinline__ void jacobi_cs(double app, double aqq, double apq, double floor_, double& c, double& s) {
  c = 1.0;
  s = 0.0;
  if (fabs(apq) <= 2.3e-16 * sqrt(fabs(app) * fabs(aqq)) || fabs(apq) <= floor_) return;
  const double theta = (aqq - app) / (2.0 * apq);
  const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
  c = 1.0 / sqrt(t * t + 1.0);
  s = t * c;
}

// Parallel two-sided cyclic Jacobi (round-robin ordering: q/2 disjoint rotations per round, q-1 rounds per sweep).
// A (q x q, leading dim lda, FULL symmetric storage; shared or global memory) is overwritten, its diagonal ends up
// holding the eigenvalues; V (q x q, leading dim ldv, global) the eigenvectors in columns.
__device__ void block_jacobi_eigh(double* A, int lda, double* V, int ldv, int q) {
  __shared__ int sp[EIG_MAX_PAIRS], sq[EIG_MAX_PAIRS];
  __shared__ double sc[EIG_MAX_PAIRS], ss[EIG_MAX_PAIRS];
  __shared__ int s_rot;
  __shared__ double s_floor;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int n2 = (q + 1) & ~1, np = n2 >> 1;
  for (int idx = tid; idx < q * q; idx += nt) V[(size_t)(idx / q) * ldv + idx % q] = (idx / q == idx % q) ? 1.0 : 0.0;
  if (tid == 0) {
    double mx = 0.0;
    for (int i = 0; i < q; ++i) mx = fmax(mx, fabs(A[(size_t)i * lda + i]));
    s_floor = 1e-22 * mx + 1e-300;
  }
  __syncthreads();
  for (int sweep = 0; sweep < EIG_MAX_SWEEPS; ++sweep) {
    if (tid == 0) s_rot = 0;
    __syncthreads();
    for (int r = 0; r < n2 - 1; ++r) {
      for (int i = tid; i < np; i += nt) {
        int a = (r + i) % (n2 - 1);
        int b = i == 0 ? n2 - 1 : (r + n2 - 1 - i) % (n2 - 1);
        if (a > b) { const int t = a; a = b; b = t; }
        double c = 1.0, s = 0.0;
        if (b < q) jacobi_cs(A[(size_t)a * lda + a], A[(size_t)b * lda + b], A[(size_t)a * lda + b], s_floor, c, s);
        sp[i] = a; sq[i] = b; sc[i] = c; ss[i] = s;
        if (s != 0.0) s_rot = 1;  // benign race: every writer stores 1
      }
      __syncthreads();
      // columns p, q of A and V:  [x_p, x_q] <- [c x_p - s x_q, s x_p + c x_q]
      for (int idx = tid; idx < np * q; idx += nt) {
        const int i = idx % np, k = idx / np;
        const double s = ss[i];
        if (s == 0.0) continue;
        const double c = sc[i];
        const int p = sp[i], qq = sq[i];
        double* ar = A + (size_t)k * lda;
        const double ap = ar[p], aq = ar[qq];
        ar[p] = c * ap - s * aq;
        ar[qq] = s * ap + c * aq;
        double* vr = V + (size_t)k * ldv;
        const double vp = vr[p], vq = vr[qq];
        vr[p] = c * vp - s * vq;
        vr[qq] = s * vp + c * vq;
      }
      __syncthreads();
      // rows p, q of A
      for (int idx = tid; idx < np * q; idx += nt) {
        const int i = idx / q, k = idx % q;
        const double s = ss[i];
        if (s == 0.0) continue;
        const double c = sc[i];
        double* rp = A + (size_t)sp[i] * lda;
        double* rq = A + (size_t)sq[i] * lda;
        const double ap = rp[k], aq = rq[k];
        rp[k] = c * ap - s * aq;
        rq[k] = s * ap + c * aq;
      }
      __syncthreads();
    }
    if (!s_rot) break;  // a whole sweep without a rotation: converged
    __syncthreads();
  }
}

// The draw through the eigen root, then sample_gp's post-processing -- the counterpart of block_sample for the whole
// batch once any element's ladder failed.  A_smem != NULL: q*q doubles of shared memory for the iteration matrix.
__global__ void __launch_bounds__(EIG_THREADS)
k_sample_eig(DevState st, int H, const double* __restrict__ eps, gpmpc_sample_opts opts, double* __restrict__ y,
             int* __restrict__ jitter_level, int a_in_smem) {
  extern __shared__ __align__(16) double eig_smem[];
  if (*(volatile int*)st.eig_flag != st.eig_epoch) return;  // nobody failed in this call: the Cholesky draw stands
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int T = st.T, q = H * T;
  const double* S = st.S + (size_t)b * q * q;
  double* V = st.C + (size_t)b * q * q;
  double* A = a_in_smem ? eig_smem : st.E + (size_t)b * q * q;
  const double* mu = st.mu + (size_t)b * q;
  const double* e = eps + (size_t)b * q;
  double* yb = y + (size_t)b * q;
  // torch.linalg.eigh reads the LOWER triangle (UPLO = 'L')
  for (int idx = tid; idx < q * q; idx += nt) {
    const int r = idx / q, s = idx % q;
    A[idx] = r >= s ? S[(size_t)r * q + s] : S[(size_t)s * q + r];
  }
  __syncthreads();
  if (q == 1) {
    if (tid == 0) { V[0] = 1.0; }
  } else {
    block_jacobi_eigh(A, q, V, q, q);
  }
  __syncthreads();
  // ascending order of the eigenvalues (rank of column j), sign convention, scaled base sample of each column:
  // coef[j] = sgn_j sqrt(max(lambda_j, 0)) eps[rank_j]   (kept in the dead upper part of the work matrix: row 0 is done with)
  __shared__ double s_coef[2 * EIG_MAX_PAIRS];
  for (int j = tid; j < q; j += nt) {
    const double lj = A[(size_t)j * q + j];
    int rank = 0;
    for (int k = 0; k < q; ++k) {
      const double lk = A[(size_t)k * q + k];
      rank += (lk < lj || (lk == lj && k < j)) ? 1 : 0;
    }
    double best = 0.0, sgn = 1.0;
    for (int k = 0; k < q; ++k) {
      const double v = V[(size_t)k * q + j];
      if (fabs(v) > best) { best = fabs(v); sgn = v < 0.0 ? -1.0 : 1.0; }
    }
    s_coef[j] = sgn * sqrt(fmax(lj, 0.0)) * e[rank];
  }
  __syncthreads();
  for (int r = tid; r < q; r += nt) {
    double acc = mu[r];
    const double* vr = V + (size_t)r * q;
    for (int j = 0; j < q; ++j) acc = fma(vr[j], s_coef[j], acc);
    yb[r] = acc;
  }
  if (tid == 0) {
    if (jitter_level) jitter_level[b] = 4;
    if (b == 0) atomicOr(st.status, GPMPC_ST_SAMPLE_EIG);
  }
  __syncthreads();
  block_postprocess(st, b, H, opts, yb);
}

// ---- T x T version for the fused step (one thread per element; rare path, kept out of line) -------------------------
// S: packed lower triangle (r(r+1)/2 + s).  R: full T x T root, R[r][k] multiplies eps[k] (ascending eigenvalues).
template <int T>
__device__ __noinline__ void eig_root_T(const double* S, double* R) {
  double a[T][T], v[T][T];
  for (int r = 0; r < T; ++r)
    for (int s = 0; s < T; ++s) {
      a[r][s] = r >= s ? S[r * (r + 1) / 2 + s] : S[s * (s + 1) / 2 + r];
      v[r][s] = r == s ? 1.0 : 0.0;
    }
  double mx = 0.0;
  for (int i = 0; i < T; ++i) mx = fmax(mx, fabs(a[i][i]));
  const double floor_ = 1e-22 * mx + 1e-300;
  for (int sweep = 0; sweep < EIG_MAX_SWEEPS; ++sweep) {
    bool rot = false;
    for (int p = 0; p < T - 1; ++p)
      for (int qq = p + 1; qq < T; ++qq) {
        double c, s;
        jacobi_cs(a[p][p], a[qq][qq], a[p][qq], floor_, c, s);
        if (s == 0.0) continue;
        rot = true;
        for (int k = 0; k < T; ++k) {
          const double ap = a[k][p], aq = a[k][qq];
          a[k][p] = c * ap - s * aq;
          a[k][qq] = s * ap + c * aq;
          const double vp = v[k][p], vq = v[k][qq];
          v[k][p] = c * vp - s * vq;
          v[k][qq] = s * vp + c * vq;
        }
        for (int k = 0; k < T; ++k) {
          const double ap = a[p][k], aq = a[qq][k];
          a[p][k] = c * ap - s * aq;
          a[qq][k] = s * ap + c * aq;
        }
      }
    if (!rot) break;
  }
  for (int j = 0; j < T; ++j) {
    const double lj = a[j][j];
    int rank = 0;
    for (int k = 0; k < T; ++k) rank += (a[k][k] < lj || (a[k][k] == lj && k < j)) ? 1 : 0;
    double best = 0.0, sgn = 1.0;
    for (int k = 0; k < T; ++k)
      if (fabs(v[k][j]) > best) { best = fabs(v[k][j]); sgn = v[k][j] < 0.0 ? -1.0 : 1.0; }
    const double sc = sgn * sqrt(fmax(lj, 0.0));
    for (int r = 0; r < T; ++r) R[r * T + rank] = v[r][j] * sc;
  }
}
