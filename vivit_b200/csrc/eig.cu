// Symmetric eigensolver for the per-group Gram matrices: two-sided block Jacobi with a
// round-robin (tournament) parallel ordering.
//
//   * the matrix is cut into 16-wide blocks; in every round the nb blocks are paired
//     (nb/2 disjoint pairs), each pair's 32x32 diagonal block is diagonalised by one CTA with
//     a parallel-order cyclic Jacobi held in shared memory (jacobi_diag_kernel), giving an
//     orthogonal Q per pair;
//   * all off-diagonal 32x32 tiles are then updated independently, tile (a, b) <- Qa^T X Qb,
//     exploiting symmetry (only a < b is computed, the mirror tile is written transposed), and
//     the eigenvector accumulator J gets J[:, b] <- J[:, b] Qb (jacobi_tile_kernel);
//   * nb-1 rounds make a sweep; sweeps repeat until a whole sweep applies no rotation.
//
// Rotations are skipped when |a_pq| <= max(eps*sqrt(|a_pp a_qq|), eps*||G||_F/sqrt(R)), which bounds
// the final off-diagonal Frobenius norm by sqrt(R)*eps*||G||_F (LAPACK-grade absolute accuracy),
// terminates on the exactly rank-deficient Grams this path produces, and avoids spending sweeps
// on diagonalising the rounding noise that fills their null space.
// Pure Jacobi: converges on the reference's LAPACK-killer fixture without a diagonal shift.
#include <cstdlib>

#include "common.cuh"

namespace vvt {

constexpr int JB = 16;        // block width
constexpr int JT = 2 * JB;    // tile width (a pair of blocks)
constexpr int kMaxSweeps = 40;
constexpr int kMaxInnerSweeps = 12;

__host__ __device__ inline void rr_pair(int n, int round, int idx, int& a, int& b) {
  const int m = n - 1;
  if (idx == 0) {
    a = round;
    b = m;
  } else {
    a = (round + idx) % m;
    b = (round - idx + m) % m;
  }
}

__device__ __forceinline__ int tile_index(int bp, int bq, int i) {
  return i < JB ? bp * JB + i : bq * JB + (i - JB);
}

template <typename T>
struct Eps;
template <>
struct Eps<float> {
  static constexpr float v = 1.1920929e-7f;
};
template <>
struct Eps<double> {
  static constexpr double v = 2.220446049250313e-16;
};

struct JacobiScalars {
  double norm2;                 // ||G||_F^2
  unsigned long long rotations; // CTAs that rotated in the current sweep
};

// A = sym(G) padded with zeros, J = I, and ||G||_F^2
template <typename T>
__global__ void jacobi_init_kernel(T* A, T* J, const T* G, int64_t R, int64_t Rp, JacobiScalars* sc) {
  const int64_t total = Rp * Rp;
  double s = 0.0;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t i = idx / Rp, j = idx % Rp;
    T v = 0;
    if (i < R && j < R) v = (i <= j) ? G[i * R + j] : G[j * R + i];  // upper triangle, as symeig(upper=True)
    A[idx] = v;
    if (J) J[idx] = (i == j) ? T(1) : T(0);
    s += double(v) * double(v);
  }
  __shared__ double red[32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(&sc->norm2, s);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
jacobi_diag_kernel(T* A, T* dlo, T* Qbuf, int* flags, int Rp, int nb, int round, int64_t R, JacobiScalars* sc) {
  __shared__ T W[JT][JT + 1];
  __shared__ T Q[JT][JT + 1];
  __shared__ T rot[JB][4];  // s, tau = s/(1+c), new a_pp, new a_qq
  __shared__ T dl[JT];      // low words of the compensated diagonal
  __shared__ int pq[JB][2];
  const int tid = threadIdx.x;
  int bp, bq;
  rr_pair(nb, round, blockIdx.x, bp, bq);
  for (int idx = tid; idx < JT * JT; idx += 256) {
    const int i = idx / JT, j = idx % JT;
    W[i][j] = A[int64_t(tile_index(bp, bq, i)) * Rp + tile_index(bp, bq, j)];
    Q[i][j] = (i == j) ? T(1) : T(0);
  }
  if (tid < JT) dl[tid] = dlo[tile_index(bp, bq, tid)];
  const T thr_abs = T(Eps<T>::v * sqrt(sc->norm2 / double(R)));
  __syncthreads();

  // Rotations are applied in Rutishauser's form  x' = x - s (y + tau x),  y' = y + s (x - tau y):
  // cos(theta) - 1 = -s tau is never rounded to zero, so thousands of small-angle rotations do not
  // inflate the Frobenius norm (c = 1/sqrt(1+t^2) rounds to 1 for |t| < sqrt(eps)).
  bool rotated_any = false;
  for (int sweep = 0; sweep < kMaxInnerSweeps; ++sweep) {
    bool rotated_sweep = false;
    for (int rr = 0; rr < JT - 1; ++rr) {
      int did = 0;
      if (tid < JB) {
        int p, q;
        rr_pair(JT, rr, tid, p, q);
        const T app = W[p][p], aqq = W[q][q], apq = W[p][q];
        T sn = 0, tau = 0, dpp = app, dqq = aqq;
        const T lim = max(thr_abs, Eps<T>::v * sqrt(fabs(app) * fabs(aqq)));
        if (fabs(apq) > lim) {
          const T theta = (aqq - app) / (T(2) * apq);
          const T t = (theta >= T(0) ? T(1) : T(-1)) / (fabs(theta) + sqrt(T(1) + theta * theta));
          const T c = T(1) / sqrt(T(1) + t * t);
          sn = t * c;
          tau = sn / (T(1) + c);
          // The diagonal is kept as an unevaluated sum hi + lo (error-free TwoSum): rotations move
          // t*a_pq from the smaller to the larger pivot, and without the low word every increment
          // below half an ulp of the large one would be dropped -- a systematic loss of trace.
          const T delta = t * apq;
          const T xp = dl[p] - delta, xq = dl[q] + delta;
          dpp = app + xp;
          dqq = aqq + xq;
          T bb = dpp - app;
          dl[p] = (app - (dpp - bb)) + (xp - bb);
          bb = dqq - aqq;
          dl[q] = (aqq - (dqq - bb)) + (xq - bb);
          did = 1;
        }
        rot[tid][0] = sn;
        rot[tid][1] = tau;
        rot[tid][2] = dpp;
        rot[tid][3] = dqq;
        pq[tid][0] = p;
        pq[tid][1] = q;
      }
      if (!__syncthreads_or(did)) continue;  // uniform
      rotated_sweep = true;
      // columns: W <- W Jrot, Q <- Q Jrot
      for (int it = tid; it < JT * JB; it += 256) {
        const int k = it & (JT - 1), t = it / JT;
        const T sn = rot[t][0], tau = rot[t][1];
        if (sn != T(0)) {
          const int p = pq[t][0], q = pq[t][1];
          const T wp = W[k][p], wq = W[k][q];
          W[k][p] = wp - sn * (wq + tau * wp);
          W[k][q] = wq + sn * (wp - tau * wq);
          const T qp = Q[k][p], qq = Q[k][q];
          Q[k][p] = qp - sn * (qq + tau * qp);
          Q[k][q] = qq + sn * (qp - tau * qq);
        }
      }
      __syncthreads();
      // rows: W <- Jrot^T W; the 2x2 pivot block gets its closed-form values
      for (int it = tid; it < JT * JB; it += 256) {
        const int k = it & (JT - 1), t = it / JT;
        const T sn = rot[t][0], tau = rot[t][1];
        if (sn != T(0)) {
          const int p = pq[t][0], q = pq[t][1];
          const T wp = W[p][k], wq = W[q][k];
          T np = wp - sn * (wq + tau * wp), nq = wq + sn * (wp - tau * wq);
          if (k == p) np = rot[t][2], nq = T(0);
          if (k == q) np = T(0), nq = rot[t][3];
          W[p][k] = np;
          W[q][k] = nq;
        }
      }
      __syncthreads();
    }
    if (!rotated_sweep) break;
    rotated_any = true;
  }
  // write back the (now diagonal up to threshold) block and this pair's rotation
  if (rotated_any) {
    for (int idx = tid; idx < JT * JT; idx += 256) {
      const int i = idx / JT, j = idx % JT;
      A[int64_t(tile_index(bp, bq, i)) * Rp + tile_index(bp, bq, j)] = W[i][j];
      Qbuf[int64_t(blockIdx.x) * JT * JT + idx] = Q[i][j];
    }
    if (tid < JT) dlo[tile_index(bp, bq, tid)] = dl[tid];
  }
  if (tid == 0) {
    flags[blockIdx.x] = rotated_any ? 1 : 0;
    if (rotated_any) atomicAdd(&sc->rotations, 1ull);
  }
}

// 32x32x32 product helpers on shared tiles (256 threads, 4 outputs each)
template <typename T>
__device__ __forceinline__ void tile_mul_nn(T (*Out)[JT + 1], const T (*X)[JT + 1], const T (*Qm)[JT + 1], int tid) {
  const int i = tid >> 3, j0 = (tid & 7) * 4;
  T a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll 8
  for (int k = 0; k < JT; ++k) {
    const T x = X[i][k];
    a0 += x * Qm[k][j0];
    a1 += x * Qm[k][j0 + 1];
    a2 += x * Qm[k][j0 + 2];
    a3 += x * Qm[k][j0 + 3];
  }
  Out[i][j0] = a0;
  Out[i][j0 + 1] = a1;
  Out[i][j0 + 2] = a2;
  Out[i][j0 + 3] = a3;
}
// Out = Qm^T X
template <typename T>
__device__ __forceinline__ void tile_mul_tn(T (*Out)[JT + 1], const T (*Qm)[JT + 1], const T (*X)[JT + 1], int tid) {
  const int i = tid >> 3, j0 = (tid & 7) * 4;
  T a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll 8
  for (int k = 0; k < JT; ++k) {
    const T q = Qm[k][i];
    a0 += q * X[k][j0];
    a1 += q * X[k][j0 + 1];
    a2 += q * X[k][j0 + 2];
    a3 += q * X[k][j0 + 3];
  }
  Out[i][j0] = a0;
  Out[i][j0 + 1] = a1;
  Out[i][j0 + 2] = a2;
  Out[i][j0 + 3] = a3;
}

template <typename T>
__global__ void __launch_bounds__(256)
jacobi_tile_kernel(T* A, T* J, const T* Qbuf, const int* flags, int Rp, int nb, int round) {
  __shared__ T X[JT][JT + 1];
  __shared__ T Qa[JT][JT + 1];
  __shared__ T Qb[JT][JT + 1];
  __shared__ T Tm[JT][JT + 1];
  const int half = nb / 2, tid = threadIdx.x;
  const int pc = blockIdx.x, y = blockIdx.y;
  int cp, cq;
  rr_pair(nb, round, pc, cp, cq);
  if (y < half) {
    const int pr = y;
    if (pr >= pc) return;
    const int fa = flags[pr], fb = flags[pc];
    if (!fa && !fb) return;
    int rp, rq;
    rr_pair(nb, round, pr, rp, rq);
    for (int idx = tid; idx < JT * JT; idx += 256) {
      const int i = idx / JT, j = idx % JT;
      X[i][j] = A[int64_t(tile_index(rp, rq, i)) * Rp + tile_index(cp, cq, j)];
      Qa[i][j] = fa ? Qbuf[int64_t(pr) * JT * JT + idx] : (i == j ? T(1) : T(0));
      Qb[i][j] = fb ? Qbuf[int64_t(pc) * JT * JT + idx] : (i == j ? T(1) : T(0));
    }
    __syncthreads();
    tile_mul_nn<T>(Tm, X, Qb, tid);
    __syncthreads();
    tile_mul_tn<T>(X, Qa, Tm, tid);
    __syncthreads();
    for (int idx = tid; idx < JT * JT; idx += 256) {
      const int i = idx / JT, j = idx % JT;
      A[int64_t(tile_index(rp, rq, i)) * Rp + tile_index(cp, cq, j)] = X[i][j];
    }
    for (int idx = tid; idx < JT * JT; idx += 256) {
      const int j = idx / JT, i = idx % JT;  // i fastest: coalesced rows of the mirror tile
      A[int64_t(tile_index(cp, cq, j)) * Rp + tile_index(rp, rq, i)] = X[i][j];
    }
  } else {
    if (!J || !flags[pc]) return;
    const int64_t r0 = int64_t(y - half) * JT;
    for (int idx = tid; idx < JT * JT; idx += 256) {
      const int i = idx / JT, j = idx % JT;
      X[i][j] = J[(r0 + i) * Rp + tile_index(cp, cq, j)];
      Qb[i][j] = Qbuf[int64_t(pc) * JT * JT + idx];
    }
    __syncthreads();
    tile_mul_nn<T>(Tm, X, Qb, tid);
    __syncthreads();
    for (int idx = tid; idx < JT * JT; idx += 256) {
      const int i = idx / JT, j = idx % JT;
      J[(r0 + i) * Rp + tile_index(cp, cq, j)] = Tm[i][j];
    }
  }
}

template <typename T>
__global__ void jacobi_rank_kernel(int* rank, T* ev, const T* A, const T* dlo, int64_t R, int64_t Rp) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= R) return;
  const T vi = A[i * Rp + i] + dlo[i];
  ev[i] = vi;
  int r = 0;
  for (int64_t j = 0; j < R; ++j) {
    const T vj = ldg(A + j * Rp + j) + ldg(dlo + j);
    r += (vj < vi) || (vj == vi && j < i) || (vj != vj && vi == vi);  // NaNs sort first, stable
  }
  if (vi != vi) {  // NaN: order among NaNs by index
    r = 0;
    for (int64_t j = 0; j < i; ++j) {
      const T vj = ldg(A + j * Rp + j) + ldg(dlo + j);
      r += (vj != vj);
    }
  }
  rank[i] = r;
}

template <typename T>
__global__ void jacobi_permute_kernel(T* evals, T* evecs, const T* ev, const T* J, const int* rank,
                                      int64_t R, int64_t Rp) {
  const int64_t total = evecs ? R * R : R;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = idx / R, i = idx % R;
    const int dst = rank[i];
    if (r == 0) evals[dst] = ev[i];
    if (evecs) evecs[r * R + dst] = J[r * Rp + i];
  }
}

struct JacobiLayout {
  int64_t Rp, nb, off_A, off_J, off_Q, off_ev, off_dlo, off_rank, off_flags, off_sc, total;
};

static JacobiLayout jacobi_layout(int64_t R, int jobz, int64_t es) {
  JacobiLayout L;
  L.Rp = align_up(R, JT);
  L.nb = L.Rp / JB;
  int64_t o = 0;
  auto take = [&](int64_t bytes) {
    const int64_t at = o;
    o += align_up(bytes, 256);
    return at;
  };
  L.off_A = take(L.Rp * L.Rp * es);
  L.off_J = jobz ? take(L.Rp * L.Rp * es) : -1;
  L.off_Q = take((L.nb / 2) * JT * JT * es);
  L.off_ev = take(L.Rp * es);
  L.off_dlo = take(L.Rp * es);
  L.off_rank = take(L.Rp * 4);
  L.off_flags = take((L.nb / 2) * 4);
  L.off_sc = take(sizeof(JacobiScalars));
  L.total = o;
  return L;
}

template <typename T>
static int syevj_impl(T* evals, T* evecs, const T* G, int64_t R, int jobz, char* ws, int* info,
                      cudaStream_t s) {
  const JacobiLayout L = jacobi_layout(R, jobz, sizeof(T));
  T* A = (T*)(ws + L.off_A);
  T* J = jobz ? (T*)(ws + L.off_J) : nullptr;
  T* Qbuf = (T*)(ws + L.off_Q);
  T* ev = (T*)(ws + L.off_ev);
  T* dlo = (T*)(ws + L.off_dlo);
  int* rank = (int*)(ws + L.off_rank);
  int* flags = (int*)(ws + L.off_flags);
  JacobiScalars* sc = (JacobiScalars*)(ws + L.off_sc);
  const int Rp = int(L.Rp), nb = int(L.nb), half = nb / 2;

  VVT_TRY(check_cuda(cudaMemsetAsync(sc, 0, sizeof(JacobiScalars), s), "vvt_syevj"));
  VVT_TRY(check_cuda(cudaMemsetAsync(dlo, 0, L.Rp * sizeof(T), s), "vvt_syevj"));
  const bool debug = getenv("VVT_SYEVJ_DEBUG") != nullptr;
  const int init_blocks = int(vmin<int64_t>(ceil_div(L.Rp * L.Rp, 256), 8 * num_sms()));
  jacobi_init_kernel<T><<<init_blocks, 256, 0, s>>>(A, J, G, R, L.Rp, sc);
  VVT_TRY(launched("vvt_syevj(init)"));

  int sweeps = 0, converged = 0;
  for (; sweeps < kMaxSweeps;) {
    VVT_TRY(check_cuda(cudaMemsetAsync(&sc->rotations, 0, sizeof(unsigned long long), s), "vvt_syevj"));
    for (int round = 0; round < nb - 1; ++round) {
      jacobi_diag_kernel<T><<<half, 256, 0, s>>>(A, dlo, Qbuf, flags, Rp, nb, round, R, sc);
      VVT_TRY(launched("vvt_syevj(diag)"));
      const unsigned gy = unsigned(half + (jobz ? Rp / JT : 0));
      if (half > 1 || jobz) {
        jacobi_tile_kernel<T><<<dim3(half, gy), 256, 0, s>>>(A, J, Qbuf, flags, Rp, nb, round);
        VVT_TRY(launched("vvt_syevj(tile)"));
      }
    }
    ++sweeps;
    unsigned long long rot = 0;
    VVT_TRY(check_cuda(cudaMemcpyAsync(&rot, &sc->rotations, sizeof(rot), cudaMemcpyDeviceToHost, s),
                       "vvt_syevj"));
    VVT_TRY(check_cuda(cudaStreamSynchronize(s), "vvt_syevj"));
    if (debug) fprintf(stderr, "[vvt_syevj] R=%lld sweep %d: %llu of %d block solves rotated\n", (long long)R, sweeps, rot, half * (nb - 1));
    if (rot == 0) {
      converged = 1;
      break;
    }
  }
  jacobi_rank_kernel<T><<<unsigned(ceil_div(R, 128)), 128, 0, s>>>(rank, ev, A, dlo, R, L.Rp);
  VVT_TRY(launched("vvt_syevj(rank)"));
  const int64_t total = jobz ? R * R : R;
  const int pblocks = int(vmin<int64_t>(ceil_div(total, 256), 8 * num_sms()));
  jacobi_permute_kernel<T><<<pblocks, 256, 0, s>>>(evals, jobz ? evecs : nullptr, ev, J, rank, R, L.Rp);
  VVT_TRY(launched("vvt_syevj(permute)"));
  if (info) {
    info[0] = sweeps;
    info[1] = converged;
  }
  return VVT_OK;
}

}  // namespace vvt

using namespace vvt;

extern "C" {

int64_t vvt_syevj_workspace_bytes(int64_t R, int jobz, int dtype) {
  if (R <= 0) return 0;
  return jacobi_layout(R, jobz, dtype == VVT_F32 ? 4 : 8).total;
}

int vvt_syevj(void* evals, void* evecs, const void* G, int64_t R, int jobz, void* workspace,
              int64_t workspace_bytes, int* info_host, int dtype, void* stream) {
  VVT_REQUIRE(R >= 0, "negative size");
  if (info_host) info_host[0] = 0, info_host[1] = 1;
  if (R == 0) return VVT_OK;
  VVT_REQUIRE(R <= (int64_t(1) << 20), "matrix too large");
  VVT_REQUIRE(evals && G && workspace && (!jobz || evecs), "null pointer");
  if (workspace_bytes < vvt_syevj_workspace_bytes(R, jobz, dtype))
    return fail(VVT_ERR_WORKSPACE, "%s: workspace too small", __func__);
  VVT_DISPATCH(dtype, {
    return syevj_impl<T>((T*)evals, (T*)evecs, (const T*)G, R, jobz, (char*)workspace, info_host,
                         as_stream(stream));
  });
}

int vvt_syevj_batched(void* const* evals, void* const* evecs, const void* const* G,
                      const int64_t* R, int64_t batch, int jobz, void* workspace,
                      int64_t workspace_bytes, int* info_host, int dtype, void* stream) {
  VVT_REQUIRE(batch >= 0, "negative size");
  VVT_REQUIRE(batch == 0 || (evals && G && R), "null pointer");
  for (int64_t b = 0; b < batch; ++b) {
    VVT_TRY(vvt_syevj(evals[b], jobz ? evecs[b] : nullptr, G[b], R[b], jobz, workspace,
                      workspace_bytes, info_host ? info_host + 2 * b : nullptr, dtype, stream));
  }
  return VVT_OK;
}

}  // extern "C"
