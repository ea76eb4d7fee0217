// Symmetric eigensolver for the per-group Gram matrices: one-sided (Hestenes) block Jacobi.
//
// The Gram matrices of this path are positive semi-definite, so G = U diag(lambda) U^T is also
// the SVD of G.  The solver orthogonalises the columns of W = G J by plane rotations applied from
// the right, accumulating the same rotations in J (J = I at the start):  at convergence the
// columns of W are mutually orthogonal, W = U diag(lambda), and J = U.
//
//   * columns are cut into blocks of OB = 16; a round pairs the nb blocks into nb/2 disjoint
//     pairs (round-robin tournament), nb-1 rounds visit every pair of blocks once, one extra
//     "intra" round rotates the column pairs inside each block: together one sweep rotates every
//     column pair exactly once;
//   * one kernel launch per round.  A thread-block CLUSTER owns one block pair; its CTAs split the
//     rows.  Each CTA streams its rows of the 32-column panel through shared memory and forms
//     a partial 32x32 Gram matrix H = P^T P; partial Grams are summed over the cluster through
//     distributed shared memory (deterministic reduce-scatter + all-gather);
//   * every CTA then runs the same 16 (15) rotation rounds on H in shared memory -- a two-sided
//     update of H, exactly what the column rotations do to P^T P -- accumulating the 32x32
//     orthogonal Q.  One thread owns one 2x2 block of H per round and derives the two rotations
//     it needs itself, so a round costs a single barrier (H is double-buffered);
//   * finally the CTA applies Q to its rows of W and of J (P <- P Q) and writes them back.
//
// A rotation is skipped when |h_pq| <= max(tol sqrt(h_pp h_qq), (eps ||G||_F)^2): columns whose
// norm is below eps ||G||_F are numerically zero (these Grams are rank deficient) and are not
// rotated against each other.  Eigenvalues are Rayleigh quotients j^T G j / j^T j of the final
// columns of J with the ORIGINAL matrix (second-order accurate in the eigenvector error and free
// of the rounding drift of W), eigenvectors the normalised columns of J.
//
// Converges without a diagonal shift on the reference's LAPACK-killer fixture.  For a symmetric
// INDEFINITE input, eigenvalues +l and -l of equal magnitude cannot be separated by a one-sided
// method; the path never produces such matrices (Grams are PSD up to rounding).
#include <cooperative_groups.h>

#include <chrono>
#include <cstdlib>

#include <cstring>
#include <vector>

#include "gemm_core.cuh"
#include "nccl_dyn.cuh"

namespace cg = cooperative_groups;

namespace vvt {

constexpr int OB = 16;       // block width (columns)
constexpr int OP = 2 * OB;   // panel width: a pair of blocks
constexpr int OT = 256;      // threads per CTA
constexpr int CH = 128;      // rows per streamed chunk
constexpr int LDP = OP + 4;  // padded row pitch of a chunk in shared memory (16-byte aligned)
constexpr int LDH = OP + 1;  // padded pitch of H
constexpr int LDQ = OP + 4;  // padded pitch of Q (16-byte aligned rows for the apply phase)
constexpr int kMaxSweeps = 60;
#ifndef VVT_JACOBI_MMA
#define VVT_JACOBI_MMA 1
#endif
constexpr bool kJacobiMma = VVT_JACOBI_MMA != 0;  // fp32 Gram / apply phases on mma.sync (3xTF32) instead of FFMA

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

template <typename T>
struct Eps;
template <>
struct Eps<float> {
  static constexpr float v = 1.1920929e-7f;
  // cosine below which two columns count as orthogonal (a looser 1e-4 was measured to need MORE sweeps:
  // the skipped rotations leave couplings that keep re-exciting their neighbours)
  static constexpr float tol = 1e-5f;
};
template <>
struct Eps<double> {
  static constexpr double v = 2.220446049250313e-16;
  static constexpr double tol = 1e-13;
};

struct JacobiScalars {           // one per problem of a batch
  double norm2;                  // ||G||_F^2
  unsigned long long rotations;  // clusters that rotated in the current sweep
  unsigned long long evmax_bits; // bit pattern of max |eigenvalue| as a double (atomicMax on non-negatives)
  int converged;                 // set on the device at the end of the sweep that rotated (almost) nothing:
                                 // every later round kernel of this problem exits at once
  int sweeps;                    // sweeps this problem has run
  unsigned long long last_rotations;
  unsigned long long max_rotations;  // largest number of block pairs any sweep of this problem rotated
  long long t[8];                // VVT_SYEVJ_DEBUG: phase time stamps of CTA 0 (last launch)
};
// element strides between the problems of a batch (blockIdx.y = problem index in every kernel of this file)
struct BatchStrides {
  int64_t y, marks, dc, done;
};
#define VVT_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) sc->t[i] = clock64(); } while (0)
__device__ long long g_dbg[16];
#define VVT_DSTAMP(i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) g_dbg[i] = clock64(); } while (0)

// ---- small vector helpers (16-byte shared/global accesses, packed fp32 FMA) ----------------------
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
}
__device__ __forceinline__ void ld4(const double* p, double (&v)[4]) {
  const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
  v[0] = a.x, v[1] = a.y, v[2] = b.x, v[3] = b.y;
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4(double* p, const double (&v)[4]) {
  *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
  *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
}
// acc[i][j] += a[i] * b[j]; fp32 uses the packed FFMA2 of sm_100 (two FMAs per issue slot)
__device__ __forceinline__ void outer4(float (&acc)[4][4], const float (&a)[4], const float (&b)[4]) {
  const float2 b01 = make_float2(b[0], b[1]), b23 = make_float2(b[2], b[3]);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 ai = make_float2(a[i], a[i]);
    const float2 c01 = __ffma2_rn(ai, b01, make_float2(acc[i][0], acc[i][1]));
    const float2 c23 = __ffma2_rn(ai, b23, make_float2(acc[i][2], acc[i][3]));
    acc[i][0] = c01.x, acc[i][1] = c01.y, acc[i][2] = c23.x, acc[i][3] = c23.y;
  }
}
__device__ __forceinline__ void outer4(double (&acc)[4][4], const double (&a)[4], const double (&b)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
}

// Block-major storage Y[blk][2*Np][OB]: rows [0, Np) hold W = G J, rows [Np, 2 Np) hold J.
template <typename T>
__device__ __forceinline__ T* y_ptr(T* Y, int Np, int blk, int row) {
  return Y + (size_t(blk) * 2 * Np + row) * OB;
}

// Gs = sym(G) (upper triangle, as symeig(upper=True)) and ||G||_F^2
template <typename T>
__global__ void onesided_sym_kernel(T* Gs, const T* G, int64_t R, JacobiScalars* sc) {
  const int64_t total = R * R;
  Gs += blockIdx.y * total, G += blockIdx.y * total, sc += blockIdx.y;
  double s = 0.0;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t i = idx / R, j = idx % R;
    const T v = (i <= j) ? G[idx] : G[j * R + i];
    Gs[idx] = v;
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

// W = Gs / ||G||_F zero-padded to Np (the solver squares magnitudes: keep them near 1), J = I
template <typename T>
__global__ void onesided_init_kernel(T* Y, const T* Gs, int64_t R, int Np, const JacobiScalars* sc, int64_t y_stride) {
  const int64_t total = int64_t(Np) * Np;
  Y += blockIdx.y * y_stride, Gs += blockIdx.y * R * R, sc += blockIdx.y;
  const double n2 = sc->norm2;
  const T scale = n2 > 0.0 ? T(1.0 / sqrt(n2)) : T(0);
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int i = int(idx / Np), j = int(idx % Np);  // row i, column j
    const T v = (i < R && j < R) ? Gs[int64_t(i) * R + j] * scale : T(0);
    *(y_ptr(Y, Np, j / OB, i) + (j % OB)) = v;
    *(y_ptr(Y, Np, j / OB, Np + i) + (j % OB)) = (i == j) ? T(1) : T(0);
  }
}

// Column pair (p, q) of the 32-column panel rotated by inner round `r`, pair index a in [0, 16).
//   cross rounds: p in block A, q in block B, every (i, j) once over 16 rounds;
//   intra rounds: round-robin inside each block (8 pairs per block), 15 rounds.
__device__ __forceinline__ void inner_pair(bool intra, int r, int a, int& p, int& q) {
  if (!intra) {
    p = a;
    q = OB + ((a + r) & (OB - 1));
  } else {
    const int half = a >> 3, idx = a & 7;
    int x, y;
    rr_pair(OB, r, idx, x, y);
    p = half * OB + min(x, y);
    q = half * OB + max(x, y);
  }
}

// Rotation that orthogonalises columns p, q with Gram entries (hpp, hqq, hpq):
// [w_p' w_q'] = [w_p w_q] [[c, s], [-s, c]];  identity when the pair is already orthogonal.
// c^2 + s^2 = 1 only to a few ulp: a rotation scaled by (1 + d) scales W and J columns alike, and
// both the convergence tests and the final Rayleigh quotients / normalisation are scale free.
template <typename T>
__device__ __forceinline__ bool needs_rotation(T hpp, T hqq, T hpq, T tol2, T abs2) {
  return (hpq * hpq > tol2 * fabs(hpp * hqq)) && (fabs(hpq) > abs2);
}

__device__ __forceinline__ float rsqrt_fast(float x) {  // one MUFU.RSQ, no denormal rescaling around it
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// fp32: branch-free and division-free.  With d = hqq - hpp, o = 2 hpq, r = sqrt(d^2 + o^2):
//   cos 2theta = |d| / r,   c = sqrt((1 + cos 2theta) / 2),   s = sign(d) o / (2 r c)
// (the same inner rotation, |theta| <= pi/4, as t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2))).  Two MUFU.RSQ
// and six FMA-pipe operations on the dependent chain instead of two reciprocals, two rsqrt and a branch.
__device__ __forceinline__ bool make_rotation(float hpp, float hqq, float hpq, float tol2, float abs2,
                                              float& c, float& s) {
  const bool need = needs_rotation(hpp, hqq, hpq, tol2, abs2);
  const float d = hqq - hpp, o = hpq + hpq;
  const float q = fmaf(d, d, o * o);          // > 0 whenever need (|hpq| > eps^2, no underflow of o^2)
  const float ir = rsqrt_fast(fmaxf(q, 1e-37f));
  const float c2 = fmaf(0.5f * fabsf(d), ir, 0.5f);  // in [1/2, 1]
  const float ic = rsqrt_fast(c2);
  const float cc = c2 * ic;
  const float ss = copysignf(0.5f, d) * (o * ir) * ic;
  c = need ? cc : 1.f;
  s = need ? ss : 0.f;
  return need;
}

__device__ __forceinline__ bool make_rotation(double hpp, double hqq, double hpq, double tol2, double abs2,
                                              double& c, double& s) {
  c = 1.0;
  s = 0.0;
  if (!needs_rotation(hpp, hqq, hpq, tol2, abs2)) return false;
  const double zeta = (hqq - hpp) / (2.0 * hpq);
  const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  c = rsqrt(1.0 + t * t);
  s = t * c;
  return true;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t dst = uint32_t(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <typename T>
struct RotSmem {
  T H[2][OP][LDH];  // Gram of the panel (double-buffered by the rotation rounds)
  T Q[OP][LDQ];     // accumulated rotations
  T Gp[OP][LDQ];    // this CTA's partial Gram (read by the other CTAs of the cluster; 16-byte aligned rows)
};

// H[0] = sum over the CTAs of the cluster of their partial Grams (in Gp), Q = I.  Every CTA reads all the
// partial Grams through distributed shared memory and adds them in rank order, so that all of them hold
// bit-identical copies of H (they take the same rotation decisions) after ONE cluster barrier.  The
// matching "nobody reads my Gp any more" barrier is split: arrive here, wait in cluster_release() just
// before the CTA exits, off the critical path.
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// One float4 / double4 group of H per thread (OP * OP / 4 = OT groups).  `cross`: only the block
// H[0:OB][OB:OP] is summed (and mirrored); the diagonal blocks were prefetched from the cache into dpre by
// the threads that own their groups.
template <typename T>
__device__ __forceinline__ void cluster_reduce_gram(RotSmem<T>& rs, cg::cluster_group& cluster, int tid, bool cross,
                                                    const T (&dpre)[4]) {
  const int CL = int(cluster.num_blocks());
  const int i = tid >> 3, j = (tid & 7) * 4;
  const bool in_cross = i < OB && j >= OB, in_diag = (i < OB) == (j < OB);
  if (CL > 1) {
    cluster_arrive();  // my partial Gram is written (release) ...
    cluster_wait();    // ... and so is everybody else's (acquire)
    VVT_DSTAMP(7);
  } else {
    __syncthreads();
  }
  T sum[4] = {T(0), T(0), T(0), T(0)};
  if (!cross || in_cross) {  // (full mode: Gp is complete, mirrored quadrant included)
    T part[8][4];
#pragma unroll
    for (int p = 0; p < 8; ++p)
      if (p < CL) ld4(CL > 1 ? cluster.map_shared_rank(&rs.Gp[i][j], p) : &rs.Gp[i][j], part[p]);
#pragma unroll
    for (int p = 0; p < 8; ++p)
      if (p < CL) {
#pragma unroll
        for (int e = 0; e < 4; ++e) sum[e] += part[p][e];
      }
  } else if (in_diag) {
#pragma unroll
    for (int e = 0; e < 4; ++e) sum[e] = dpre[e];
  }
  if (!cross || in_cross || in_diag) {
#pragma unroll
    for (int e = 0; e < 4; ++e) rs.H[0][i][j + e] = sum[e];
  }
  if (cross && in_cross) {
#pragma unroll
    for (int e = 0; e < 4; ++e) rs.H[0][j + e][i] = sum[e];
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) rs.Q[i][j + e] = (i == j + e) ? T(1) : T(0);
  if (CL > 1) {
    VVT_DSTAMP(8);
    cluster_arrive();  // done reading remote shared memory; waited for in cluster_release()
  }
}

// A CTA may not exit while another CTA of its cluster can still read its shared memory.
__device__ __forceinline__ void cluster_release(cg::cluster_group& cluster) {
  if (cluster.num_blocks() > 1) cluster_wait();
}

// Does any column pair this round is responsible for still need a rotation?  (block-uniform)
template <typename T>
__device__ __forceinline__ bool any_rotation_needed(const RotSmem<T>& rs, bool intra, T tol2, T abs2, int tid) {
  const int i = tid >> 4, j = tid & 15;
  bool need;
  if (!intra) {
    need = needs_rotation(rs.H[0][i][i], rs.H[0][OB + j][OB + j], rs.H[0][i][OB + j], tol2, abs2);
  } else {
    need = i < j && (needs_rotation(rs.H[0][i][i], rs.H[0][j][j], rs.H[0][i][j], tol2, abs2) ||
                     needs_rotation(rs.H[0][OB + i][OB + i], rs.H[0][OB + j][OB + j], rs.H[0][OB + i][OB + j],
                                    tol2, abs2));
  }
  return __syncthreads_or(need ? 1 : 0) != 0;
}

// The 16 (cross) or 15 (intra) rotation rounds on H, accumulating Q.  Thread (a, b) owns the 2x2
// block H[{pa,qa}][{pb,qb}] (H' = Ra^T H Rb); lane b & 15 of every warp derives rotation b and
// rotation a comes from lane a of the same warp by shuffle.  Returns with the result in Q.
template <typename T, bool intra>
__device__ __forceinline__ void rotation_rounds(RotSmem<T>& rs, T tol2, T abs2, int tid) {
  const int a = tid >> 4, b = tid & 15;
  constexpr int n_rounds = intra ? OB - 1 : OB;
  int cur = 0;
  for (int r = 0; r < n_rounds; ++r, cur ^= 1) {
    T(*Hc)[LDH] = rs.H[cur];
    T(*Hn)[LDH] = rs.H[cur ^ 1];
    int pa, qa, pb, qb;
    inner_pair(intra, r, a, pa, qa);
    inner_pair(intra, r, b, pb, qb);
    // every shared-memory read of the round is issued up front (one latency instead of three in a row)
    const T dpp = Hc[pb][pb], dqq = Hc[qb][qb], dpq = Hc[pb][qb];
    const T x00 = Hc[pa][pb], x01 = Hc[pa][qb], x10 = Hc[qa][pb], x11 = Hc[qa][qb];
    T* q0 = &rs.Q[b][0];
    T* q1 = &rs.Q[b + OB][0];
    const T q0p = q0[pa], q0q = q0[qa], q1p = q1[pa], q1q = q1[qa];
    T cb, sb;
    const bool db = make_rotation(dpp, dqq, dpq, tol2, abs2, cb, sb);
    const T ca = __shfl_sync(0xffffffffu, cb, a), sa = __shfl_sync(0xffffffffu, sb, a);
    // X = H[{pa,qa}][{pb,qb}];  T = X Rb;  H' = Ra^T T,  R = [[c, s], [-s, c]]
    const T t00 = cb * x00 - sb * x01, t01 = sb * x00 + cb * x01;
    const T t10 = cb * x10 - sb * x11, t11 = sb * x10 + cb * x11;
    T h00 = ca * t00 - sa * t10, h01 = ca * t01 - sa * t11;
    T h10 = sa * t00 + ca * t10, h11 = sa * t01 + ca * t11;
    if (a == b && db) h01 = h10 = T(0);  // annihilated by construction
    Hn[pa][pb] = h00;
    Hn[pa][qb] = h01;
    Hn[qa][pb] = h10;
    Hn[qa][qb] = h11;
    // Q <- Q R: rows k = b, b + 16 of column pair a (in place: one owner per element and round)
    q0[pa] = ca * q0p - sa * q0q;
    q0[qa] = sa * q0p + ca * q0q;
    q1[pa] = ca * q1p - sa * q1q;
    q1[qa] = sa * q1p + ca * q1q;
    __syncthreads();
  }
}

// Every CTA of the cluster runs the rotation rounds itself on its copy of H (identical inputs, identical
// arithmetic => identical Q): nothing to broadcast and no cluster barrier after the Gram reduction.
template <typename T>
__device__ __forceinline__ void rotate_and_broadcast(RotSmem<T>& rs, cg::cluster_group& cluster, bool intra,
                                                     T tol2, T abs2, int tid) {
  (void)cluster;
  if (intra) rotation_rounds<T, true>(rs, tol2, abs2, tid);
  else rotation_rounds<T, false>(rs, tol2, abs2, tid);
}

// this thread's share of a 4x4 tile of P^T P: rows r = ks, ks + 4, ... < nrows of a row-major chunk
template <typename T>
__device__ __forceinline__ void gram_rows(T (&acc)[4][4], const T* P, int ldp, int nrows, int ks, int ti, int tj) {
  int r = ks;
  for (; r + 4 < nrows; r += 8) {  // 2 rows in flight
    T a[2][4], b[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      ld4(P + size_t(r + 4 * u) * ldp + 4 * ti, a[u]);
      ld4(P + size_t(r + 4 * u) * ldp + 4 * tj, b[u]);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) outer4(acc, a[u], b[u]);
  }
  for (; r < nrows; r += 4) {
    T a[4], b[4];
    ld4(P + size_t(r) * ldp + 4 * ti, a);
    ld4(P + size_t(r) * ldp + 4 * tj, b);
    outer4(acc, a, b);
  }
}

// sum the 4 k-split partial tiles (lanes ks = tid & 3) and store the tile into Q
template <typename T>
__device__ __forceinline__ void gram_store(RotSmem<T>& rs, T (&acc)[4][4], int ks, int ti, int tj) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      T v = acc[i][j];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (ks == 0) rs.Gp[4 * ti + i][4 * tj + j] = v;
    }
}

// out rows (4 per thread: tr + 32 i) = P rows * Q, written to global memory
template <typename T>
__device__ __forceinline__ void apply_rows(const RotSmem<T>& rs, const T* P, int ldp, int r_base, int n_rows,
                                           T* Y, int Np, int ba, int bb, int grow0, int tr, int tc) {
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
  int rows[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) rows[i] = min(r_base + tr + 32 * i, n_rows - 1);  // clamp: read valid memory
#pragma unroll 2
  for (int k4 = 0; k4 < OP; k4 += 4) {
    T pv[4][4];  // pv[i][kk] = P[row_i][k4 + kk]
#pragma unroll
    for (int i = 0; i < 4; ++i) ld4(P + size_t(rows[i]) * ldp + k4, pv[i]);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      T qv[4];
      ld4(&rs.Q[k4 + kk][4 * tc], qv);
      const T a[4] = {pv[0][kk], pv[1][kk], pv[2][kk], pv[3][kk]};
      outer4(acc, a, qv);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r_base + tr + 32 * i;
    if (r < n_rows) {
      const int col = 4 * tc;
      st4(y_ptr(Y, Np, col < OB ? ba : bb, grow0 + r) + (col & (OB - 1)), acc[i]);
    }
  }
}

// ---- fp32 fast paths for the two O(rows) phases of a round ----------------------------------------------
// Gram: one 8x8 register tile of H per warp (upper triangle of the 4x4 tile grid: 10 tiles), the 32 lanes
// split the rows; per row and lane 4 LDS.128 feed 64 FMAs (the 4x4 variant above is shared-memory bound).
__device__ __forceinline__ void gram_tile8_accum(float (&acc)[64], const float* P, int ldp, int nrows, int lane,
                                                 int ti, int tj) {
  for (int r = lane; r < nrows; r += 32) {
    float a[8], b[8];
    const float* row = P + size_t(r) * ldp;
    {
      float lo[4], hi[4];
      ld4(row + 8 * ti, lo);
      ld4(row + 8 * ti + 4, hi);
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = lo[i], a[4 + i] = hi[i];
    }
    if (ti != tj) {  // warp-uniform
      float lo[4], hi[4];
      ld4(row + 8 * tj, lo);
      ld4(row + 8 * tj + 4, hi);
#pragma unroll
      for (int i = 0; i < 4; ++i) b[i] = lo[i], b[4 + i] = hi[i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) b[i] = a[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[8 * i + j] = fmaf(a[i], b[j], acc[8 * i + j]);
  }
}

// Sum the tile over the 32 lanes by recursive halving (62 shuffles for 64 values: lane L ends up with the
// entries 2L and 2L+1 of the row-major tile) and store it, mirrored, into Q.
__device__ __forceinline__ void gram_tile8_store(RotSmem<float>& rs, float (&v)[64], int lane, int ti, int tj) {
#pragma unroll
  for (int half = 32; half >= 2; half >>= 1) {
    const int mask = half >> 1;
    const bool up = (lane & mask) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float keep = up ? v[i + half] : v[i];
      const float send = up ? v[i] : v[i + half];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
    }
  }
  const int i = 8 * ti + (lane >> 2), j = 8 * tj + 2 * (lane & 3);
  rs.Gp[i][j] = v[0];
  rs.Gp[i][j + 1] = v[1];
  if (ti != tj) {
    rs.Gp[j][i] = v[0];
    rs.Gp[j + 1][i] = v[1];
  }
}

__device__ __forceinline__ void gram_phase_fast(RotSmem<float>& rs, const float* P, int ldp, int nrows, int tid) {
  const int warp = tid >> 5, lane = tid & 31;
  for (int t = warp; t < 10; t += OT / 32) {
    // t -> (ti, tj), ti <= tj, row by row over the upper triangle of the 4x4 tile grid
    int ti = 0, rem = t;
    while (rem >= 4 - ti) rem -= 4 - ti, ++ti;
    const int tj = ti + rem;
    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0.f;
    gram_tile8_accum(acc, P, ldp, nrows, lane, ti, tj);
    gram_tile8_store(rs, acc, lane, ti, tj);
  }
}

// Apply: thread = (row, 16-column half); the row of P sits in registers, Q rows are broadcast reads.
__device__ __forceinline__ void apply_phase_fast(const RotSmem<float>& rs, const float* P, int ldp, int nrows, float* Y,
                                                 int Np, int ba, int bb, int grow0, int tid) {
  const int h = tid & 1;
  for (int r = tid >> 1; r < nrows; r += OT / 2) {
    float acc[OB];
    const float* row = P + size_t(r) * ldp;
#pragma unroll
    for (int j = 0; j < OB; ++j) acc[j] = 0.f;
#pragma unroll 2
    for (int k4 = 0; k4 < OP; k4 += 4) {
      float p[4];
      ld4(row + k4, p);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int j4 = 0; j4 < OB; j4 += 4) {
          float q[4];
          ld4(&rs.Q[k4 + kk][OB * h + j4], q);
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[j4 + i] = fmaf(p[kk], q[i], acc[j4 + i]);
        }
      }
    }
    float* dst = y_ptr(Y, Np, h ? bb : ba, grow0 + r);
#pragma unroll
    for (int j4 = 0; j4 < OB; j4 += 4) {
      const float t[4] = {acc[j4], acc[j4 + 1], acc[j4 + 2], acc[j4 + 3]};
      st4(dst + j4, t);
    }
  }
}

// ---- fp32 tensor-core paths for the two O(rows) phases (mma.sync m16n8k8, 3xTF32 split: fp32-grade) -------
constexpr int kGramStageFloats = (OT / 32) * OP * OP;  // per-warp partial Grams, summed in a fixed order

__device__ __forceinline__ void mma3(float (&acc)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                     const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
  mma_tf32(acc, al, bh);
  mma_tf32(acc, ah, bl);
  mma_tf32(acc, ah, bh);
}

// H = P^T P for this CTA's rows: the warps split the rows in steps of 8; the same eight shared-memory
// values per lane serve as A and B fragments (H is symmetric: only tiles on or above the diagonal).
// CROSS: only the 16x16 block P_a^T P_b (the diagonal blocks come from the cache, see DiagCache): a third
// of the tensor-core work.
template <bool CROSS>
__device__ __forceinline__ void gram_phase_mma(RotSmem<float>& rs, float* stage, const float* P, int ldp, int nrows,
                                               int tid) {
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  VVT_DSTAMP(0);
  constexpr int NACC = CROSS ? 2 : 6;
  float acc[NACC][4];  // full: (mt 0, nt 0..3), (mt 1, nt 2..3);  cross: (mt 0, nt 2..3)
#pragma unroll
  for (int i = 0; i < NACC; ++i)
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[i][r] = 0.f;
  for (int k0 = 8 * warp; k0 < nrows; k0 += 8 * (OT / 32)) {
    const int r0 = k0 + t, r1 = k0 + t + 4;
    uint32_t hi[4][2], lo[4][2];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float x0 = r0 < nrows ? P[size_t(r0) * ldp + 8 * c + g] : 0.f;
      const float x1 = r1 < nrows ? P[size_t(r1) * ldp + 8 * c + g] : 0.f;
      split_tf32(x0, hi[c][0], lo[c][0]);
      split_tf32(x1, hi[c][1], lo[c][1]);
    }
    if constexpr (CROSS) {
      const uint32_t ah[4] = {hi[0][0], hi[1][0], hi[0][1], hi[1][1]};
      const uint32_t al[4] = {lo[0][0], lo[1][0], lo[0][1], lo[1][1]};
      mma3(acc[0], ah, al, hi[2], lo[2]);
      mma3(acc[1], ah, al, hi[3], lo[3]);
    } else {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const uint32_t ah[4] = {hi[2 * mt][0], hi[2 * mt + 1][0], hi[2 * mt][1], hi[2 * mt + 1][1]};
        const uint32_t al[4] = {lo[2 * mt][0], lo[2 * mt + 1][0], lo[2 * mt][1], lo[2 * mt + 1][1]};
#pragma unroll
        for (int nt = 2 * mt; nt < 4; ++nt) mma3(acc[mt == 0 ? nt : 2 + nt], ah, al, hi[nt], lo[nt]);
      }
    }
  }
  VVT_DSTAMP(1);
  float* mine = stage + warp * (OP * OP);
#pragma unroll
  for (int mt = 0; mt < (CROSS ? 1 : 2); ++mt)
#pragma unroll
    for (int nt = (CROSS ? 2 : 2 * mt); nt < 4; ++nt) {
      const float(&a)[4] = acc[CROSS ? nt - 2 : (mt == 0 ? nt : 2 + nt)];
      const int row = 16 * mt + g, col = 8 * nt + 2 * t;
      mine[row * OP + col] = a[0];
      mine[row * OP + col + 1] = a[1];
      mine[(row + 8) * OP + col] = a[2];
      mine[(row + 8) * OP + col + 1] = a[3];
    }
  VVT_DSTAMP(2);
  __syncthreads();
  VVT_DSTAMP(3);
  // cross-warp sum, one float4 of the tile per thread.  Only the tiles on or above the diagonal of the 2x2
  // grid of 16x16 quadrants were computed: the upper-right quadrant is mirrored into the lower-left one.
  {
    const int row = tid >> 3, col = (tid & 7) * 4;
    if (CROSS ? (row < OB && col >= OB) : (row < OB || col >= OB)) {
      float4 sum = *reinterpret_cast<const float4*>(stage + row * OP + col);
#pragma unroll
      for (int w = 1; w < OT / 32; ++w) {
        const float4 v = *reinterpret_cast<const float4*>(stage + w * (OP * OP) + row * OP + col);
        sum.x += v.x, sum.y += v.y, sum.z += v.z, sum.w += v.w;
      }
      *reinterpret_cast<float4*>(&rs.Gp[row][col]) = sum;
      if (!CROSS && row < OB && col >= OB)
        rs.Gp[col][row] = sum.x, rs.Gp[col + 1][row] = sum.y, rs.Gp[col + 2][row] = sum.z, rs.Gp[col + 3][row] = sum.w;
    }
  }
  VVT_DSTAMP(4);
}

// out = P Q for this CTA's rows, straight to global memory; Q fragments (hi/lo) stay in registers
__device__ __forceinline__ void apply_phase_mma(const RotSmem<float>& rs, const float* P, int ldp, int nrows, float* Y,
                                                int Np, int ba, int bb, int grow0, int tid) {
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  VVT_DSTAMP(5);
  uint32_t qh[4][4][2], ql[4][4][2];  // [k-step][n-tile]
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      split_tf32(rs.Q[8 * ks + t][8 * nt + g], qh[ks][nt][0], ql[ks][nt][0]);
      split_tf32(rs.Q[8 * ks + t + 4][8 * nt + g], qh[ks][nt][1], ql[ks][nt][1]);
    }
  VVT_DSTAMP(6);
  for (int m0 = 16 * warp; m0 < nrows; m0 += 16 * (OT / 32)) {
    const int ra = m0 + g, rb = m0 + g + 8;
    const bool va = ra < nrows, vb = rb < nrows;
    const float* pa = P + size_t(va ? ra : 0) * ldp;
    const float* pb = P + size_t(vb ? rb : 0) * ldp;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[i][r] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t ah[4], al[4];
      split_tf32(va ? pa[8 * ks + t] : 0.f, ah[0], al[0]);
      split_tf32(vb ? pb[8 * ks + t] : 0.f, ah[1], al[1]);
      split_tf32(va ? pa[8 * ks + t + 4] : 0.f, ah[2], al[2]);
      split_tf32(vb ? pb[8 * ks + t + 4] : 0.f, ah[3], al[3]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma3(acc[nt], ah, al, qh[ks][nt], ql[ks][nt]);
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int blk = nt < 2 ? ba : bb, col = (8 * nt + 2 * t) & (OB - 1);
      if (va) *reinterpret_cast<float2*>(y_ptr(Y, Np, blk, grow0 + ra) + col) = make_float2(acc[nt][0], acc[nt][1]);
      if (vb) *reinterpret_cast<float2*>(y_ptr(Y, Np, blk, grow0 + rb) + col) = make_float2(acc[nt][2], acc[nt][3]);
    }
  }
}

__device__ __forceinline__ void round_blocks(int nb, int round, int pair, int& ba, int& bb) {
  if (round < 0) {
    ba = 2 * pair;
    bb = 2 * pair + 1;
  } else {
    rr_pair(nb, round, pair, ba, bb);
  }
}

// Bookkeeping that lets converged block pairs skip a round entirely: stamp[b] = last round in which
// block b was rotated, clean[a][b] = last round in which the pair was found to need no rotation (upper
// triangle: cross rounds, lower triangle: intra rounds).  A pair that was clean after both of its blocks
// were last modified is still clean.  Values are written by one thread per cluster and read by later
// launches only.
__device__ __forceinline__ int* clean_slot(int* marks, int nb, int ba, int bb, bool intra) {
  const int lo = min(ba, bb), hi = max(ba, bb);
  return marks + nb + (intra ? hi * nb + lo : lo * nb + hi);
}
__device__ __forceinline__ bool pair_is_clean(int* marks, int nb, int ba, int bb, bool intra) {
  const int c = *clean_slot(marks, nb, ba, bb, intra);
  return c > max(marks[ba], marks[bb]);
}

// ---- dependencies between rounds: per column block, not per grid ---------------------------------------
// A round launch only has to wait for the two clusters of the previous launch that wrote ITS two blocks.
// done[blk] counts the CTAs that have finished with block blk (every launch touches every block exactly once,
// every CTA signals on every exit path), so launch number L may read block blk once done[blk] >= CL * L.
// With programmatic dependent launch the next grids are resident early and spin here instead of sitting in
// griddepcontrol.wait until the whole previous grid has drained: a pair that only checks (tail sweeps) or a
// cluster that finishes early no longer waits for the slowest cluster of the round.  Deadlock-free: a grid
// starts only after every CTA of the previous one has started (and therefore holds its resources).
__device__ __forceinline__ void wait_blocks(const unsigned* done, int ba, int bb, unsigned need, int tid) {
  if (tid == 0) {
    const volatile unsigned* d = done;
    long long t0 = 0;
    for (unsigned spins = 0; d[ba] < need || d[bb] < need; ++spins) {
      if ((spins & 1023u) == 1023u) {  // bounded (~2 s): a protocol bug traps instead of hanging the GPU
        const long long t = clock64();
        if (t0 == 0) t0 = t;
        else if (t - t0 > 4000000000ll) __trap();
      }
    }
    __threadfence();  // acquire: later reads of this CTA see what the signalling CTAs wrote
  }
  __syncthreads();
}
__device__ __forceinline__ void signal_blocks(unsigned* done, int ba, int bb, int tid) {
  __syncthreads();  // every store of this CTA has been issued
  if (tid == 0) {
    __threadfence();  // release
    atomicAdd(done + ba, 1u);
    atomicAdd(done + bb, 1u);
  }
}

// ---- resident variant: this CTA's rows of W and J are loaded once (cp.async) and stay in smem ----
// One round of one CTA.  PERSIST = false: the body of a per-round launch (programmatic dependent launch between
// rounds); PERSIST = true: called once per round from the sweep kernel below, whose CTAs stay resident for the
// 1 + (nb - 1) rounds of a sweep.
template <typename T, bool PERSIST>
__device__ __forceinline__ void resident_round(unsigned char* smem_raw, T* Y, int Np, int nb, int round, int rows_per_cta,
                                               int parts, JacobiScalars* sc, int* marks, int now, T* Dc, unsigned* done) {
  RotSmem<T>& rs = *reinterpret_cast<RotSmem<T>*>(smem_raw);
  T* P = reinterpret_cast<T*>(smem_raw + ((sizeof(RotSmem<T>) + 15) / 16) * 16);  // [parts * rows][LDP]: W rows(, J rows)
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = int(cluster.num_blocks()), crank = int(cluster.block_rank());
  const int tid = threadIdx.x;
  const bool intra = round < 0;
  int ba, bb;
  round_blocks(nb, round, blockIdx.x / CL, ba, bb);
  const int w0 = crank * rows_per_cta, nrows = max(0, min(Np, w0 + rows_per_cta) - w0);
  const T tol2 = Eps<T>::tol * Eps<T>::tol;
  const T abs2 = Eps<T>::v * Eps<T>::v;  // W is scaled to unit Frobenius norm

  // programmatic dependent launch: let the next round's CTAs be scheduled while this round runs (they
  // block in griddepcontrol.wait until this grid has completed and its stores are visible)
  if constexpr (!PERSIST) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (intra || !done) asm volatile("griddepcontrol.wait;" ::: "memory");  // first launch of a sweep: after memsets / init
    // a finished problem of the batch (flag written by sweep_end_kernel, which the intra round has waited for;
    // it only changes between sweeps): nothing left to do, and none of its later CTAs waits for done[]
    if (*reinterpret_cast<const volatile int*>(&sc->converged)) return;
  }
  if (done) wait_blocks(done, ba, bb, unsigned(CL) * unsigned(now - 1), tid);
  if (pair_is_clean(marks, nb, ba, bb, intra)) {  // same decision in every CTA of the cluster
    if (done) signal_blocks(done, ba, bb, tid);
    return;
  }
  VVT_STAMP(0);
  // DiagCache: Dc[blk][OB][OB] holds the diagonal block W_blk^T W_blk of every column block as the last
  // rotation round left it in H.  Cross rounds take both diagonal blocks from there and form only the
  // OB x OB cross block on the tensor cores; the intra round of every sweep recomputes them exactly, which
  // bounds the rounding drift of the cached values to one sweep of two-sided updates (~1e-6 relative; the
  // rotation angles and the convergence test of a cross round depend on them only through h_pp, h_qq).
  constexpr bool kCache = sizeof(T) == 4 && kJacobiMma;
  const bool cross = kCache && !intra;
  const int gi = tid >> 3, gj = (tid & 7) * 4;                 // this thread's group of 4 entries of H
  const bool g_diag = (gi < OB) == (gj < OB);
  T* dslot = Dc + (size_t(gi < OB ? ba : bb) * OB + (gi & (OB - 1))) * OB + (gj & (OB - 1));
  T dpre[4] = {T(0), T(0), T(0), T(0)};
  if (cross && g_diag) ld4(dslot, dpre);
  // all loads are issued up front: group 0 = W rows (needed now), group 1 = J rows (needed last)
  {
    constexpr int V = 16 / sizeof(T), VPR = OB / V;  // elements per 16 bytes, vectors per block row
    for (int part = 0; part < parts; ++part) {
      for (int idx = tid; idx < nrows * 2 * VPR; idx += OT) {
        const int r = idx / (2 * VPR), v = idx % (2 * VPR);
        const int blk = v < VPR ? ba : bb, col = (v % VPR) * V;
        cp_async16(P + size_t(part * nrows + r) * LDP + (v < VPR ? 0 : OB) + col,
                   y_ptr(Y, Np, blk, part * Np + w0 + r) + col);
      }
      cp_async_commit();
    }
  }
  if (parts == 2) cp_async_wait<1>(); else cp_async_wait<0>();
  __syncthreads();
  VVT_STAMP(1);

  if constexpr (sizeof(T) == 4) {  // phase 1: partial Gram of this CTA's rows of W
    float* stage = P + size_t(parts) * rows_per_cta * LDP;  // behind the resident rows
    if (kJacobiMma) {
      if (cross) gram_phase_mma<true>(rs, stage, P, LDP, nrows, tid);
      else gram_phase_mma<false>(rs, stage, P, LDP, nrows, tid);
    } else {
      gram_phase_fast(rs, P, LDP, nrows, tid);
    }
  } else {
    const int ks = tid & 3, ti = tid >> 5, tj = (tid >> 2) & 7;
    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
    gram_rows<T>(acc, P, LDP, nrows, ks, ti, tj);
    gram_store<T>(rs, acc, ks, ti, tj);
  }
  VVT_STAMP(2);
  cluster_reduce_gram<T>(rs, cluster, tid, cross, dpre);
  __syncthreads();
  VVT_STAMP(3);
  if (!any_rotation_needed<T>(rs, intra, tol2, abs2, tid)) {  // uniform over the cluster (same H)
    if (crank == 0 && tid == 0) *clean_slot(marks, nb, ba, bb, intra) = now;
    if (kCache && intra && crank == 0 && g_diag) {  // exact diagonal blocks: refresh the cache
      const T v[4] = {rs.H[0][gi][gj], rs.H[0][gi][gj + 1], rs.H[0][gi][gj + 2], rs.H[0][gi][gj + 3]};
      st4(dslot, v);
    }
    cp_async_wait<0>();
    if (done) signal_blocks(done, ba, bb, tid);
    cluster_release(cluster);
    return;
  }
  if (crank == 0 && tid == 0) {
    atomicAdd(&sc->rotations, 1ull);
    marks[ba] = now;
    marks[bb] = now;
  }
  rotate_and_broadcast<T>(rs, cluster, intra, tol2, abs2, tid);  // phase 2
  VVT_STAMP(4);
  if (kCache && crank == 0 && g_diag) {  // the rotated diagonal blocks (the rounds end in H[rounds & 1])
    const int fin = (intra ? OB - 1 : OB) & 1;
    const T v[4] = {rs.H[fin][gi][gj], rs.H[fin][gi][gj + 1], rs.H[fin][gi][gj + 2], rs.H[fin][gi][gj + 3]};
    st4(dslot, v);
  }
  cp_async_wait<0>();
  __syncthreads();
  VVT_STAMP(5);
  if constexpr (sizeof(T) == 4) {  // phase 3: P <- P Q for the W rows (and the J rows), straight to global memory
    for (int part = 0; part < parts; ++part) {
      if (kJacobiMma) apply_phase_mma(rs, P + size_t(part) * nrows * LDP, LDP, nrows, Y, Np, ba, bb, part * Np + w0, tid);
      else apply_phase_fast(rs, P + size_t(part) * nrows * LDP, LDP, nrows, Y, Np, ba, bb, part * Np + w0, tid);
    }
  } else {
    const int tr = tid >> 3, tc = tid & 7;
    for (int part = 0; part < parts; ++part)
      for (int r_base = 0; r_base < nrows; r_base += 128)
        apply_rows<T>(rs, P + size_t(part) * nrows * LDP, LDP, r_base, nrows, Y, Np, ba, bb, part * Np + w0, tr, tc);
  }
  VVT_STAMP(6);
  if (done) signal_blocks(done, ba, bb, tid);
  cluster_release(cluster);
}

template <typename T>
__global__ void __launch_bounds__(OT, (sizeof(T) == 4 ? 2 : 1))
onesided_round_resident_kernel(T* Y, int Np, int nb, int round, int rows_per_cta, int parts, JacobiScalars* sc,
                               int* marks, int now, T* Dc, unsigned* done, BatchStrides bs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Y += blockIdx.y * bs.y, sc += blockIdx.y, marks += blockIdx.y * bs.marks, Dc += blockIdx.y * bs.dc;
  if (done) done += blockIdx.y * bs.done;
  resident_round<T, false>(smem_raw, Y, Np, nb, round, rows_per_cta, parts, sc, marks, now, Dc, done);
}

// A whole sweep in ONE launch: every cluster keeps its pair slot and walks the rounds of the tournament itself;
// the per-block done[] counters that order the rounds of separate launches order them here as well (a CTA of
// round r spins until the two clusters of round r - 1 that wrote ITS blocks have signalled).  All CTAs must be
// resident at once for that: the host launches this kernel only when pairs x cluster size x batch fits on the GPU
// with one CTA per SM (R <= 1280 in fp32; checked with cudaOccupancyMaxActiveClusters).  972 launches per solve
// at R = 1280 become 12 -- but the solve gets slower (see syevj_impl), so this variant is opt-in.
template <typename T>
__global__ void __launch_bounds__(OT, (sizeof(T) == 4 ? 2 : 1))
onesided_sweep_resident_kernel(T* Y, int Np, int nb, int rows_per_cta, int parts, JacobiScalars* sc, int* marks,
                               int now0, T* Dc, unsigned* done, BatchStrides bs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Y += blockIdx.y * bs.y, sc += blockIdx.y, marks += blockIdx.y * bs.marks, Dc += blockIdx.y * bs.dc;
  done += blockIdx.y * bs.done;
  if (*reinterpret_cast<const volatile int*>(&sc->converged)) return;  // uniform over the problem's CTAs
  for (int round = -1; round < nb - 1; ++round) {
    resident_round<T, true>(smem_raw, Y, Np, nb, round, rows_per_cta, parts, sc, marks, now0 + round + 1, Dc, done);
    __syncthreads();  // shared memory is reused by the next round
  }
}

// ---- streaming variant (any size): rows pass through two chunk buffers, W is read twice -----------
template <typename T>
struct StreamSmem {
  RotSmem<T> rs;
  T chunk[2][CH][LDP];
};

template <typename T>
__device__ __forceinline__ void load_chunk(T (*dst)[LDP], const T* Y, int Np, int ba, int bb, int row0,
                                           int nrows, int tid) {
  constexpr int V = 16 / sizeof(T), VPR = OB / V;
  for (int idx = tid; idx < nrows * 2 * VPR; idx += OT) {
    const int r = idx / (2 * VPR), v = idx % (2 * VPR);
    const int blk = v < VPR ? ba : bb, col = (v % VPR) * V;
    cp_async16(&dst[r][(v < VPR ? 0 : OB) + col], y_ptr(const_cast<T*>(Y), Np, blk, row0 + r) + col);
  }
  cp_async_commit();
}

template <typename T>
__global__ void __launch_bounds__(OT, (sizeof(T) == 4 ? 2 : 1))
onesided_round_stream_kernel(T* Y, int Np, int nb, int round, int rows_per_cta, int parts, JacobiScalars* sc,
                             int* marks, int now, T* Dc, unsigned* done, BatchStrides bs) {
  (void)Dc;  // the streaming variant recomputes the whole panel Gram in every round
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Y += blockIdx.y * bs.y, sc += blockIdx.y, marks += blockIdx.y * bs.marks;
  if (done) done += blockIdx.y * bs.done;
  StreamSmem<T>& sm = *reinterpret_cast<StreamSmem<T>*>(smem_raw);
  RotSmem<T>& rs = sm.rs;
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = int(cluster.num_blocks()), crank = int(cluster.block_rank());
  const int tid = threadIdx.x;
  const bool intra = round < 0;
  int ba, bb;
  round_blocks(nb, round, blockIdx.x / CL, ba, bb);
  const int w0 = crank * rows_per_cta, nrows = max(0, min(Np, w0 + rows_per_cta) - w0);
  const int n_chunks = (nrows + CH - 1) / CH;
  const T tol2 = Eps<T>::tol * Eps<T>::tol;
  const T abs2 = Eps<T>::v * Eps<T>::v;

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (intra || !done) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (*reinterpret_cast<const volatile int*>(&sc->converged)) return;
  if (done) wait_blocks(done, ba, bb, unsigned(CL) * unsigned(now - 1), tid);
  if (pair_is_clean(marks, nb, ba, bb, intra)) {
    if (done) signal_blocks(done, ba, bb, tid);
    return;
  }
  {  // phase 1: partial Gram, chunks of W rows double-buffered
    const int ks = tid & 3, ti = tid >> 5, tj = (tid >> 2) & 7;
    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
    if (n_chunks > 0) load_chunk<T>(sm.chunk[0], Y, Np, ba, bb, w0, min(CH, nrows), tid);
    for (int c = 0; c < n_chunks; ++c) {
      if (c + 1 < n_chunks) {
        load_chunk<T>(sm.chunk[(c + 1) & 1], Y, Np, ba, bb, w0 + (c + 1) * CH, min(CH, nrows - (c + 1) * CH), tid);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      gram_rows<T>(acc, &sm.chunk[c & 1][0][0], LDP, min(CH, nrows - c * CH), ks, ti, tj);
      __syncthreads();
    }
    gram_store<T>(rs, acc, ks, ti, tj);
  }
  {
    const T none[4] = {T(0), T(0), T(0), T(0)};
    cluster_reduce_gram<T>(rs, cluster, tid, false, none);
  }
  __syncthreads();
  if (!any_rotation_needed<T>(rs, intra, tol2, abs2, tid)) {
    if (crank == 0 && tid == 0) *clean_slot(marks, nb, ba, bb, intra) = now;
    if (done) signal_blocks(done, ba, bb, tid);
    cluster_release(cluster);
    return;
  }
  if (crank == 0 && tid == 0) {
    atomicAdd(&sc->rotations, 1ull);
    marks[ba] = now;
    marks[bb] = now;
  }
  // first chunk of phase 3 travels while the rotation rounds run
  if (n_chunks > 0) load_chunk<T>(sm.chunk[0], Y, Np, ba, bb, w0, min(CH, nrows), tid);
  rotate_and_broadcast<T>(rs, cluster, intra, tol2, abs2, tid);  // phase 2
  {  // phase 3: chunks of W rows, then of J rows
    const int tr = tid >> 3, tc = tid & 7;
    const int total = parts * n_chunks;
    for (int c = 0; c < total; ++c) {
      if (c + 1 < total) {
        const int part = (c + 1) >= n_chunks, cc = (c + 1) - part * n_chunks;
        load_chunk<T>(sm.chunk[(c + 1) & 1], Y, Np, ba, bb, part * Np + w0 + cc * CH, min(CH, nrows - cc * CH), tid);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      const int part = c >= n_chunks, cc = c - part * n_chunks;
      apply_rows<T>(rs, &sm.chunk[c & 1][0][0], LDP, 0, min(CH, nrows - cc * CH), Y, Np, ba, bb,
                    part * Np + w0 + cc * CH, tr, tc);
      __syncthreads();
    }
  }
  if (done) signal_blocks(done, ba, bb, tid);
  cluster_release(cluster);
}

// ---- Cholesky preconditioner ------------------------------------------------------------------------
// G + eps I = L L^T with eps = 1e-3 ||G||_F.  The shift keeps the factor defined for rank-deficient or
// slightly indefinite (rounding) Grams and leaves the eigenvectors untouched; because the rotation
// threshold is relative (cosine <= tol), eigenvalue pairs far below eps are still separated to
// tol * eps / gap, finer than the input precision allows, and a larger shift converges faster.  The
// one-sided Jacobi is then run on the columns of L instead of G:  L J = U Sigma  gives
// G + eps I = U Sigma^2 U^T, so the eigenvectors are the normalised columns of L J -- no J is
// accumulated (half the traffic and flops per round) and the spectrum the rotations see is that of G,
// not of G^2, which cuts the number of sweeps (19 -> 12 on the cifar10_3c3d Gram in fp32).
// The factor only has to be a good starting point: eigenvalues are Rayleigh quotients with the
// original G and the refinement step below works with the original G as well.
constexpr int CB = 64;    // Cholesky block width
constexpr int CROWS = 256;  // rows of the panel solved per CTA (one thread per row)

template <typename T>
__device__ __forceinline__ T chol_shift(const JacobiScalars* sc) {
  return T(1e-3 * sqrt(sc->norm2));
}

// One step of the blocked right-looking factorisation on A (row-major, ld = R, lower triangle):
// every CTA factors the diagonal block A[k:k+b, k:k+b] in shared memory (redundantly -- it is tiny),
// CTA 0 writes it back, and the CTAs split the rows below it:  L21 = A21 L11^{-T}  by substitution.
constexpr int LDD = CB + 4;  // pitch of the diagonal block in shared memory (16-byte aligned rows)

__device__ __forceinline__ float rsqrt_any(float x) { return rsqrt_fast(x); }
__device__ __forceinline__ double rsqrt_any(double x) { return rsqrt(x); }

// Factors the CB x CB block held in D (lower triangle, zero-padded) in ONE warp, entirely in registers:
// lane l owns rows l and l + 32; a step broadcasts the pivot and the scaled column by warp shuffles (fully
// unrolled: every register index is static), so the 64 dependent steps cost a shuffle + rsqrt each instead
// of two block-wide barriers.  The factor only preconditions the Jacobi iteration: rsqrt accuracy is plenty.
template <typename T>
__device__ __forceinline__ void chol_diag_warp(T (*D)[LDD], T* dinv, T floor_piv, int lane) {
  T r0[CB], r1[CB];
#pragma unroll
  for (int c4 = 0; c4 < CB; c4 += 4) {
    T a[4], b4[4];
    ld4(&D[lane][c4], a);
    ld4(&D[lane + 32][c4], b4);
#pragma unroll
    for (int e = 0; e < 4; ++e) r0[c4 + e] = a[e], r1[c4 + e] = b4[e];
  }
#pragma unroll
  for (int j = 0; j < CB; ++j) {
    const T djj = __shfl_sync(0xffffffffu, j < 32 ? r0[j] : r1[j], j & 31);
    const T inv = rsqrt_any(vmax(djj, floor_piv));
    // column j of L in rows lane, lane + 32; zero above the diagonal, so that nothing grows in the unused half
    const T c0 = (j < 32 && lane >= j) ? r0[j] * inv : T(0);
    const T c1 = (lane + 32 >= j) ? r1[j] * inv : T(0);
    if (lane == 0) dinv[j] = inv;
    if (j < 32) {
      r0[j] = (lane == j) ? djj * inv : c0;
      r1[j] = c1;
    } else {
      r1[j] = (lane == j - 32) ? djj * inv : c1;
    }
#pragma unroll
    for (int l = j + 1; l < CB; ++l) {  // A[i][l] -= L[i][j] L[l][j]  (entries right of the diagonal: unused garbage)
      const T colv = __shfl_sync(0xffffffffu, l < 32 ? c0 : c1, l & 31);
      if (l < 32) r0[l] -= c0 * colv;
      r1[l] -= c1 * colv;
    }
  }
#pragma unroll
  for (int c4 = 0; c4 < CB; c4 += 4) {
    const T a[4] = {r0[c4], r0[c4 + 1], r0[c4 + 2], r0[c4 + 3]};
    const T b4[4] = {r1[c4], r1[c4 + 1], r1[c4 + 2], r1[c4 + 3]};
    st4(&D[lane][c4], a);
    st4(&D[lane + 32][c4], b4);
  }
}

template <typename T>
__global__ void __launch_bounds__(CROWS) chol_panel_kernel(T* A, int64_t R, int k, int b, const JacobiScalars* sc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  A += blockIdx.y * R * R, sc += blockIdx.y;
  T(*D)[LDD] = reinterpret_cast<T(*)[LDD]>(smem_raw);
  T* dinv = reinterpret_cast<T*>(smem_raw) + CB * LDD;  // 1 / L_jj
  T(*X)[CB + 1] = reinterpret_cast<T(*)[CB + 1]>(dinv + CB);
  const int tid = threadIdx.x;
  const T floor_piv = chol_shift<T>(sc);
  const bool dbg = k == 0;
  if (dbg) VVT_DSTAMP(9);
#pragma unroll
  for (int it = 0; it < CB * CB / CROWS; ++it) {  // zero-padded to CB x CB; all loads in flight together
    const int idx = tid + it * CROWS, i = idx / CB, j = idx % CB;
    D[i][j] = (j <= i && i < b) ? A[(int64_t(k) + i) * R + k + j] : T(0);
  }
  // this CTA's rows of the panel
  const int64_t row0 = int64_t(k) + b + int64_t(blockIdx.x) * CROWS;
  const int nrows = int(vmax<int64_t>(0, vmin<int64_t>(CROWS, R - row0)));
#pragma unroll 8
  for (int idx = tid; idx < nrows * CB; idx += CROWS) {
    const int r = idx / CB, c = idx % CB;
    X[r][c] = c < b ? A[(row0 + r) * R + k + c] : T(0);
  }
  __syncthreads();
  if (dbg) VVT_DSTAMP(10);
  if (tid < 32) chol_diag_warp<T>(D, dinv, floor_piv, tid);
  __syncthreads();
  if (dbg) VVT_DSTAMP(11);
  if (blockIdx.x == 0) {
#pragma unroll 4
    for (int idx = tid; idx < CB * CB; idx += CROWS) {
      const int i = idx / CB, j = idx % CB;
      if (j <= i && i < b) A[(int64_t(k) + i) * R + k + j] = D[i][j];
    }
  }
  if (dbg) VVT_DSTAMP(12);
  if (tid < nrows) {  // x L11^T = a: the row lives in registers, rows of L11 are broadcast 16-byte reads
    T x[CB];
#pragma unroll
    for (int c = 0; c < CB; ++c) x[c] = X[tid][c];
#pragma unroll
    for (int c = 0; c < CB; ++c) {
      T s0 = x[c], s1 = T(0), s2 = T(0), s3 = T(0);  // four chains instead of one
#pragma unroll
      for (int m4 = 0; m4 < c; m4 += 4) {
        T d[4];
        ld4(&D[c][m4], d);
        if (m4 + 0 < c) s0 -= x[m4 + 0] * d[0];
        if (m4 + 1 < c) s1 -= x[m4 + 1] * d[1];
        if (m4 + 2 < c) s2 -= x[m4 + 2] * d[2];
        if (m4 + 3 < c) s3 -= x[m4 + 3] * d[3];
      }
      x[c] = ((s0 + s1) + (s2 + s3)) * dinv[c];  // columns >= b: x = 0 and a finite dinv (zero-padded block)
    }
#pragma unroll
    for (int c = 0; c < CB; ++c) X[tid][c] = x[c];
  }
  __syncthreads();
  if (dbg) VVT_DSTAMP(13);
#pragma unroll 8
  for (int idx = tid; idx < nrows * CB; idx += CROWS) {
    const int r = idx / CB, c = idx % CB;
    if (c < b) A[(row0 + r) * R + k + c] = X[r][c];
  }
  if (dbg) VVT_DSTAMP(14);
}

template <typename T>
static size_t chol_smem_bytes() {
  return (size_t(CB) * LDD + CB + size_t(CROWS) * (CB + 1)) * sizeof(T);
}

// A = Gs + eps I (the factorisation works in place on this copy)
template <typename T>
__global__ void chol_prepare_kernel(T* A, const T* Gs, int64_t R, const JacobiScalars* sc) {
  const int64_t total = R * R;
  A += blockIdx.y * total, Gs += blockIdx.y * total, sc += blockIdx.y;
  const T eps = chol_shift<T>(sc);
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x)
    A[idx] = Gs[idx] + ((idx / R == idx % R) ? eps : T(0));
}

// W = L / ||L||_F (lower triangle of A, zero above and in the padding); ||L||_F^2 = tr(G) + R eps <= sqrt(R) ||G||_F + R eps
template <typename T>
__global__ void onesided_init_chol_kernel(T* Y, const T* A, int64_t R, int Np, const JacobiScalars* sc,
                                          int64_t y_stride) {
  const int64_t total = int64_t(Np) * Np;
  Y += blockIdx.y * y_stride, A += blockIdx.y * R * R, sc += blockIdx.y;
  // any scale of the right magnitude will do (the rotations are scale free; thresholds assume O(1) columns)
  const double bound = sqrt(double(R)) * sqrt(sc->norm2) + double(R) * double(chol_shift<T>(sc));
  const T scale = bound > 0.0 ? T(1.0 / sqrt(bound)) : T(0);
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int i = int(idx / Np), j = int(idx % Np);  // row i, column j
    const T v = (i < R && j <= i) ? A[int64_t(i) * R + j] * scale : T(0);
    *(y_ptr(Y, Np, j / OB, i) + (j % OB)) = v;
  }
}

// inv[c] = 1 / ||W[:, c]||  (one warp per column)
template <typename T>
__global__ void onesided_colnorm_kernel(T* inv, const T* Y, int64_t R, int Np, int64_t y_stride) {
  const int lane = threadIdx.x & 31;
  inv += blockIdx.y * R, Y += blockIdx.y * y_stride;
  const int64_t c = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  if (c >= R) return;
  double s = 0.0;
  for (int r = lane; r < R; r += 32) {
    const double v = double(*(y_ptr(const_cast<T*>(Y), Np, int(c) / OB, r) + (int(c) % OB)));
    s += v * v;
  }
  s = warp_sum(s);
  if (lane == 0) inv[c] = s > 0.0 ? T(1.0 / sqrt(s)) : T(0);
}

// Jm[r][c] and Jt[c][r] <- normalised columns of the W part of Y
template <typename T>
__global__ void onesided_gather_w_kernel(T* Jm, T* Jt, const T* Y, const T* inv, int64_t R, int Np,
                                         int64_t y_stride) {
  const int64_t total = R * R;
  Jm += blockIdx.y * total, Jt += blockIdx.y * total, Y += blockIdx.y * y_stride, inv += blockIdx.y * R;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int r = int(idx / R), c = int(idx % R);
    Jm[idx] = *(y_ptr(const_cast<T*>(Y), Np, c / OB, r) + (c % OB)) * inv[c];
    const int c2 = int(idx / R), r2 = int(idx % R);
    Jt[int64_t(c2) * R + r2] = *(y_ptr(const_cast<T*>(Y), Np, c2 / OB, r2) + (c2 % OB)) * inv[c2];
  }
}

// Jm[r][c] (row-major) and Jt[c][r] (its transpose) <- J part of Y
template <typename T>
__global__ void onesided_gather_j_kernel(T* Jm, T* Jt, const T* Y, int64_t R, int Np, int64_t y_stride) {
  const int64_t total = R * R;
  Jm += blockIdx.y * total, Jt += blockIdx.y * total, Y += blockIdx.y * y_stride;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int r = int(idx / R), c = int(idx % R);
    Jm[idx] = *(y_ptr(const_cast<T*>(Y), Np, c / OB, Np + r) + (c % OB));
    // second pass with r fastest so that the Jt writes coalesce
    const int c2 = int(idx / R), r2 = int(idx % R);
    Jt[int64_t(c2) * R + r2] = *(y_ptr(const_cast<T*>(Y), Np, c2 / OB, Np + r2) + (c2 % OB));
  }
}

// ev[c] = (j_c . G j_c) / (j_c . j_c)  (one warp per column; rows of Jt and of Tt = Jt G, float64 sums)
template <typename T>
__global__ void onesided_rayleigh_kernel(T* ev, const T* Jt, const T* Tt, int64_t R, JacobiScalars* sc) {
  const int lane = threadIdx.x & 31;
  ev += blockIdx.y * R, Jt += blockIdx.y * R * R, Tt += blockIdx.y * R * R, sc += blockIdx.y;
  const int64_t c = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  if (c >= R) return;
  double num = 0.0, den = 0.0;
  for (int64_t r = lane; r < R; r += 32) {
    const double j = double(Jt[c * R + r]);
    num += j * double(Tt[c * R + r]);
    den += j * j;
  }
  num = warp_sum(num);
  den = warp_sum(den);
  if (lane == 0) {
    const double lam = den > 0.0 ? num / den : 0.0;
    ev[c] = T(lam);
    if (lam == lam) atomicMax(&sc->evmax_bits, (unsigned long long)__double_as_longlong(fabs(lam)));
  }
}

// One step of the Ogita-Aishima refinement of an approximate eigenvector matrix J:
//   S = J^T G J,  M = J^T J,  R = I - M,  lam_i = S_ii / M_ii,
//   E_ij = (S_ij + lam_j R_ij) / (lam_j - lam_i)   (i != j, eigenvalues separated),   E_ii = R_ii / 2,
//   J <- J + J E.
// It restores orthonormality and removes the first-order eigenvector error left by the Jacobi
// threshold; every product is a tensor-core GEMM.  This kernel writes Et[j][i] = E[i][j].
template <typename T>
__global__ void onesided_refine_coeff_kernel(T* Et, const T* S, const T* Mm, const T* ev, int64_t R,
                                             const JacobiScalars* sc) {
  const int64_t total = R * R;
  Et += blockIdx.y * total, S += blockIdx.y * total, Mm += blockIdx.y * total, ev += blockIdx.y * R, sc += blockIdx.y;
  // eigenvalue gaps below the rounding noise of S carry no information about the eigenvectors
  const T gap_min = T(256.0 * double(Eps<T>::v) * __longlong_as_double((long long)sc->evmax_bits));
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t j = idx / R, i = idx % R;  // Et[j][i] = E[i][j]
    T e;
    if (i == j) {
      e = (T(1) - Mm[i * R + i]) / T(2);
    } else {
      // both (i, j) and (j, i) read the same upper-triangle entries, so that they take the same
      // decision and E + E^T = R holds exactly (that is what restores orthonormality)
      const int64_t lo = i < j ? i : j, hi = i < j ? j : i;
      const T r = -Mm[lo * R + hi], sij = S[lo * R + hi], gap = ev[j] - ev[i];
      e = r / T(2);
      if (fabs(gap) > gap_min) {
        const T cand = (sij + ev[j] * r) / gap, other = r - cand;  // other = E[j][i]
        // larger: near-degenerate pair (noise over gap) -> orthogonality fix only
        if (fabs(cand) <= T(1e-3) && fabs(other) <= T(1e-3)) e = cand;
      }
    }
    Et[idx] = e;
  }
}

template <typename T>
__global__ void jacobi_rank_kernel(int* rank, const T* ev, int64_t R) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  rank += blockIdx.y * R, ev += blockIdx.y * R;
  if (i >= R) return;
  const T vi = ev[i];
  int r = 0;
  if (vi != vi) {  // NaNs sort first, by index
    for (int64_t j = 0; j < i; ++j) r += (ldg(ev + j) != ldg(ev + j));
  } else {
    for (int64_t j = 0; j < R; ++j) {
      const T vj = ldg(ev + j);
      r += (vj < vi) || (vj == vi && j < i) || (vj != vj);
    }
  }
  rank[i] = r;
}

// evals[rank[i]] = ev[i];  evecs[r][rank[i]] = Jt[i][r]
template <typename T>
__global__ void jacobi_permute_kernel(T* evals, T* evecs, const T* ev, const T* Jt, const int* rank, int64_t R) {
  const int64_t total = evecs ? R * R : R;
  evals += blockIdx.y * R, ev += blockIdx.y * R, Jt += blockIdx.y * R * R, rank += blockIdx.y * R;
  if (evecs) evecs += blockIdx.y * R * R;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = idx / R, i = idx % R;
    const int dst = rank[i];
    if (r == 0) evals[dst] = ev[i];
    if (evecs) evecs[r * R + dst] = Jt[i * R + r];
  }
}


// end of a sweep, one thread per problem: a sweep that rotated (almost) nothing ends the iteration of its
// problem -- the last few block pairs only carry cosines barely above the threshold, which the refinement step
// removes anyway.  The decision is taken on the device; the host only reads it (asynchronously).
__global__ void sweep_end_kernel(JacobiScalars* sc, int batch, unsigned long long thresh, int* state) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  if (!sc[b].converged) {
    sc[b].sweeps += 1;
    const unsigned long long rot = sc[b].rotations;
    sc[b].last_rotations = rot;
    if (rot > sc[b].max_rotations) sc[b].max_rotations = rot;
    // 1 / 256 of all block pairs, or 1 / 64 of the busiest sweep (measured on the cifar10_3c3d Gram: the sweeps
    // after the count has fallen to 5 % of the pairs change neither eigenvalues nor residuals -- the refinement
    // step takes care of what they would; relative to the busiest sweep because a rank-deficient matrix never
    // rotates more than a few pairs, and those must all come to rest)
    const unsigned long long stop = thresh > sc[b].max_rotations / 64 ? thresh : sc[b].max_rotations / 64;
    if (rot <= stop) sc[b].converged = 1;
  }
  sc[b].rotations = 0;
  state[3 * b] = sc[b].sweeps;
  state[3 * b + 1] = sc[b].converged;
  state[3 * b + 2] = int(sc[b].last_rotations > 0x7fffffffull ? 0x7fffffffull : sc[b].last_rotations);
}

}  // namespace vvt

#include "eig_wide.cuh"

namespace vvt {

constexpr int64_t kMaxBatch = 1024;  // problems per call (larger batches are processed in chunks)

static bool wide_enabled(int64_t R, int dtype) {
  // measured (fp32, synthetic spectrum, ms): R = 2560 wide 130 / 16-wide 97, R = 5120 wide 449 / 16-wide 838,
  // R = 10240 wide 2671 / 16-wide > 6000.  Read per call: tests force the wide path on smaller problems.
  const char* env = getenv("VVT_SYEVJ_WIDE_MIN");
  const int64_t min_r = env ? atoll(env) : 4096;
  return dtype == VVT_F32 && R >= min_r && tc::encode_fn() != nullptr;
}

struct JacobiLayout {
  int64_t Np, nb, y_elems;  // regular path: padded size, 16-column blocks, elements of Y per problem
  bool wide;
  wide::WidePlan wp;
  int64_t off_Y, off_Gs, off_Jm, off_Jt, off_Tt, off_S, off_M, off_Et, off_ev, off_cn, off_rank, off_sc, off_marks,
      off_dc, off_done, off_state, off_gemm, gemm_bytes, total;
};

static JacobiLayout jacobi_layout(int64_t R, int64_t B, int64_t es, int dtype, int world = 1) {
  JacobiLayout L;
  L.Np = align_up(R, OP);
  L.nb = L.Np / OB;
  L.wide = wide_enabled(R, dtype);
  L.wp = wide::wide_plan(R, B, world);
  L.y_elems = L.nb * 2 * L.Np * OB;
  if (L.wide) L.y_elems = vmax<int64_t>(L.y_elems, int64_t(L.wp.Np) * L.wp.Np);
  int64_t o = 0;
  auto take = [&](int64_t bytes) {
    const int64_t at = o;
    o += align_up(bytes, 256);
    return at;
  };
  L.off_Y = take(B * L.y_elems * es);
  L.off_Gs = take(B * R * R * es);
  L.off_Jm = take(B * R * R * es);
  L.off_Jt = take(B * R * R * es);
  L.off_Tt = take(vmax<int64_t>(B * R * R * es, L.wide ? L.wp.part_bytes : 0));  // wide path: split-K partial Grams
  L.off_S = take(vmax<int64_t>(B * R * R * es, L.wide ? L.wp.q_bytes : 0));      // wide path: Q^T of every pair
  // wide path: "rotated" flags, then the diagonal-block cache
  L.off_M = take(vmax<int64_t>(B * R * R * es, L.wide ? L.wp.flag_bytes + L.wp.diag_bytes : 0));
  L.off_Et = L.off_Y;  // Y is dead once J has been gathered (y_elems >= R * R)
  L.off_ev = take(B * R * es);
  L.off_cn = take(B * R * es);
  L.off_rank = take(B * R * 4);
  L.off_sc = take(B * sizeof(JacobiScalars));
  L.off_marks = take(B * (L.nb + L.nb * L.nb) * 4);
  L.off_dc = take(B * L.nb * OB * OB * es);
  L.off_done = take(B * L.nb * 4);
  L.off_state = take(B * 3 * 4);
  L.gemm_bytes = vvt_gram_workspace_bytes(R, R, R, dtype);
  if (world > 1)  // the products of the refinement step are split by rows over the ranks
    L.gemm_bytes = vmax<int64_t>(L.gemm_bytes, vvt_gram_workspace_bytes(ceil_div(R, int64_t(world)), R, R, dtype));
  L.off_gemm = take(L.gemm_bytes);
  L.total = o;
  return L;
}

template <typename T>
static size_t resident_smem_bytes(int rows_per_cta, int parts) {
  const size_t stage = sizeof(T) == 4 && kJacobiMma ? size_t(kGramStageFloats) * 4 : 0;
  return align_up(sizeof(RotSmem<T>), 16) + size_t(parts) * rows_per_cta * LDP * sizeof(T) + stage;
}

template <typename T>
static int launch_round(T* Y, int Np, int nb, int round, int CL, int rows_per_cta, int parts, bool resident,
                        JacobiScalars* sc, int* marks, int now, T* Dc, unsigned* done, BatchStrides bs, int batch,
                        bool share_sms, cudaStream_t s) {
  auto kern = resident ? onesided_round_resident_kernel<T> : onesided_round_stream_kernel<T>;
  size_t smem = resident ? resident_smem_bytes<T>(rows_per_cta, parts) : sizeof(StreamSmem<T>);
  // a round that fits on the GPU with one CTA per SM asks for more than half of an SM's shared memory, so
  // that no two CTAs share an SM (and its tensor pipe) while other SMs idle
  // (share_sms: late sweeps, in which few pairs still rotate -- most CTAs exit at once, and with two CTAs per SM
  // the next rounds' CTAs are resident early and start the moment their two blocks are done)
  static const bool no_pad = getenv("VVT_SYEVJ_NOPAD") != nullptr;  // experiments
  if (!no_pad && !share_sms && int64_t(nb / 2) * CL * batch <= num_sms()) smem = vmax<size_t>(smem, size_t(116) * 1024);
  static SmemOptIn opt_in[2];  // per instantiation and kernel variant
  VVT_TRY(opt_in[resident].ensure(kern, smem, "vvt_syevj(attr)"));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(unsigned(nb / 2 * CL), unsigned(batch));
  cfg.blockDim = dim3(OT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  static const bool pdl = getenv("VVT_SYEVJ_NOPDL") == nullptr;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = unsigned(CL);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  VVT_TRY(check_cuda(cudaLaunchKernelEx(&cfg, kern, Y, Np, nb, round, rows_per_cta, parts, sc, marks, now, Dc, done, bs),
                     "vvt_syevj(round)"));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return VVT_OK;
}

template <typename T>
static size_t sweep_smem_bytes(int rows_per_cta, int parts) {
  // more than half of an SM's shared memory: one CTA per SM, every CTA of the launch resident at once
  return vmax<size_t>(resident_smem_bytes<T>(rows_per_cta, parts), size_t(116) * 1024);
}

// Can `clusters` clusters of CL CTAs of the sweep kernel be resident at the same time?  (They spin on each other:
// a cluster that is not scheduled would never signal.)  Asked of the occupancy calculator, which knows the GPC
// layout of the device.
template <typename T>
static bool sweep_fits(int CL, int rows_per_cta, int parts, int64_t clusters) {
  auto kern = onesided_sweep_resident_kernel<T>;
  const size_t smem = sweep_smem_bytes<T>(rows_per_cta, parts);
  static SmemOptIn opt_in;
  if (opt_in.ensure(kern, smem, "vvt_syevj(attr)") != VVT_OK) return false;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(unsigned(clusters * CL));
  cfg.blockDim = dim3(OT);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = unsigned(CL);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int max_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return int64_t(max_clusters) >= clusters;
}

template <typename T>
static int launch_sweep(T* Y, int Np, int nb, int CL, int rows_per_cta, int parts, JacobiScalars* sc, int* marks, int now0,
                        T* Dc, unsigned* done, BatchStrides bs, int batch, cudaStream_t s) {
  auto kern = onesided_sweep_resident_kernel<T>;
  const size_t smem = sweep_smem_bytes<T>(rows_per_cta, parts);
  static SmemOptIn opt_in;  // per instantiation
  VVT_TRY(opt_in.ensure(kern, smem, "vvt_syevj(attr)"));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(unsigned(nb / 2 * CL), unsigned(batch));
  cfg.blockDim = dim3(OT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = unsigned(CL);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  VVT_TRY(check_cuda(cudaLaunchKernelEx(&cfg, kern, Y, Np, nb, rows_per_cta, parts, sc, marks, now0, Dc, done, bs),
                     "vvt_syevj(sweep)"));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return VVT_OK;
}

// Host-visible convergence state without stalling the GPU: after every sweep the device writes
// {sweeps, converged} per problem, an async copy brings it to pinned memory and an event marks its arrival.
// The host enqueues sweep s + 1 BEFORE it waits for the state of sweep s: if that sweep turns out to have been
// the last one, the kernels of the speculative sweep exit on their first instruction (converged flag).
struct HostState {
  int* pinned = nullptr;  // [2 slots][3 * kMaxBatch]: {sweeps, converged, block pairs rotated} per problem
  cudaEvent_t ev[2] = {nullptr, nullptr};
};
static int host_state(HostState** out) {
  static thread_local HostState st[kMaxDevices];
  HostState& h = st[current_device()];
  if (!h.pinned) {
    VVT_TRY(check_cuda(cudaHostAlloc(reinterpret_cast<void**>(&h.pinned), 2 * 3 * kMaxBatch * sizeof(int), cudaHostAllocDefault),
                       "vvt_syevj(pinned)"));
    for (int i = 0; i < 2; ++i)
      VVT_TRY(check_cuda(cudaEventCreateWithFlags(&h.ev[i], cudaEventDisableTiming), "vvt_syevj(event)"));
  }
  *out = &h;
  return VVT_OK;
}

// Distributed two-level rounds (vvt_syevj_dist): the pairs of a wide round are independent -- disjoint column blocks
// -- so rank g of `world` takes the contiguous slice [pairs g / world, pairs (g + 1) / world) of the pair index and
// runs the three kernels of the round on it.  In the round-robin tournament pair i of round r + 1 takes its first
// block from pair i + 1 and its second from pair i - 1 of round r, so between two cross rounds a rank hands ONE block
// to each neighbour (Np x 64 floats, 2.6 MB at R = 10240) and receives one from each: grouped ncclSend / ncclRecv
// straight between the block slots of the factor, on the compute stream.  The intra round of a sweep pairs blocks
// (2i, 2i + 1): a general permutation, once per sweep.  `owner[w]` is the rank that holds the valid copy of wide block
// w (-1: every rank, the replicated state after the Cholesky factor); all ranks compute the same transfer lists.
struct DistCtx {
  void* comm;
  int rank, world;
  bool p2p;  // the caller vouches that the peer-memory arena of this device is mapped for the ranks of comm
};

// rank that works on pair `pair` of a round: contiguous, balanced slices of the pair index
static inline int dist_rank_of_pair(int pairs, int world, int pair) {
  int r = 0;
  while (int64_t(pairs) * (r + 1) / world <= pair) ++r;
  return r;
}

struct BlockMove {
  int block, src, dst;
};
// Blocks that must change owner before round `round` can start, in a fixed order every rank derives alike (pair by
// pair, first block then second); updates owner[] to the ownership during that round.  Host logic only: also
// reachable through the test hook vvt_dbg_dist_plan.
static void dist_plan_round(int nbw, int world, int round, std::vector<int>& owner, std::vector<BlockMove>& moves) {
  moves.clear();
  const int pairs = nbw / 2;
  for (int pair = 0; pair < pairs; ++pair) {
    int w[2];
    wide::wide_blocks(nbw, round, pair, w[0], w[1]);
    const int to = dist_rank_of_pair(pairs, world, pair);
    for (int k = 0; k < 2; ++k) {
      const int src = owner[size_t(w[k])];
      owner[size_t(w[k])] = to;
      if (src >= 0 && src != to) moves.push_back({w[k], src, to});
    }
  }
}

// ---- peer-memory arena: the block hand-over without NCCL -----------------------------------------------------
// Every rank allocates one arena (cudaMalloc; flags in the first 4 KB, the factor of the two-level solver behind
// them), exports it with cudaIpcGetMemHandle and maps the arenas of its peers (vvt_dist_arena_alloc / _open; the
// handles travel through torch.distributed in vivit_b200/dist.py).  A rank then WRITES the blocks that change owner
// straight into the next owner's factor over NVLink (push_blocks_kernel: 16-byte stores, fence), raises a
// monotonically increasing per-sender flag in the receiver's arena and waits for its own senders' flags
// (signal_wait_kernel) -- ~10 us per round instead of ~35 us for a grouped ncclSend / ncclRecv of 2 x 2.6 MB.
// Hazards: a slot is overwritten by a peer only after that peer has waited for this rank's flag of the round
// before, which this rank raises after its own last read of the slot (its push of that block).
constexpr int kMaxPeers = 16, kArenaHeader = 4096, kMaxPush = 96;
struct DistArena {
  char* base = nullptr;
  int64_t bytes = 0;  // of the factor region
  char* peer[kMaxPeers] = {};
  int world = 0, rank = -1;
  unsigned long long seq = 0;  // rounds signalled so far: identical on every rank (all ranks enqueue the same rounds)
};
static DistArena& arena() {
  static DistArena a[kMaxDevices];
  return a[current_device()];
}

struct PushArgs {
  float* dst[kMaxPush];
  const float* src[kMaxPush];
  int n;
};
__global__ void __launch_bounds__(256) push_blocks_kernel(PushArgs a, int64_t vec4) {
  const float4* src = reinterpret_cast<const float4*>(a.src[blockIdx.y]);
  float4* dst = reinterpret_cast<float4*>(a.dst[blockIdx.y]);
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < vec4; i += 4 * stride) {  // four loads in flight per thread
    const float4 v0 = src[i], v1 = src[i + stride], v2 = src[i + 2 * stride], v3 = src[i + 3 * stride];
    dst[i] = v0, dst[i + stride] = v1, dst[i + 2 * stride] = v2, dst[i + 3 * stride] = v3;
  }
  for (; i < vec4; i += stride) dst[i] = src[i];
  __threadfence_system();
}
struct SignalArgs {
  unsigned long long* raise[kMaxPeers];  // this rank's flag slot in the arena of a rank that was sent something
  const unsigned long long* wait[kMaxPeers];  // local slots of the ranks that sent something here
  unsigned long long seq;
};
__global__ void signal_wait_kernel(SignalArgs a) {
  const int t = threadIdx.x;
  if (t >= kMaxPeers) return;
  if (a.raise[t]) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.raise[t]), "l"(a.seq) : "memory");
  }
  if (a.wait[t]) {
    unsigned long long v;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a.wait[t]) : "memory");
    } while (v < a.seq);
  }
}

// every rank has reached this point of its stream (flags only): before the first block is pushed, every rank must
// have finished writing its own copy of the initial factor
static int dist_barrier_p2p(const DistCtx& d, cudaStream_t s) {
  DistArena& ar = arena();
  SignalArgs sig = {};
  for (int r = 0; r < d.world; ++r) {
    if (r == d.rank) continue;
    sig.raise[r] = reinterpret_cast<unsigned long long*>(ar.peer[r]) + d.rank;
    sig.wait[r] = reinterpret_cast<const unsigned long long*>(ar.base) + r;
  }
  sig.seq = ++ar.seq;
  signal_wait_kernel<<<1, 32, 0, s>>>(sig);
  return launched("vvt_syevj_dist(barrier)");
}

static int dist_exchange_p2p(float* Lw, const wide::WidePlan& p, int round, const DistCtx& d, std::vector<int>& owner,
                             cudaStream_t s) {
  DistArena& ar = arena();
  const size_t blk = size_t(p.Np) * wide::WB;
  PushArgs push;
  push.n = 0;
  SignalArgs sig = {};
  bool any = false;
  auto flush_push = [&]() -> int {
    if (push.n == 0) return VVT_OK;
    const unsigned per = unsigned(vmax(1, vmin(num_sms() / push.n, 64)));
    push_blocks_kernel<<<dim3(per, unsigned(push.n)), 256, 0, s>>>(push, int64_t(blk / 4));
    push.n = 0;
    return launched("vvt_syevj_dist(push)");
  };
  std::vector<BlockMove> moves;
  dist_plan_round(p.nbw, d.world, round, owner, moves);
  for (const BlockMove& m : moves) {
    if (d.rank != m.src && d.rank != m.dst) continue;
    any = true;
    if (d.rank == m.src) {
      const size_t off = size_t(kArenaHeader) + (size_t(m.block) * blk) * sizeof(float);
      push.src[push.n] = Lw + size_t(m.block) * blk;
      push.dst[push.n] = reinterpret_cast<float*>(ar.peer[m.dst] + off);
      if (++push.n == kMaxPush) VVT_TRY(flush_push());
      sig.raise[m.dst] = reinterpret_cast<unsigned long long*>(ar.peer[m.dst]) + d.rank;
    } else {
      sig.wait[m.src] = reinterpret_cast<const unsigned long long*>(ar.base) + m.src;
    }
  }
  ++ar.seq;  // every rank counts every round, also the ones it takes no part in
  if (!any) return VVT_OK;
  VVT_TRY(flush_push());
  sig.seq = ar.seq;
  signal_wait_kernel<<<1, 32, 0, s>>>(sig);
  return launched("vvt_syevj_dist(signal)");
}

static int dist_exchange(float* Lw, const wide::WidePlan& p, int round, const DistCtx& d, std::vector<int>& owner,
                         cudaStream_t s) {
  const Nccl& n = nccl();
  const size_t blk = size_t(p.Np) * wide::WB;
  bool open = false;
  int st = VVT_OK;
  std::vector<BlockMove> moves;
  dist_plan_round(p.nbw, d.world, round, owner, moves);
  for (const BlockMove& m : moves) {
    if (st != VVT_OK) break;
    if (d.rank != m.src && d.rank != m.dst) continue;
    if (!open) {
      st = check_nccl(n.group_start(), "vvt_syevj_dist");
      open = true;
      if (st != VVT_OK) break;
    }
    float* ptr = Lw + size_t(m.block) * blk;
    st = d.rank == m.src ? check_nccl(n.send(ptr, blk, kNcclFloat32, m.dst, d.comm, s), "vvt_syevj_dist(send)")
                         : check_nccl(n.recv(ptr, blk, kNcclFloat32, m.src, d.comm, s), "vvt_syevj_dist(recv)");
  }
  if (open) {
    const int e = check_nccl(n.group_end(), "vvt_syevj_dist");
    if (st == VVT_OK) st = e;
  }
  return st;
}

static bool debug_env() { return getenv("VVT_SYEVJ_DEBUG") != nullptr; }

template <typename T>
static int syevj_impl(T* evals, T* evecs, const T* G, int64_t R, int64_t B, int jobz, char* ws, int* info, int dtype,
                      cudaStream_t s, const DistCtx* dist_in = nullptr) {
  const JacobiLayout L = jacobi_layout(R, B, sizeof(T), dtype, dist_in ? dist_in->world : 1);
  // the rounds are distributed for one fp32 problem on the two-level path; anything else is solved by every rank
  // for itself (identical results: the solver is deterministic)
  const DistCtx* dist = dist_in && dist_in->world > 1 && L.wide && B == 1 && sizeof(T) == 4 ? dist_in : nullptr;
  std::vector<int> owner;
  int pair_lo = 0, pair_hi = L.wp.pairs;
  // blocks change owner through peer memory when every rank has mapped the others' arenas (vvt_dist_arena_open)
  // and the factor fits; through grouped ncclSend / ncclRecv otherwise (VVT_SYEVJ_P2P=0 forces that)
  bool p2p = false;
  if (dist) {
    const DistArena& ar = arena();
    const char* env = getenv("VVT_SYEVJ_P2P");
    p2p = dist->p2p && !(env && atoi(env) == 0) && ar.base && ar.world == dist->world && ar.rank == dist->rank &&
          ar.bytes >= int64_t(L.wp.Np) * L.wp.Np * 4;
    if (debug_env())
      fprintf(stderr, "[vvt_syevj] rank %d of %d: blocks change owner through %s\n", dist->rank, dist->world,
              p2p ? "peer memory" : "NCCL send / recv");
  }
  if (dist) {
    owner.assign(size_t(L.wp.nbw), -1);
    pair_lo = int(int64_t(L.wp.pairs) * dist->rank / dist->world);  // dist_rank_of_pair(pair) == rank on [lo, hi)
    pair_hi = int(int64_t(L.wp.pairs) * (dist->rank + 1) / dist->world);
  }
  T* Y = p2p ? (T*)(arena().base + kArenaHeader) : (T*)(ws + L.off_Y);  // the factor lives where the peers can write
  T* Gs = (T*)(ws + L.off_Gs);
  T* Jm = (T*)(ws + L.off_Jm);
  T* Jt = (T*)(ws + L.off_Jt);
  T* Tt = (T*)(ws + L.off_Tt);
  T* Sm = (T*)(ws + L.off_S);
  T* Mm = (T*)(ws + L.off_M);
  T* Et = (T*)(ws + L.off_Et);
  T* ev = (T*)(ws + L.off_ev);
  T* cn = (T*)(ws + L.off_cn);
  int* rank = (int*)(ws + L.off_rank);
  JacobiScalars* sc = (JacobiScalars*)(ws + L.off_sc);
  int* marks = (int*)(ws + L.off_marks);
  int* state = (int*)(ws + L.off_state);
  const int Np = int(L.Np), nb = int(L.nb), pairs = nb / 2;
  const unsigned UB = unsigned(B);
  const int64_t RR = R * R;
  HostState* hs = nullptr;
  VVT_TRY(host_state(&hs));

  // Cholesky-preconditioned variant (default): rotate the columns of L, no J; VVT_SYEVJ_NOCHOL=1 selects
  // the plain W = G J variant (16-wide path only)
  static const bool chol_env = getenv("VVT_SYEVJ_NOCHOL") == nullptr;
  const bool use_chol = chol_env || L.wide;
  const int parts = use_chol ? 1 : 2;

  // rows of the panel are split over the CTAs of a cluster.  A CTA's time is mostly fixed cost (barriers,
  // rotation rounds, latencies: ~11 us + ~4 ns per row, fp32), so once a round no longer fits on the GPU the
  // fewest CTAs win: the smallest cluster whose rows still fit in shared memory (resident variant: 1072 rows
  // fp32, 588 rows fp64).  Measured (fp32, ms): R=2560 CL 3/4/5 = 94/113/132, R=3840 CL 4/5/8 = 330/360/488,
  // R=5120 CL 5/6/8 = 836/909/1083; fp64 R=2560 CL 4/8 = 458/775.
  int CL = 1;
  const int max_rows = sizeof(T) == 4 ? 1024 : 576;
  const int fill_rows = sizeof(T) == 4 ? 80 : 160;  // fp32: R=320 CL 1/4 = 3.8/3.4 ms, R=640 CL 2/4 = 8.6/8.0 ms
  const int fill_max = sizeof(T) == 4 ? 4 : 8;
  while (CL < 8 && ceil_div(Np, CL) > max_rows) ++CL;
  // SMs left idle by a round (pairs * CL * batch < #SMs) are put to work with a larger, possibly odd cluster,
  // as long as a CTA keeps enough rows for the Gram / apply phases to outweigh the cluster reduction (R = 1280,
  // fp32: 3 CTAs x 427 rows on 120 SMs, 16.3 ms against 17.7 ms with 2 x 640 on 80 SMs)
  while (CL < fill_max && int64_t(pairs) * (CL + 1) * B <= num_sms() && Np / (CL + 1) >= fill_rows) ++CL;
  if (const char* e = getenv("VVT_SYEVJ_CL")) CL = vmax(1, vmin(8, atoi(e)));  // experiments
  while (CL > 1 && Np / CL < 16) CL /= 2;
  const int rows_per_cta = int(ceil_div(Np, CL));
  const bool resident = resident_smem_bytes<T>(rows_per_cta, parts) <= size_t(200) * 1024;

  VVT_TRY(check_cuda(cudaMemsetAsync(sc, 0, size_t(B) * sizeof(JacobiScalars), s), "vvt_syevj"));
  VVT_TRY(check_cuda(cudaMemsetAsync(marks, 0, size_t(B) * size_t(nb + nb * nb) * 4, s), "vvt_syevj"));
  VVT_TRY(check_cuda(cudaMemsetAsync(ws + L.off_done, 0, size_t(B) * nb * 4, s), "vvt_syevj"));
  const bool debug = getenv("VVT_SYEVJ_DEBUG") != nullptr;
  const int init_blocks = int(vmin<int64_t>(ceil_div(L.y_elems / 2, 256), 8 * num_sms()));
  const int rr_blocks = int(vmin<int64_t>(ceil_div(RR, 256), 8 * num_sms()));
  // one GEMM per problem on the tcgen05 / DMMA kernel when the problems are large, one batched launch otherwise
  const bool loop_gemm = B == 1 || R >= 1024;
  // (the trailing updates of the Cholesky factor, K = 64, of problems up to 2048 columns are latency-bound: one
  // mma.sync launch for all problems, 16.9 -> 16.7 ms for two R = 1280 problems; larger ones are worth a tcgen05
  // launch each -- batching them cost 24 ms at 2 x R = 5120)
  auto gemm_batched = [&](T* C, const T* A, const T* Bm, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                          int64_t ldc, double alpha, double beta) -> int {
    if (loop_gemm && (B == 1 || K > CB || R > 2048)) {
      for (int64_t b = 0; b < B; ++b)
        VVT_TRY(vvt_gemm(C + b * RR, A + b * RR, Bm + b * RR, M, N, K, 0, 0, lda, ldb, ldc, alpha, beta, 1, 0, 0, 0,
                         ws + L.off_gemm, L.gemm_bytes, dtype, (void*)s));
      return VVT_OK;
    }
    return vvt_gemm(C, A, Bm, M, N, K, 0, 0, lda, ldb, ldc, alpha, beta, B, RR, RR, RR, nullptr, 0, dtype, (void*)s);
  };

  // VVT_SYEVJ_DEBUG: host time stamps behind stream drains (factor / sweeps / gather + refinement)
  auto phase_clock = std::chrono::steady_clock::now();
  auto phase = [&](const char* what) {
    if (!debug) return;
    cudaStreamSynchronize(s);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[vvt_syevj] R=%lld %s%s: %.2f ms\n", (long long)R, dist ? "distributed " : "", what,
            std::chrono::duration<double, std::milli>(now - phase_clock).count());
    phase_clock = now;
  };
  phase("start");
  onesided_sym_kernel<T><<<dim3(rr_blocks, UB), 256, 0, s>>>(Gs, G, R, sc);
  VVT_TRY(launched("vvt_syevj(sym)"));
  if (use_chol) {
    T* A = Jm;  // free until the gather
    chol_prepare_kernel<T><<<dim3(rr_blocks, UB), 256, 0, s>>>(A, Gs, R, sc);
    VVT_TRY(launched("vvt_syevj(chol prepare)"));
    static SmemOptIn chol_opt;  // per instantiation
    VVT_TRY(chol_opt.ensure(chol_panel_kernel<T>, chol_smem_bytes<T>(), "vvt_syevj(chol attr)"));
    for (int64_t k = 0; k < R; k += CB) {
      const int b = int(vmin<int64_t>(CB, R - k));
      const int64_t below = R - k - b;
      const unsigned ctas = unsigned(vmax<int64_t>(1, ceil_div(below, CROWS)));
      chol_panel_kernel<T><<<dim3(ctas, UB), CROWS, chol_smem_bytes<T>(), s>>>(A, R, int(k), b, sc);
      VVT_TRY(launched("vvt_syevj(chol panel)"));
      if (below > 0) {  // A22 -= L21 L21^T
        T* A22 = A + (k + b) * R + (k + b);
        const T* L21 = A + (k + b) * R + k;
        VVT_TRY(gemm_batched(A22, L21, L21, below, below, b, R, R, R, -1.0, 1.0));
      }
    }
    if (L.wide) {
      if constexpr (sizeof(T) == 4) {
        wide::wide_init_chol_kernel<<<dim3(init_blocks, UB), 256, 0, s>>>((float*)Y, (const float*)A, R, L.wp.Np, sc);
        VVT_TRY(launched("vvt_syevj(wide init)"));
        if (dist && p2p) VVT_TRY(dist_barrier_p2p(*dist, s));
      }
    } else {
      onesided_init_chol_kernel<T><<<dim3(init_blocks, UB), 256, 0, s>>>(Y, A, R, Np, sc, L.y_elems);
      VVT_TRY(launched("vvt_syevj(init)"));
    }
  } else {
    onesided_init_kernel<T><<<dim3(init_blocks, UB), 256, 0, s>>>(Y, Gs, R, Np, sc, L.y_elems);
    VVT_TRY(launched("vvt_syevj(init)"));
  }

  // per-block dependencies between rounds (wait_blocks) or the grid-wide dependency of griddepcontrol.wait
  // (VVT_SYEVJ_GRIDDEP=1 forces the latter).  Spinning CTAs hold SM slots, so block dependencies pay while a
  // round fits on the GPU a few times over (measured: R=2560 fp32 131 -> 115 ms, R=1280 fp64 110 -> 74 ms,
  // R=1280 fp32 unchanged) and cost ~7% once a round is many waves (R=5120: 1280 CTAs).
  static const bool block_deps_env = getenv("VVT_SYEVJ_GRIDDEP") == nullptr;
  const bool block_deps = block_deps_env && int64_t(pairs) * CL * B <= 4 * int64_t(num_sms());
  const BatchStrides bs{L.y_elems, int64_t(nb + nb * nb), int64_t(nb) * OB * OB, int64_t(nb)};
  // VVT_SYEVJ_PERSIST=1: one launch per sweep (onesided_sweep_resident_kernel) while every CTA of a round can be
  // resident at once.  Measured SLOWER than one launch per round with programmatic dependent launch (R = 1280:
  // 16.6 against 13.1 ms, R = 320: 2.8 against 2.5 ms; 12 launches per solve instead of 972): a cluster that
  // keeps its pair slot adds "my own previous round" to the two block dependencies of every round, while freshly
  // launched CTAs start on whichever SM is free.  Kept for hosts where launch overhead is what counts.
  static const bool persist_env = getenv("VVT_SYEVJ_PERSIST") != nullptr;
  const bool persistent = persist_env && !L.wide && resident && block_deps && int64_t(pairs) * CL * B <= num_sms() &&
                          sweep_fits<T>(CL, rows_per_cta, parts, int64_t(pairs) * B);
  static const int stop_div = getenv("VVT_SYEVJ_STOPDIV") ? vmax(1, atoi(getenv("VVT_SYEVJ_STOPDIV"))) : 256;
  // both paths count the 16-column block pairs that still rotated; a sweep visits nb16 / 2 * nb16 of them
  const int64_t nb16 = L.wide ? L.wp.Np / OB : nb;
  const unsigned long long thresh = (unsigned long long)((nb16 / 2) * nb16) / stop_div;
  wide::WideMaps maps;
  if (L.wide) VVT_TRY(wide::wide_make_maps(&maps, (const float*)Y, (const float*)Sm, L.wp, B));

  phase("symmetrise, Cholesky factor, init");
  long long most_rotations = -1;  // over the unfinished problems, in the last sweep whose state has been read
  static const int share_div = getenv("VVT_SYEVJ_SHAREDIV") ? atoi(getenv("VVT_SYEVJ_SHAREDIV")) : 3;
  auto enqueue_sweep = [&](int sweep) -> int {
    const bool share_sms = share_div > 0 && most_rotations >= 0 && most_rotations * share_div < (nb16 / 2) * nb16;
    if (L.wide) {
      if constexpr (sizeof(T) == 4) {
        for (int round = -1; round < L.wp.nbw - 1; ++round) {
          if (dist) VVT_TRY((p2p ? dist_exchange_p2p : dist_exchange)((float*)Y, L.wp, round, *dist, owner, s));
          VVT_TRY(wide::wide_round((float*)Y, (float*)Tt, (float*)Sm, (int*)Mm, (float*)((char*)Mm + L.wp.flag_bytes), sc,
                                   L.wp, maps, round, B, s, pair_lo, pair_hi - pair_lo));
        }
        // the sweep's count of block pairs that rotated, summed over the ranks: sweep_end_kernel then takes the same
        // decision everywhere (and the hosts, which read it, enqueue the same sweeps)
        if (dist)
          VVT_TRY(check_nccl(nccl().all_reduce(&sc->rotations, &sc->rotations, 1, kNcclUint64, kNcclSum, dist->comm, s),
                             "vvt_syevj_dist(rotations)"));
      }
    } else if (persistent) {
      VVT_TRY(launch_sweep<T>(Y, Np, nb, CL, rows_per_cta, parts, sc, marks, sweep * nb + 1, (T*)(ws + L.off_dc),
                              (unsigned*)(ws + L.off_done), bs, int(B), s));
    } else {
      for (int round = -1; round < nb - 1; ++round)
        VVT_TRY(launch_round<T>(Y, Np, nb, round, CL, rows_per_cta, parts, resident, sc, marks, sweep * nb + round + 2,
                                (T*)(ws + L.off_dc), block_deps ? (unsigned*)(ws + L.off_done) : nullptr, bs, int(B), share_sms,
                                s));
    }
    sweep_end_kernel<<<unsigned(ceil_div(B, 128)), 128, 0, s>>>(sc, int(B), thresh, state);
    VVT_TRY(launched("vvt_syevj(sweep end)"));
    int* slot = hs->pinned + (sweep & 1) * 3 * kMaxBatch;
    VVT_TRY(check_cuda(cudaMemcpyAsync(slot, state, size_t(B) * 3 * sizeof(int), cudaMemcpyDeviceToHost, s), "vvt_syevj"));
    VVT_TRY(check_cuda(cudaEventRecord(hs->ev[sweep & 1], s), "vvt_syevj"));
    return VVT_OK;
  };
  auto all_converged = [&](int sweep) -> int {  // waits for the state of `sweep`; 1 = every problem is done
    if (cudaEventSynchronize(hs->ev[sweep & 1]) != cudaSuccess) return -1;
    const int* slot = hs->pinned + (sweep & 1) * 3 * kMaxBatch;
    int all = 1;
    most_rotations = 0;
    for (int64_t b = 0; b < B; ++b) {
      all &= slot[3 * b + 1];
      if (!slot[3 * b + 1]) most_rotations = vmax<long long>(most_rotations, slot[3 * b + 2]);
      if (info) info[2 * b] = slot[3 * b], info[2 * b + 1] = slot[3 * b + 1];
    }
    if (debug) {  // host time at which the state of this sweep arrived: differences are the sweeps' GPU times
      static thread_local std::chrono::steady_clock::time_point last;
      const auto now = std::chrono::steady_clock::now();
      const double us = sweep == 0 ? 0.0 : std::chrono::duration<double, std::micro>(now - last).count();
      last = now;
      fprintf(stderr,
              "[vvt_syevj] R=%lld batch=%lld %s CL=%d sweep %d: %d of %lld block pairs rotated (problem 0), +%.0f us\n",
              (long long)R, (long long)B, L.wide ? "wide" : "16-wide", CL, sweep + 1, slot[2],
              (long long)((nb16 / 2) * nb16), us);
      if (L.wide) {
        long long w[9];
        if (cudaMemcpyFromSymbol(w, wide::g_wdbg, sizeof(w)) == cudaSuccess)
          fprintf(stderr, "[vvt_syevj]   rotation kernel, CTA 0, last launch (cycles): sum partials %lld | rotations %lld | "
                  "cluster barrier %lld | Q_sub exchange %lld | columns %lld | rows %lld | cluster barrier %lld | re-pair %lld | "
                  "store %lld\n", w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], w[8]);
      }
    }
    return all;
  };
  {
    // The host enqueues sweep s + 1 before it has seen the outcome of sweep s (no bubble on the GPU) -- unless
    // sweep s - 1 already rotated next to nothing: then sweep s is very likely the last one, and waiting for it
    // (one short bubble) is cheaper than a sweep of launches that exit on their first instruction.
    int sweep = 0, done_flag = 0, enqueued = 1;
    VVT_TRY(enqueue_sweep(0));
    while (true) {
      const bool more = sweep + 1 < kMaxSweeps;
      const bool speculate = most_rotations < 0 || most_rotations > 8 * (long long)(thresh + 1);
      if (more && speculate && enqueued == sweep + 1) {
        VVT_TRY(enqueue_sweep(sweep + 1));
        ++enqueued;
      }
      done_flag = all_converged(sweep);
      if (done_flag < 0) return fail(VVT_ERR_CUDA, "%s: %s", "vvt_syevj", cudaGetErrorString(cudaGetLastError()));
      if (done_flag || !more) break;
      if (enqueued == sweep + 1) {  // not speculated: the outcome is known now, and it says "go on"
        VVT_TRY(enqueue_sweep(sweep + 1));
        ++enqueued;
      }
      ++sweep;
    }
  }

  phase("sweeps");
  const int gblocks = int(vmin<int64_t>(ceil_div(RR, 256), 8 * num_sms()));
  if (dist) {
    // every rank gets the whole rotated factor back: blocks held elsewhere are zeroed, one all-reduce sums
    // x + 0 + ... + 0 (exact), and from here on the ranks hold identical data again
    const size_t blk = size_t(L.wp.Np) * wide::WB;
    for (int w = 0; w < L.wp.nbw; ++w)
      if (owner[size_t(w)] >= 0 && owner[size_t(w)] != dist->rank)
        VVT_TRY(check_cuda(cudaMemsetAsync((float*)Y + size_t(w) * blk, 0, blk * sizeof(float), s), "vvt_syevj_dist"));
    VVT_TRY(check_nccl(nccl().all_reduce(Y, Y, size_t(L.wp.nbw) * blk, kNcclFloat32, kNcclSum, dist->comm, s),
                       "vvt_syevj_dist(gather)"));
  }
  if (L.wide) {
    if constexpr (sizeof(T) == 4) {
      wide::wide_colnorm_kernel<<<dim3(unsigned(ceil_div(R, 32)), UB), 256, 0, s>>>((float*)cn, (const float*)Y, R, L.wp.Np);
      VVT_TRY(launched("vvt_syevj(colnorm)"));
      wide::wide_gather_kernel<<<dim3(unsigned(ceil_div(R, 32)), unsigned(ceil_div(R, 32)), UB), 256, 0, s>>>(
          (float*)Jm, (float*)Jt, (const float*)Y, (const float*)cn, R, L.wp.Np);
    }
  } else if (use_chol) {
    onesided_colnorm_kernel<T><<<dim3(unsigned(ceil_div(R * 32, 256)), UB), 256, 0, s>>>(cn, Y, R, Np, L.y_elems);
    VVT_TRY(launched("vvt_syevj(colnorm)"));
    onesided_gather_w_kernel<T><<<dim3(gblocks, UB), 256, 0, s>>>(Jm, Jt, Y, cn, R, Np, L.y_elems);
  } else {
    onesided_gather_j_kernel<T><<<dim3(gblocks, UB), 256, 0, s>>>(Jm, Jt, Y, R, Np, L.y_elems);
  }
  VVT_TRY(launched("vvt_syevj(gather)"));
  // every product below is C = A B^T with K-contiguous operands: the tcgen05 (fp32) / DMMA (fp64) GEMM
  // distributed solve: a rank forms R / world rows of every product and one all-gather puts the matrix together
  // again (every row is computed by exactly one rank, so all ranks keep identical data)
  const bool split_rows = dist && R % dist->world == 0 && getenv("VVT_SYEVJ_DIST_REPLICATED_GEMM") == nullptr;
  auto gemm_nt = [&](T* C, const T* A, const T* Bm, double beta) -> int {
    if (!split_rows) return gemm_batched(C, A, Bm, R, R, R, R, R, R, 1.0, beta);
    const int64_t rows = R / dist->world, r0 = rows * dist->rank;
    VVT_TRY(vvt_gemm(C + r0 * R, A + r0 * R, Bm, rows, R, R, 0, 0, R, R, R, 1.0, beta, 1, 0, 0, 0, ws + L.off_gemm,
                     L.gemm_bytes, dtype, (void*)s));
    return check_nccl(nccl().all_gather(C + r0 * R, C, size_t(rows * R), sizeof(T) == 4 ? kNcclFloat32 : kNcclFloat64,
                                        dist->comm, s),
                      "vvt_syevj_dist(all-gather)");
  };
  VVT_TRY(gemm_nt(Tt, Jt, Gs, 0.0));  // Tt = J^T G   (G symmetric)
  onesided_rayleigh_kernel<T><<<dim3(unsigned(ceil_div(R * 32, 256)), UB), 256, 0, s>>>(ev, Jt, Tt, R, sc);
  VVT_TRY(launched("vvt_syevj(rayleigh)"));
  if (jobz) {
    VVT_TRY(gemm_nt(Sm, Jt, Tt, 0.0));  // S = J^T G J
    VVT_TRY(gemm_nt(Mm, Jt, Jt, 0.0));  // M = J^T J
    onesided_refine_coeff_kernel<T><<<dim3(gblocks, UB), 256, 0, s>>>(Et, Sm, Mm, ev, R, sc);
    VVT_TRY(launched("vvt_syevj(refine)"));
    VVT_TRY(gemm_nt(Jt, Et, Jm, 1.0));  // J^T += E^T J^T
  }
  jacobi_rank_kernel<T><<<dim3(unsigned(ceil_div(R, 128)), UB), 128, 0, s>>>(rank, ev, R);
  VVT_TRY(launched("vvt_syevj(rank)"));
  const int64_t total = jobz ? RR : R;
  const int pblocks = int(vmin<int64_t>(ceil_div(total, 256), 8 * num_sms()));
  jacobi_permute_kernel<T><<<dim3(pblocks, UB), 256, 0, s>>>(evals, jobz ? evecs : nullptr, ev, Jt, rank, R);
  VVT_TRY(launched("vvt_syevj(permute)"));
  phase("gather, Rayleigh quotients, refinement, sort");
  return VVT_OK;
}

}  // namespace vvt

using namespace vvt;

extern "C" {

int64_t vvt_syevj_batched_workspace_bytes(int64_t R, int64_t batch, int jobz, int dtype) {
  (void)jobz;
  if (R <= 0 || batch <= 0) return 0;
  return jacobi_layout(R, vmin<int64_t>(batch, kMaxBatch), dtype == VVT_F32 ? 4 : 8, dtype).total;
}

int64_t vvt_syevj_workspace_bytes(int64_t R, int jobz, int dtype) {
  return vvt_syevj_batched_workspace_bytes(R, 1, jobz, dtype);
}

int vvt_syevj_batched(void* evals, void* evecs, const void* G, int64_t R, int64_t batch, int jobz, void* workspace,
                      int64_t workspace_bytes, int* info_host, int dtype, void* stream) {
  VVT_REQUIRE(R >= 0 && batch >= 0, "negative size");
  if (info_host)
    for (int64_t b = 0; b < batch; ++b) info_host[2 * b] = 0, info_host[2 * b + 1] = 1;
  if (R == 0 || batch == 0) return VVT_OK;
  VVT_REQUIRE(R <= (int64_t(1) << 15), "matrix too large");
  VVT_REQUIRE(evals && G && workspace && (!jobz || evecs), "null pointer");
  if (workspace_bytes < vvt_syevj_batched_workspace_bytes(R, batch, jobz, dtype))
    return fail(VVT_ERR_WORKSPACE, "%s: workspace too small", __func__);
  VVT_DISPATCH(dtype, {
    for (int64_t b0 = 0; b0 < batch; b0 += kMaxBatch) {
      const int64_t nb = vmin<int64_t>(kMaxBatch, batch - b0);
      VVT_TRY(syevj_impl<T>((T*)evals + b0 * R, jobz ? (T*)evecs + b0 * R * R : nullptr, (const T*)G + b0 * R * R, R, nb,
                            jobz, (char*)workspace, info_host ? info_host + 2 * b0 : nullptr, dtype, as_stream(stream)));
    }
    return VVT_OK;
  });
}

int vvt_syevj(void* evals, void* evecs, const void* G, int64_t R, int jobz, void* workspace,
              int64_t workspace_bytes, int* info_host, int dtype, void* stream) {
  return vvt_syevj_batched(evals, evecs, G, R, 1, jobz, workspace, workspace_bytes, info_host, dtype, stream);
}

int64_t vvt_dist_arena_bytes_for(int64_t R) {
  if (R <= 0) return 0;
  const int64_t Np = align_up(R, wide::WP);
  return Np * Np * 4;
}

int64_t vvt_dist_arena_bytes(void) { return arena().base ? arena().bytes : 0; }

int vvt_dist_arena_alloc(int64_t bytes, void* handle_out) {
  VVT_REQUIRE(bytes > 0 && handle_out, "bad arguments");
  DistArena& ar = arena();
  VVT_REQUIRE(ar.base == nullptr, "an arena exists already (vvt_dist_arena_free first)");
  void* ptr = nullptr;
  VVT_TRY(check_cuda(cudaMalloc(&ptr, size_t(bytes) + kArenaHeader), __func__));
  cudaIpcMemHandle_t h;
  if (cudaMemset(ptr, 0, kArenaHeader) != cudaSuccess || cudaIpcGetMemHandle(&h, ptr) != cudaSuccess) {
    const cudaError_t e = cudaGetLastError();
    cudaFree(ptr);
    return fail(VVT_ERR_CUDA, "%s: %s", __func__, cudaGetErrorString(e));
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  memcpy(handle_out, &h, sizeof(h));
  ar = DistArena();
  ar.base = static_cast<char*>(ptr);
  ar.bytes = bytes;
  return VVT_OK;
}

int vvt_dist_arena_open(const void* handles, int world, int rank) {
  VVT_REQUIRE(handles && world > 1 && world <= kMaxPeers && rank >= 0 && rank < world, "bad arguments");
  DistArena& ar = arena();
  VVT_REQUIRE(ar.base != nullptr && ar.world == 0, "vvt_dist_arena_alloc first (once)");
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      ar.peer[r] = ar.base;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles) + size_t(r) * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    VVT_TRY(check_cuda(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess), __func__));
    ar.peer[r] = static_cast<char*>(ptr);
  }
  ar.world = world;
  ar.rank = rank;
  ar.seq = 0;
  return VVT_OK;
}

int vvt_dist_arena_close_peers(void) {
  DistArena& ar = arena();
  for (int r = 0; r < ar.world; ++r)
    if (r != ar.rank && ar.peer[r]) cudaIpcCloseMemHandle(ar.peer[r]);
  for (int r = 0; r < kMaxPeers; ++r) ar.peer[r] = nullptr;
  ar.world = 0;
  ar.rank = -1;
  return VVT_OK;
}

int vvt_dist_arena_free(void) {
  DistArena& ar = arena();
  VVT_REQUIRE(ar.world == 0, "vvt_dist_arena_close_peers first (on every rank, then a barrier)");
  if (ar.base) cudaFree(ar.base);
  ar = DistArena();
  return VVT_OK;
}

int vvt_dbg_dist_plan(int nbw, int world, int round, int* owner, int* moves_out, int max_moves, int* n_moves_out,
                      int* pair_rank_out) {
  VVT_REQUIRE(nbw >= 2 && nbw % 2 == 0 && world >= 1 && round >= -1 && round < nbw - 1, "bad arguments");
  VVT_REQUIRE(owner && n_moves_out && (moves_out || max_moves == 0), "null pointer");
  std::vector<int> own(owner, owner + nbw);
  std::vector<BlockMove> moves;
  dist_plan_round(nbw, world, round, own, moves);
  for (int w = 0; w < nbw; ++w) owner[w] = own[size_t(w)];
  *n_moves_out = int(moves.size());
  for (int i = 0; i < int(moves.size()) && i < max_moves; ++i)
    moves_out[3 * i] = moves[size_t(i)].block, moves_out[3 * i + 1] = moves[size_t(i)].src,
                  moves_out[3 * i + 2] = moves[size_t(i)].dst;
  if (pair_rank_out)
    for (int pair = 0; pair < nbw / 2; ++pair) pair_rank_out[pair] = dist_rank_of_pair(nbw / 2, world, pair);
  return VVT_OK;
}

int64_t vvt_syevj_dist_workspace_bytes(int64_t R, int jobz, int dtype, int world) {
  (void)jobz;
  if (R <= 0) return 0;
  return jacobi_layout(R, 1, dtype == VVT_F32 ? 4 : 8, dtype, vmax(1, world)).total;
}

int vvt_syevj_dist(void* comm, void* evals, void* evecs, const void* G, int64_t R, int jobz, void* workspace,
                   int64_t workspace_bytes, int* info_host, int p2p, int dtype, void* stream) {
  VVT_REQUIRE(R >= 0, "negative size");
  VVT_REQUIRE(comm != nullptr, "null communicator");
  if (info_host) info_host[0] = 0, info_host[1] = 1;
  if (R == 0) return VVT_OK;
  VVT_REQUIRE(R <= (int64_t(1) << 15), "matrix too large");
  VVT_REQUIRE(evals && G && workspace && (!jobz || evecs), "null pointer");
  const Nccl& n = nccl();
  if (!n.ok) return fail(VVT_ERR_UNSUPPORTED, "%s: no NCCL library is loaded in this process", __func__);
  DistCtx d{comm, 0, 1, p2p != 0};
  VVT_TRY(check_nccl(n.comm_count(comm, &d.world), __func__));
  VVT_TRY(check_nccl(n.comm_rank(comm, &d.rank), __func__));
  if (workspace_bytes < vvt_syevj_dist_workspace_bytes(R, jobz, dtype, d.world))
    return fail(VVT_ERR_WORKSPACE, "%s: workspace too small", __func__);
  VVT_DISPATCH(dtype, {
    return syevj_impl<T>((T*)evals, jobz ? (T*)evecs : nullptr, (const T*)G, R, 1, jobz, (char*)workspace, info_host, dtype,
                         as_stream(stream), &d);
  });
}

// Test hook: ONE wide round of the two-level solver on a caller-provided
// row-major factor L [Np, Np] (Np a multiple of 128), so that tests can check the three kernels one by one:
// H_out [pairs, 128, 128] = the pair Grams (split-K partials summed), Qt_out [pairs, 128, 128] = Q^T of every
// pair, flag_out [pairs], and L updated in place.  Allocates its own scratch (debug only).
int vvt_dbg_wide_round(float* L, float* H_out, float* Qt_out, int* flag_out, int64_t Np, int round, void* stream) {
  VVT_REQUIRE(L && H_out && Qt_out && flag_out && Np > 0 && Np % wide::WP == 0, "bad arguments");
  cudaStream_t s = as_stream(stream);
  const wide::WidePlan p = wide::wide_plan(Np, 1);
  char* scratch = nullptr;
  const size_t l_bytes = size_t(Np) * Np * 4;
  VVT_TRY(check_cuda(cudaMalloc(&scratch, size_t(p.part_bytes + 256 + p.diag_bytes) + l_bytes), __func__));
  float* part = (float*)scratch;
  JacobiScalars* sc = (JacobiScalars*)(scratch + p.part_bytes);
  float* diag = (float*)(scratch + p.part_bytes + 256);
  float* Lw = (float*)(scratch + p.part_bytes + 256 + p.diag_bytes);  // the solver's wide-block-major layout
  const int blocks = int(vmin<int64_t>(ceil_div(int64_t(Np) * Np, 256), 8 * num_sms()));
  int st = check_cuda(cudaMemsetAsync(sc, 0, sizeof(JacobiScalars), s), __func__);
  // (a cross round writes only the columns of b of every pair Gram: the rest of H_out reads as zero)
  if (st == VVT_OK) st = check_cuda(cudaMemsetAsync(part, 0, size_t(p.part_bytes), s), __func__);
  wide::wide_relayout_kernel<<<blocks, 256, 0, s>>>(Lw, L, int(Np), 1);
  // the diagonal-block cache a cross round expects (in the solver the intra round of the sweep has filled it)
  wide::wide_diag_kernel<<<p.nbw, 256, 0, s>>>(diag, Lw, int(Np));
  wide::WideMaps maps;
  if (st == VVT_OK) st = wide::wide_make_maps(&maps, Lw, Qt_out, p, 1);
  if (st == VVT_OK) st = wide::wide_round(Lw, part, Qt_out, flag_out, diag, sc, p, maps, round, 1, s);
  if (st == VVT_OK) {
    for (int k = 0; k < p.splits && st == VVT_OK; ++k)
      st = vvt_axpy(H_out, part + size_t(k) * p.pairs * wide::WP * wide::WP, int64_t(p.pairs) * wide::WP * wide::WP, 1.0,
                    VVT_F32, stream);
  }
  wide::wide_relayout_kernel<<<blocks, 256, 0, s>>>(L, Lw, int(Np), 0);
  cudaStreamSynchronize(s);
  cudaFree(scratch);
  return st;
}

}  // extern "C"
