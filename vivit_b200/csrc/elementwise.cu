// Bandwidth-bound kernels: loss-Hessian factors, element-wise / pooling Jacobians,
// bias emit, dense back-transform, Newton step, filters.  All grid-stride, coalesced
// along the innermost (feature / parameter) dimension.
#include "common.cuh"

namespace vvt {

thread_local char g_err[512] = {0};
char* last_error_buffer() { return g_err; }
std::atomic<int64_t> g_launches{0};

static inline int ew_blocks(int64_t total, int per_thread = 1) {
  return int(vmax<int64_t>(1, vmin<int64_t>(ceil_div(total, 256 * per_thread), 16 * num_sms())));
}

#define GRID_STRIDE(i, total)                                                      \
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < (total); \
       i += int64_t(gridDim.x) * blockDim.x)

// ---- loss factors -----------------------------------------------------------
// one warp per sample: softmax in registers (strided over C), then write C x C block
template <typename T, bool MC>
__global__ void ce_factor_kernel(T* S, const T* logits, const int64_t* sub, const int64_t* cls,
                                 int64_t n_sub, int64_t C, int64_t M, T scale) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t n = warp; n < n_sub; n += nwarps) {
    const T* row = logits + (sub ? sub[n] : n) * C;
    T mx = -INFINITY;
    for (int64_t c = lane; c < C; c += 32) mx = max(mx, row[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    T sum = 0;
    for (int64_t c = lane; c < C; c += 32) sum += exp(row[c] - mx);
    sum = warp_sum(sum);
    const T inv = T(1) / sum;
    if (MC) {
      // S[m, n, c] = (p_c - [cls[m,n] == c]) * scale
      for (int64_t m = 0; m < M; ++m) {
        const int64_t y = cls[m * n_sub + n];
        for (int64_t c = lane; c < C; c += 32) {
          const T p = exp(row[c] - mx) * inv;
          S[(m * n_sub + n) * C + c] = (p - (c == y ? T(1) : T(0))) * scale;
        }
      }
    } else {
      // S[v, n, c] = tau_c (delta_vc - tau_v tau_c) * scale
      for (int64_t v = 0; v < C; ++v) {
        const T tv = sqrt(exp(row[v] - mx) * inv);
        for (int64_t c = lane; c < C; c += 32) {
          const T tc = sqrt(exp(row[c] - mx) * inv);
          S[(v * n_sub + n) * C + c] = tc * ((c == v ? T(1) : T(0)) - tv * tc) * scale;
        }
      }
    }
  }
}

template <typename T>
__global__ void mse_factor_kernel(T* S, int64_t n_sub, int64_t C, T scale) {
  const int64_t total = C * n_sub * C;
  GRID_STRIDE(i, total) {
    const int64_t c = i % C, v = i / (C * n_sub);
    S[i] = (c == v) ? scale : T(0);
  }
}

template <typename T>
__global__ void scale_kernel(T* t, int64_t n, T alpha) {
  GRID_STRIDE(i, n) t[i] *= alpha;
}

template <typename T>
__global__ void axpy_kernel(T* y, const T* x, int64_t n, T alpha) {
  GRID_STRIDE(i, n) y[i] += alpha * x[i];
}

// ---- element-wise Jacobians -------------------------------------------------
template <typename T>
__device__ __forceinline__ T act_derivative(T r, int act, T scale) {
  switch (act) {
    case VVT_ACT_RELU: return r > T(0) ? T(1) : T(0);
    case VVT_ACT_SIGMOID: return r * (T(1) - r);
    case VVT_ACT_TANH: return T(1) - r * r;
    case VVT_ACT_DROPOUT: return r != T(0) ? scale : T(0);
    case VVT_ACT_LEAKY_RELU: return r > T(0) ? T(1) : scale;
    case VVT_ACT_ELU: return r > T(0) ? T(1) : scale * exp(r);
    case VVT_ACT_SELU:
      return T(1.0507009873554804934193349852946) * (r > T(0) ? T(1) : T(1.6732632423543772848170429916717) * exp(r));
    case VVT_ACT_LOGSIGMOID: return T(1) / (T(1) + exp(r));
    default: return r;
  }
}

template <typename T>
__global__ void act_kernel(T* out, const T* S, const T* ref, int64_t V, int64_t nf, int act, T scale) {
  const int64_t total = V * nf;
  GRID_STRIDE(i, total) out[i] = S[i] * act_derivative(ldg(ref + i % nf), act, scale);
}

// Streaming form (16-byte aligned rows): a thread owns one 16-byte group of features, evaluates the derivative
// ONCE and streams the V rows of the factor through it, UNROLL rows in flight: `ref` is read once instead of V
// times, no division per element, 16-byte accesses.
template <typename T, int UNROLL>
__global__ void __launch_bounds__(256) act_rows_kernel(T* out, const T* S, const T* ref, int64_t V, int64_t nf, int act,
                                                       T scale) {
  constexpr int VW = 16 / int(sizeof(T));
  const int64_t groups = nf / VW;
  for (int64_t gidx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; gidx < groups;
       gidx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t f = gidx * VW;
    T d[VW];
    {
      const float4 r4 = __ldg(reinterpret_cast<const float4*>(ref + f));
      const T* r = reinterpret_cast<const T*>(&r4);
#pragma unroll
      for (int e = 0; e < VW; ++e) d[e] = act_derivative(r[e], act, scale);
    }
    const int64_t v_lo = blockIdx.y * int64_t(UNROLL);
    for (int64_t v0 = v_lo; v0 < V; v0 += int64_t(gridDim.y) * UNROLL) {
      float4 x[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (v0 + u < V) x[u] = __ldcs(reinterpret_cast<const float4*>(S + (v0 + u) * nf + f));
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (v0 + u < V) {
          T* xe = reinterpret_cast<T*>(&x[u]);
#pragma unroll
          for (int e = 0; e < VW; ++e) xe[e] *= d[e];
          *reinterpret_cast<float4*>(out + (v0 + u) * nf + f) = x[u];
        }
    }
  }
}

// Arg-max positions of a max-pool forward pass, recomputed from the layer input for the factor back-propagation
// ([BackPACK] re-runs the pooling with return_indices=True for its Jacobian; torch's rule is restated here so the
// same input position is chosen among tied values -- ReLU zeros tie all the time): windows are scanned row by row,
// a value replaces the running maximum when it is GREATER (the first of equal values wins) or NaN; padding is
// skipped, the index is flat into the h_in x w_in plane.  One thread per output position.
template <typename T>
__global__ void __launch_bounds__(256)
maxpool_argmax_kernel(int64_t* idx, const T* x, int64_t total, int ho, int wo, int hi, int wi, int kh, int kw, int sh,
                      int sw, int ph, int pw, int dh, int dw) {
  const int hw_out = ho * wo;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t plane = i / hw_out;
    const int q = int(i - plane * hw_out), oy = q / wo, ox = q - oy * wo;
    int y0 = oy * sh - ph, x0 = ox * sw - pw;
    const int y1 = min(y0 + (kh - 1) * dh + 1, hi), x1 = min(x0 + (kw - 1) * dw + 1, wi);
    while (y0 < 0) y0 += dh;
    while (x0 < 0) x0 += dw;
    const T* xp = x + plane * int64_t(hi) * wi;
    int best = y0 * wi + x0;
    T vmax_ = -INFINITY;
    for (int y = y0; y < y1; y += dh)
      for (int xx = x0; xx < x1; xx += dw) {
        const T v = xp[y * wi + xx];
        if (v > vmax_ || v != v) vmax_ = v, best = y * wi + xx;
      }
    idx[i] = best;
  }
}

// gather form of the max-pool scatter: every input position sums the outputs that chose it.
// One block per (row, channel) plane (grid-stride): the plane decomposition is done once per block and
// the strides are compile-time constants for the common cases (STRIDE = 1, 2; 0 = run-time), so the
// inner loop has no 64-bit and no run-time divisions.
template <typename T, int STRIDE>
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(T* out, const T* S, const int64_t* argmax, int64_t planes, int64_t N, int64_t ch, int ho,
                   int wo, int hi, int wi, int kh, int kw, int sh_rt, int sw_rt, int ph, int pw, int dh, int dw) {
  const int sh = STRIDE ? STRIDE : sh_rt, sw = STRIDE ? STRIDE : sw_rt;
  const int hw_in = hi * wi, hw_out = ho * wo;
  // blockDim = (positions, planes per block): small maps share a block
  for (int64_t pl = int64_t(blockIdx.x) * blockDim.y + threadIdx.y; pl < planes; pl += int64_t(gridDim.x) * blockDim.y) {
    const int64_t r = pl / ch, c = pl - r * ch, n = r % N;
    const T* s = S + pl * hw_out;
    const int64_t* am = argmax + (n * ch + c) * hw_out;
    T* o = out + pl * hw_in;
    for (int pos = threadIdx.x; pos < hw_in; pos += blockDim.x) {
      const int y = pos / wi, x = pos - y * wi;
      // candidate windows: oy*sh <= y+ph <= oy*sh + (kh-1)*dh
      const int ay = y + ph, ax = x + pw;
      const int oy_hi = min(ho - 1, ay / sh), ox_hi = min(wo - 1, ax / sw);
      const int by = ay - (kh - 1) * dh, bx = ax - (kw - 1) * dw;
      const int oy_lo = by > 0 ? (by + sh - 1) / sh : 0, ox_lo = bx > 0 ? (bx + sw - 1) / sw : 0;
      T acc = 0;
      for (int oy = oy_lo; oy <= oy_hi; ++oy) {
        if (dh > 1 && (ay - oy * sh) % dh) continue;
        for (int ox = ox_lo; ox <= ox_hi; ++ox) {
          if (dw > 1 && (ax - ox * sw) % dw) continue;
          if (int(am[oy * wo + ox]) == pos) acc += s[oy * wo + ox];
        }
      }
      o[pos] = acc;
    }
  }
}

// Fast path: the argmax map is shared by the V rows of a sample (S is [V, N, ch, ho, wo]), so one block takes one
// (n, c) plane, resolves for every input position which of its (at most 2 x 2) candidate windows chose it ONCE,
// and then streams the V output planes.  The pooled planes of a chunk of rows are brought to shared memory with
// coalesced loads (every pooled value is wanted by exactly one input position, so reading them from global
// memory position by position fetched a 32-byte sector per 4-byte value); the scatter then reads shared memory
// and the stores are coalesced along the position.  Windows with ceil(k / stride) <= 2, no dilation, input
// planes of at most PMAX * 256 positions.
constexpr int kPoolRows = 8;  // rows of the factor staged per pass
template <typename T, int PMAX>
__global__ void __launch_bounds__(256)
maxpool_bwd_shared_kernel(T* out, const T* S, const int64_t* argmax, int64_t V, int64_t N, int64_t ch, int ho, int wo,
                          int hi, int wi, int kh, int kw, int sh, int sw, int ph, int pw) {
  extern __shared__ __align__(16) unsigned char pool_smem[];
  T* sS = reinterpret_cast<T*>(pool_smem);  // [kPoolRows][hw_out]
  const int hw_in = hi * wi, hw_out = ho * wo;
  const int64_t planes = N * ch;
  for (int64_t pl = blockIdx.x; pl < planes; pl += gridDim.x) {
    const int64_t* am = argmax + pl * hw_out;
    int off[PMAX][4];
#pragma unroll
    for (int p = 0; p < PMAX; ++p) {
      const int pos = threadIdx.x + p * 256;
#pragma unroll
      for (int j = 0; j < 4; ++j) off[p][j] = -1;
      if (pos < hw_in) {
        const int y = pos / wi, x = pos - y * wi;
        const int ay = y + ph, ax = x + pw;
        const int oy_hi = min(ho - 1, ay / sh), ox_hi = min(wo - 1, ax / sw);
        const int by = ay - (kh - 1), bx = ax - (kw - 1);
        const int oy_lo = by > 0 ? (by + sh - 1) / sh : 0, ox_lo = bx > 0 ? (bx + sw - 1) / sw : 0;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const int oy = oy_lo + dy, ox = ox_lo + dx;
            const bool hit = oy <= oy_hi && ox <= ox_hi && int(am[oy * wo + ox]) == pos;
            off[p][2 * dy + dx] = hit ? oy * wo + ox : -1;
          }
      }
    }
    const int64_t s_step = planes * hw_out, o_step = planes * hw_in;
    for (int64_t v0 = 0; v0 < V; v0 += kPoolRows) {
      const int vn = int(vmin<int64_t>(kPoolRows, V - v0));
      __syncthreads();  // the previous chunk has been consumed
      for (int idx = threadIdx.x; idx < vn * hw_out; idx += 256) {
        const int u = idx / hw_out, i = idx - u * hw_out;
        sS[u * hw_out + i] = __ldcs(S + (v0 + u) * s_step + pl * hw_out + i);
      }
      __syncthreads();
#pragma unroll
      for (int p = 0; p < PMAX; ++p) {
        const int pos = threadIdx.x + p * 256;
        if (pos < hw_in) {
          T* o = out + (v0 * planes + pl) * int64_t(hw_in) + pos;
          for (int u = 0; u < vn; ++u) {
            T acc = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (off[p][j] >= 0) acc += sS[u * hw_out + off[p][j]];
            o[u * o_step] = acc;
          }
        }
      }
    }
  }
}

template <typename T>
static void launch_maxpool_bwd(T* out, const T* S, const int64_t* argmax, int64_t planes, int64_t N, int64_t ch, int ho,
                               int wo, int hi, int wi, int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw,
                               cudaStream_t s) {
  const size_t pool_smem = size_t(kPoolRows) * ho * wo * sizeof(T);
  if (dh == 1 && dw == 1 && kh <= 2 * sh && kw <= 2 * sw && N * ch > 0 && hi * wi <= 4 * 256 && pool_smem <= 48 * 1024) {
    const int64_t V = planes / (N * ch);
    const unsigned blocks = unsigned(vmin<int64_t>(N * ch, int64_t(16) * num_sms()));
    if (hi * wi <= 256)
      maxpool_bwd_shared_kernel<T, 1><<<blocks, 256, pool_smem, s>>>(out, S, argmax, V, N, ch, ho, wo, hi, wi, kh, kw, sh, sw,
                                                                     ph, pw);
    else if (hi * wi <= 512)
      maxpool_bwd_shared_kernel<T, 2><<<blocks, 256, pool_smem, s>>>(out, S, argmax, V, N, ch, ho, wo, hi, wi, kh, kw, sh, sw,
                                                                     ph, pw);
    else
      maxpool_bwd_shared_kernel<T, 4><<<blocks, 256, pool_smem, s>>>(out, S, argmax, V, N, ch, ho, wo, hi, wi, kh, kw, sh, sw,
                                                                     ph, pw);
    return;
  }
  const int tx = int(vmin<int64_t>(256, align_up(int64_t(hi) * wi, 32))), ty = 256 / tx;
  const dim3 threads(tx, ty);
  const unsigned blocks = unsigned(vmin<int64_t>(ceil_div(planes, ty), int64_t(32) * num_sms()));
  if (sh == sw && sh == 2)
    maxpool_bwd_kernel<T, 2><<<blocks, threads, 0, s>>>(out, S, argmax, planes, N, ch, ho, wo, hi, wi, kh, kw, sh, sw, ph,
                                                       pw, dh, dw);
  else if (sh == sw && sh == 1)
    maxpool_bwd_kernel<T, 1><<<blocks, threads, 0, s>>>(out, S, argmax, planes, N, ch, ho, wo, hi, wi, kh, kw, sh, sw, ph,
                                                       pw, dh, dw);
  else
    maxpool_bwd_kernel<T, 0><<<blocks, threads, 0, s>>>(out, S, argmax, planes, N, ch, ho, wo, hi, wi, kh, kw, sh, sw, ph,
                                                       pw, dh, dw);
}

template <typename T>
__global__ void avgpool_bwd_kernel(T* out, const T* S, int64_t rows, int64_t ch, int ho, int wo, int hi,
                                   int wi, int kh, int kw, int sh, int sw, int ph, int pw) {
  const int64_t total = rows * ch * hi * wi;
  const T inv = T(1) / T(kh * kw);
  GRID_STRIDE(i, total) {
    const int x = int(i % wi), y = int((i / wi) % hi);
    const int64_t rc = i / (int64_t(wi) * hi);
    const T* s = S + rc * int64_t(ho) * wo;
    T acc = 0;
    for (int ky = 0; ky < kh; ++ky) {
      const int ty = y + ph - ky;
      if (ty < 0 || ty % sh) continue;
      const int oy = ty / sh;
      if (oy >= ho) continue;
      for (int kx = 0; kx < kw; ++kx) {
        const int tx = x + pw - kx;
        if (tx < 0 || tx % sw) continue;
        const int ox = tx / sw;
        if (ox >= wo) continue;
        acc += s[oy * wo + ox];
      }
    }
    out[i] = acc * inv;
  }
}

// ---- V emit ---------------------------------------------------------------
// Vt[r, o] = sum_x S[r, o, x]: one warp per (r, o)
template <typename T>
__global__ void bias_emit_kernel(T* Vt, const T* S, int64_t ro, int64_t spatial) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t w = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5; w < ro; w += nwarps) {
    const T* s = S + w * spatial;
    T acc = 0;
    for (int64_t x = lane; x < spatial; x += 32) acc += s[x];
    acc = warp_sum(acc);
    if (lane == 0) Vt[w] = acc;
  }
}

template <typename T>
__global__ void linear_emit_kernel(T* Vt, const T* S, const T* Z, int64_t V, int64_t N, int64_t n_out,
                                   int64_t n_in) {
  const int64_t total = V * N * n_out * n_in;
  GRID_STRIDE(i, total) {
    const int64_t ii = i % n_in, o = (i / n_in) % n_out, rn = i / (n_in * n_out);
    Vt[i] = ldg(S + rn * n_out + o) * ldg(Z + (rn % N) * n_in + ii);
  }
}

// ---- dense back-transform: E[K, D] = U[K, R] V[R, D] -------------------------
// Streams V exactly once per KT directions.  A CTA owns 32 x VW consecutive columns (one 16-byte load per
// row and lane, coalesced across the warp); its 8 warps split the rows (warp w takes rows w*8 .. w*8+7 of
// every 64-row chunk, all 8 loads in flight), so that narrow factors (biases, small kernels) still spread
// their rows over 8 warps and wide ones keep 2048 threads x 128 bytes in flight per SM.  U is staged
// through shared memory in 64-row chunks (double-buffered: one barrier per chunk); the 8 partial sums per
// column are added in a fixed order through shared memory (deterministic), squared norms fused.
template <typename T, int KT, int VW>
__global__ void __launch_bounds__(256)
backtransform_dense_kernel(T* E, double* norm2, const T* U, const T* V, int64_t K, int64_t R, int64_t D,
                           int64_t k0) {
  constexpr int RC = 64, RB = 8, NW = 8;  // rows per U chunk, rows in flight per thread, warps
  __shared__ __align__(16) T Us[2][RC][KT];
  __shared__ __align__(16) T red[NW][32][VW];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t d0 = (blockIdx.x * int64_t(32) + lane) * VW;
  const int kn = int(vmin<int64_t>(KT, K - k0));
  const bool vec = (D % VW == 0) && d0 + VW <= D && (reinterpret_cast<uintptr_t>(V) & 15) == 0;  // aligned full vector
  T acc[KT][VW];
#pragma unroll
  for (int k = 0; k < KT; ++k)
#pragma unroll
    for (int v = 0; v < VW; ++v) acc[k][v] = 0;
  auto stage_u = [&](int buf, int64_t r0) {
    const int rn = int(vmin<int64_t>(RC, R - r0));
    for (int i = threadIdx.x; i < KT * RC; i += 256) {
      const int k = i / RC, r = i % RC;
      Us[buf][r][k] = (k < kn && r < rn) ? U[(k0 + k) * R + r0 + r] : T(0);
    }
  };
  stage_u(0, 0);
  int buf = 0;
  for (int64_t r0 = 0; r0 < R; r0 += RC, buf ^= 1) {
    __syncthreads();  // chunk `buf` is staged; everybody is done with the other buffer
    if (r0 + RC < R) stage_u(buf ^ 1, r0 + RC);
    const int rn = int(vmin<int64_t>(RC, R - r0));
    const int rb = warp * RB;
    if (d0 < D && rb < rn) {
      const T* vp = V + (r0 + rb) * D + d0;
      T x[RB][VW];
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        if (rb + i < rn) {
          if (vec) {
            if constexpr (sizeof(T) * VW == 16) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(vp + int64_t(i) * D));
              const T* tp = reinterpret_cast<const T*>(&t);
#pragma unroll
              for (int v = 0; v < VW; ++v) x[i][v] = tp[v];
            } else {
#pragma unroll
              for (int v = 0; v < VW; ++v) x[i][v] = ldg(vp + int64_t(i) * D + v);
            }
          } else {
#pragma unroll
            for (int v = 0; v < VW; ++v) x[i][v] = d0 + v < D ? ldg(vp + int64_t(i) * D + v) : T(0);
          }
        } else {
#pragma unroll
          for (int v = 0; v < VW; ++v) x[i][v] = T(0);
        }
      }
#pragma unroll
      for (int i = 0; i < RB; ++i) {  // rows past rn are staged as zero
        if constexpr (KT % 4 == 0 && sizeof(T) == 4) {  // 16-byte broadcast reads of the U row
#pragma unroll
          for (int k4 = 0; k4 < KT; k4 += 4) {
            const float4 u4 = *reinterpret_cast<const float4*>(&Us[buf][rb + i][k4]);
            const T u[4] = {T(u4.x), T(u4.y), T(u4.z), T(u4.w)};
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
              for (int v = 0; v < VW; ++v) acc[k4 + e][v] += u[e] * x[i][v];
          }
        } else {
#pragma unroll
          for (int k = 0; k < KT; ++k) {
            const T u = Us[buf][rb + i][k];
#pragma unroll
            for (int v = 0; v < VW; ++v) acc[k][v] += u * x[i][v];
          }
        }
      }
    }
  }
  // cross-warp sum, direction by direction: warp (k mod 8) finishes direction k
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    __syncthreads();
#pragma unroll
    for (int v = 0; v < VW; ++v) red[warp][lane][v] = acc[k][v];
    __syncthreads();
    if (warp == (k & (NW - 1)) && k < kn) {
      T sum[VW];
#pragma unroll
      for (int v = 0; v < VW; ++v) sum[v] = red[0][lane][v];
#pragma unroll
      for (int w = 1; w < NW; ++w)
#pragma unroll
        for (int v = 0; v < VW; ++v) sum[v] += red[w][lane][v];
      double sq = 0.0;
#pragma unroll
      for (int v = 0; v < VW; ++v)
        if (d0 + v < D) {
          E[(k0 + k) * D + d0 + v] = sum[v];
          sq += double(sum[v]) * double(sum[v]);
        }
      if (norm2) {
        sq = warp_sum(sq);
        if (lane == 0) atomicAdd(norm2 + k0 + k, sq);
      }
    }
  }
}

template <typename T>
__global__ void scale_rows_rsqrt_kernel(T* E, const double* norm2, int64_t K, int64_t D) {
  const int64_t total = K * D;
  GRID_STRIDE(i, total) { E[i] = T(double(E[i]) / sqrt(norm2[i / D])); }
}

// out[n, d] = g[n, d] - mean_m g[m, d]  (CenteredBatchGrad / CenteredGramBatchGrad prologue).  A block owns a strip
// of 32 * VEC columns: its 8 warps split the rows for the column sums (16-byte loads when VEC > 1), the partial sums
// are added in a fixed order (deterministic), and every warp subtracts the mean from the rows it has just read, so
// the second read comes from L2 and HBM sees one read and one write of the matrix.  out may alias g.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) center_rows_kernel(T* out, const T* g, int64_t N, int64_t D) {
  __shared__ T part[8][32 * VEC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t strips = (D + 32 * VEC - 1) / (32 * VEC);
  for (int64_t strip = blockIdx.x; strip < strips; strip += gridDim.x) {
    const int64_t d0 = strip * (32 * VEC) + int64_t(lane) * VEC;
    const bool live = d0 < D;  // D % VEC == 0 whenever VEC > 1
    T acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = T(0);
    if (live) {
      for (int64_t n = warp; n < N; n += 8) {
        T x[VEC];
        if constexpr (VEC > 1) {
          *reinterpret_cast<int4*>(x) = *reinterpret_cast<const int4*>(g + n * D + d0);
        } else {
          x[0] = g[n * D + d0];
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] += x[v];
      }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) part[warp][lane * VEC + v] = acc[v];
    __syncthreads();
    T mean[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      T s = T(0);
#pragma unroll
      for (int w = 0; w < 8; ++w) s += part[w][lane * VEC + v];
      mean[v] = s / T(N);
    }
    if (live) {
      for (int64_t n = warp; n < N; n += 8) {
        T x[VEC];
        if constexpr (VEC > 1) {
          *reinterpret_cast<int4*>(x) = *reinterpret_cast<const int4*>(g + n * D + d0);
        } else {
          x[0] = g[n * D + d0];
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) x[v] -= mean[v];
        if constexpr (VEC > 1) {
          *reinterpret_cast<int4*>(out + n * D + d0) = *reinterpret_cast<const int4*>(x);
        } else {
          out[n * D + d0] = x[0];
        }
      }
    }
    __syncthreads();  // part is reused by the next strip
  }
}

template <typename T>
__global__ void filter_nonzero_kernel(uint8_t* mask, const T* ev, int64_t R, T atol, T rtol,
                                      unsigned long long* count) {
  GRID_STRIDE(i, R) {
    // torch.isclose(ev, 0): |ev - 0| <= atol + rtol * |0|
    const bool keep = !(fabs(ev[i]) <= atol + rtol * T(0));
    mask[i] = keep;
    if (keep && count) atomicAdd(count, 1ull);
  }
}

// coef[k] = -mean(gammas[:,k]) / (mean(lambdas[:,k]) + deltas[k]) / sqrt(evals[k]);  v = corr * U coef
// One row of U per thread; the coefficients pass through shared memory in tiles of kCoefTile, so that any
// number of directions fits.
constexpr int kCoefTile = 2048;
template <typename T>
__global__ void newton_coeff_kernel(T* v, const T* U, const T* gam, const T* lam, const T* del,
                                    const T* ev, int64_t R, int64_t K, int64_t n_g, int64_t n_l, T corr) {
  __shared__ T coef[kCoefTile];
  const int64_t r = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  T acc = 0;
  for (int64_t k0 = 0; k0 < K; k0 += kCoefTile) {
    const int64_t kn = vmin<int64_t>(kCoefTile, K - k0);
    __syncthreads();
    for (int64_t kk = threadIdx.x; kk < kn; kk += blockDim.x) {
      const int64_t k = k0 + kk;
      T g = 0, l = 0;
      for (int64_t m = 0; m < n_g; ++m) g += gam[m * K + k];
      for (int64_t n = 0; n < n_l; ++n) l += lam[n * K + k];
      g /= T(n_g);
      l /= T(n_l);
      coef[kk] = -g / (l + del[k]) / sqrt(ev[k]);
    }
    __syncthreads();
    if (r < R)
      for (int64_t kk = 0; kk < kn; ++kk) acc += U[r * K + k0 + kk] * coef[kk];
  }
  if (r < R) v[r] = acc * corr;
}

}  // namespace vvt

using namespace vvt;

extern "C" {

int vvt_abi_version(void) { return 2; }
const char* vvt_last_error(void) { return last_error_buffer(); }
int64_t vvt_launch_count(void) { return g_launches.load(); }

int vvt_loss_sqrt_hessian_ce(void* S, const void* logits, const int64_t* sub, int64_t n_total,
                             int64_t n_sub, int64_t C, double scale, int dtype, void* stream) {
  VVT_REQUIRE(n_total >= 0 && n_sub >= 0 && C >= 0, "negative size");
  if (n_sub == 0 || C == 0) return VVT_OK;
  VVT_REQUIRE(S && logits, "null pointer");
  VVT_DISPATCH(dtype, {
    ce_factor_kernel<T, false><<<ew_blocks(n_sub * 32), 256, 0, as_stream(stream)>>>(
        (T*)S, (const T*)logits, sub, nullptr, n_sub, C, 0, T(scale));
    return launched(__func__);
  });
}

int vvt_loss_sqrt_hessian_ce_mc(void* S, const void* logits, const int64_t* sub,
                                const int64_t* class_ids, int64_t n_total, int64_t n_sub, int64_t C,
                                int64_t M, double scale, int dtype, void* stream) {
  VVT_REQUIRE(n_total >= 0 && n_sub >= 0 && C >= 0 && M >= 0, "negative size");
  if (n_sub == 0 || C == 0 || M == 0) return VVT_OK;
  VVT_REQUIRE(S && logits && class_ids, "null pointer");
  VVT_DISPATCH(dtype, {
    ce_factor_kernel<T, true><<<ew_blocks(n_sub * 32), 256, 0, as_stream(stream)>>>(
        (T*)S, (const T*)logits, sub, class_ids, n_sub, C, M, T(scale));
    return launched(__func__);
  });
}

int vvt_loss_sqrt_hessian_mse(void* S, int64_t n_sub, int64_t C, double scale, int dtype,
                              void* stream) {
  VVT_REQUIRE(n_sub >= 0 && C >= 0, "negative size");
  if (n_sub == 0 || C == 0) return VVT_OK;
  VVT_REQUIRE(S, "null pointer");
  VVT_DISPATCH(dtype, {
    mse_factor_kernel<T><<<ew_blocks(C * n_sub * C), 256, 0, as_stream(stream)>>>((T*)S, n_sub, C, T(scale));
    return launched(__func__);
  });
}

int vvt_scale(void* t, int64_t numel, double alpha, int dtype, void* stream) {
  VVT_REQUIRE(numel >= 0, "negative size");
  if (numel == 0) return VVT_OK;
  VVT_REQUIRE(t, "null pointer");
  VVT_DISPATCH(dtype, {
    scale_kernel<T><<<ew_blocks(numel), 256, 0, as_stream(stream)>>>((T*)t, numel, T(alpha));
    return launched(__func__);
  });
}

int vvt_axpy(void* y, const void* x, int64_t numel, double alpha, int dtype, void* stream) {
  VVT_REQUIRE(numel >= 0, "negative size");
  if (numel == 0) return VVT_OK;
  VVT_REQUIRE(y && x, "null pointer");
  VVT_DISPATCH(dtype, {
    axpy_kernel<T><<<ew_blocks(numel), 256, 0, as_stream(stream)>>>((T*)y, (const T*)x, numel, T(alpha));
    return launched(__func__);
  });
}

int vvt_sqrt_backprop_elementwise(void* out, const void* S, const void* ref, int64_t V,
                                  int64_t n_feat, int act, double scale, int dtype, void* stream) {
  VVT_REQUIRE(V >= 0 && n_feat >= 0, "negative size");
  VVT_REQUIRE(act >= 0 && act <= VVT_ACT_LOGSIGMOID, "unknown activation");
  if (V * n_feat == 0) return VVT_OK;
  VVT_REQUIRE(out && S && ref, "null pointer");
  VVT_DISPATCH(dtype, {
    constexpr int VW = 16 / int(sizeof(T));
    const bool aligned = ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(S) | reinterpret_cast<uintptr_t>(ref)) & 15) == 0;
    if (aligned && n_feat % VW == 0 && V > 1) {
      constexpr int UNROLL = 5;
      const int64_t groups = n_feat / VW;
      const unsigned bx = unsigned(vmax<int64_t>(1, vmin<int64_t>(ceil_div(groups, 256), 8 * num_sms())));
      // enough CTAs to fill the GPU: split the rows too when there are few feature groups
      const unsigned by = unsigned(vmax<int64_t>(1, vmin<int64_t>(ceil_div(V, UNROLL), ceil_div(4 * int64_t(num_sms()), bx))));
      act_rows_kernel<T, UNROLL><<<dim3(bx, by), 256, 0, as_stream(stream)>>>((T*)out, (const T*)S, (const T*)ref, V, n_feat,
                                                                              act, T(scale));
    } else {
      act_kernel<T><<<ew_blocks(V * n_feat), 256, 0, as_stream(stream)>>>((T*)out, (const T*)S, (const T*)ref, V, n_feat, act,
                                                                         T(scale));
    }
    return launched(__func__);
  });
}

int vvt_maxpool2d_argmax(int64_t* argmax, const void* x, int64_t planes, int64_t h_out, int64_t w_out, int64_t h_in,
                         int64_t w_in, int64_t kh, int64_t kw, int64_t stride_h, int64_t stride_w, int64_t pad_h,
                         int64_t pad_w, int64_t dil_h, int64_t dil_w, int dtype, void* stream) {
  VVT_REQUIRE(planes >= 0 && h_out >= 0 && w_out >= 0 && h_in >= 0 && w_in >= 0, "negative size");
  VVT_REQUIRE(stride_h > 0 && stride_w > 0 && dil_h > 0 && dil_w > 0 && kh > 0 && kw > 0 && pad_h >= 0 && pad_w >= 0,
              "bad window");
  VVT_REQUIRE(h_in * w_in < (int64_t(1) << 31) && h_out * w_out < (int64_t(1) << 31), "plane too large");
  const int64_t total = planes * h_out * w_out;
  if (total == 0) return VVT_OK;
  VVT_REQUIRE(argmax && x, "null pointer");
  // every window must contain at least one input position (torch's output-size rule guarantees it)
  VVT_REQUIRE((h_out - 1) * stride_h - pad_h < h_in && (w_out - 1) * stride_w - pad_w < w_in, "window outside the input");
  VVT_DISPATCH(dtype, {
    maxpool_argmax_kernel<T><<<ew_blocks(total), 256, 0, as_stream(stream)>>>(
        argmax, (const T*)x, total, int(h_out), int(w_out), int(h_in), int(w_in), int(kh), int(kw), int(stride_h),
        int(stride_w), int(pad_h), int(pad_w), int(dil_h), int(dil_w));
    return launched(__func__);
  });
}

int vvt_sqrt_backprop_maxpool2d(void* out, const void* S, const int64_t* argmax, int64_t V,
                                int64_t N, int64_t ch, int64_t h_out, int64_t w_out, int64_t h_in,
                                int64_t w_in, int64_t kh, int64_t kw, int64_t stride_h,
                                int64_t stride_w, int64_t pad_h, int64_t pad_w, int64_t dil_h,
                                int64_t dil_w, int dtype, void* stream) {
  VVT_REQUIRE(V >= 0 && N >= 0 && ch >= 0, "negative size");
  VVT_REQUIRE(stride_h > 0 && stride_w > 0 && dil_h > 0 && dil_w > 0 && kh > 0 && kw > 0, "bad window");
  const int64_t total = V * N * ch * h_in * w_in;
  if (total == 0) return VVT_OK;
  VVT_REQUIRE(out && S && argmax, "null pointer");
  VVT_DISPATCH(dtype, {
    launch_maxpool_bwd<T>((T*)out, (const T*)S, argmax, V * N * ch, N, ch, int(h_out), int(w_out), int(h_in),
                          int(w_in), int(kh), int(kw), int(stride_h), int(stride_w), int(pad_h), int(pad_w),
                          int(dil_h), int(dil_w), as_stream(stream));
    return launched(__func__);
  });
}

int vvt_sqrt_backprop_avgpool2d(void* out, const void* S, int64_t rows, int64_t ch, int64_t h_out,
                                int64_t w_out, int64_t h_in, int64_t w_in, int64_t kh, int64_t kw,
                                int64_t stride_h, int64_t stride_w, int64_t pad_h, int64_t pad_w,
                                int dtype, void* stream) {
  VVT_REQUIRE(rows >= 0 && ch >= 0, "negative size");
  VVT_REQUIRE(stride_h > 0 && stride_w > 0 && kh > 0 && kw > 0, "bad window");
  const int64_t total = rows * ch * h_in * w_in;
  if (total == 0) return VVT_OK;
  VVT_REQUIRE(out && S, "null pointer");
  VVT_DISPATCH(dtype, {
    avgpool_bwd_kernel<T><<<ew_blocks(total), 256, 0, as_stream(stream)>>>(
        (T*)out, (const T*)S, rows, ch, int(h_out), int(w_out), int(h_in), int(w_in), int(kh), int(kw),
        int(stride_h), int(stride_w), int(pad_h), int(pad_w));
    return launched(__func__);
  });
}

int vvt_v_emit_bias(void* Vt, const void* S, int64_t rows, int64_t c_out, int64_t spatial,
                    int dtype, void* stream) {
  VVT_REQUIRE(rows >= 0 && c_out >= 0 && spatial >= 0, "negative size");
  if (rows * c_out == 0) return VVT_OK;
  VVT_REQUIRE(Vt && S, "null pointer");
  VVT_DISPATCH(dtype, {
    bias_emit_kernel<T><<<ew_blocks(rows * c_out * 32), 256, 0, as_stream(stream)>>>(
        (T*)Vt, (const T*)S, rows * c_out, spatial);
    return launched(__func__);
  });
}

int vvt_v_emit_linear(void* Vt, const void* S, const void* Z, int64_t V, int64_t N, int64_t n_out,
                      int64_t n_in, int dtype, void* stream) {
  VVT_REQUIRE(V >= 0 && N >= 0 && n_out >= 0 && n_in >= 0, "negative size");
  const int64_t total = V * N * n_out * n_in;
  if (total == 0) return VVT_OK;
  VVT_REQUIRE(Vt && S && Z, "null pointer");
  VVT_DISPATCH(dtype, {
    linear_emit_kernel<T><<<ew_blocks(total), 256, 0, as_stream(stream)>>>(
        (T*)Vt, (const T*)S, (const T*)Z, V, N, n_out, n_in);
    return launched(__func__);
  });
}

int vvt_backtransform_dense(void* E, void* norm2, const void* U, const void* V, int64_t K,
                            int64_t R, int64_t D, int dtype, void* stream) {
  VVT_REQUIRE(K >= 0 && R >= 0 && D >= 0, "negative size");
  if (K == 0 || D == 0) return VVT_OK;
  VVT_REQUIRE(E && U && V, "null pointer");
  VVT_DISPATCH(dtype, {
    constexpr int VW = 16 / int(sizeof(T));  // one 16-byte load per row and thread
    const unsigned blocks = unsigned(ceil_div(D, 32 * VW));
    int64_t k0 = 0;
    while (k0 < K) {  // 16 directions per pass over V while that many remain, then 12 / 8 / 4 / 1
      const int64_t left = K - k0;
      if (left > 12) {
        backtransform_dense_kernel<T, 16, VW><<<blocks, 256, 0, as_stream(stream)>>>(
            (T*)E, (double*)norm2, (const T*)U, (const T*)V, K, R, D, k0);
        k0 += 16;
      } else if (left > 8) {
        backtransform_dense_kernel<T, 12, VW><<<blocks, 256, 0, as_stream(stream)>>>(
            (T*)E, (double*)norm2, (const T*)U, (const T*)V, K, R, D, k0);
        k0 += 12;
      } else if (left > 4) {
        backtransform_dense_kernel<T, 8, VW><<<blocks, 256, 0, as_stream(stream)>>>(
            (T*)E, (double*)norm2, (const T*)U, (const T*)V, K, R, D, k0);
        k0 += 8;
      } else if (left > 1) {
        backtransform_dense_kernel<T, 4, VW><<<blocks, 256, 0, as_stream(stream)>>>(
            (T*)E, (double*)norm2, (const T*)U, (const T*)V, K, R, D, k0);
        k0 += 4;
      } else {
        backtransform_dense_kernel<T, 1, VW><<<blocks, 256, 0, as_stream(stream)>>>(
            (T*)E, (double*)norm2, (const T*)U, (const T*)V, K, R, D, k0);
        k0 += 1;
      }
      VVT_TRY(launched(__func__));
    }
    return VVT_OK;
  });
}

int vvt_v_apply_dense(void* step, const void* v, const void* V, int64_t R, int64_t D, int dtype,
                      void* stream) {
  VVT_REQUIRE(R >= 0 && D >= 0, "negative size");
  if (D == 0) return VVT_OK;
  VVT_REQUIRE(step && v && V, "null pointer");
  VVT_DISPATCH(dtype, {
    constexpr int VW = 16 / int(sizeof(T));
    backtransform_dense_kernel<T, 1, VW><<<unsigned(ceil_div(D, 32 * VW)), 256, 0, as_stream(stream)>>>(
        (T*)step, nullptr, (const T*)v, (const T*)V, 1, R, D, 0);
    return launched(__func__);
  });
}

int vvt_scale_rows_rsqrt(void* E, const void* norm2, int64_t K, int64_t D, int dtype, void* stream) {
  VVT_REQUIRE(K >= 0 && D >= 0, "negative size");
  if (K * D == 0) return VVT_OK;
  VVT_REQUIRE(E && norm2, "null pointer");
  VVT_DISPATCH(dtype, {
    scale_rows_rsqrt_kernel<T><<<ew_blocks(K * D), 256, 0, as_stream(stream)>>>((T*)E, (const double*)norm2, K, D);
    return launched(__func__);
  });
}

int vvt_center_rows(void* out, const void* g, int64_t N, int64_t D, int dtype, void* stream) {
  VVT_REQUIRE(N >= 0 && D >= 0, "negative size");
  if (N * D == 0) return VVT_OK;
  VVT_REQUIRE(out && g, "null pointer");
  VVT_DISPATCH(dtype, {
    constexpr int VEC = 16 / int(sizeof(T));
    const bool vec = D % VEC == 0 && (reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(g)) % 16 == 0;
    const int64_t strips = ceil_div(D, vec ? 32 * VEC : 32);
    const unsigned blocks = unsigned(vmin<int64_t>(strips, 16 * int64_t(num_sms())));
    if (vec)
      center_rows_kernel<T, VEC><<<blocks, 256, 0, as_stream(stream)>>>((T*)out, (const T*)g, N, D);
    else
      center_rows_kernel<T, 1><<<blocks, 256, 0, as_stream(stream)>>>((T*)out, (const T*)g, N, D);
    return launched(__func__);
  });
}

int vvt_filter_nonzero(uint8_t* mask, const void* evals, int64_t R, double atol, double rtol,
                       int64_t* count_host, int dtype, void* stream) {
  VVT_REQUIRE(R >= 0, "negative size");
  if (R == 0) {
    if (count_host) *count_host = 0;
    return VVT_OK;
  }
  VVT_REQUIRE(mask && evals, "null pointer");
  cudaStream_t s = as_stream(stream);
  unsigned long long* dcount = nullptr;
  if (count_host) {
    VVT_TRY(check_cuda(cudaMallocAsync((void**)&dcount, sizeof(unsigned long long), s), __func__));
    VVT_TRY(check_cuda(cudaMemsetAsync(dcount, 0, sizeof(unsigned long long), s), __func__));
  }
  VVT_DISPATCH(dtype, {
    filter_nonzero_kernel<T><<<ew_blocks(R), 256, 0, s>>>(mask, (const T*)evals, R, T(atol), T(rtol), dcount);
    VVT_TRY(launched(__func__));
  });
  if (count_host) {
    unsigned long long h = 0;
    VVT_TRY(check_cuda(cudaMemcpyAsync(&h, dcount, sizeof(h), cudaMemcpyDeviceToHost, s), __func__));
    VVT_TRY(check_cuda(cudaStreamSynchronize(s), __func__));
    VVT_TRY(check_cuda(cudaFreeAsync(dcount, s), __func__));
    *count_host = int64_t(h);
  }
  return VVT_OK;
}

int vvt_newton_coeff(void* v, const void* U, const void* gammas, const void* lambdas,
                     const void* deltas, const void* evals, int64_t R, int64_t K, int64_t n_g,
                     int64_t N_ggn, double corr, int dtype, void* stream) {
  VVT_REQUIRE(R >= 0 && K >= 0 && n_g > 0 && N_ggn > 0, "bad size");
  if (R == 0) return VVT_OK;
  VVT_REQUIRE(v && U && gammas && lambdas && deltas && evals, "null pointer");
  VVT_DISPATCH(dtype, {
    const unsigned blocks = unsigned(ceil_div(R, 256));
    newton_coeff_kernel<T><<<blocks, 256, 0, as_stream(stream)>>>(
        (T*)v, (const T*)U, (const T*)gammas, (const T*)lambdas, (const T*)deltas, (const T*)evals, R,
        K, n_g, N_ggn, T(corr));
    return launched(__func__);
  });
}

}  // extern "C"
