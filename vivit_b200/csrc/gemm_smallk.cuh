// (header) Persistent tcgen05 NT-GEMM for SHORT contractions:  C[b][m, n] = sum_k A[b][m, k] B[b][n, k],  K <= 160.
//
// The conv factor emit (K = spatial positions of a small feature map) and the conv data gradient (K = output
// channels) are GEMMs with two to five k-blocks per 128 x 128 output tile.  With one CTA per tile (gram_tc_kernel)
// such a tile is prologue and epilogue -- TMEM allocation, barrier set-up, pipeline fill, a 64 KB tile drained through
// shared memory -- around 24 ... 60 MMAs: measured 19 - 21 TFLOP/s on the data gradient of cifar10_3c3d.  Here
// a CTA walks a contiguous range of (batch, column tile, row tile) items, row tiles fastest:
//   * the B tile (hi and lo of the 3xTF32 split, all its k-blocks) stays in shared memory while batch entry and
//     column tile do not change;
//   * the rows of A stream through a 4-deep ring of k-blocks; the converter warps hand them (hi and lo) to the tensor
//     core through TENSOR MEMORY, so an MMA reads only B from shared memory;
//   * two TMEM accumulators alternate between the MMA warp and four epilogue warps, which pass the finished tile to
//     the store functor straight from registers (a thread owns one row of the tile).
//   warp 0: TMA producer   warp 1: MMA issuer   warps 2-5: converters   warps 6-9: epilogue
// The structure is that of wide_apply_kernel (eig_wide.cuh), which is the K = 128, one-B-tile-per-pair special case.
//
// Store functor ST:  row_offset(batch, m) -> int64 (negative: skip the row);
//                    store_chunk(off, m, col0, ncols, const uint32_t (&r)[32])  -- ncols valid values of row m from col0.
// A4D: rows of an A tile are a 32 x 4 box of two row indices (see gram_tc_kernel).
#pragma once
#include "gemm_tc.cuh"

namespace vvt {
namespace smallk {

constexpr int MAX_KB = 5;  // k-blocks of 32 (K <= 160)
constexpr int A_STAGES = 4;
constexpr int THREADS = 320;
constexpr int TMEM_COLS_ALL = 512;  // two accumulators (2 x 128) + A_STAGES x (32 hi + 32 lo) columns of A

static inline size_t smem_bytes(int nkb) {
  return size_t(nkb) * 2 * tc::TILE_BYTES + size_t(A_STAGES) * tc::TILE_BYTES + 1024 + 256;
}

struct Args {
  int64_t M, N;       // rows / columns of one batch entry (M may be a padded count, see A4D)
  int tiles_m, tiles_n, batch, nkb, a4_chunks;
};

template <typename ST, bool A4D>
__global__ void __launch_bounds__(THREADS, 1)
smallk_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, ST st, Args a) {
  using namespace tc;
  extern __shared__ unsigned char smem_raw[];
  const int nkb = a.nkb;
  const int64_t total = int64_t(a.batch) * a.tiles_n * a.tiles_m;
  const int64_t per = (total + gridDim.x - 1) / gridDim.x;
  const int64_t it0 = blockIdx.x * per, it1 = vmin<int64_t>(total, it0 + per);

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t b_bytes = uint32_t(nkb) * 2 * TILE_BYTES;
  const uint32_t bbase = base, abase = base + b_bytes;
  const uint32_t bars = abase + A_STAGES * TILE_BYTES;
  auto bar_a_tma = [&](int s) { return bars + 8u * s; };
  auto bar_a_conv = [&](int s) { return bars + 8u * (A_STAGES + s); };
  auto bar_a_empty = [&](int s) { return bars + 8u * (2 * A_STAGES + s); };
  const uint32_t bar_b_tma = bars + 8u * (3 * A_STAGES), bar_b_conv = bar_b_tma + 8, bar_b_empty = bar_b_tma + 16;
  auto bar_acc_full = [&](int b) { return bars + 8u * (3 * A_STAGES + 3 + b); };
  auto bar_acc_empty = [&](int b) { return bars + 8u * (3 * A_STAGES + 5 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + b_bytes + A_STAGES * TILE_BYTES + 8 * (3 * A_STAGES + 7));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < A_STAGES; ++s) {
      mbar_init(bar_a_tma(s), 1);
      mbar_init(bar_a_conv(s), 128);
      mbar_init(bar_a_empty(s), 1);
    }
    mbar_init(bar_b_tma, 1);
    mbar_init(bar_b_conv, 128);
    mbar_init(bar_b_empty, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc_full(b), 1);
      mbar_init(bar_acc_empty(b), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS_ALL) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  auto a_cols = [&](int s) { return tmem_base + uint32_t(TMEM_COLS + 64 * s); };  // hi at +0, lo at +32
  auto b_hi = [&](int j) { return bbase + uint32_t(j) * 2 * TILE_BYTES; };
  // item -> (batch entry, column tile, row tile), row tiles fastest
  auto decode = [&](int64_t it, int& b, int& tn, int& tm) {
    tm = int(it % a.tiles_m);
    const int64_t t = it / a.tiles_m;
    tn = int(t % a.tiles_n);
    b = int(t / a.tiles_n);
  };

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int kbc = 0, bgen = 0, last_key = -1;
      for (int64_t it = it0; it < it1; ++it) {
        int b, tn, tm;
        decode(it, b, tn, tm);
        const int key = b * a.tiles_n + tn;
        if (key != last_key) {
          if (bgen > 0) mbar_wait(bar_b_empty, (bgen - 1) & 1);  // every MMA that read the old B tile has completed
          mbar_expect_tx(bar_b_tma, uint32_t(nkb) * TILE_BYTES);
          for (int j = 0; j < nkb; ++j) tma_load_3d(b_hi(j), &mapB, bar_b_tma, j * BK, tn * BN, b);
          ++bgen;
          last_key = key;
        }
        for (int j = 0; j < nkb; ++j, ++kbc) {
          const int s = kbc % A_STAGES, use = kbc / A_STAGES;
          if (use > 0) mbar_wait(bar_a_empty(s), (use - 1) & 1);
          mbar_expect_tx(bar_a_tma(s), TILE_BYTES);
          const uint32_t dst = abase + uint32_t(s) * TILE_BYTES;
          if constexpr (A4D) tma_load_4d(dst, &mapA, bar_a_tma(s), j * BK, (tm % a.a4_chunks) * 32, b, (tm / a.a4_chunks) * 4);
          else tma_load_3d(dst, &mapA, bar_a_tma(s), j * BK, tm * BM, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: A (hi / lo) from tensor memory, B (hi / lo) from shared memory =====
    if (lane == 0) {
      int kbc = 0, bgen = 0, last_key = -1, n = 0;
      for (int64_t it = it0; it < it1; ++it) {
        int b, tn, tm;
        decode(it, b, tn, tm);
        const int key = b * a.tiles_n + tn;
        if (key != last_key) {
          if (last_key >= 0) umma_commit(bar_b_empty);  // fires when every MMA issued so far has completed
          mbar_wait(bar_b_tma, bgen & 1);
          mbar_wait(bar_b_conv, bgen & 1);
          tcgen05_fence_after();
          ++bgen;
          last_key = key;
        }
        const int buf = n & 1;
        if (n >= 2) {
          mbar_wait(bar_acc_empty(buf), ((n >> 1) - 1) & 1);
          tcgen05_fence_after();
        }
        const uint32_t acc = tmem_base + uint32_t(buf * BN);
        for (int j = 0; j < nkb; ++j, ++kbc) {
          const int s = kbc % A_STAGES, use = kbc / A_STAGES;
          const uint32_t bh = b_hi(j), bl = bh + TILE_BYTES;
          mbar_wait(bar_a_conv(s), use & 1);
          tcgen05_fence_after();
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint32_t a_hi = a_cols(s) + 8 * k, a_lo = a_hi + 32;
            umma_tf32_ts(acc, a_hi, make_desc(bh + 32 * k), kIdesc, !(j == 0 && k == 0));
            umma_tf32_ts(acc, a_hi, make_desc(bl + 32 * k), kIdesc, 1);
            umma_tf32_ts(acc, a_lo, make_desc(bh + 32 * k), kIdesc, 1);
          }
          umma_commit(bar_a_empty(s));
        }
        umma_commit(bar_acc_full(buf));
        ++n;
      }
    }
  } else if (warp < 6) {
    // ===== converters: lo of the B tile (shared memory, once per tile), hi / lo of every k-block of A (tensor memory) =====
    const int ct = threadIdx.x - 64;  // 0..127
    const int lane_grp = warp & 3, m = lane_grp * 32 + lane;
    int kbc = 0, bgen = 0, last_key = -1;
    for (int64_t it = it0; it < it1; ++it) {
      int b, tn, tm;
      decode(it, b, tn, tm);
      const int key = b * a.tiles_n + tn;
      if (key != last_key) {
        mbar_wait(bar_b_tma, bgen & 1);
        for (int j = 0; j < nkb; ++j) {
          unsigned char* src = base_ptr + size_t(j) * 2 * TILE_BYTES;
#pragma unroll 4
          for (int v = ct; v < TILE_BYTES / 16; v += 128) {
            const float4 x = *reinterpret_cast<const float4*>(src + size_t(v) * 16);
            const float e[4] = {x.x, x.y, x.z, x.w};
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float hi = __uint_as_float(__float_as_uint(e[q]) & 0xFFFFE000u);
              o[q] = __uint_as_float(__float_as_uint(e[q] - hi) + 0x1000u);
            }
            *reinterpret_cast<float4*>(src + TILE_BYTES + size_t(v) * 16) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
        fence_proxy_async();
        mbar_arrive(bar_b_conv);
        ++bgen;
        last_key = key;
      }
      for (int j = 0; j < nkb; ++j, ++kbc) {
        const int s = kbc % A_STAGES, use = kbc / A_STAGES;
        const unsigned char* stg = base_ptr + b_bytes + size_t(s) * TILE_BYTES;
        mbar_wait(bar_a_tma(s), use & 1);
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {  // row m of the k-block: chunk c sits at chunk c ^ (m & 7) (128-byte swizzle)
          const float4 x = *reinterpret_cast<const float4*>(stg + size_t(m) * 128 + size_t((c ^ (m & 7)) * 16));
          const float e[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t bits = __float_as_uint(e[q]);
            hi[4 * c + q] = bits;
            lo[4 * c + q] = __float_as_uint(e[q] - __uint_as_float(bits & 0xFFFFE000u)) + 0x1000u;
          }
        }
        const uint32_t ta = a_cols(s) + (uint32_t(lane_grp * 32) << 16);
        tmem_st32(ta, hi);
        tmem_st32(ta + 32, lo);
        tmem_st_wait();
        tcgen05_fence_before();
        mbar_arrive(bar_a_conv(s));
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> store functor (thread = one row of the tile) =====
    const int lane_grp = warp & 3;  // a warp may only touch TMEM lanes 32 * (warp % 4) .. + 31
    const int row = lane_grp * 32 + lane;
    int n = 0;
    for (int64_t it = it0; it < it1; ++it) {
      int b, tn, tm;
      decode(it, b, tn, tm);
      const int buf = n & 1;
      mbar_wait(bar_acc_full(buf), (n >> 1) & 1);
      tcgen05_fence_after();
      const int64_t mrow = int64_t(tm) * BM + row;
      const int64_t off = mrow < a.M ? st.row_offset(b, mrow) : int64_t(-1);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_base + (uint32_t(lane_grp * 32) << 16) + uint32_t(buf * BN + 32 * c), r);
        const int64_t col0 = int64_t(tn) * BN + 32 * c;
        const int ncols = int(vmin<int64_t>(32, a.N - col0));
        if (off >= 0 && ncols > 0) st.store_chunk(off, mrow, col0, ncols, r);
      }
      tcgen05_fence_before();
      mbar_arrive(bar_acc_empty(buf));
      ++n;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS_ALL) : "memory");
  }
}

// Is the persistent form applicable / worthwhile?  Short contraction, enough tiles to keep every SM busy for a while.
static inline bool worthwhile(int64_t M, int64_t N, int64_t K, int64_t batch) {
  static const bool off = getenv("VVT_NO_SMALLK") != nullptr;
  if (off || K > int64_t(MAX_KB) * tc::BK) return false;
  // (a CTA loads and converts a B tile before its first MMA: below about six tiles per SM the per-tile kernel,
  // which spreads the tiles over more CTAs, is as fast or faster -- measured on the All-CNN-C layers)
  const int64_t tiles = ceil_div(M, tc::BM) * ceil_div(N, tc::BN) * batch;
  return tiles >= 6 * int64_t(num_sms());
}

template <typename ST, bool A4D>
static inline int launch(const CUtensorMap& mapA, const CUtensorMap& mapB, ST st, int64_t M, int64_t N, int64_t K, int64_t batch,
                         int tiles_m, int a4_chunks, cudaStream_t stream, const char* what) {
  auto kern = smallk_kernel<ST, A4D>;
  const int nkb = int(vmax<int64_t>(1, ceil_div(K, tc::BK)));
  static SmemOptIn opt_in;  // per instantiation
  VVT_TRY(opt_in.ensure(kern, smem_bytes(MAX_KB), what));
  Args a;
  a.M = M, a.N = N;
  a.tiles_m = tiles_m, a.tiles_n = int(ceil_div(N, tc::BN)), a.batch = int(batch), a.nkb = nkb, a.a4_chunks = a4_chunks;
  const int64_t total = int64_t(a.batch) * a.tiles_n * a.tiles_m;
  const unsigned ctas = unsigned(vmax<int64_t>(1, vmin<int64_t>(total, num_sms())));
  kern<<<ctas, THREADS, smem_bytes(nkb), stream>>>(mapA, mapB, st, a);
  return launched(what);
}

}  // namespace smallk
}  // namespace vvt
