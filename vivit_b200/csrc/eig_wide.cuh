// Two-level ("wide") rounds of the one-sided block Jacobi eigensolver for large fp32 problems (R >= 2048).
//
// With 16-column blocks a sweep moves the whole R x R factor R / 16 times and every block pair pays the fixed
// latency chain of a round; beyond a few thousand columns that is what the solver's time is made of
// (R = 5120: 449 ms, R = 10240: 5.96 s in round 1).  Here the columns are cut into WIDE blocks of 64; a round
// pairs the wide blocks (the same round-robin tournament) and is three launches over ALL pairs and all
// problems of a batch:
//
//   1. wide_tc_kernel<GRAM>   H = P^T P of every 128-column pair panel P = [L_a | L_b]   (128 x 128 x R)
//        tcgen05.mma kind::tf32 with MN-MAJOR operands: rows are the contraction index and are not contiguous
//        (the factor is stored wide-block-major, [wide block][row][64 columns]); a k-block is
//        four TMA boxes of 32 rows x 32 columns (128-byte swizzle with 32-byte atoms) which land exactly in
//        the canonical MN-major SWIZZLE_128B_BASE32B layout, the only one the tensor core takes for MN-major
//        tf32 (see make_desc_mn).  3xTF32 split by the converter warps, split-K over the rows, raw
//        partial sums to global memory (summed in a fixed order by the next kernel: deterministic).
//   2. wide_rot_kernel        one CTA (1024 threads) per pair: sums the partials, then diagonalises H by
//        cyclic-by-blocks two-sided Jacobi on its eight 16-column sub-blocks: four groups of 256 threads run
//        the scalar rotation rounds of the 16-wide kernel (rotation_rounds, named barriers) on four disjoint
//        sub-block pairs at a time, then all threads apply the four 32 x 32 rotations to H (both sides) and to
//        the accumulated Q (128 x 128).  Cross rounds rotate every column of a against every column of b;
//        the intra round (once per sweep) rotates inside the two wide blocks.  Emits Q^T and a "rotated" flag.
//   3. wide_tc_kernel<APPLY>  P <- P Q in place: tcgen05, K-major operands (rows of the factor / rows of
//        Q^T), one CTA per 128 rows x pair, staged coalesced stores.
//
// A sweep moves the factor R / 64 times instead of R / 16, all O(R) work is tensor-core GEMM, and the scalar
// part works on 128 x 128 matrices in shared memory.  Thresholds, rotation formulas and the convergence
// statistic (16-column block pairs that still rotated) are those of the 16-wide kernel.
#pragma once
#include "gemm_tc.cuh"

namespace vvt {
namespace wide {

constexpr int WB = 64;             // wide block (columns)
constexpr int WP = 2 * WB;         // pair panel
constexpr int NSUB = WP / OB;      // 16-column sub-blocks of a panel (8)
constexpr int LDW = WP + 1;        // padded pitch of the 128 x 128 matrices in shared memory
constexpr int RT = 1024;           // threads of the rotation kernel
constexpr int NGRP = RT / OT;      // groups of 256 threads (4)
enum { GRAM = 0, APPLY = 1 };

__host__ __device__ inline void wide_blocks(int nbw, int round, int pair, int& wa, int& wb) {
  if (round < 0) {
    wa = 2 * pair;
    wb = 2 * pair + 1;
  } else {
    rr_pair(nbw, round, pair, wa, wb);
  }
}

// The three kernels of a round are launched with programmatic stream serialization: a kernel lets its successor be
// scheduled at once (launch_dependents) and blocks in griddepcontrol.wait until its predecessor has completed and
// flushed, so that the launch latency of a kernel (3 - 5 us, three per round, ~1100 rounds per solve at R = 5120)
// hides behind the tail of the one before it.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

struct WideArgs {
  float* out;               // GRAM: partial sums [problem][split][pair][128][128];  APPLY: the factor
  const int* flag;          // APPLY: [problem][pair], 0 = the pair's Gram was already diagonal
  const JacobiScalars* sc;  // [problem]
  int64_t l_stride;         // elements between the factors of two problems
  int Np, nbw, round, pairs, splits, kblocks_total, kblocks_per_split;
  int pair0, pairs_local;  // this launch covers pairs pair0 .. pair0 + pairs_local - 1 of the round (all of them on
                           // one GPU; a rank's share when the pairs of a round are distributed, vvt_syevj_dist)
  int variant;  // experiments (VVT_WIDE_DESC): descriptor encodings of the MN-major Gram operands
  int cross_only;  // GRAM, cross rounds: only the 128 x 64 block H[:, b] (the diagonal block of a comes from a cache)
};

// MN-major fp32 (tf32) operands have ONE legal shared-memory layout on sm_100: 128-byte swizzle with 32-byte
// atomicity (UMMA layout type SWIZZLE_128B_BASE32B = 1, TMA swizzle CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; with
// the plain SWIZZLE_128B type the tensor core returns zeros -- measured).  Canonical form
// ((4,8,m),(4,k)) : ((1,4,LBO),(32,SBO)) in elements: 32 contiguous M/N elements (128 bytes) x 4 K rows of 128
// bytes per 512-byte atom, 32-byte chunks XOR-ed with the row index mod 4.  A TMA box of 32 rows x 32 columns
// of the row-major factor lands in exactly this form: K groups of 4 rows 512 bytes apart (SBO), the four
// 32-column chunks of a panel are separate boxes 4096 bytes apart (LBO).  One MMA (K = 8) spans two K groups;
// the next k-step starts 1024 bytes further.  variant: experiments only.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, int variant = 0) {
  const uint32_t lbo = variant == 1 ? 512 : 4096, sbo = variant == 1 ? 4096 : (variant == 4 ? 1024 : 512);
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFF) >> 4);  // start address
  d |= uint64_t(lbo >> 4) << 16;              // leading byte offset: 32-element chunks along M/N
  d |= uint64_t(sbo >> 4) << 32;              // stride byte offset: 4-row groups along K
  d |= uint64_t(1) << 46;                     // descriptor version (Blackwell)
  d |= uint64_t(1) << 61;                     // SWIZZLE_128B_BASE32B
  return d;
}
constexpr uint32_t kIdescMN = tc::kIdesc | (1u << 15) | (1u << 16);  // A and B MN-major
constexpr uint32_t kIdescBMN = tc::kIdesc | (1u << 16);               // A from tensor memory, B MN-major
constexpr uint32_t kIdescBMN64 = (kIdescBMN & ~(0x3Fu << 17)) | (uint32_t(WB >> 3) << 17);  // the same with N = 64

template <int MODE>
__global__ void __launch_bounds__(tc::THREADS, 1)
wide_tc_kernel(const __grid_constant__ CUtensorMap mapL, const __grid_constant__ CUtensorMap mapQ, WideArgs a) {
  using namespace tc;
  extern __shared__ unsigned char smem_raw[];
  pdl_enter();
  const int prob = blockIdx.z;
  if (*reinterpret_cast<const volatile int*>(&a.sc[prob].converged)) return;  // uniform over the CTA
  const int pair = a.pair0 + int(MODE == GRAM ? blockIdx.x : blockIdx.y);
  const int split = MODE == GRAM ? blockIdx.y : 0;
  const int rtile = MODE == GRAM ? 0 : blockIdx.x;
  if (MODE == APPLY && a.flag[prob * a.pairs + pair] == 0) return;
  int wa, wb;
  wide_blocks(a.nbw, a.round, pair, wa, wb);
  constexpr bool diag = MODE == GRAM;  // one operand tile serves as A and B
  // Cross rounds need only H[:, b] = P^T P_b (H_ab and H_bb): H_aa is kept up to date by the rotation kernel (the
  // diagonal-block cache, as in the 16-wide path) and H_ba = H_ab^T.  N = 64 instead of 128: half the MMA time and
  // half the shared-memory operand traffic; the operand loads from global memory stay (A = the whole panel).
  const bool cross = MODE == GRAM && a.cross_only != 0 && a.round >= 0;
  const int bn = cross ? WB : BN;
  const int kb0 = MODE == GRAM ? split * a.kblocks_per_split : 0;
  const int nkb = MODE == GRAM ? min(a.kblocks_total, kb0 + a.kblocks_per_split) - kb0 : WP / BK;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + STAGES * STAGE_BYTES;
  auto bar_tma = [&](int s) { return bars + 8u * s; };
  auto bar_conv = [&](int s) { return bars + 8u * (STAGES + s); };
  auto bar_empty = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  auto bar_acc_full = [&](int b) { return bars + 8u * (3 * STAGES + b); };
  auto bar_acc_empty = [&](int b) { return bars + 8u * (3 * STAGES + 2 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + STAGES * STAGE_BYTES + 8 * (3 * STAGES + 4));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_tma(s), 1);
      mbar_init(bar_conv(s), 128);
      mbar_init(bar_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc_full(b), 1);
      mbar_init(bar_acc_empty(b), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // GRAM: the A operand (hi and lo) lives in tensor memory behind the two accumulators, 64 columns per stage
  // (see gram_tc_kernel<.., TS = true>: the MMAs then fetch only B from shared memory)
  constexpr bool TS = MODE == GRAM;
  constexpr int kTmemCols = TS ? 512 : TMEM_COLS;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  auto a_cols = [&](int s) { return tmem_base + uint32_t(TMEM_COLS + 64 * s); };  // hi at +0, lo at +32

  // the chunk-th group of 32 panel columns: wide block (third TMA coordinate) and first column inside it
  auto chunk_blk = [&](int chunk) { return prob * a.nbw + (chunk < 2 ? wa : wb); };
  auto chunk_c0 = [&](int chunk) { return (chunk & 1) * 32; };

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES, use = i / STAGES;
        if (use > 0) mbar_wait(bar_empty(s), (use - 1) & 1);
        const uint32_t stage = base + s * STAGE_BYTES;
        if (MODE == GRAM) {  // 32 rows of the factor x the 128 panel columns: four boxes of 32 x 32
          mbar_expect_tx(bar_tma(s), TILE_BYTES);
          const int r0 = (kb0 + i) * BK;
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_3d(stage + j * 4096, &mapL, bar_tma(s), chunk_c0(j), r0, chunk_blk(j));
        } else {  // 128 rows x 32 panel columns of the factor; 128 rows x 32 columns of Q^T
          mbar_expect_tx(bar_tma(s), 2 * TILE_BYTES);
          tma_load_3d(stage, &mapL, bar_tma(s), chunk_c0(i), rtile * BM, chunk_blk(i));
          tma_load_3d(stage + TILE_BYTES, &mapQ, bar_tma(s), i * BK, 0, prob * a.pairs + pair);
        }
      }
    }
  } else if (warp == 1 && TS) {
    // ===== MMA issuer, GRAM: A from tensor memory, B = the MN-major panel tile (raw / lo) in shared memory =====
    if (lane == 0) {
      const int vr = a.variant;
      const uint32_t idesc = cross ? kIdescBMN64 : kIdescBMN;
      const uint32_t b_off = cross ? 2u * 4096u : 0u;  // the two 32-column chunks of b
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES, use = i / STAGES;
        const int grp = i / PROMOTE, first = (i % PROMOTE) == 0;
        const uint32_t acc = tmem_base + uint32_t((grp & 1) * BN);
        const uint32_t b_hi = base + s * STAGE_BYTES + b_off, b_lo = b_hi + 2 * TILE_BYTES;
        if (first && grp >= 2) {
          mbar_wait(bar_acc_empty(grp & 1), ((grp >> 1) - 1) & 1);
          tcgen05_fence_after();
        }
        mbar_wait(bar_conv(s), use & 1);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {  // one MMA (K = 8) = two 4-row groups of B, next k-step 1 KB further
          const uint32_t a_hi = a_cols(s) + 8 * k, a_lo = a_hi + 32;
          umma_tf32_ts(acc, a_hi, make_desc_mn(b_hi + 1024 * k, vr), idesc, !(first && k == 0));
          umma_tf32_ts(acc, a_hi, make_desc_mn(b_lo + 1024 * k, vr), idesc, 1);
          umma_tf32_ts(acc, a_lo, make_desc_mn(b_hi + 1024 * k, vr), idesc, 1);
        }
        umma_commit(bar_empty(s));
        if ((i % PROMOTE) == PROMOTE - 1 || i == nkb - 1) umma_commit(bar_acc_full(grp & 1));
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer, per-tile APPLY (software-pipelined by one k-block, see gemm_tc.cuh) =====
    if (lane == 0) {
      auto stage_of = [&](int i) { return base + (i % STAGES) * STAGE_BYTES; };
      auto issue_hi = [&](int i) {
        const int s = i % STAGES, use = i / STAGES;
        const int grp = i / PROMOTE, first = (i % PROMOTE) == 0;
        const uint32_t acc = tmem_base + uint32_t((grp & 1) * BN);
        const uint32_t a_hi = stage_of(i), b_hi = a_hi + TILE_BYTES;
        if (first && grp >= 2) {
          mbar_wait(bar_acc_empty(grp & 1), ((grp >> 1) - 1) & 1);
          tcgen05_fence_after();
        }
        mbar_wait(bar_tma(s), use & 1);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < BK / 8; ++k)
          umma_tf32(acc, make_desc(a_hi + 32 * k), make_desc(b_hi + 32 * k), kIdesc, !(first && k == 0));
      };
      issue_hi(0);
      for (int i = 0; i < nkb; ++i) {
        if (i + 1 < nkb) issue_hi(i + 1);
        const int s = i % STAGES, use = i / STAGES;
        const int grp = i / PROMOTE;
        const uint32_t acc = tmem_base + uint32_t((grp & 1) * BN);
        const uint32_t a_hi = stage_of(i), b_hi = a_hi + TILE_BYTES;
        const uint32_t a_lo = a_hi + 2 * TILE_BYTES, b_lo = a_hi + 3 * TILE_BYTES;
        mbar_wait(bar_conv(s), use & 1);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          umma_tf32(acc, make_desc(a_hi + 32 * k), make_desc(b_lo + 32 * k), kIdesc, 1);
          umma_tf32(acc, make_desc(a_lo + 32 * k), make_desc(b_hi + 32 * k), kIdesc, 1);
        }
        umma_commit(bar_empty(s));
        if ((i % PROMOTE) == PROMOTE - 1 || i == nkb - 1) umma_commit(bar_acc_full(grp & 1));
      }
    }
  } else {
    // ===== converters (lo tiles), promotion of finished accumulator groups, epilogue =====
    const int ct = threadIdx.x - 64;  // 0..127
    constexpr int n_vec = (diag ? 1 : 2) * (TILE_BYTES / 16);
    const int lane_grp = warp & 3;
    const int n_groups = (nkb + PROMOTE - 1) / PROMOTE;
    float total[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) total[j] = 0.f;
    auto drain = [&](int grp) {
      mbar_wait(bar_acc_full(grp & 1), (grp >> 1) & 1);
      tcgen05_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (c0 < bn) {  // uniform
          uint32_t r[32];
          tmem_ld32(tmem_base + (uint32_t(lane_grp * 32) << 16) + uint32_t((grp & 1) * BN + c0), r);
#pragma unroll
          for (int j = 0; j < 32; ++j) total[c0 + j] += __uint_as_float(r[j]);
        }
      }
      tcgen05_fence_before();
      mbar_arrive(bar_acc_empty(grp & 1));
    };
    for (int i = 0; i < nkb; ++i) {
      const int s = i % STAGES, use = i / STAGES;
      unsigned char* stage = base_ptr + s * STAGE_BYTES;
      mbar_wait(bar_tma(s), use & 1);
      if constexpr (TS) {
        // A role: panel column m = this thread's TMEM lane; its 32 rows (K) of the MN-major tile: chunk m / 32 at
        // 4096 bytes, 4-row group k / 4 at 512, row k % 4 at 128, 32-byte chunk (m % 32) / 8 XOR k % 4
        const int m = lane_grp * 32 + lane;
        const unsigned char* col = stage + size_t(m >> 5) * 4096 + size_t(m & 7) * 4;
        const int c32 = (m & 31) >> 3;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float x = *reinterpret_cast<const float*>(col + (k >> 2) * 512 + (k & 3) * 128 + ((c32 ^ (k & 3)) << 5));
          const uint32_t bits = __float_as_uint(x);
          hi[k] = bits;
          lo[k] = __float_as_uint(x - __uint_as_float(bits & 0xFFFFE000u)) + 0x1000u;
        }
        const uint32_t ta = a_cols(s) + (uint32_t(lane_grp * 32) << 16);
        tmem_st32(ta, hi);
        tmem_st32(ta + 32, lo);
        tmem_st_wait();
      }
#pragma unroll 4
      for (int v = (cross ? n_vec / 2 : 0) + ct; v < n_vec; v += 128) {  // (cross: only the chunks of b are a B operand)
        const float4 x = *reinterpret_cast<const float4*>(stage + size_t(v) * 16);
        const float e[4] = {x.x, x.y, x.z, x.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float hi = __uint_as_float(__float_as_uint(e[j]) & 0xFFFFE000u);
          o[j] = __uint_as_float(__float_as_uint(e[j] - hi) + 0x1000u);
        }
        *reinterpret_cast<float4*>(stage + 2 * TILE_BYTES + size_t(v) * 16) = make_float4(o[0], o[1], o[2], o[3]);
      }
      fence_proxy_async();
      if constexpr (TS) tcgen05_fence_before();
      mbar_arrive(bar_conv(s));
      if ((i % PROMOTE) == PROMOTE - 1 && i / PROMOTE >= 1) drain(i / PROMOTE - 1);
    }
    for (int g = (nkb % PROMOTE == 0) ? n_groups - 1 : vmax(0, n_groups - 2); g < n_groups; ++g) drain(g);
    // epilogue: every MMA and every TMA load of this CTA has completed; the stages hold the staged tile
    float* tile = reinterpret_cast<float*>(base_ptr);
    const int trow = lane_grp * 32 + lane;
#pragma unroll
    for (int j = 0; j < BN; ++j) tile[trow * (BN + 1) + j] = total[j];
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int cw = warp - 2;  // 0..3: this warp stores rows cw, cw + 4, ...; lanes = consecutive columns
    if (MODE == GRAM) {
      float* out = a.out + ((size_t(prob) * a.splits + split) * a.pairs + pair) * size_t(WP * WP);
      if (cross) {  // columns 64 .. 127 of the pair Gram
        for (int r = cw; r < BM; r += 4)
#pragma unroll
          for (int q = 0; q < WB / 32; ++q) out[r * WP + WB + lane + 32 * q] = tile[r * (BN + 1) + lane + 32 * q];
      } else {
        for (int r = cw; r < BM; r += 4)
#pragma unroll
          for (int q = 0; q < BN / 32; ++q) out[r * WP + lane + 32 * q] = tile[r * (BN + 1) + lane + 32 * q];
      }
    } else {
      for (int r = cw; r < BM; r += 4)
#pragma unroll
        for (int q = 0; q < BN / 32; ++q)
          a.out[(size_t(chunk_blk(q)) * a.Np + size_t(rtile) * BM + r) * WB + chunk_c0(q) + lane] = tile[r * (BN + 1) + lane + 32 * q];
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// ---- persistent apply kernel ----------------------------------------------------------------------------
// P <- P Q for all pairs of a round: K = 128 only, so a CTA per output tile (wide_tc_kernel<APPLY>) spends its
// life in prologue and epilogue (measured: 8.4 us per 128 x 128 tile, 1.6 us of which are MMAs).  Here a CTA
// walks a contiguous range of (pair, row tile) work items: Q^T (hi and lo, 128 KB) stays in shared memory while
// the pair does not change, the rows of the factor stream through a 4-deep ring of k-blocks, two TMEM
// accumulators alternate between the MMA warp and four epilogue warps that store straight from registers
// (every thread owns one output row: 4 x 128 contiguous bytes).  The A operand (rows of P, hi and lo) is handed to
// the tensor core through TENSOR MEMORY by the converter warps, so the MMAs read only Q^T from shared memory.
//   warp 0: TMA producer   warp 1: MMA issuer   warps 2-5: converters (lo tiles)   warps 6-9: epilogue
constexpr int AP_STAGES = 4;
constexpr int AP_THREADS = 320;
constexpr int AP_Q_BYTES = 8 * tc::TILE_BYTES;  // 4 k-blocks x (hi | lo)
constexpr int AP_STAGE_BYTES = tc::TILE_BYTES;  // one k-block of P, raw (hi and lo of it go to tensor memory)
constexpr int AP_TMEM_COLS = 512;               // two accumulators (2 x 128) + AP_STAGES x (32 hi + 32 lo) columns of P
constexpr size_t AP_SMEM = size_t(AP_Q_BYTES) + AP_STAGES * AP_STAGE_BYTES + 1024 + 256;

__global__ void __launch_bounds__(AP_THREADS, 1)
wide_apply_kernel(const __grid_constant__ CUtensorMap mapL, const __grid_constant__ CUtensorMap mapQ, WideArgs a) {
  using namespace tc;
  extern __shared__ unsigned char smem_raw[];
  pdl_enter();
  const int prob = blockIdx.y;
  if (*reinterpret_cast<const volatile int*>(&a.sc[prob].converged)) return;
  // items are (pair, 128-row tile) with the pair index of the whole round; this launch owns pairs_local of them
  const int row_tiles = a.Np / BM, first = a.pair0 * row_tiles, total = first + a.pairs_local * row_tiles;
  const int per = (a.pairs_local * row_tiles + int(gridDim.x) - 1) / int(gridDim.x);
  const int it0 = first + blockIdx.x * per, it1 = min(total, it0 + per);
  const int* flag = a.flag + prob * a.pairs;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t qbase = base, abase = base + AP_Q_BYTES;
  const uint32_t bars = abase + AP_STAGES * AP_STAGE_BYTES;
  auto bar_a_tma = [&](int s) { return bars + 8u * s; };
  auto bar_a_conv = [&](int s) { return bars + 8u * (AP_STAGES + s); };
  auto bar_a_empty = [&](int s) { return bars + 8u * (2 * AP_STAGES + s); };
  const uint32_t bar_q_tma = bars + 8u * (3 * AP_STAGES), bar_q_conv = bar_q_tma + 8, bar_q_empty = bar_q_tma + 16;
  auto bar_acc_full = [&](int b) { return bars + 8u * (3 * AP_STAGES + 3 + b); };
  auto bar_acc_empty = [&](int b) { return bars + 8u * (3 * AP_STAGES + 5 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + AP_Q_BYTES + AP_STAGES * AP_STAGE_BYTES + 8 * (3 * AP_STAGES + 7));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < AP_STAGES; ++s) {
      mbar_init(bar_a_tma(s), 1);
      mbar_init(bar_a_conv(s), 128);
      mbar_init(bar_a_empty(s), 1);
    }
    mbar_init(bar_q_tma, 1);
    mbar_init(bar_q_conv, 128);
    mbar_init(bar_q_empty, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc_full(b), 1);
      mbar_init(bar_acc_empty(b), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(AP_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  auto p_cols = [&](int s) { return tmem_base + uint32_t(TMEM_COLS + 64 * s); };  // hi at +0, lo at +32

  auto q_hi = [&](int j) { return qbase + uint32_t(j) * 2 * TILE_BYTES; };
  auto a_raw = [&](int s) { return abase + uint32_t(s) * AP_STAGE_BYTES; };
  auto blk_of = [&](int pair, int chunk) {  // wide block (third TMA coordinate) of the chunk-th 32 panel columns
    int wa, wb;
    wide_blocks(a.nbw, a.round, pair, wa, wb);
    return prob * a.nbw + (chunk < 2 ? wa : wb);
  };

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int kbc = 0, qgen = 0, last_pair = -1;
      for (int it = it0; it < it1; ++it) {
        const int pair = it / row_tiles, rtile = it % row_tiles;
        if (!flag[pair]) continue;
        if (pair != last_pair) {
          if (qgen > 0) mbar_wait(bar_q_empty, (qgen - 1) & 1);  // every MMA that read the old Q has completed
          mbar_expect_tx(bar_q_tma, 4 * TILE_BYTES);
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_3d(q_hi(j), &mapQ, bar_q_tma, j * BK, 0, prob * a.pairs + pair);
          ++qgen;
          last_pair = pair;
        }
        for (int j = 0; j < 4; ++j, ++kbc) {
          const int s = kbc % AP_STAGES, use = kbc / AP_STAGES;
          if (use > 0) mbar_wait(bar_a_empty(s), (use - 1) & 1);
          mbar_expect_tx(bar_a_tma(s), TILE_BYTES);
          tma_load_3d(a_raw(s), &mapL, bar_a_tma(s), (j & 1) * 32, rtile * BM, blk_of(pair, j));
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int kbc = 0, qgen = 0, last_pair = -1, n = 0;
      for (int it = it0; it < it1; ++it) {
        const int pair = it / row_tiles;
        if (!flag[pair]) continue;
        if (pair != last_pair) {
          if (last_pair >= 0) umma_commit(bar_q_empty);  // fires when every MMA issued so far has completed
          mbar_wait(bar_q_tma, qgen & 1);
          mbar_wait(bar_q_conv, qgen & 1);
          tcgen05_fence_after();
          ++qgen;
          last_pair = pair;
        }
        const int buf = n & 1;
        if (n >= 2) {
          mbar_wait(bar_acc_empty(buf), ((n >> 1) - 1) & 1);
          tcgen05_fence_after();
        }
        const uint32_t acc = tmem_base + uint32_t(buf * BN);
        for (int j = 0; j < 4; ++j, ++kbc) {
          const int s = kbc % AP_STAGES, use = kbc / AP_STAGES;
          const uint32_t qh = q_hi(j), ql = qh + TILE_BYTES;
          mbar_wait(bar_a_conv(s), use & 1);  // hi and lo of the k-block are in tensor memory
          tcgen05_fence_after();
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint32_t p_hi = p_cols(s) + 8 * k, p_lo = p_hi + 32;
            umma_tf32_ts(acc, p_hi, make_desc(qh + 32 * k), kIdesc, !(j == 0 && k == 0));
            umma_tf32_ts(acc, p_hi, make_desc(ql + 32 * k), kIdesc, 1);
            umma_tf32_ts(acc, p_lo, make_desc(qh + 32 * k), kIdesc, 1);
          }
          umma_commit(bar_a_empty(s));  // frees the stage: shared memory and the columns in tensor memory
        }
        umma_commit(bar_acc_full(buf));
        ++n;
      }
    }
  } else if (warp < 6) {
    // ===== converters: lo = rna_tf32(x - trunc_tf32(x)) of Q^T (once per pair) and of every k-block of P =====
    const int ct = threadIdx.x - 64;  // 0..127
    auto convert = [&](unsigned char* src, unsigned char* dst, int n_vec) {
#pragma unroll 4
      for (int v = ct; v < n_vec; v += 128) {
        const float4 x = *reinterpret_cast<const float4*>(src + size_t(v) * 16);
        const float e[4] = {x.x, x.y, x.z, x.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float hi = __uint_as_float(__float_as_uint(e[j]) & 0xFFFFE000u);
          o[j] = __uint_as_float(__float_as_uint(e[j] - hi) + 0x1000u);
        }
        *reinterpret_cast<float4*>(dst + size_t(v) * 16) = make_float4(o[0], o[1], o[2], o[3]);
      }
    };
    int kbc = 0, qgen = 0, last_pair = -1;
    for (int it = it0; it < it1; ++it) {
      const int pair = it / row_tiles;
      if (!flag[pair]) continue;
      if (pair != last_pair) {
        mbar_wait(bar_q_tma, qgen & 1);
        for (int j = 0; j < 4; ++j)
          convert(base_ptr + size_t(j) * 2 * TILE_BYTES, base_ptr + size_t(j) * 2 * TILE_BYTES + TILE_BYTES, TILE_BYTES / 16);
        fence_proxy_async();
        mbar_arrive(bar_q_conv);
        ++qgen;
        last_pair = pair;
      }
      for (int j = 0; j < 4; ++j, ++kbc) {
        const int s = kbc % AP_STAGES, use = kbc / AP_STAGES;
        const unsigned char* st = base_ptr + AP_Q_BYTES + size_t(s) * AP_STAGE_BYTES;
        mbar_wait(bar_a_tma(s), use & 1);
        // this thread's row of the k-block (128-byte swizzle: chunk c of row m sits at chunk c ^ (m & 7)) -> TMEM
        const int lane_grp = warp & 3, m = lane_grp * 32 + lane;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 x = *reinterpret_cast<const float4*>(st + size_t(m) * 128 + size_t((c ^ (m & 7)) * 16));
          const float e[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t bits = __float_as_uint(e[q]);
            hi[4 * c + q] = bits;
            lo[4 * c + q] = __float_as_uint(e[q] - __uint_as_float(bits & 0xFFFFE000u)) + 0x1000u;
          }
        }
        const uint32_t ta = p_cols(s) + (uint32_t(lane_grp * 32) << 16);
        tmem_st32(ta, hi);
        tmem_st32(ta + 32, lo);
        tmem_st_wait();
        tcgen05_fence_before();
        mbar_arrive(bar_a_conv(s));
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global (thread = one output row of the tile) =====
    const int lane_grp = warp & 3;  // a warp may only touch TMEM lanes 32 * (warp % 4) .. + 31
    const int row = lane_grp * 32 + lane;
    int n = 0;
    for (int it = it0; it < it1; ++it) {
      const int pair = it / row_tiles, rtile = it % row_tiles;
      if (!flag[pair]) continue;
      const int buf = n & 1;
      mbar_wait(bar_acc_full(buf), (n >> 1) & 1);
      tcgen05_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_base + (uint32_t(lane_grp * 32) << 16) + uint32_t(buf * BN + 32 * c), r);
        float* dst = a.out + (size_t(blk_of(pair, c)) * a.Np + size_t(rtile) * BM + row) * WB + (c & 1) * 32;
#pragma unroll
        for (int q = 0; q < 4; ++q)  // 256-bit stores: every thread writes whole 32-byte sectors of its row
          asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 8 * q), "r"(r[8 * q]),
                       "r"(r[8 * q + 1]), "r"(r[8 * q + 2]), "r"(r[8 * q + 3]), "r"(r[8 * q + 4]), "r"(r[8 * q + 5]),
                       "r"(r[8 * q + 6]), "r"(r[8 * q + 7])
                       : "memory");
      }
      tcgen05_fence_before();
      mbar_arrive(bar_acc_empty(buf));
      ++n;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(AP_TMEM_COLS) : "memory");
  }
}

// ---- rotation kernel ----------------------------------------------------------------------------------
template <int BAR>
__device__ __forceinline__ void group_sync() {  // the 256 threads of one group
  asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(OT) : "memory");
}
__device__ __forceinline__ void group_sync_dyn(int bar) { asm volatile("bar.sync %0, %1;" ::"r"(bar), "n"(OT) : "memory"); }

struct WideRotSmem {
  float H[WP][LDW];
  float Q[WP][LDW];
  RotSmem<float> rs[NGRP];
  int sub_a[NGRP], sub_b[NGRP], rot[NGRP];
  int warp_need[RT / 32];
  int any;
};

// sub-block pair of group g in inner round ir (0..3) and whether it is an intra round
__device__ __forceinline__ void sub_pair(bool intra_round, int ir, int g, int& sa, int& sb, bool& intra) {
  intra = false;
  if (!intra_round) {  // every sub-block of a against every sub-block of b
    sa = g;
    sb = NSUB / 2 + ((g + ir) & 3);
  } else if (ir < 3) {  // tournaments inside a (groups 0, 1) and inside b (groups 2, 3)
    int x, y;
    rr_pair(NSUB / 2, ir, g & 1, x, y);
    sa = (g >> 1) * (NSUB / 2) + min(x, y);
    sb = (g >> 1) * (NSUB / 2) + max(x, y);
  } else {  // column pairs inside each sub-block
    sa = 2 * g;
    sb = 2 * g + 1;
    intra = true;
  }
}

__device__ __forceinline__ int panel_index(int sa, int sb, int k) {  // column of the panel for index k of a sub-pair
  return k < OB ? sa * OB + k : sb * OB + (k - OB);
}

// rotation rounds of one group on its 32 x 32 sub-Gram (same arithmetic as the 16-wide kernel; barriers are
// the group's named barrier instead of __syncthreads)
template <bool intra>
__device__ __forceinline__ void group_rotation_rounds(RotSmem<float>& rs, float tol2, float abs2, int gt, int bar) {
  const int a = gt >> 4, b = gt & 15;
  constexpr int n_rounds = intra ? OB - 1 : OB;
  int cur = 0;
  for (int r = 0; r < n_rounds; ++r, cur ^= 1) {
    float(*Hc)[LDH] = rs.H[cur];
    float(*Hn)[LDH] = rs.H[cur ^ 1];
    int pa, qa, pb, qb;
    inner_pair(intra, r, a, pa, qa);
    inner_pair(intra, r, b, pb, qb);
    const float dpp = Hc[pb][pb], dqq = Hc[qb][qb], dpq = Hc[pb][qb];
    const float x00 = Hc[pa][pb], x01 = Hc[pa][qb], x10 = Hc[qa][pb], x11 = Hc[qa][qb];
    float* q0 = &rs.Q[b][0];
    float* q1 = &rs.Q[b + OB][0];
    const float q0p = q0[pa], q0q = q0[qa], q1p = q1[pa], q1q = q1[qa];
    float cb, sb;
    const bool db = make_rotation(dpp, dqq, dpq, tol2, abs2, cb, sb);
    const float ca = __shfl_sync(0xffffffffu, cb, a), sa = __shfl_sync(0xffffffffu, sb, a);
    const float t00 = cb * x00 - sb * x01, t01 = sb * x00 + cb * x01;
    const float t10 = cb * x10 - sb * x11, t11 = sb * x10 + cb * x11;
    float h00 = ca * t00 - sa * t10, h01 = ca * t01 - sa * t11;
    float h10 = sa * t00 + ca * t10, h11 = sa * t01 + ca * t11;
    if (a == b && db) h01 = h10 = 0.f;
    Hn[pa][pb] = h00;
    Hn[pa][qb] = h01;
    Hn[qa][pb] = h10;
    Hn[qa][qb] = h11;
    q0[pa] = ca * q0p - sa * q0q;
    q0[qa] = sa * q0p + ca * q0q;
    q1[pa] = ca * q1p - sa * q1q;
    q1[qa] = sa * q1p + ca * q1q;
    group_sync_dyn(bar);
  }
}

// The two update phases are separate NON-INLINED functions: the kernel runs 1024 threads (64 registers each)
// and keeps a dozen values live across its inner rounds; inlined, the 32 accumulators of a phase ended up in
// local memory (measured: 570 us per launch instead of ~40).  Inside a call they have the register file to
// themselves.
//
// Columns {sa, sb} of H and of Q:  X[:, cols] <- X[:, cols] Q_sub.  Per sub-pair 2 matrices x 128 rows x 32
// columns of output; a thread owns 4 rows (lane, lane + 32, ...) x 8 columns, a warp one (matrix, column group):
// the Q_sub reads are broadcasts, the X reads conflict-free.  The products stay in registers until every thread
// has read its inputs (barrier inside).
__device__ __noinline__ void update_columns(WideRotSmem& sm, int tid) {
  const int ug = tid >> 8, u = tid & 255, v = u & 127, cg = v >> 5, ln = v & 31;
  const bool on = sm.rot[ug] != 0;
  float(*X)[LDW] = (u >> 7) ? sm.Q : sm.H;
  const int sa_u = sm.sub_a[ug], sb_u = sm.sub_b[ug];
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  if (on) {
    const float* q = &sm.rs[ug].Q[0][8 * cg];
#pragma unroll 2
    for (int k = 0; k < OP; ++k) {
      const int pc = panel_index(sa_u, sb_u, k);
      const float4 q0 = *reinterpret_cast<const float4*>(q + k * LDQ);
      const float4 q1 = *reinterpret_cast<const float4*>(q + k * LDQ + 4);
      const float2 qa = make_float2(q0.x, q0.y), qb = make_float2(q0.z, q0.w);
      const float2 qc = make_float2(q1.x, q1.y), qd = make_float2(q1.z, q1.w);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float x = X[ln + 32 * i][pc];
        const float2 xx = make_float2(x, x);
        const float2 r0 = __ffma2_rn(xx, qa, make_float2(acc[i][0], acc[i][1]));
        const float2 r1 = __ffma2_rn(xx, qb, make_float2(acc[i][2], acc[i][3]));
        const float2 r2 = __ffma2_rn(xx, qc, make_float2(acc[i][4], acc[i][5]));
        const float2 r3 = __ffma2_rn(xx, qd, make_float2(acc[i][6], acc[i][7]));
        acc[i][0] = r0.x, acc[i][1] = r0.y, acc[i][2] = r1.x, acc[i][3] = r1.y;
        acc[i][4] = r2.x, acc[i][5] = r2.y, acc[i][6] = r3.x, acc[i][7] = r3.y;
      }
    }
  }
  __syncthreads();
  if (on) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int pc = panel_index(sa_u, sb_u, 8 * cg + j);
#pragma unroll
      for (int i = 0; i < 4; ++i) X[ln + 32 * i][pc] = acc[i][j];
    }
  }
}

// Rows {sa, sb} of H:  H[rows, :] <- Q_sub^T H[rows, :].  Per sub-pair 32 rows x 128 columns of output; threads
// 0..127 of the group own 8 rows x 4 columns (lane, lane + 32, ...), a warp one row group.
__device__ __noinline__ void update_rows(WideRotSmem& sm, int tid) {
  const int ug = tid >> 8, u = tid & 255, pg = (u >> 5) & 3, ln = u & 31;
  const bool on = sm.rot[ug] != 0 && u < 128;
  const int sa_u = sm.sub_a[ug], sb_u = sm.sub_b[ug];
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  if (on) {
    const float* q = &sm.rs[ug].Q[0][8 * pg];
#pragma unroll 2
    for (int k = 0; k < OP; ++k) {
      const int pr = panel_index(sa_u, sb_u, k);
      const float4 q0 = *reinterpret_cast<const float4*>(q + k * LDQ);
      const float4 q1 = *reinterpret_cast<const float4*>(q + k * LDQ + 4);
      const float qv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
      const float2 h01 = make_float2(sm.H[pr][ln], sm.H[pr][ln + 32]);
      const float2 h23 = make_float2(sm.H[pr][ln + 64], sm.H[pr][ln + 96]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 qq = make_float2(qv[i], qv[i]);
        const float2 r0 = __ffma2_rn(qq, h01, make_float2(acc[i][0], acc[i][1]));
        const float2 r1 = __ffma2_rn(qq, h23, make_float2(acc[i][2], acc[i][3]));
        acc[i][0] = r0.x, acc[i][1] = r0.y, acc[i][2] = r1.x, acc[i][3] = r1.y;
      }
    }
  }
  __syncthreads();  // every read of the old rows precedes the first write
  if (on) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int pr = panel_index(sa_u, sb_u, 8 * pg + i);
#pragma unroll
      for (int j = 0; j < 4; ++j) sm.H[pr][ln + 32 * j] = acc[i][j];
    }
  }
}

__global__ void __launch_bounds__(RT, 1)
wide_rot_kernel(float* Qt, int* flag, const float* part, int nbw, int round, int pairs, int splits, JacobiScalars* sc,
                int pair0) {
  extern __shared__ __align__(16) unsigned char wide_smem[];
  WideRotSmem& sm = *reinterpret_cast<WideRotSmem*>(wide_smem);
  const int pair = pair0 + int(blockIdx.x), prob = blockIdx.y, tid = threadIdx.x;
  sc += prob;
  if (*reinterpret_cast<const volatile int*>(&sc->converged)) return;
  const bool intra_round = round < 0;
  (void)nbw;
  const float tol2 = Eps<float>::tol * Eps<float>::tol, abs2 = Eps<float>::v * Eps<float>::v;
  // H = sum of the split-K partial sums in a fixed order, Q = I
  {
    const float* p0 = part + (size_t(prob) * splits * pairs + pair) * size_t(WP * WP);
    const size_t split_stride = size_t(pairs) * WP * WP;
    for (int idx = tid; idx < WP * WP / 4; idx += RT) {
      float4 s4 = *reinterpret_cast<const float4*>(p0 + size_t(idx) * 4);
      for (int k = 1; k < splits; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(p0 + k * split_stride + size_t(idx) * 4);
        s4.x += v.x, s4.y += v.y, s4.z += v.z, s4.w += v.w;
      }
      const int r = (idx * 4) / WP, c = (idx * 4) % WP;
      sm.H[r][c] = s4.x, sm.H[r][c + 1] = s4.y, sm.H[r][c + 2] = s4.z, sm.H[r][c + 3] = s4.w;
      sm.Q[r][c] = r == c ? 1.f : 0.f;
      sm.Q[r][c + 1] = r == c + 1 ? 1.f : 0.f;
      sm.Q[r][c + 2] = r == c + 2 ? 1.f : 0.f;
      sm.Q[r][c + 3] = r == c + 3 ? 1.f : 0.f;
    }
    if (tid == 0) sm.any = 0;
  }
  __syncthreads();
  for (int idx = tid; idx < WP * WP; idx += RT) {  // the two triangles differ in the last bits
    const int r = idx / WP, c = idx % WP;
    if (r < c) {
      const float m = 0.5f * (sm.H[r][c] + sm.H[c][r]);
      sm.H[r][c] = m;
      sm.H[c][r] = m;
    }
  }
  __syncthreads();

  const int g = tid >> 8, gt = tid & (OT - 1), bar = 1 + g;  // named barriers 1..4 (0 = __syncthreads)
  RotSmem<float>& rs = sm.rs[g];
  for (int ir = 0; ir < 4; ++ir) {
    int sa, sb;
    bool intra;
    sub_pair(intra_round, ir, g, sa, sb, intra);
    // ---- this group's 32 x 32 sub-Gram, Q_sub = I
    {
      const int i = gt >> 3, j = (gt & 7) * 4;
      const int pi = panel_index(sa, sb, i);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        rs.H[0][i][j + e] = sm.H[pi][panel_index(sa, sb, j + e)];
        rs.Q[i][j + e] = (i == j + e) ? 1.f : 0.f;
      }
    }
    group_sync_dyn(bar);
    // does any column pair of this visit still need a rotation?  (uniform over the group)
    {
      const int i = gt >> 4, j = gt & 15;
      bool need;
      if (!intra) {
        need = needs_rotation(rs.H[0][i][i], rs.H[0][OB + j][OB + j], rs.H[0][i][OB + j], tol2, abs2);
      } else {
        need = i < j && (needs_rotation(rs.H[0][i][i], rs.H[0][j][j], rs.H[0][i][j], tol2, abs2) ||
                         needs_rotation(rs.H[0][OB + i][OB + i], rs.H[0][OB + j][OB + j], rs.H[0][OB + i][OB + j],
                                        tol2, abs2));
      }
      const unsigned w = __ballot_sync(0xffffffffu, need);
      if ((tid & 31) == 0) sm.warp_need[tid >> 5] = w != 0;
    }
    group_sync_dyn(bar);
    bool rotate = false;
#pragma unroll
    for (int w = 0; w < OT / 32; ++w) rotate |= sm.warp_need[g * (OT / 32) + w] != 0;
    if (rotate) {
      if (intra) group_rotation_rounds<true>(rs, tol2, abs2, gt, bar);
      else group_rotation_rounds<false>(rs, tol2, abs2, gt, bar);
    }
    if (gt == 0) {
      sm.sub_a[g] = sa, sm.sub_b[g] = sb, sm.rot[g] = rotate;
      if (rotate) {
        atomicAdd(&sc->rotations, 1ull);
        sm.any = 1;
      }
    }
    __syncthreads();
    update_columns(sm, tid);
    __syncthreads();
    update_rows(sm, tid);
    __syncthreads();
  }
  // Q^T to global memory (the apply kernel's B operand: rows = output columns, K-contiguous)
  const int any = sm.any;
  if (tid == 0) flag[prob * pairs + pair] = any;
  if (any) {
    float* dst = Qt + (size_t(prob) * pairs + pair) * size_t(WP * WP);
    for (int idx = tid; idx < WP * WP; idx += RT) {
      const int n = idx / WP, k = idx % WP;
      dst[idx] = sm.Q[k][n];
    }
  }
}

// ---- rotation kernel, one CLUSTER of four CTAs per pair -------------------------------------------------------
// wide_rot_kernel above keeps the whole 128 x 128 problem in one SM: measured 106 us per launch, 70 % of that
// SM's shared-memory bandwidth (the two-sided update of H and the update of Q), 40 of 148 SMs busy at R = 5120.
// Here the four sub-block pairs that rotate concurrently belong to four CTAs of a cluster:
//   * CTA g owns the 32 ROWS of H of its current sub-pair (all 128 columns) and, for good, rows 32 g .. 32 g + 31
//     of the accumulated Q.  Its 32 x 32 sub-Gram is a slice of its own rows; the scalar rotation rounds are the
//     16-wide kernel's (rotation_rounds, the CTA has its SM's barrier to itself);
//   * after the rotations the four Q_sub travel through distributed shared memory (one cluster barrier); every
//     CTA applies all four to the COLUMNS of its rows of H and of Q (rows are independent) and its own Q_sub to
//     its ROWS of H -- a quarter of the update work per SM;
//   * between inner rounds the sub-blocks re-pair: a CTA fetches the 2 x 16 rows of its next sub-pair from the
//     two CTAs that hold them (double-buffered: one more cluster barrier per inner round).
constexpr int CR = 4;

struct RotCta {
  float Hrows[2][OP][LDW];     // rows of H of the current sub-pair (16 of sub-block sa, 16 of sb), double-buffered
  float Qown[OP][LDW];         // rows 32 g .. 32 g + 31 of the accumulated Q
  RotSmem<float> rs;           // rotation rounds: H_sub (double-buffered), Q_sub
  float Qsubs[CR][OP][LDQ];    // the four Q_sub of the inner round
  int my_rot;                  // did my sub-pair rotate?  (read by the other CTAs)
  int rot[CR];
};

__device__ __forceinline__ void holder_of(bool intra_round, int ir, int x, int& h, int& half) {  // who holds sub-block x?
  h = 0, half = 0;
#pragma unroll
  for (int g = 0; g < CR; ++g) {
    int sa, sb;
    bool intra;
    sub_pair(intra_round, ir, g, sa, sb, intra);
    if (sa == x) h = g, half = 0;
    if (sb == x) h = g, half = 1;
  }
}

__device__ long long g_wdbg[16];

__global__ void __launch_bounds__(OT, 2)
wide_rot_cluster_kernel(float* Qt, int* flag, const float* part, int nbw, int round, int pairs, int splits,
                        JacobiScalars* sc, float* diag, int cross_only, int pair0) {
  extern __shared__ __align__(16) unsigned char wide_smem[];
  RotCta& sm = *reinterpret_cast<RotCta*>(wide_smem);
  cg::cluster_group cluster = cg::this_cluster();
  pdl_enter();
  const int g = int(cluster.block_rank()), pair = pair0 + int(blockIdx.x) / CR, prob = blockIdx.y, tid = threadIdx.x;
  sc += prob;
  if (*reinterpret_cast<const volatile int*>(&sc->converged)) return;  // uniform over the cluster
  const bool intra_round = round < 0;
  // Diagonal-block cache (cross_only): diag[problem][wide block][64][64] holds P_w^T P_w of every wide block, written
  // here after every rotation (the diagonal blocks of Q^T H Q) and recomputed from the factor by the intra round of
  // every sweep.  In cross rounds the Gram kernel then forms only H[:, b]; H_aa comes from the cache, H_ba = H_ab^T.
  const bool cached = cross_only != 0 && !intra_round;
  int wa, wb;
  wide_blocks(nbw, round, pair, wa, wb);
  float* const diag_a = diag + (size_t(prob) * nbw + wa) * size_t(WB * WB);
  float* const diag_b = diag + (size_t(prob) * nbw + wb) * size_t(WB * WB);
  const float tol2 = Eps<float>::tol * Eps<float>::tol, abs2 = Eps<float>::v * Eps<float>::v;
  const int warp = tid >> 5, lane = tid & 31;
  // VVT_SYEVJ_DEBUG: cycles of CTA 0 per phase, summed over the inner rounds (g_wdbg, read by the host)
  const bool stamp = blockIdx.x == 0 && blockIdx.y == 0 && tid == 0;
  long long t_last = stamp ? clock64() : 0, t_acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  auto lap = [&](int i) {
    if (stamp) {
      const long long now = clock64();
      t_acc[i] += now - t_last;
      t_last = now;
    }
  };

  // ---- my rows of H (sum of the split-K partial sums in a fixed order, symmetrised), my rows of Q = I
  {
    int sa, sb;
    bool intra;
    sub_pair(intra_round, 0, g, sa, sb, intra);
    const float* p0 = part + (size_t(prob) * splits * pairs + pair) * size_t(WP * WP);
    const size_t split_stride = size_t(pairs) * WP * WP;
    // (rows only, coalesced: the tensor core forms both triangles of P^T P from the same operands, they differ in
    // the last bits at most -- the order of the hi*lo / lo*hi terms -- and the rotation angles do not care;
    // reading the mirrored entries as well, column-wise, cost 18 % of this kernel)
    if (!cached) {
      // (a thread's four vectors of one split are loaded before any of them is added: one L2 round trip per split
      // instead of four)
      constexpr int NV = OP * WP / 4 / OT;  // 4
      float4 acc4[NV];
      size_t off[NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int idx = tid + j * OT;
        off[j] = size_t(panel_index(sa, sb, (idx * 4) / WP)) * WP + (idx * 4) % WP;
        acc4[j] = *reinterpret_cast<const float4*>(p0 + off[j]);
      }
      for (int k = 1; k < splits; ++k) {
        float4 v[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) v[j] = *reinterpret_cast<const float4*>(p0 + k * split_stride + off[j]);
#pragma unroll
        for (int j = 0; j < NV; ++j) acc4[j].x += v[j].x, acc4[j].y += v[j].y, acc4[j].z += v[j].z, acc4[j].w += v[j].w;
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int idx = tid + j * OT;
        const int i = (idx * 4) / WP, c = (idx * 4) % WP;
        const float4 a = acc4[j];
        sm.Hrows[0][i][c] = a.x, sm.Hrows[0][i][c + 1] = a.y, sm.Hrows[0][i][c + 2] = a.z, sm.Hrows[0][i][c + 3] = a.w;
      }
    } else {
      // cross round (sa in a, sb in b): columns of b for all 32 rows from the partial sums (two vectors per thread);
      // columns of a: my 16 rows of a from the cache, my 16 rows of b as the transposed entries H[c][r] of H_ab
      float4 right[2], left;
      size_t off_r[2], off_l;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int idx = tid + j * OT;  // 32 rows x 16 vectors
        off_r[j] = size_t(panel_index(sa, sb, idx / 16)) * WP + WB + (idx % 16) * 4;
        right[j] = *reinterpret_cast<const float4*>(p0 + off_r[j]);
      }
      // threads 0 .. 127: 16 rows x 8 pieces (8 floats) of the cache; threads 128 .. 255: 64 columns x 2 pieces of H_ab
      const bool from_cache = tid < OT / 2;
      const int t2 = tid - OT / 2;
      off_l = from_cache ? size_t((sa * OB + tid / 8) * WB + (tid % 8) * 8)
                         : size_t(t2 / 2) * WP + size_t(sb * OB + (t2 % 2) * 8);
      float4 left2;
      if (from_cache) {
        left = *reinterpret_cast<const float4*>(diag_a + off_l);
        left2 = *reinterpret_cast<const float4*>(diag_a + off_l + 4);
      } else {
        left = *reinterpret_cast<const float4*>(p0 + off_l);
        left2 = *reinterpret_cast<const float4*>(p0 + off_l + 4);
      }
      for (int k = 1; k < splits; ++k) {
        float4 v[2], w = make_float4(0.f, 0.f, 0.f, 0.f), w2 = w;
#pragma unroll
        for (int j = 0; j < 2; ++j) v[j] = *reinterpret_cast<const float4*>(p0 + k * split_stride + off_r[j]);
        if (!from_cache) {
          w = *reinterpret_cast<const float4*>(p0 + k * split_stride + off_l);
          w2 = *reinterpret_cast<const float4*>(p0 + k * split_stride + off_l + 4);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) right[j].x += v[j].x, right[j].y += v[j].y, right[j].z += v[j].z, right[j].w += v[j].w;
        left.x += w.x, left.y += w.y, left.z += w.z, left.w += w.w;
        left2.x += w2.x, left2.y += w2.y, left2.z += w2.z, left2.w += w2.w;
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int idx = tid + j * OT;
        const int i = idx / 16, c = WB + (idx % 16) * 4;
        sm.Hrows[0][i][c] = right[j].x, sm.Hrows[0][i][c + 1] = right[j].y;
        sm.Hrows[0][i][c + 2] = right[j].z, sm.Hrows[0][i][c + 3] = right[j].w;
      }
      if (from_cache) {
        const int i = tid / 8, c = (tid % 8) * 8;
        sm.Hrows[0][i][c] = left.x, sm.Hrows[0][i][c + 1] = left.y, sm.Hrows[0][i][c + 2] = left.z, sm.Hrows[0][i][c + 3] = left.w;
        sm.Hrows[0][i][c + 4] = left2.x, sm.Hrows[0][i][c + 5] = left2.y;
        sm.Hrows[0][i][c + 6] = left2.z, sm.Hrows[0][i][c + 7] = left2.w;
      } else {  // column c of a, eight of my rows of b
        const int c = t2 / 2, i0 = OB + (t2 % 2) * 8;
        sm.Hrows[0][i0][c] = left.x, sm.Hrows[0][i0 + 1][c] = left.y, sm.Hrows[0][i0 + 2][c] = left.z, sm.Hrows[0][i0 + 3][c] = left.w;
        sm.Hrows[0][i0 + 4][c] = left2.x, sm.Hrows[0][i0 + 5][c] = left2.y;
        sm.Hrows[0][i0 + 6][c] = left2.z, sm.Hrows[0][i0 + 7][c] = left2.w;
      }
    }
    for (int idx = tid; idx < OP * WP; idx += OT) sm.Qown[idx / WP][idx % WP] = (32 * g + idx / WP == idx % WP) ? 1.f : 0.f;
  }
  int any = 0, cur = 0;
  __syncthreads();
  lap(0);

  for (int ir = 0; ir < 4; ++ir, cur ^= 1) {
    int sa, sb;
    bool intra;
    sub_pair(intra_round, ir, g, sa, sb, intra);
    float(*H)[LDW] = sm.Hrows[cur];
    // ---- 32 x 32 sub-Gram = my rows at the columns of my sub-pair; Q_sub = I
    {
      const int i = tid >> 3, j = (tid & 7) * 4;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        sm.rs.H[0][i][j + e] = H[i][panel_index(sa, sb, j + e)];
        sm.rs.Q[i][j + e] = (i == j + e) ? 1.f : 0.f;
      }
    }
    __syncthreads();
    const bool rotate = any_rotation_needed<float>(sm.rs, intra, tol2, abs2, tid);
    if (rotate) {
      if (intra) rotation_rounds<float, true>(sm.rs, tol2, abs2, tid);
      else rotation_rounds<float, false>(sm.rs, tol2, abs2, tid);
    }
    if (tid == 0) {
      sm.my_rot = rotate;
      if (rotate) atomicAdd(&sc->rotations, 1ull);
    }
    // ---- the four Q_sub of the round, through distributed shared memory
    lap(1);
    cluster_arrive();
    cluster_wait();
    lap(2);
    {
      // (every remote read is issued before the first local write: the compiler cannot prove that the two do not
      // alias and would otherwise wait out one distributed-shared-memory round trip per vector)
      constexpr int QV = OP * LDQ / 4;             // float4 per Q_sub (288)
      constexpr int QN = (CR * QV + OT - 1) / OT;  // per thread (5)
      float4 v[QN];
#pragma unroll
      for (int j = 0; j < QN; ++j) {
        const int idx = tid + j * OT;
        if (idx < CR * QV) {
          const RotCta* remote = cluster.map_shared_rank(&sm, idx / QV);
          v[j] = reinterpret_cast<const float4*>(&remote->rs.Q[0][0])[idx % QV];
        }
      }
      int rflag = 0;
      if (tid < CR) rflag = cluster.map_shared_rank(&sm, tid)->my_rot;
#pragma unroll
      for (int j = 0; j < QN; ++j) {
        const int idx = tid + j * OT;
        if (idx < CR * QV) reinterpret_cast<float4*>(&sm.Qsubs[idx / QV][0][0])[idx % QV] = v[j];
      }
      if (tid < CR) sm.rot[tid] = rflag;
    }
    __syncthreads();
    lap(3);
#pragma unroll
    for (int u = 0; u < CR; ++u) any |= sm.rot[u];
    // ---- columns: X[:, cols(u)] <- X[:, cols(u)] Q_sub(u) for my rows of H and of Q.  A warp owns one
    //      (sub-pair, matrix): 32 rows x 32 columns, a thread 4 rows x 8 columns of it.
    {
      const int u = warp >> 1, rg = lane >> 2, cgp = lane & 3;
      float(*X)[LDW] = (warp & 1) ? sm.Qown : H;
      int ua, ub;
      bool dummy;
      sub_pair(intra_round, ir, u, ua, ub, dummy);
      if (sm.rot[u]) {  // warp-uniform
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        const float* q = &sm.Qsubs[u][0][8 * cgp];
#pragma unroll 2
        for (int k = 0; k < OP; ++k) {
          const int pc = panel_index(ua, ub, k);
          const float4 q0 = *reinterpret_cast<const float4*>(q + k * LDQ);
          const float4 q1 = *reinterpret_cast<const float4*>(q + k * LDQ + 4);
          const float2 qa = make_float2(q0.x, q0.y), qb = make_float2(q0.z, q0.w);
          const float2 qc = make_float2(q1.x, q1.y), qd = make_float2(q1.z, q1.w);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float x = X[4 * rg + i][pc];
            const float2 xx = make_float2(x, x);
            const float2 r0 = __ffma2_rn(xx, qa, make_float2(acc[i][0], acc[i][1]));
            const float2 r1 = __ffma2_rn(xx, qb, make_float2(acc[i][2], acc[i][3]));
            const float2 r2 = __ffma2_rn(xx, qc, make_float2(acc[i][4], acc[i][5]));
            const float2 r3 = __ffma2_rn(xx, qd, make_float2(acc[i][6], acc[i][7]));
            acc[i][0] = r0.x, acc[i][1] = r0.y, acc[i][2] = r1.x, acc[i][3] = r1.y;
            acc[i][4] = r2.x, acc[i][5] = r2.y, acc[i][6] = r3.x, acc[i][7] = r3.y;
          }
        }
        __syncwarp();  // the warp owns these 32 x 32 entries: all reads precede the writes
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int pc = panel_index(ua, ub, 8 * cgp + j);
#pragma unroll
          for (int i = 0; i < 4; ++i) X[4 * rg + i][pc] = acc[i][j];
        }
      }
    }
    __syncthreads();
    lap(4);
    // ---- rows: H[my rows, :] <- Q_sub(g)^T H[my rows, :]; a thread owns 4 rows x 4 columns
    {
      const bool on = sm.rot[g] != 0;
      float acc[4][4];
      if (on) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        const float* q = &sm.Qsubs[g][0][4 * warp];
#pragma unroll 4
        for (int k = 0; k < OP; ++k) {
          const float4 qv = *reinterpret_cast<const float4*>(q + k * LDQ);
          const float qq[4] = {qv.x, qv.y, qv.z, qv.w};
          const float2 h01 = make_float2(H[k][lane], H[k][lane + 32]);
          const float2 h23 = make_float2(H[k][lane + 64], H[k][lane + 96]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 q2 = make_float2(qq[i], qq[i]);
            const float2 r0 = __ffma2_rn(q2, h01, make_float2(acc[i][0], acc[i][1]));
            const float2 r1 = __ffma2_rn(q2, h23, make_float2(acc[i][2], acc[i][3]));
            acc[i][0] = r0.x, acc[i][1] = r0.y, acc[i][2] = r1.x, acc[i][3] = r1.y;
          }
        }
      }
      __syncthreads();  // every read of the old rows precedes the first write
      if (on) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) H[4 * warp + i][lane + 32 * j] = acc[i][j];
      }
    }
    // ---- re-pair: fetch the rows of my next sub-pair from the CTAs that hold them
    lap(5);
    cluster_arrive();
    cluster_wait();
    lap(6);
    if (ir < 3) {
      int na, nb_;
      bool dummy;
      sub_pair(intra_round, ir + 1, g, na, nb_, dummy);
      constexpr int RN = OB * WP / OT;  // 8 elements per thread and sub-block
      float v[2][RN];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {  // all remote reads first (see the Q_sub exchange)
        int h, half;
        holder_of(intra_round, ir, hh ? nb_ : na, h, half);
        const RotCta* remote = cluster.map_shared_rank(&sm, h);
#pragma unroll
        for (int j = 0; j < RN; ++j) {
          const int idx = tid + j * OT;
          v[hh][j] = remote->Hrows[cur][OB * half + idx / WP][idx % WP];
        }
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int j = 0; j < RN; ++j) {
          const int idx = tid + j * OT;
          sm.Hrows[cur ^ 1][OB * hh + idx / WP][idx % WP] = v[hh][j];
        }
      __syncthreads();
    }
    lap(7);
  }
  // ---- diagonal-block cache: the rows I hold at the end (sub-pair of the last inner round) are final rows of
  //      Q^T H Q; their entries inside their own wide block are P_w^T P_w after the apply.  (Nothing rotated and no
  //      fresh Gram: the cache is still right.)
  if (cross_only != 0 && (any || intra_round)) {
    int sa, sb;
    bool dummy;
    sub_pair(intra_round, 3, g, sa, sb, dummy);
    const float(*H)[LDW] = sm.Hrows[cur ^ 1];  // the buffer of inner round 3
    for (int idx = tid; idx < OP * WB; idx += OT) {
      const int i = idx / WB, c = idx % WB;
      const int r = panel_index(sa, sb, i);  // panel row: wide block r / 64, local row r % 64
      float* dst = r < WB ? diag_a : diag_b;
      dst[(r % WB) * WB + c] = H[i][(r / WB) * WB + c];
    }
  }
  // ---- Q^T of my rows to global memory (rows of Q^T = output columns n; my slice is k = 32 g ..)
  if (g == 0 && tid == 0) flag[prob * pairs + pair] = any;
  if (any) {
    float* dst = Qt + (size_t(prob) * pairs + pair) * size_t(WP * WP);
    for (int idx = tid; idx < OP * WP; idx += OT) {
      const int n = idx / OP, i = idx % OP;
      dst[size_t(n) * WP + 32 * g + i] = sm.Qown[i][n];
    }
  }
  lap(8);
  if (stamp)
    for (int i = 0; i < 9; ++i) g_wdbg[i] = t_acc[i];
  // nobody may exit while a neighbour can still read its shared memory: the last remote reads (Q_sub of the
  // fourth inner round) precede the cluster barrier that follows them
}

// ---- layout conversion at both ends ---------------------------------------------------------------------
// The factor is stored WIDE-BLOCK-MAJOR: Lw[problem][wide block][row][64 columns].  A panel operation touches
// two wide blocks, each one contiguous piece of memory (row-major storage made every 128-byte row segment of a
// TMA box its own DRAM page: measured 2.0 - 2.4 TB/s for the Gram and apply kernels at R = 5120).
__device__ __forceinline__ size_t wide_index(int Np, int nbw, int prob, int row, int col) {
  return ((size_t(prob) * nbw + col / WB) * Np + row) * WB + (col % WB);
}

// Lw <- lower triangle of the Cholesky factor A [R][R], scaled, zero elsewhere
__global__ void wide_init_chol_kernel(float* Lw, const float* A, int64_t R, int Np, const JacobiScalars* sc) {
  const int64_t total = int64_t(Np) * Np;
  const int prob = blockIdx.y, nbw = Np / WB;
  A += prob * R * R, sc += prob;
  const double bound = sqrt(double(R)) * sqrt(sc->norm2) + double(R) * double(chol_shift<float>(sc));
  const float scale = bound > 0.0 ? float(1.0 / sqrt(bound)) : 0.f;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    // idx enumerates the destination: (wide block, row, column in block)
    const int w = int(idx / (int64_t(Np) * WB)), rem = int(idx % (int64_t(Np) * WB));
    const int i = rem / WB, j = w * WB + rem % WB;
    Lw[size_t(prob) * total + idx] = (i < R && j <= i) ? A[int64_t(i) * R + j] * scale : 0.f;
  }
}

// row-major [Np][Np] <-> wide-block-major (test hook only)
__global__ void wide_relayout_kernel(float* dst, const float* src, int Np, int to_blocks) {
  const int64_t total = int64_t(Np) * Np;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int r = int(idx / Np), c = int(idx % Np);
    const size_t b = wide_index(Np, Np / WB, 0, r, c);
    if (to_blocks) dst[b] = src[idx];
    else dst[idx] = src[b];
  }
}

// inv[c] = 1 / ||Lw[:, c]||: a block owns 32 columns, its 8 warps split the rows (coalesced), fixed-order sum
__global__ void __launch_bounds__(256) wide_colnorm_kernel(float* inv, const float* Lw, int64_t R, int Np) {
  __shared__ double red[8][32];
  const int prob = blockIdx.y, nbw = Np / WB;
  inv += prob * R;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t c = int64_t(blockIdx.x) * 32 + lane;
  double s = 0.0;
  if (c < R)
    for (int64_t r = w; r < R; r += 8) {
      const double v = double(Lw[wide_index(Np, nbw, prob, int(r), int(c))]);
      s += v * v;
    }
  red[w][lane] = s;
  __syncthreads();
  if (w == 0 && c < R) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][lane];
    inv[c] = t > 0.0 ? float(1.0 / sqrt(t)) : 0.f;
  }
}

// Jm[r][c] = Lw[r][c] * inv[c] and its transpose Jt[c][r] (32 x 32 tiles through shared memory)
__global__ void __launch_bounds__(256) wide_gather_kernel(float* Jm, float* Jt, const float* Lw, const float* inv,
                                                          int64_t R, int Np) {
  __shared__ float tile[32][33];
  const int prob = blockIdx.z, nbw = Np / WB;
  inv += prob * R, Jm += prob * R * R, Jt += prob * R * R;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c0 = int64_t(blockIdx.x) * 32, r0 = int64_t(blockIdx.y) * 32;
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + i, c = c0 + tx;
    float v = 0.f;
    if (r < R && c < R) {
      v = Lw[wide_index(Np, nbw, prob, int(r), int(c))] * inv[c];
      Jm[r * R + c] = v;
    }
    tile[i][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int64_t c = c0 + i, r = r0 + tx;
    if (r < R && c < R) Jt[c * R + r] = tile[tx][i];
  }
}

// diag[w] = P_w^T P_w of every wide block, plain FFMA (test hook only: the solver's intra rounds fill the cache)
__global__ void __launch_bounds__(256) wide_diag_kernel(float* diag, const float* Lw, int Np) {
  const int w = blockIdx.x, i0 = (threadIdx.x / 16) * 4, j0 = (threadIdx.x % 16) * 4;
  const float* P = Lw + size_t(w) * Np * WB;
  float acc[4][4] = {};
  for (int r = 0; r < Np; ++r) {
    const float4 a = *reinterpret_cast<const float4*>(P + size_t(r) * WB + i0);
    const float4 b = *reinterpret_cast<const float4*>(P + size_t(r) * WB + j0);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) diag[(size_t(w) * WB + i0 + i) * WB + j0 + j] = acc[i][j];
}

// ---- host side ------------------------------------------------------------------------------------------
struct WidePlan {
  int Np, nbw, pairs, splits, kblocks, kblocks_per_split;
  int64_t part_bytes, q_bytes, flag_bytes, diag_bytes;  // diag: the diagonal-block cache, behind the flags
};

// `world` > 1: the pairs of a round are distributed over that many GPUs (vvt_syevj_dist); the split-K plan is the
// one of a rank's share, the buffers stay indexed by the pair index of the whole round
static inline WidePlan wide_plan(int64_t R, int64_t batch, int world = 1) {
  WidePlan p;
  p.Np = int(align_up(R, WP));
  p.nbw = p.Np / WB;
  p.pairs = p.nbw / 2;
  p.kblocks = p.Np / tc::BK;
  // split-K over the rows: one CTA per SM is resident, so the kernel runs in waves of num_sms CTAs; take the
  // split count (at least 4 k-blocks per CTA) that minimises waves x (k-blocks per CTA + prologue / epilogue)
  {
    const int64_t sms = num_sms(), tiles = ceil_div(int64_t(p.pairs), int64_t(vmax(1, world))) * batch;
    int64_t best = 1, best_cost = INT64_MAX;
    for (int64_t sp = 1; sp <= vmax<int64_t>(1, p.kblocks / 4); ++sp) {
      const int64_t per = ceil_div(p.kblocks, sp);
      const int64_t ctas = tiles * ceil_div(p.kblocks, per);
      const int64_t cost = ceil_div(ctas, sms) * (per + 8);
      if (cost < best_cost) best_cost = cost, best = sp;
    }
    p.kblocks_per_split = int(ceil_div(p.kblocks, best));
    p.splits = int(ceil_div(p.kblocks, p.kblocks_per_split));
  }
  p.part_bytes = align_up(batch * p.splits * p.pairs * int64_t(WP * WP) * 4, 256);
  p.q_bytes = align_up(batch * p.pairs * int64_t(WP * WP) * 4, 256);
  p.flag_bytes = align_up(batch * p.pairs * 4, 256);
  p.diag_bytes = align_up(batch * p.nbw * int64_t(WB * WB) * 4, 256);
  return p;
}

static inline bool make_map_box(CUtensorMap* map, const float* ptr, int64_t rows, int64_t cols, int64_t ld, int64_t batch,
                                int64_t batch_stride, int box_rows, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  tc::EncodeTiledFn fn = tc::encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[3] = {cuuint64_t(cols), cuuint64_t(rows), cuuint64_t(batch)};
  const cuuint64_t strides[2] = {cuuint64_t(ld) * 4, cuuint64_t(batch_stride) * 4};
  const cuuint32_t box[3] = {tc::BK, cuuint32_t(box_rows), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct WideMaps {
  CUtensorMap gram, apply, q;
};

static inline int wide_make_maps(WideMaps* m, const float* Lw, const float* Qt, const WidePlan& p, int64_t batch) {
  const int variant = getenv("VVT_WIDE_DESC") ? atoi(getenv("VVT_WIDE_DESC")) : 0;
  // the factor as [problem * wide block][row][64 columns]
  const int64_t blocks = batch * p.nbw, bs = int64_t(p.Np) * WB;
  if (!make_map_box(&m->gram, Lw, p.Np, WB, WB, blocks, bs, 32,
                    variant >= 2 && variant != 4 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
      !make_map_box(&m->apply, Lw, p.Np, WB, WB, blocks, bs, tc::BM) ||
      !make_map_box(&m->q, Qt, WP, WP, WP, batch * p.pairs, int64_t(WP) * WP, tc::BN))
    return fail(VVT_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed", "vvt_syevj(wide)");
  return VVT_OK;
}

// one round (all pairs, all problems): Gram, rotations, apply
static inline int wide_round(float* Lw, float* part, float* Qt, int* flag, float* diag, JacobiScalars* sc, const WidePlan& p,
                             const WideMaps& m, int round, int64_t batch, cudaStream_t s, int pair0 = 0,
                             int pairs_local = -1) {
  if (pairs_local < 0) pairs_local = p.pairs;
  if (pairs_local == 0) return VVT_OK;
  static SmemOptIn opt_g, opt_a, opt_r;
  VVT_TRY(opt_g.ensure(wide_tc_kernel<GRAM>, tc::SMEM_BYTES, "vvt_syevj(wide gram)"));
  VVT_TRY(opt_a.ensure(wide_tc_kernel<APPLY>, tc::SMEM_BYTES, "vvt_syevj(wide apply)"));
  VVT_TRY(opt_r.ensure(wide_rot_kernel, sizeof(WideRotSmem), "vvt_syevj(wide rot)"));
  WideArgs a;
  a.flag = flag;
  a.sc = sc;
  a.l_stride = int64_t(p.Np) * p.Np;
  a.Np = p.Np, a.nbw = p.nbw, a.round = round, a.pairs = p.pairs, a.splits = p.splits;
  a.kblocks_total = p.kblocks, a.kblocks_per_split = p.kblocks_per_split;
  a.pair0 = pair0, a.pairs_local = pairs_local;
  a.variant = getenv("VVT_WIDE_DESC") ? atoi(getenv("VVT_WIDE_DESC")) : 0;
  a.out = part;
  static const bool one_cta = getenv("VVT_WIDE_ROT_ONE_CTA") != nullptr;  // experiments: the single-CTA kernel
  // VVT_WIDE_CROSS_GRAM=1: cross rounds form only H[:, b] (N = 64) and take H_aa from the diagonal-block cache.
  // Measured: the Gram kernel is bound by its converter warps and the operand loads, not by the MMAs -- 2.5 % per
  // sweep at R = 5120, 6 % at R = 10240, and one sweep more on both (24 / 27 instead of 23 / 26) -- so it is off.
  static const bool cross_gram = getenv("VVT_WIDE_CROSS_GRAM") != nullptr;
  const int cross_only = !one_cta && cross_gram;
  a.cross_only = cross_only;
  static const bool pdl = getenv("VVT_WIDE_NOPDL") == nullptr;
  cudaLaunchAttribute pdl_attr[2];
  pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(pairs_local), unsigned(p.splits), unsigned(batch));
    cfg.blockDim = dim3(tc::THREADS);
    cfg.dynamicSmemBytes = tc::SMEM_BYTES;
    cfg.stream = s;
    cfg.attrs = pdl_attr;
    cfg.numAttrs = pdl ? 1 : 0;
    VVT_TRY(check_cuda(cudaLaunchKernelEx(&cfg, wide_tc_kernel<GRAM>, m.gram, m.q, a), "vvt_syevj(wide gram)"));
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  if (one_cta) {
    wide_rot_kernel<<<dim3(unsigned(pairs_local), unsigned(batch)), RT, sizeof(WideRotSmem), s>>>(Qt, flag, part, p.nbw, round,
                                                                                                  p.pairs, p.splits, sc, pair0);
    VVT_TRY(launched("vvt_syevj(wide rot)"));
  } else {
    static SmemOptIn opt_c;
    VVT_TRY(opt_c.ensure(wide_rot_cluster_kernel, sizeof(RotCta), "vvt_syevj(wide rot)"));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(pairs_local * CR), unsigned(batch));
    cfg.blockDim = dim3(OT);
    cfg.dynamicSmemBytes = sizeof(RotCta);
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CR;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1] = pdl_attr[0];
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 2 : 1;
    VVT_TRY(check_cuda(cudaLaunchKernelEx(&cfg, wide_rot_cluster_kernel, Qt, flag, part, p.nbw, round, p.pairs, p.splits, sc,
                                          diag, cross_only, pair0),
                       "vvt_syevj(wide rot)"));
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  a.out = Lw;
  static const bool per_tile = getenv("VVT_WIDE_APPLY_PER_TILE") != nullptr;  // experiments: one CTA per output tile
  if (per_tile) {
    wide_tc_kernel<APPLY><<<dim3(unsigned(p.Np / tc::BM), unsigned(pairs_local), unsigned(batch)), tc::THREADS, tc::SMEM_BYTES, s>>>(
        m.apply, m.q, a);
  } else {
    static SmemOptIn opt_p;
    VVT_TRY(opt_p.ensure(wide_apply_kernel, AP_SMEM, "vvt_syevj(wide apply)"));
    const int64_t items = int64_t(pairs_local) * (p.Np / tc::BM);
    const unsigned ctas = unsigned(vmax<int64_t>(1, vmin<int64_t>(items, num_sms() / vmin<int64_t>(batch, num_sms()))));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas, unsigned(batch));
    cfg.blockDim = dim3(AP_THREADS);
    cfg.dynamicSmemBytes = AP_SMEM;
    cfg.stream = s;
    cfg.attrs = pdl_attr;
    cfg.numAttrs = pdl ? 1 : 0;
    VVT_TRY(check_cuda(cudaLaunchKernelEx(&cfg, wide_apply_kernel, m.apply, m.q, a), "vvt_syevj(wide apply)"));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return VVT_OK;
  }
  VVT_TRY(launched("vvt_syevj(wide apply)"));
  return VVT_OK;
}

}  // namespace wide
}  // namespace vvt
