// K3: tcgen05 tile GEMM with the SAGE update epilogue fused in.
//
//   out[m, 0:512] = epilogue( sum_s A_s[m, :] . B_s^T )       (up to 6 K-segments)
//
// The SAGE update is the two segments (agg, lin_l.weight) and (x, lin_r.weight), i.e.
// the concatenated [lin_l | lin_r] GEMM with K = 1024 -- replaces `lin_l(agg) +
// lin_r(x)`, `F.normalize`, eval `BatchNorm1d`, `ReLU` and the skip connection of the
// reference layer loop (Models/BuckGNN.py:449-457 + PyG SAGEConv.forward).  The same
// kernel runs the encoders' 128 -> 512 Linear (:73, :81) and every Linear of the EA-GNN
// block (:528-566), whose gathered operands arrive as epilogue addends.
//
// One persistent CTA pair per SM pair (tcgen05 cta_group::2), 12 warps per CTA:
//   warp 0     TMA producer: this CTA's 128 activation rows x 128 B of K and its 256 of
//              the 512 weight rows into a 4-stage ring of 128B-swizzled K-major tiles
//   warp 1     allocates TMEM; in the leader CTA one elected lane issues tcgen05.mma
//              (M = 256 over the pair, N = 256, two N halves -> the 512 fp32 columns of
//              TMEM hold one full output row per lane: L2-normalize needs the whole row)
//   warps 2-3  idle (register donors: setmaxnreg 40 for warps 0-3, 232 for the epilogue)
//   warps 4-11 epilogue, two groups of 4 warps = two 256-column halves; one TMEM lane
//              (= output row) per thread.
//              16-bit output: ONE pass over TMEM (double-buffered tcgen05.ld) adds the bias,
//              accumulates the row's sum of squares and stashes the row as packed 16-bit
//              pairs in 128 registers; TMEM is released right after, so the next tile's
//              MMAs overlap pass 2, which applies normalize / BN / ReLU / skip from the stash.
//              fp32 output (tf32 / 3xTF32 modes): two passes over TMEM (no stash).
//              Global I/O of pass 2 is warp-cooperative and coalesced: a warp moves 128-byte
//              row chunks with 8 lanes per row (4 full lines per instruction) between
//              global memory and a 4 KB swizzled staging tile of its own, where each thread
//              picks up / drops its row.  Skip rows and gathered addends are prefetched one
//              chunk ahead.  (An earlier version staged whole 128-row tiles for TMA stores;
//              the store-read latency of the 2-deep ring, ~3k cycles, made every K <= 512
//              GEMM epilogue-bound -- profiles/r01_gemm_role_cycles_v3.txt.)
// Tensor-core bound for K = 1024: 2*K*512 flops per row; algorithmic bytes per row
// = (K_total + 512 [+512 skip]) * elem_size.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace bg {

constexpr int kTileM = 128;                       // rows per CTA
constexpr int kStageKBytes = 128;                 // one 128B swizzle row of K per stage
constexpr int kATileBytes = kTileM * kStageKBytes;        // 16 KB
constexpr int kGemmThreads = 384;                 // 12 warps
constexpr int kGatherWarps = 4;                   // kFuse: warps 12-15 (one warpgroup: setmaxnreg is per warpgroup) gather the aggregate operand
constexpr int kGemmThreadsFused = 512;            // 16 warps
constexpr int kEpiFirstWarp = 4;
constexpr int kEpiWarps = 8;
constexpr int kEpiStageBytes = 32 * 128;          // per-warp staging tile: 32 rows x 128 B

template <int kCg> struct GemmCfg {
  static constexpr int kBRows = kHidden / kCg;                  // weight rows held by one CTA
  static constexpr int kBTileBytes = kBRows * kStageKBytes;     // 64 KB / 32 KB
  static constexpr int kStageBytes = kATileBytes + kBTileBytes; // 80 KB / 48 KB
#ifndef BG_GEMM_STAGES
#define BG_GEMM_STAGES 4
#endif
  static constexpr int kStages = (kCg == 1) ? 2 : BG_GEMM_STAGES;
  static constexpr int kMiscBytes = 128;                        // mbarriers + TMEM address slot
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kEpiWarps * kEpiStageBytes + kMiscBytes;
};

static_assert(GemmCfg<2>::kSmemBytes <= 232448, "exceeds the 227 KB of dynamic shared memory a CTA may have");
static_assert((2 * GemmCfg<2>::kStages + 2) * 8 <= 96, "mbarriers overlap the TMEM address slot");

struct alignas(64) GemmSeg { CUtensorMap a; CUtensorMap b; };

struct alignas(64) GemmParams {
  GemmSeg seg[BG_MAX_GEMM_SEGMENTS];
  int32_t kblocks[BG_MAX_GEMM_SEGMENTS];
  int32_t n_seg;
  int32_t k_elems_per_block;        // 64 (16-bit operands) or 32 (tf32)
  uint32_t a_fmt, b_fmt;            // UMMA operand formats: 0 f16, 1 bf16, 2 tf32
  int32_t n_tiles;                  // row tiles of 128*cg rows
  int32_t normalize, relu;
  int32_t n_gather;                 // 0..2 gathered pre-activation addends: v += G_k[gidx_k[m], :]
  int64_t m;
  void* out;                        // [M, 512] of the output dtype, leading dimension ldo (elements)
  int64_t ldo;
  const void* residual;             // [M, 512] skip rows of the output dtype (ld = ldr), or null
  int64_t ldr;
  const void* gather[2];            // [*, 512] row-major matrices of the output dtype, ld = gather_ld
  const int32_t* gidx[2];           // [M] row index into gather[k]
  int64_t gather_ld;
  float* inv_norm_out;              // [M] f32 or null: 1 / max(|row|, eps) of normalized rows (saved for the backward pass)
  float* pool_sums;                 // kPool: [ceil(M/32), 512] f32 column sums of every 32-row block of the OUTPUT rows
  const uint8_t* pool_keep;         // kPool: [ceil(M/32)] only blocks with a non-zero flag are also stored to `out`
  int32_t b_group_tiles;            // 0: one B for all rows; t: row tile i multiplies B rows [(i / t) * 512, +512)
  int32_t mn_major;                 // 1: weight-gradient mode.  seg[0].a / .b map the ROW-MAJOR [n_rows, 512] matrices
                                    // dz and act; out[(s*512 + o), i] = sum over the nodes of chunk s of dz[n,o] act[n,i]:
                                    // both operands are MN-major (the reduction runs over matrix rows), so no
                                    // transposed copies are needed.  Tile t = (chunk t/2, output rows (t%2)*256..+256).
  int64_t chunk_k;                  // nodes per chunk (mn_major)
  // kFuse (fused SAGE layer): segment 0's A operand -- the neighbourhood aggregate of the rows of `fuse_x` -- is not
  // read from memory but PRODUCED inside the kernel by gather warps (see fused_gather below); seg[0].a is unused
  const void* fuse_x;               // [*, 512] 16-bit rows of the layer input (the same matrix seg[1].a maps), ld = fuse_ldx
  int64_t fuse_ldx;
  const int32_t* fuse_rowptr;       // CSR by destination row (bg_csr_build)
  const int32_t* fuse_col;
  const void* fuse_hub_agg;         // [n_big, 512] aggregates of the hub rows (degree > kBigRowThreshold), slot = index in big_rows
  const int32_t* fuse_big_rows;
  int32_t fuse_n_big;
  int32_t fuse_mean;                // 1: mean, 0: sum
  float bias[kHidden];              // epilogue vectors by value -> constant bank, broadcast reads
  float scale[kHidden];             // 1 when there is no BN
  float shift[kHidden];             // 0 when there is no BN
  uint32_t scale_h2[kHidden / 2];   // the same two vectors as packed fp16 pairs (fp16 outputs: pass 2 runs on
  uint32_t shift_h2[kHidden / 2];   // HMUL2 / HFMA2 / HMNMX2 / HADD2, two columns per instruction)
};

// Optional role-level cycle accounting (build with -DBG_PROFILE, see tools/gemm_bench.py):
// per CTA: [0] producer wait-empty, [1] mma wait-full, [2] mma wait-tmem-empty, [3] mma total,
// [4] epi wait-tmem-full, [5] epi pass 1, [6] epi pass 2, [7] epi total
#ifdef BG_PROFILE
__device__ unsigned long long g_gemm_prof[296 * 8];
#define BG_PROF_DECL long long _pt0 = 0, _pacc_a = 0, _pacc_b = 0, _pacc_c = 0; const long long _pstart = clock64();
#define BG_PROF_T0() _pt0 = clock64()
#define BG_PROF_ADD(acc) acc += clock64() - _pt0
#define BG_PROF_STORE(slot, v) g_gemm_prof[blockIdx.x * 8 + (slot)] = (unsigned long long)(v)
#else
#define BG_PROF_DECL
#define BG_PROF_T0()
#define BG_PROF_ADD(acc)
#define BG_PROF_STORE(slot, v)
#endif

#ifndef BG_GEMM_P1_WIDE
#define BG_GEMM_P1_WIDE 0          // 1: pass 1 of the epilogue reads TMEM 32 columns per load.  Measured SLOWER (r02:
                                   // pass 1 4.9k -> 8.1k cycles per tile, spills): the drain runs at the TMEM read rate
                                   // (~57 B/cycle/SM: 256 KB in 4.5k cycles), not at load latency, so wider loads cannot help
#endif

// Epilogue global I/O experiments (r02, tools/gemm_bench.py, K = 1024 fp16 with skip rows; baseline 1.085 ms per launch):
// the MMA phase of a tile is shared-memory-bandwidth bound (operand reads 64 B/cycle + TMA fill ~39 B/cycle of the SM's
// 128 B/cycle) and the previous tile's pass 2 runs beside it, so moving the epilogue's global I/O off the staging tile
// looked attractive.  With one ROW per thread it is not: a thread reading the skip pieces of its own row (eight 16-byte
// loads per 128-byte line) 1.74 ms; plus 16-byte stores of its own row 1.91 ms, and the K = 128 encoder GEMM 0.30 ->
// 0.61 ms.  32 distinct lines per warp instruction cost far more in L1/L2 transactions than the STS + LDS round trip
// through the warp's swizzled tile (code removed; git 1543111).  What does pay: skip rows never enter shared memory at
// all -- they are added to the coalesced output pieces on their way out (kSkipAtStore below).
// 1 (default): pass 2 of the normalize / BN epilogue on packed fp32 pairs with the ReLU and the skip add on the packed
// 16-bit result.  0: the scalar fp32 sequence (skip rows added in fp32 before the one rounding).
#ifndef BG_EPI_F32X2
#define BG_EPI_F32X2 1
#endif

enum : uint32_t { kTagEmpty = 1, kTagFull = 2, kTagTmemEmpty = 3, kTagTmemFull = 4 };

BG_DEVINL void named_bar_sync(uint32_t id, uint32_t threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
BG_DEVINL uint4 lds_v4(uint32_t addr) {
  uint4 r;
  asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
  return r;
}
BG_DEVINL void sts_v4(uint32_t addr, uint4 v) {
  asm volatile("st.volatile.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
template <int kRegs> BG_DEVINL void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <int kRegs> BG_DEVINL void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }


struct EpiCtx {
  uint32_t tmem_base, stage_u32;    // stage_u32: this warp's 4 KB staging tile
  float* my_tile;                   // generic pointers to this warp's tile and to the tile of the warp that owns
  const float* partner_tile;        // the other 256 columns of the same rows (row sum-of-squares exchange)
  uint32_t tmem_full_bar, tmem_empty_bar, rank;
  int tile0, tile_stride;
};

// TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns
BG_DEVINL void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}

// what pass 2 adds to the accumulator besides bias / BN: nothing, skip rows after the activation, or
// gathered rows before it.  A template parameter of the kernel: the epilogue is fully unrolled
// straight-line code (the register stash forbids loops), and dead variants would only evict live
// code from the instruction cache (v4 of this kernel, 13k SASS instructions, spent most issue slots
// stalled on instruction fetch -- sm__icc_request_hit_rate 52 %).
enum : int { kAddNone = 0, kAddResidual = 1, kAddGather = 2 };

// One epilogue warp: 32 TMEM lanes = 32 output rows, the 256 columns starting at g*256.
// `g` is a run-time value (both column halves share ONE copy of the code); the per-column
// vectors are read from the constant bank as float4 at g*256 + immediate.
// kPlain: no L2-normalize and no BatchNorm scale/shift (every EA-GNN Linear, the encoders, the training step's
// input-gradient GEMMs).  With a 16-bit output, pass 2 then is "stash (+ gathered rows) -> ReLU (+ skip rows)" on
// packed pairs: HADD2 / HMNMX2 round exactly like the fp32 path does (the sum of two 16-bit values is exact in
// fp32, so both round the exact sum once), at ~1.5 instead of ~6.5 instructions per column.
// kPool (last GraphSAGE layer of a graph-level model, 16-bit output, normalize epilogue, no addends): the layer's
// output is only read by the pooling layer (Models/BuckGNN.py:515), which is linear in the rows, so the epilogue sums
// the rows it produces instead of storing them -- per 32-row block (one warp's rows) and column, fp32, fixed order --
// and writes [M/32, 512] sums (6 % of the bytes).  Blocks holding the first or last row of a graph (pool_keep) are
// stored as well: bg_pool_head_blocks takes those rows from `out`, so graph boundaries inside a block, super-node
// pooling variants and graphs smaller than a block all stay exact.  Saves the 1 GB write here and the 1 GB read of
// the pooling pass (cfg 2).
template <int kCg, typename TOut, int kAdd, bool kPlain, bool kPool = false>
BG_DEVINL void epilogue_warp(const GemmParams& p, const EpiCtx& cx, const int g) {
  static_assert(!kPool || (sizeof(TOut) == 2 && kAdd == kAddNone && !kPlain), "pool-fused epilogue: 16-bit normalize epilogue only");
  constexpr bool kOut16 = sizeof(TOut) == 2;
  // -DBG_GEMM_PACKED_EPILOGUE: pass 2 of fp16 outputs as packed half2 math also WITH normalize / BN.  Measured
  // (r01): SAGE update 1.04 -> 1.00 ms, but the prediction error vs the fp32 oracle rises from 6e-5 to 2.8e-4
  // (u = v * inv and the BN scale are rounded to 11 bits before the affine step).  Off by default: 4 % is not
  // worth a 5x smaller accuracy margin.
#ifdef BG_GEMM_PACKED_EPILOGUE
  constexpr bool kPacked = sizeof(TOut) == 2 && !is_bf16<TOut>::value && !kPlain;
#else
  constexpr bool kPacked = false;
#endif
  constexpr bool kPlainPacked = kPlain && kOut16;
  constexpr bool kF32x2 = kOut16 && !kPlain && !kPacked && kAdd != kAddGather && BG_EPI_F32X2;
  // skip rows are added to the packed 16-bit activation AFTER the transpose through the staging tile, in the coalesced
  // layout they were loaded in: 256 KB less shared-memory traffic per tile (the MMA phase is shared-memory-bound)
  constexpr bool kSkipAtStore = kAdd == kAddResidual && (kPlainPacked || kF32x2);
  constexpr int kChunkCols = 128 / (int)sizeof(TOut);           // columns per 128-byte row chunk: 64 / 32
  constexpr int kChunks = 256 / kChunkCols;                     // 4 / 8
  constexpr int kPer = 16 / (int)sizeof(TOut);                  // 8 or 4 columns per 16-byte piece
  // (warp index through a shuffle: the compiler then KNOWS it is warp-uniform, so g-indexed reads of the per-column
  // vectors become uniform-datapath constant loads feeding FFMA2 / FADD2 directly instead of per-thread indexed LDCs,
  // which were half of pass 2's time -- r02 probe -DBG_EPI_NO_CONST)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int q = warp & 3;                                       // TMEM lane quarter this warp may read
  const int cb = g * 256;
  const uint32_t taddr = cx.tmem_base + ((uint32_t)(q * 32) << 16) + cb;
#ifdef BG_EPI_NO_CONST                                          // timing probe (wrong results): per-column vectors as immediates
  struct FakeVec { float a; BG_DEVINL float4 operator[](int) const { return make_float4(a, a + 0.25f, a + 0.5f, a + 0.75f); } };
  const FakeVec bias4{0.5f}, scale4{2.f}, shift4{1.f};
#else
  const float4* bias4 = reinterpret_cast<const float4*>(p.bias + cb);
  const float4* scale4 = reinterpret_cast<const float4*>(p.scale + cb);
  const float4* shift4 = reinterpret_cast<const float4*>(p.shift + cb);
#endif
  // per-thread view of the staging tile (own row = lane) and cooperative view (8 lanes per row, 4 rows per pass;
  // the swizzle of row 4t + r4 only depends on the parity of t)
  const uint32_t my_row = cx.stage_u32 + (uint32_t)lane * 128u;
  const uint32_t my_sw = (uint32_t)(lane & 7);
  const int r4 = lane >> 3, piece = lane & 7;
  const uint32_t coop_even = cx.stage_u32 + (uint32_t)r4 * 128u + (((uint32_t)piece ^ (uint32_t)r4) << 4);
  const uint32_t coop_odd = cx.stage_u32 + (uint32_t)(4 + r4) * 128u + (((uint32_t)piece ^ (uint32_t)(4 + r4)) << 4);
  constexpr size_t esz = sizeof(TOut);
  const size_t out_row_bytes = (size_t)p.ldo * esz;
  uint32_t it = 0;
  BG_PROF_DECL
  for (int tile = cx.tile0; tile < p.n_tiles; tile += cx.tile_stride, ++it) {
    const int64_t warp_row0 = (int64_t)tile * (kTileM * kCg) + (int64_t)cx.rank * kTileM + q * 32;
    const int64_t m_row = warp_row0 + lane;
    const int rows_here = (int)min((int64_t)32, p.m - warp_row0);                 // <= 0: nothing to store
    [[maybe_unused]] bool keep_rows = true;
    if constexpr (kPool) keep_rows = rows_here > 0 && p.pool_keep[warp_row0 >> 5] != 0;   // warp-uniform
    [[maybe_unused]] int32_t gi0 = 0, gi1 = 0;
    if constexpr (kAdd == kAddGather) {
      if (m_row < p.m) { gi0 = p.gidx[0][m_row]; if (p.n_gather > 1) gi1 = p.gidx[1][m_row]; }
    }
    // addend rows of one 128-byte chunk for this warp's 32 rows: pass t covers rows 4t..4t+3, 8 lanes x 16 B each
    [[maybe_unused]] uint4 pre[kAdd == kAddNone ? 1 : 8];
    [[maybe_unused]] const char* add_base = nullptr;
    if constexpr (kAdd == kAddResidual)
      add_base = reinterpret_cast<const char*>(p.residual) + (size_t)(warp_row0 + r4) * p.ldr * esz + (size_t)cb * esz + piece * 16;
    if constexpr (kAdd == kAddGather)
      add_base = reinterpret_cast<const char*>(p.gather[0]) + (size_t)cb * esz + piece * 16;
    auto fetch = [&](int ch) {
      if constexpr (kAdd == kAddResidual) {
#pragma unroll
        for (int t = 0; t < 8; ++t)
          pre[t] = (t * 4 + r4 < rows_here) ? (kSkipAtStore ? ldg_volatile_v4(add_base + (size_t)t * 4 * p.ldr * esz + ch * 128)
                                                            : ldg_nc_v4(add_base + (size_t)t * 4 * p.ldr * esz + ch * 128))
                                            : make_uint4(0u, 0u, 0u, 0u);
      } else if constexpr (kAdd == kAddGather) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int32_t n0 = __shfl_sync(0xffffffffu, gi0, t * 4 + r4);
          pre[t] = ldg_v4(add_base + (size_t)n0 * p.gather_ld * esz + ch * 128);
        }
      }
    };
    // first addend chunk: skip rows that join at the store are first needed after chunk 0's math, so their loads are
    // issued after pass 1 and the row-norm exchange (32 fewer live registers in pass 1; and a CTA barrier waits for
    // volatile loads in flight).  Staged addends are needed at once: in flight while we wait for the MMAs.
    constexpr bool kFetchLate = kSkipAtStore || BG_GEMM_P1_WIDE;
    if constexpr (!kFetchLate) fetch(0);

    BG_PROF_T0();
    mbar_wait(cx.tmem_full_bar, it & 1u, kTagTmemFull);
    BG_PROF_ADD(_pacc_a);
    tc_fence_after();
    BG_PROF_T0();

    // ---- pass 1: drain the accumulator.  This is the one part of the epilogue the next tile's MMAs cannot overlap
    // (the 128 x 512 fp32 accumulator is all of TMEM), and it is issue-bound, not TMEM-bound: the bare drain takes
    // ~0.7k cycles per tile, every instruction per column pair adds ~1k (r02 probe, -DBG_P1_DRAIN_ONLY: 0.93 ms per
    // SAGE update against 1.04 ms with bias + sum of squares + pack in here).  So the normalize / BN epilogue
    // (kDeferBias) only PACKS the raw accumulator here; the bias is added when the stash is unpacked (one mixed-precision
    // add per column either way) and the row's sum of squares is taken from the stash after TMEM is released.
    // Plain epilogues (no normalize) add the bias here and never need the sum.
    constexpr bool kDeferBias = kF32x2;
    [[maybe_unused]] uint32_t stash[kOut16 ? 128 : 1];
    float ss = 0.f;
    uint64_t ss_a = 0ull, ss_b = 0ull;                          // four independent partial sums of squares (two packed pairs)
    if (kOut16 || p.normalize) {
      auto consume4 = [&](const uint32_t* r, int c4) {          // c4: index of the 4-column group
#ifdef BG_P1_DRAIN_ONLY                                         // timing probe (wrong results): TMEM -> registers only
        if constexpr (kOut16) { stash[2 * c4] = r[0] ^ r[1]; stash[2 * c4 + 1] = r[2] ^ r[3]; return; }
#endif
        if constexpr (kDeferBias) {
          stash[2 * c4] = Pack16<TOut>::pack(__uint_as_float(r[0]), __uint_as_float(r[1]));
          stash[2 * c4 + 1] = Pack16<TOut>::pack(__uint_as_float(r[2]), __uint_as_float(r[3]));
          return;
        }
        // packed fp32 pairs (FADD2 / FFMA2); the sum of squares is four independent chains
        const float4 b = bias4[c4];
        const uint64_t v01 = f2_add(f2_pack(__uint_as_float(r[0]), __uint_as_float(r[1])), f2_pack(b.x, b.y));
        const uint64_t v23 = f2_add(f2_pack(__uint_as_float(r[2]), __uint_as_float(r[3])), f2_pack(b.z, b.w));
        if constexpr (!kPlain) {
          ss_a = f2_fma(v01, v01, ss_a);
          ss_b = f2_fma(v23, v23, ss_b);
        }
        if constexpr (kOut16) {
          float v0, v1, v2, v3;
          f2_unpack(v01, v0, v1); f2_unpack(v23, v2, v3);
          stash[2 * c4] = Pack16<TOut>::pack(v0, v1);
          stash[2 * c4 + 1] = Pack16<TOut>::pack(v2, v3);
        }
      };
#if BG_GEMM_P1_WIDE
      uint32_t ra[32], rb[32];
      tmem_ld_32x32(taddr, ra);
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        tmem_ld_wait();
        tmem_ld_32x32(taddr + (c + 1) * 32, rb);
#pragma unroll
        for (int i = 0; i < 8; ++i) consume4(ra + 4 * i, c * 8 + i);
        tmem_ld_wait();
        if (c + 2 < 8) tmem_ld_32x32(taddr + (c + 2) * 32, ra);
#pragma unroll
        for (int i = 0; i < 8; ++i) consume4(rb + 4 * i, (c + 1) * 8 + i);
      }
#else
      uint32_t ra[16], rb[16];
      tmem_ld_32x16(taddr, ra);
#pragma unroll
      for (int c = 0; c < 16; c += 2) {
        tmem_ld_wait();
        tmem_ld_32x16(taddr + (c + 1) * 16, rb);
#pragma unroll
        for (int i = 0; i < 4; ++i) consume4(ra + 4 * i, c * 4 + i);
        tmem_ld_wait();
        if (c + 2 < 16) tmem_ld_32x16(taddr + (c + 2) * 16, ra);
#pragma unroll
        for (int i = 0; i < 4; ++i) consume4(rb + 4 * i, (c + 1) * 4 + i);
      }
#endif
    }
    if constexpr (kOut16) {                                     // accumulator fully read: release TMEM now
      tc_fence_before();
      if (kCg == 1 || cx.rank == 0) mbar_arrive(cx.tmem_empty_bar);
      else mbar_arrive_cluster(cx.tmem_empty_bar, 0);
    }
    if constexpr (kDeferBias) {
      if (p.normalize) {                                        // sum of squares of (stash + bias), beside the next tile's MMAs
        uint64_t ss_c = 0ull, ss_d = 0ull;
#pragma unroll
        for (int c4 = 0; c4 < 64; c4 += 2) {
          const float4 b0 = bias4[c4], b1 = bias4[c4 + 1];
          using P = Pack16<TOut>;
          const uint64_t p0 = f2_pack(P::add_lo(stash[2 * c4], b0.x), P::add_hi(stash[2 * c4], b0.y));
          const uint64_t p1 = f2_pack(P::add_lo(stash[2 * c4 + 1], b0.z), P::add_hi(stash[2 * c4 + 1], b0.w));
          const uint64_t p2 = f2_pack(P::add_lo(stash[2 * c4 + 2], b1.x), P::add_hi(stash[2 * c4 + 2], b1.y));
          const uint64_t p3 = f2_pack(P::add_lo(stash[2 * c4 + 3], b1.z), P::add_hi(stash[2 * c4 + 3], b1.w));
          ss_a = f2_fma(p0, p0, ss_a); ss_b = f2_fma(p1, p1, ss_b);
          ss_c = f2_fma(p2, p2, ss_c); ss_d = f2_fma(p3, p3, ss_d);
        }
        ss_a = f2_add(ss_a, ss_c); ss_b = f2_add(ss_b, ss_d);
      }
    }
    {
      float s0, s1, s2, s3;
      f2_unpack(ss_a, s0, s1); f2_unpack(ss_b, s2, s3);
      ss = (s0 + s1) + (s2 + s3);
    }
    float inv = 1.f;
    if (p.normalize) {                                          // exchange through the (idle) staging tiles
      cx.my_tile[lane] = ss;
      named_bar_sync(1, 256);
      inv = 1.f / fmaxf(sqrtf(ss + cx.partner_tile[lane]), 1e-12f);
      named_bar_sync(1, 256);                                   // partner has read before pass 2 reuses the tile
      if (p.inv_norm_out != nullptr && g == 0 && m_row < p.m) p.inv_norm_out[m_row] = inv;
    }
    [[maybe_unused]] const __half2 inv2 = __float2half2_rn(inv);
    if constexpr (kFetchLate) fetch(0);
    BG_PROF_ADD(_pacc_b);
    BG_PROF_T0();

    // ---- pass 2: normalize / BN / ReLU / addends, 128-byte row chunks through the warp's staging tile
    char* out_base = reinterpret_cast<char*>(p.out) + (size_t)(warp_row0 + r4) * out_row_bytes + (size_t)cb * esz + piece * 16;
#pragma unroll
    for (int ch = 0; ch < kChunks; ++ch) {
      if constexpr (kAdd != kAddNone && !kSkipAtStore) {
        if constexpr (kAdd == kAddGather) {
          if (p.n_gather > 1) {                                 // second gathered matrix, summed on the way in
            const char* g1_base = reinterpret_cast<const char*>(p.gather[1]) + (size_t)cb * esz + piece * 16 + ch * 128;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              const int32_t n1 = __shfl_sync(0xffffffffu, gi1, t * 4 + r4);
              const uint4 g1 = ldg_v4(g1_base + (size_t)n1 * p.gather_ld * esz);
              if constexpr (kOut16) {
                pre[t].x = Pack16<TOut>::hadd2(pre[t].x, g1.x); pre[t].y = Pack16<TOut>::hadd2(pre[t].y, g1.y);
                pre[t].z = Pack16<TOut>::hadd2(pre[t].z, g1.z); pre[t].w = Pack16<TOut>::hadd2(pre[t].w, g1.w);
              } else {
                pre[t].x = __float_as_uint(__uint_as_float(pre[t].x) + __uint_as_float(g1.x));
                pre[t].y = __float_as_uint(__uint_as_float(pre[t].y) + __uint_as_float(g1.y));
                pre[t].z = __float_as_uint(__uint_as_float(pre[t].z) + __uint_as_float(g1.z));
                pre[t].w = __float_as_uint(__uint_as_float(pre[t].w) + __uint_as_float(g1.w));
              }
            }
          }
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) sts_v4(((t & 1) ? coop_odd : coop_even) + (uint32_t)(t >> 1) * 1024u, pre[t]);
        __syncwarp();
        if (ch + 1 < kChunks) fetch(ch + 1);                    // next chunk's addends fly during the math below
      }
      // addend piece j of this chunk for the thread's own row (staged above)
      auto addend = [&](int j, uint32_t addr) -> uint4 { return lds_v4(addr); };
      auto emit = [&](int j, uint32_t addr, uint4 o) { sts_v4(addr, o); };
      [[maybe_unused]] uint32_t r[32];
      if constexpr (!kOut16) {
        tmem_ld_32x32(taddr + ch * 32, r);
        tmem_ld_wait();
        if (ch == kChunks - 1) {                                // last TMEM read of the tile
          tc_fence_before();
          if (kCg == 1 || cx.rank == 0) mbar_arrive(cx.tmem_empty_bar);
          else mbar_arrive_cluster(cx.tmem_empty_bar, 0);
        }
      }
      if constexpr (kPlainPacked) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t addr = my_row + (((uint32_t)j ^ my_sw) << 4);
          uint32_t o[4] = {stash[ch * 32 + j * 4], stash[ch * 32 + j * 4 + 1], stash[ch * 32 + j * 4 + 2], stash[ch * 32 + j * 4 + 3]};
          if constexpr (kAdd != kAddNone && !kSkipAtStore) {
            const uint4 ad = addend(j, addr);
            const uint32_t au[4] = {ad.x, ad.y, ad.z, ad.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if constexpr (kAdd == kAddGather) o[e] = Pack16<TOut>::hadd2(o[e], au[e]);      // addends before the activation
              if (p.relu) o[e] = Pack16<TOut>::relu2(o[e]);
              if constexpr (kAdd == kAddResidual) o[e] = Pack16<TOut>::hadd2(o[e], au[e]);    // skip rows after it
            }
          } else if (p.relu) {
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = Pack16<TOut>::relu2(o[e]);
          }
          emit(j, addr, make_uint4(o[0], o[1], o[2], o[3]));
        }
      } else if constexpr (kPacked) {
        // fp16 output: the stash already holds the row as fp16 pairs, so normalize / BN / ReLU / addends run as
        // packed half2 math -- 4 instructions per TWO columns instead of ~11 (unpack, FMUL, FFMA, FMNMX, FHADD,
        // pack).  The epilogue warps are latency-bound (2 per scheduler, ~3.8 k dependent-ish instructions per
        // tile), so the instruction count is what sets the tile rate of every K <= 1024 GEMM here.
        const uint2* sc2 = reinterpret_cast<const uint2*>(p.scale_h2 + cb / 2 + ch * 32);
        const uint2* sh2 = reinterpret_cast<const uint2*>(p.shift_h2 + cb / 2 + ch * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t addr = my_row + (((uint32_t)j ^ my_sw) << 4);
          [[maybe_unused]] uint4 ad;
          if constexpr (kAdd != kAddNone) ad = addend(j, addr);
          const uint32_t au[4] = {kAdd != kAddNone ? ad.x : 0u, kAdd != kAddNone ? ad.y : 0u,
                                  kAdd != kAddNone ? ad.z : 0u, kAdd != kAddNone ? ad.w : 0u};
          const uint2 sca = sc2[2 * j], scb = sc2[2 * j + 1], sha = sh2[2 * j], shb = sh2[2 * j + 1];
          const uint32_t scv[4] = {sca.x, sca.y, scb.x, scb.y}, shv[4] = {sha.x, sha.y, shb.x, shb.y};
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            __half2 v = *reinterpret_cast<const __half2*>(&stash[ch * 32 + j * 4 + e]);
            if constexpr (kAdd == kAddGather) v = __hadd2(v, *reinterpret_cast<const __half2*>(&au[e]));
            v = __hmul2(v, inv2);
            v = __hfma2(v, *reinterpret_cast<const __half2*>(&scv[e]), *reinterpret_cast<const __half2*>(&shv[e]));
            if (p.relu) v = __hmax2(v, __half2(__half(0.f), __half(0.f)));
            if constexpr (kAdd == kAddResidual) v = __hadd2(v, *reinterpret_cast<const __half2*>(&au[e]));
            o[e] = *reinterpret_cast<const uint32_t*>(&v);
          }
          emit(j, addr, make_uint4(o[0], o[1], o[2], o[3]));
        }
      } else if constexpr (kF32x2) {
        // SAGE update, 16-bit output: per column pair  stash + bias (2) . FMUL2 (x 1/|row|) . FFMA2 (BN scale, shift) .
        // pack (1) . packed ReLU (1) . packed skip add (1)  = 7 instructions instead of 11.  ReLU commutes with the
        // rounding to 16 bits (both monotonic, 0 is exact); the skip row is added to the ROUNDED activation, one more
        // round-to-nearest of the stored value than the fp32 add had (|error| <= 2^-12 of the activation, unbiased).
        const uint64_t inv_pair = f2_pack(inv, inv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t addr = my_row + (((uint32_t)j ^ my_sw) << 4);
          uint32_t o[4];
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2) {
            const float4 sc = scale4[(ch * kChunkCols + j * kPer) / 4 + e2];
            const float4 sh = shift4[(ch * kChunkCols + j * kPer) / 4 + e2];
            const float4 bb = bias4[(ch * kChunkCols + j * kPer) / 4 + e2];
            using P = Pack16<TOut>;                                           // stash + bias: one mixed-precision add per column
            const uint32_t u0 = stash[ch * 32 + j * 4 + 2 * e2], u1 = stash[ch * 32 + j * 4 + 2 * e2 + 1];
            uint64_t t0 = f2_mul(f2_pack(P::add_lo(u0, bb.x), P::add_hi(u0, bb.y)), inv_pair);
            uint64_t t1 = f2_mul(f2_pack(P::add_lo(u1, bb.z), P::add_hi(u1, bb.w)), inv_pair);
            t0 = f2_fma(t0, f2_pack(sc.x, sc.y), f2_pack(sh.x, sh.y));
            t1 = f2_fma(t1, f2_pack(sc.z, sc.w), f2_pack(sh.z, sh.w));
            float a0, a1, b0, b1;
            f2_unpack(t0, a0, a1); f2_unpack(t1, b0, b1);
            o[2 * e2] = Pack16<TOut>::pack(a0, a1);
            o[2 * e2 + 1] = Pack16<TOut>::pack(b0, b1);
          }
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (p.relu) o[e] = Pack16<TOut>::relu2(o[e]);
          emit(j, addr, make_uint4(o[0], o[1], o[2], o[3]));                // (the skip row joins at the coalesced store)
        }
      } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {                             // 16-byte pieces of this thread's 128-byte row chunk
        const uint32_t addr = my_row + (((uint32_t)j ^ my_sw) << 4);
        float v[kPer];
        if constexpr (kOut16) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t u = stash[ch * 32 + j * 4 + e];
            v[2 * e] = Pack16<TOut>::lo(u);
            v[2 * e + 1] = Pack16<TOut>::hi(u);
          }
        } else {
          const float4 b = bias4[ch * 8 + j];
          v[0] = __uint_as_float(r[j * 4]) + b.x; v[1] = __uint_as_float(r[j * 4 + 1]) + b.y;
          v[2] = __uint_as_float(r[j * 4 + 2]) + b.z; v[3] = __uint_as_float(r[j * 4 + 3]) + b.w;
        }
        [[maybe_unused]] uint4 ad;
        if constexpr (kAdd != kAddNone) ad = addend(j, addr);
        if constexpr (kAdd == kAddGather) {                     // gathered addends enter before the activation
          const uint32_t au[4] = {ad.x, ad.y, ad.z, ad.w};
          if constexpr (kOut16) {
#pragma unroll
            for (int e = 0; e < 4; ++e) Pack16<TOut>::add2(v[2 * e], v[2 * e + 1], au[e]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] += __uint_as_float(au[e]);
          }
        }
#pragma unroll
        for (int e4 = 0; e4 < kPer / 4; ++e4) {
          const float4 sc = scale4[(ch * kChunkCols + j * kPer) / 4 + e4];
          const float4 sh = shift4[(ch * kChunkCols + j * kPer) / 4 + e4];
          v[4 * e4] = fmaf(v[4 * e4] * inv, sc.x, sh.x); v[4 * e4 + 1] = fmaf(v[4 * e4 + 1] * inv, sc.y, sh.y);
          v[4 * e4 + 2] = fmaf(v[4 * e4 + 2] * inv, sc.z, sh.z); v[4 * e4 + 3] = fmaf(v[4 * e4 + 3] * inv, sc.w, sh.w);
        }
        if (p.relu) {
#pragma unroll
          for (int e = 0; e < kPer; ++e) v[e] = fmaxf(v[e], 0.f);
        }
        if constexpr (kAdd == kAddResidual) {                   // skip rows enter after it
          const uint32_t au[4] = {ad.x, ad.y, ad.z, ad.w};
          if constexpr (kOut16) {
#pragma unroll
            for (int e = 0; e < 4; ++e) Pack16<TOut>::add2(v[2 * e], v[2 * e + 1], au[e]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] += __uint_as_float(au[e]);
          }
        }
        uint4 o;
        if constexpr (kOut16) {
          o.x = Pack16<TOut>::pack(v[0], v[1]); o.y = Pack16<TOut>::pack(v[2], v[3]);
          o.z = Pack16<TOut>::pack(v[4], v[5]); o.w = Pack16<TOut>::pack(v[6], v[7]);
        } else {
          o.x = __float_as_uint(v[0]); o.y = __float_as_uint(v[1]); o.z = __float_as_uint(v[2]); o.w = __float_as_uint(v[3]);
        }
        emit(j, addr, o);
      }
      }
      __syncwarp();
      // coalesced store: 4 rows x 128 B per instruction
      [[maybe_unused]] float cs[8];                             // kPool: this lane's 8 columns summed over its 8 rows
      if constexpr (kPool) {
#pragma unroll
        for (int e = 0; e < 8; ++e) cs[e] = 0.f;
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        uint4 o = lds_v4(((t & 1) ? coop_odd : coop_even) + (uint32_t)(t >> 1) * 1024u);
        if constexpr (kSkipAtStore) {                           // skip rows: global -> registers -> global, never staged
          o.x = Pack16<TOut>::hadd2(o.x, pre[t].x); o.y = Pack16<TOut>::hadd2(o.y, pre[t].y);
          o.z = Pack16<TOut>::hadd2(o.z, pre[t].z); o.w = Pack16<TOut>::hadd2(o.w, pre[t].w);
        }
        if constexpr (kPool) {
          if (t * 4 + r4 < rows_here) {                         // one mixed-precision add per column (the stored value)
            using P = Pack16<TOut>;
            cs[0] = P::add_lo(o.x, cs[0]); cs[1] = P::add_hi(o.x, cs[1]); cs[2] = P::add_lo(o.y, cs[2]); cs[3] = P::add_hi(o.y, cs[3]);
            cs[4] = P::add_lo(o.z, cs[4]); cs[5] = P::add_hi(o.z, cs[5]); cs[6] = P::add_lo(o.w, cs[6]); cs[7] = P::add_hi(o.w, cs[7]);
            if (keep_rows) stg_v4(out_base + (size_t)t * 4 * out_row_bytes + ch * 128, o);
          }
        } else {
          if (t * 4 + r4 < rows_here) stg_v4(out_base + (size_t)t * 4 * out_row_bytes + ch * 128, o);
        }
      }
      if constexpr (kPool) {                                    // rows 4t + r4: add the four lane groups (butterfly)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 8);
          cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 16);
        }
        // every lane now holds the block's sums of its 8 columns; lane groups 0 and 1 store the two halves in ONE
        // instruction, so every 32-byte sector is written whole.  (In the forward this launch is ~0.09 ms slower than
        // the unfused last layer -- 64 mixed adds + 16 shuffles + 16 adds per chunk and warp -- against 0.15 ms saved
        // in the pooling pass: profiles/r02_pool_fused_probe.txt)
        if (r4 < 2 && rows_here > 0) {
          float4* dst = reinterpret_cast<float4*>(p.pool_sums + (size_t)(warp_row0 >> 5) * kHidden + cb + ch * kChunkCols + piece * kPer);
          dst[r4] = r4 == 0 ? make_float4(cs[0], cs[1], cs[2], cs[3]) : make_float4(cs[4], cs[5], cs[6], cs[7]);
        }
      }
      __syncwarp();
      if constexpr (kSkipAtStore) {                             // (after the warp barrier: it waits for loads in flight)
        if (ch + 1 < kChunks) fetch(ch + 1);                    // in flight during the next chunk's math
      }
    }
    BG_PROF_ADD(_pacc_c);
  }
#ifdef BG_PROFILE
  if (threadIdx.x == kEpiFirstWarp * 32) {
    BG_PROF_STORE(4, _pacc_a); BG_PROF_STORE(5, _pacc_b); BG_PROF_STORE(6, _pacc_c); BG_PROF_STORE(7, clock64() - _pstart);
  }
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// kFuse: the fused SAGE layer (north_star; reference Models/BuckGNN.py:449-457 = SAGEConv.propagate + lin_l / lin_r).
// The aggregate operand never exists in global memory.  Four gather warps per CTA build the [128 rows x 64 columns]
// K block of mean_j x[j] for the CTA's rows straight into the A half of the pipeline stage the MMA will read, in the
// 128-byte-swizzled K-major layout TMA would have produced: 8 lanes per row (one 16-byte chunk each, 4 rows per warp
// pass), up to 8 neighbour rows in flight per lane, fp32 accumulation in CSR order, the same reciprocal multiply and
// rounding as k_aggregate_rows -- so the operand is bit-identical to the unfused one.  The weight half of the stage
// still arrives by TMA; a stage's "full" barrier counts the producer and the gather warps of both CTAs of the pair.
// Neighbour rows come from L2 (the x tiles of the ~74 row tiles in flight were just read by the pairs next door);
// hub rows (the super node) are copied from a small side buffer filled by k_aggregate_hubs beforehand.
BG_DEVINL void mbar_arrive_cluster_release(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 r;\n\t"
      "mapa.shared::cluster.u32 r, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [r];\n\t}" ::"r"(bar), "r"(cta) : "memory");
}

template <int kCg, typename T, int kStages, int kStageBytes>
BG_DEVINL void fused_gather(const GemmParams& p, const int gw, const uint32_t stages_u32, const uint32_t bars_u32,
                            const uint32_t rank, const int tile0, const int tile_stride) {
  constexpr uint32_t kFull = 0xffffffffu;
  constexpr uint32_t kRowBytes = kHidden * (uint32_t)sizeof(T);          // rows of fuse_x are contiguous (ldx == 512)
  const int lane = threadIdx.x & 31;
  const int grp = lane >> 3, c = lane & 7;
  const char* xb = reinterpret_cast<const char*>(p.fuse_x) + c * 16;
  const char* hb = reinterpret_cast<const char*>(p.fuse_hub_agg) + c * 16;
  const int nkb = p.kblocks[0];
  int per_tile = 0;
  for (int s = 0; s < p.n_seg; ++s) per_tile += p.kblocks[s];
  // offsets, degree and first neighbour id (lane c holds neighbour c) of this lane group's row in pass `pass`;
  // a hub row gets its slot in the side buffer instead.  (Re-read for every K block: the 3 KB of indices of a tile
  // stay in L1, and per-tile register arrays for all passes would have to be unrolled -- the first version's 9 k
  // instructions of gather code evicted the epilogue from the instruction cache.)
  auto row_info = [&](int64_t row0, int pass, int32_t& b, int32_t& d, int32_t& v) {
    const int64_t r = row0 + 4 * pass + grp;
    b = 0;
    int32_t e = 0;
    if (pass < 32 && r < p.m) { b = p.fuse_rowptr[r]; e = p.fuse_rowptr[r + 1]; }
    d = e - b;
    const bool is_hub = d > kBigRowThreshold;
    v = (!is_hub && c < d) ? p.fuse_col[b + c] : 0;
    if (__any_sync(kFull, is_hub)) {                                    // rare: find the row in big_rows (8 lanes per probe)
      for (int32_t k = 0; k < p.fuse_n_big; k += 8) {
        const bool hit = is_hub && k + c < p.fuse_n_big && p.fuse_big_rows[k + c] == (int32_t)r;
        const uint32_t bits = (__ballot_sync(kFull, hit) >> (grp * 8)) & 0xffu;
        if (bits) v = k + __ffs(bits) - 1;
      }
    }
  };
  uint32_t cnt0 = 0;                                                    // pipeline slot of this tile's first K block
  for (int tile = tile0; tile < p.n_tiles; tile += tile_stride, cnt0 += (uint32_t)per_tile) {
    const int64_t row0 = (int64_t)tile * (kTileM * kCg) + (int64_t)rank * kTileM;
    for (int kb = 0; kb < nkb; ++kb) {
      const uint32_t cnt = cnt0 + (uint32_t)kb;
      const uint32_t stage = cnt % kStages, parity = (cnt / kStages) & 1u;
      int32_t b_n, d_n, v_n;
      row_info(row0, gw, b_n, d_n, v_n);                                // (before the wait: overlaps it)
      mbar_wait(bars_u32 + 8u * (kStages + stage), parity ^ 1u, kTagEmpty);          // the MMAs that read this slot retired
      const uint32_t sa = stages_u32 + stage * kStageBytes;
      const uint32_t koff = (uint32_t)kb * kStageKBytes;
#pragma unroll 1
      for (int pass = gw; pass < 32; pass += kGatherWarps) {
        const int32_t beg = b_n, d = d_n, nb = v_n;
        const int R = 4 * pass + grp;
        const bool is_hub = d > kBigRowThreshold;
        const int32_t dn = is_hub ? 0 : d;
        uint4 q[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t n = (uint32_t)__shfl_sync(kFull, nb, (lane & 24) + j);
          q[j] = make_uint4(0u, 0u, 0u, 0u);
          if (j < dn) q[j] = ldg_v4(xb + ((size_t)n * kRowBytes + koff));
        }
        row_info(row0, pass + kGatherWarps, b_n, d_n, v_n);             // next pass's indices fly with this pass's rows
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        auto add = [&](const uint4& qq) {
          Pack16<T>::add2(acc[0], acc[1], qq.x); Pack16<T>::add2(acc[2], acc[3], qq.y);
          Pack16<T>::add2(acc[4], acc[5], qq.z); Pack16<T>::add2(acc[6], acc[7], qq.w);
        };
#pragma unroll
        for (int j = 0; j < 8; ++j) add(q[j]);
        int32_t wmax = dn;                                              // more than 8 neighbours somewhere in the warp?
        wmax = max(wmax, __shfl_xor_sync(kFull, wmax, 8));
        wmax = max(wmax, __shfl_xor_sync(kFull, wmax, 16));
        for (int32_t base = 8; base < wmax; base += 8) {
          const int32_t idx = (base + c < dn) ? p.fuse_col[beg + base + c] : 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t n = (uint32_t)__shfl_sync(kFull, idx, (lane & 24) + j);
            q[j] = make_uint4(0u, 0u, 0u, 0u);
            if (base + j < dn) q[j] = ldg_v4(xb + ((size_t)n * kRowBytes + koff));
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) add(q[j]);
        }
        uint4 o;
        if (is_hub) {
          o = ldg_v4(hb + ((size_t)(uint32_t)nb * kRowBytes + koff));
        } else {
          if (p.fuse_mean) {
            const float rd = 1.f / (float)max(d, 1);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] *= rd;
          }
          o.x = Pack16<T>::pack(acc[0], acc[1]); o.y = Pack16<T>::pack(acc[2], acc[3]);
          o.z = Pack16<T>::pack(acc[4], acc[5]); o.w = Pack16<T>::pack(acc[6], acc[7]);
        }
        sts_v4(sa + (uint32_t)R * 128u + (((uint32_t)c ^ ((uint32_t)R & 7u)) << 4), o);
      }
      fence_proxy_async_smem();                                         // generic-proxy stores -> visible to the tensor core's reads
      named_bar_sync(2, kGatherWarps * 32);
      if (gw == 0 && lane == 0) {
        if (kCg == 1 || rank == 0) mbar_arrive(bars_u32 + 8u * stage);
        else mbar_arrive_cluster_release(bars_u32 + 8u * stage, 0);
      }
    }
    // The other segments' slots are filled by the producer alone, but the gather warps take part in every use of
    // every slot all the same -- wait for `empty`, meet, one of them arrives on `full`: a parity wait can only tell the
    // barrier's current phase from the previous one, so an agent that skips uses can take the phase of a skipped use for
    // its own (round 2, first build: fast gathers overwrote tiles the tensor core had not read), and one that merely
    // watches them without being waited for can fall two phases behind and wait forever.  With its arrival required
    // for every phase, neither can happen under any scheduling (host model: tests/test_ring_protocol_cpu.py).
    for (int kb = nkb; kb < per_tile; ++kb) {
      const uint32_t cnt = cnt0 + (uint32_t)kb;
      const uint32_t stage = cnt % kStages;
      mbar_wait(bars_u32 + 8u * (kStages + stage), ((cnt / kStages) & 1u) ^ 1u, kTagEmpty);
      named_bar_sync(2, kGatherWarps * 32);
      if (gw == 0 && lane == 0) {
        if (kCg == 1 || rank == 0) mbar_arrive(bars_u32 + 8u * stage);
        else mbar_arrive_cluster_release(bars_u32 + 8u * stage, 0);
      }
    }
  }
}

template <int kCg, typename TOut, int kAdd, bool kPlain, bool kPool = false, bool kFuse = false>
__global__ void __launch_bounds__(kFuse ? kGemmThreadsFused : kGemmThreads, 1)
k_gemm512(const __grid_constant__ GemmParams p) {
  static_assert(!kFuse || (sizeof(TOut) == 2 && kCg == 2), "fused SAGE layer: 16-bit activations, CTA pairs");
  using Cfg = GemmCfg<kCg>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t gemm_smem_raw[];
  const uint32_t smem_base = (smem_u32(gemm_smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B wants 1024 B
  uint8_t* smem_gen = gemm_smem_raw + (smem_base - smem_u32(gemm_smem_raw));
  const uint32_t stages_u32 = smem_base;
  const uint32_t epi_u32 = stages_u32 + kStages * Cfg::kStageBytes;
  constexpr int kEpiBytes = kEpiWarps * kEpiStageBytes;
  uint8_t* epi_gen = smem_gen + kStages * Cfg::kStageBytes;
  const uint32_t bars_u32 = epi_u32 + kEpiBytes;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_gen + kEpiBytes + 96);
  auto full_bar = [&](int s) { return bars_u32 + 8u * s; };
  auto empty_bar = [&](int s) { return bars_u32 + 8u * (kStages + s); };
  const uint32_t tmem_full_bar = bars_u32 + 8u * (2 * kStages);
  const uint32_t tmem_empty_bar = bars_u32 + 8u * (2 * kStages + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);    // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (kCg == 2) ? cluster_ctarank() : 0u;
  const int tile0 = (kCg == 2) ? (blockIdx.x >> 1) : blockIdx.x;
  const int tile_stride = (kCg == 2) ? (gridDim.x >> 1) : gridDim.x;

  // ---- one-time setup
  if (warp == 0 && elect_one()) {
    for (int s = 0; s < p.n_seg; ++s) { if (!(kFuse && s == 0)) tma_prefetch_desc(&p.seg[s].a); tma_prefetch_desc(&p.seg[s].b); }
  }
  if (kCg == 2) cluster_sync();                // both CTAs resident before the paired TMEM alloc
  if (warp == 1) {
    if (elect_one()) {
      // (kFuse: a stage is full when the producer AND the gather warps of both CTAs have delivered -- the gather warps
      //  arrive on every use of every slot, also where they have nothing to write: see fused_gather)
      for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), kFuse ? 2 * kCg : kCg); mbar_init(empty_bar(s), 1); }
      mbar_init(tmem_full_bar, 1);
      mbar_init(tmem_empty_bar, kCg * 256);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<kCg>(smem_u32(tmem_slot), 512);
  }
  tc_fence_before();
  if (kCg == 2) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (kFuse && warp >= kEpiFirstWarp + kEpiWarps) {
    // ================================================================ gather warps 12..15 (fused SAGE layer)
    // register budget per warpgroup (setmaxnreg acts on whole warpgroups): 128 * (40 + 192 + 192 + 88) = 65536
    if constexpr (kFuse) {
      setmaxnreg_dec<88>();
      fused_gather<kCg, TOut, kStages, Cfg::kStageBytes>(p, warp - (kEpiFirstWarp + kEpiWarps), stages_u32, bars_u32, rank,
                                                         tile0, tile_stride);
    }
  } else if (warp < kEpiFirstWarp) {
    setmaxnreg_dec<40>();
    if (warp == 0) {
      // ================================================================ TMA producer
      if (elect_one()) {
        BG_PROF_DECL
        uint32_t stage = 0, phase = 0;
        for (int tile = tile0; tile < p.n_tiles; tile += tile_stride) {
          const int32_t row0 = tile * (kTileM * kCg) + (int32_t)rank * kTileM;
          const int32_t brow0 = p.b_group_tiles ? (tile / p.b_group_tiles) * kHidden : 0;   // split-K groups
          if (kCg == 2 && p.mn_major) {
            const void* map_a = &p.seg[0].a;
            const void* map_b = &p.seg[0].b;
            const int32_t cols_per_box = p.k_elems_per_block == 64 ? 64 : 32;       // 128 B of the MN dimension
            const int32_t boxes = 128 / cols_per_box;                                 // boxes per 128 MN values
            const uint32_t box_bytes = (uint32_t)p.k_elems_per_block * 128u;
            const int32_t a_col0 = (tile & 1) * 256 + (int32_t)rank * 128;
            const int64_t node_base = (int64_t)(tile >> 1) * p.chunk_k;
            for (int kb = 0; kb < p.kblocks[0]; ++kb) {
              mbar_wait(empty_bar(stage), phase ^ 1u, kTagEmpty);
              const uint32_t sa = stages_u32 + stage * Cfg::kStageBytes;
              const uint32_t sb = sa + kATileBytes;
              const int32_t node0 = (int32_t)(node_base + (int64_t)kb * p.k_elems_per_block);
              if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
              else mbar_arrive_cluster(full_bar(stage), 0);
              for (int j = 0; j < boxes; ++j)
                tma_load_2d_cg2(sa + j * box_bytes, map_a, full_bar(stage), a_col0 + j * cols_per_box, node0);
              for (int h = 0; h < 2; ++h)
                for (int j = 0; j < boxes; ++j)
                  tma_load_2d_cg2(sb + h * (Cfg::kBTileBytes / 2) + j * box_bytes, map_b, full_bar(stage),
                                  h * 256 + (int32_t)rank * 128 + j * cols_per_box, node0);
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
            continue;
          }
          for (int s = 0; s < p.n_seg; ++s) {
            const void* map_a = &p.seg[s].a;
            const void* map_b = &p.seg[s].b;
            for (int kb = 0; kb < p.kblocks[s]; ++kb) {
              BG_PROF_T0();
              mbar_wait(empty_bar(stage), phase ^ 1u, kTagEmpty);
              BG_PROF_ADD(_pacc_a);
              const uint32_t sa = stages_u32 + stage * Cfg::kStageBytes;
              const uint32_t sb = sa + kATileBytes;
              const int32_t k0 = kb * p.k_elems_per_block;
              if constexpr (kCg == 1) {
                mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
                tma_load_2d(sa, map_a, full_bar(stage), k0, row0);
#pragma unroll
                for (int j = 0; j < Cfg::kBRows / 128; ++j)
                  tma_load_2d(sb + j * 16384, map_b, full_bar(stage), k0, brow0 + j * 128);
              } else {
                // timing probes (wrong results): -DBG_PROBE_NO_B / -DBG_PROBE_NO_A load the weight / activation stages
                // of a CTA's FIRST tile only -- what the kernel would cost if that operand's L2 (and HBM) traffic were free
#ifdef BG_PROBE_NO_B
                const bool load_b = tile == tile0;
#else
                constexpr bool load_b = true;
#endif
#ifdef BG_PROBE_NO_A
                const bool load_a = tile == tile0;
#else
                constexpr bool load_a = true;
#endif
                const bool gathered = kFuse && s == 0;                  // this stage's A half comes from the gather warps
                if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * ((load_a && !gathered ? kATileBytes : 0) + (load_b ? Cfg::kBTileBytes : 0)));
                else mbar_arrive_cluster(full_bar(stage), 0);
                if (load_a && !gathered) tma_load_2d_cg2(sa, map_a, full_bar(stage), k0, row0);
                if (load_b) {
#pragma unroll
                  for (int j = 0; j < Cfg::kBRows / 128; ++j)      // N half j: weight rows j*256 + rank*128
                    tma_load_2d_cg2(sb + j * 16384, map_b, full_bar(stage), k0, brow0 + j * 256 + (int32_t)rank * 128);
                }
              }
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
          }
        }
        BG_PROF_STORE(0, _pacc_a);
      }
      __syncwarp();
    } else if (warp == 1 && rank == 0) {
      // ================================================================ MMA issuer (leader CTA)
      const uint32_t idesc = umma_idesc(p.a_fmt, p.b_fmt, kTileM * kCg, 256) | (p.mn_major ? (3u << 15) : 0u);   // a_major, b_major
      const bool tf32 = p.a_fmt == 2;
      const uint32_t mn_lbo = (uint32_t)p.k_elems_per_block * 128u;       // bytes between 128-byte MN column blocks
      const uint32_t mn_kstep = tf32 ? 1024u : 2048u;                     // one UMMA K step = 8 / 16 K rows of 128 B
      BG_PROF_DECL
      uint32_t stage = 0, phase = 0, it = 0;
      for (int tile = tile0; tile < p.n_tiles; tile += tile_stride, ++it) {
        BG_PROF_T0();
        mbar_wait(tmem_empty_bar, (it & 1u) ^ 1u, kTagTmemEmpty);   // epilogue drained the accumulator
        BG_PROF_ADD(_pacc_b);
        tc_fence_after();
        uint32_t first = 1;
        for (int s = 0; s < p.n_seg; ++s) {
          const int nkb = p.kblocks[s];
          for (int kb = 0; kb < nkb; ++kb) {
            BG_PROF_T0();
            mbar_wait(full_bar(stage), phase, kTagFull);
            BG_PROF_ADD(_pacc_a);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t sa = stages_u32 + stage * Cfg::kStageBytes;
              const uint32_t sb = sa + kATileBytes;
#pragma unroll
              for (int k = 0; k < kStageKBytes / 32; ++k) {
                const uint64_t da = p.mn_major ? umma_smem_desc_mn(sa + k * mn_kstep, mn_lbo, tf32) : umma_smem_desc(sa + k * 32);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const uint64_t db = p.mn_major ? umma_smem_desc_mn(sb + h * (Cfg::kBTileBytes / 2) + k * mn_kstep, mn_lbo, tf32)
                                                 : umma_smem_desc(sb + h * (Cfg::kBTileBytes / 2) + k * 32);
                  const uint32_t acc = (first && k == 0) ? 0u : 1u;
                  if (tf32) umma<kCg, true>(tmem_base + h * 256, da, db, idesc, acc);
                  else umma<kCg, false>(tmem_base + h * 256, da, db, idesc, acc);
                }
              }
              umma_commit<kCg>(empty_bar(stage));                   // smem slot reusable once these MMAs retire
              if (s == p.n_seg - 1 && kb == nkb - 1) umma_commit<kCg>(tmem_full_bar);
            }
            __syncwarp();
            first = 0;
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
#ifdef BG_PROFILE
      if (lane == 0) { BG_PROF_STORE(1, _pacc_a); BG_PROF_STORE(2, _pacc_b); BG_PROF_STORE(3, clock64() - _pstart); }
#endif
      __syncwarp();
    }
  } else if (warp < kEpiFirstWarp + kEpiWarps) {
    // ================================================================ epilogue warps 4..11
    if constexpr (kFuse) setmaxnreg_inc<192>(); else setmaxnreg_inc<232>();
    const int ew = warp - kEpiFirstWarp;
    const EpiCtx cx{tmem_base, epi_u32 + (uint32_t)ew * kEpiStageBytes,
                    reinterpret_cast<float*>(epi_gen + ew * kEpiStageBytes),
                    reinterpret_cast<const float*>(epi_gen + (ew ^ 4) * kEpiStageBytes),
                    tmem_full_bar, tmem_empty_bar, rank, tile0, tile_stride};
    epilogue_warp<kCg, TOut, kAdd, kPlain, kPool>(p, cx, ew >> 2);
  }
  static_assert(kEpiFirstWarp + kEpiWarps + kGatherWarps == kGemmThreadsFused / 32, "fused warp roles");

  // ---- teardown
  tc_fence_before();
  if (kCg == 2) cluster_sync(); else __syncthreads();
  if (warp == 1) tmem_dealloc<kCg>(tmem_base, 512);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                             const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_tensorMapEncodeTiled get_tensor_map_encoder();
int gemm_sm_count();

#ifndef BG_TMA_L2_PROMOTION
#define BG_TMA_L2_PROMOTION 3      // CU_TENSOR_MAP_L2_PROMOTION_L2_256B (0 none, 1 64 B, 2 128 B): an operand box is 128 B of
#endif                             // each of 128 rows; the next K block reads the adjacent 128 B of the same rows
// [rows, k] row-major matrix, box = 128 rows x 128 bytes of K, 128B swizzle, zero fill out of bounds
// (the same geometry serves the operand tiles and the epilogue's out / residual staging tiles)
// fmt: UMMA operand format (0 f16, 1 bf16, 2 tf32/f32)
static inline int make_operand_map(CUtensorMap* map, const void* base, int64_t rows, int64_t k, int64_t ld,
                                   uint32_t fmt) {
  PFN_tensorMapEncodeTiled enc = get_tensor_map_encoder();
  if (!enc) return BG_ERR_CUDA;
  const bool tf32 = fmt == 2;
  const int esz = tf32 ? 4 : 2;
  const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                      : (fmt == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
  cuuint32_t box[2] = {(cuuint32_t)(kStageKBytes / esz), 128u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(map, dt, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, (CUtensorMapL2promotion)BG_TMA_L2_PROMOTION,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? BG_OK : BG_ERR_CUDA;
}

// [n_rows, 512] row-major matrix read as an MN-major operand: box = 128 B of columns x `kblk` rows, 128B swizzle,
// rows beyond n_rows read as zero (the node dimension needs no padding)
static inline int make_mn_operand_map(CUtensorMap* map, const void* base, int64_t n_rows, int64_t n_cols, int64_t ld,
                                      uint32_t fmt, int kblk) {
  PFN_tensorMapEncodeTiled enc = get_tensor_map_encoder();
  if (!enc) return BG_ERR_CUDA;
  const bool tf32 = fmt == 2;
  const int esz = tf32 ? 4 : 2;
  const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                      : (fmt == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
  cuuint64_t dims[2] = {(cuuint64_t)n_cols, (cuuint64_t)n_rows};     // columns beyond n_cols read as zero too
  cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)kblk};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   tf32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? BG_OK : BG_ERR_CUDA;
}

template <int kCg, typename TOut, int kAdd, bool kPlain, bool kPool = false, bool kFuse = false>
static int launch_gemm512(const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<kCg>;
  auto kern = k_gemm512<kCg, TOut, kAdd, kPlain, kPool, kFuse>;
  static bool attr_set = false;
  if (!attr_set) {
    BG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int sms = gemm_sm_count();
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  if (kCg == 1) {
    cfg.gridDim = dim3((unsigned)min(p.n_tiles, sms));
    cfg.numAttrs = 0;
  } else {
    cfg.gridDim = dim3((unsigned)(2 * min(p.n_tiles, sms / 2)));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.numAttrs = 1;
  }
  cfg.attrs = attr;
  cfg.blockDim = dim3(kFuse ? kGemmThreadsFused : kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  BG_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p));
  return BG_OK;
}

}  // namespace bg
