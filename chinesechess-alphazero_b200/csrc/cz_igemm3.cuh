// cz_igemm3.cuh — CTA-pair implicit-GEMM 3x3 convolution (same mainloop as igemm::k_igemm2) with an ALL-TMA epilogue.
//
// Why: ncu of k_igemm2 on three shapes (profiles/r02a_*) fits one model: the SM's 128 B/cycle shared-memory data pipe is shared by
// the TMA operand fills, the UMMA operand reads AND every LSU wavefront of the epilogue.
//     per M-tile of one CTA:   C=256 conv1  18.4k (operands) + 2.5k (epilogue) + ~2.6k fixed  ~ 21.0k cycles measured
//                              C=256 conv2  18.4k + 6-8k (fp32 skip in, fp32 + fp16 out)       ~ 24.4k
//                              C=128 conv1   6.9k + 2-3k                                       ~ 12.1k  (tensor pipe 37 %)
// The operand traffic is what the MMA needs; the epilogue's share is ours to cut.  k_igemm2 moved every element of the tile
// through the pipe up to 8 times (coalesced ldg -> sts -> lds row -> sts -> lds -> stg fp32, lds -> stg fp16, with 30 % bank
// conflicts on top).  Here each element passes at most 4 times and no LSU global access is left:
//     skip stream  : TMA load  -> swizzled smem tile -> ONE conflict-free lds per thread-row
//     fp32 output  : written in place over the skip tile -> TMA store
//     fp16 output  : packed into a swizzled smem tile   -> TMA store
// Per epilogue warp (8 per CTA: TMEM lane quarter x column half) and 16-column chunk: a ring of 32x16 fp32 tiles F (64-byte
// rows, SWIZZLE_64B) and two 32x16 fp16 tiles H (32-byte rows, SWIZZLE_32B): the stores of chunk i overlap the math of chunk
// i+1 and the skip tiles are requested nf-2 chunks ahead.  The tiles are small on purpose: the operand ring needs every
// kilobyte (a C=128 k-block is consumed in 256 cycles, so hiding a ~2k-cycle TMA fill takes 8 stages; ncu on the first
// 32-column version showed the MMA starving with 6).  Thread r of the warp owns accumulator row r (tcgen05.ld 32x32b), i.e. one
// smem row: with the TMA swizzle a warp-wide 16-byte access touches every bank group exactly 4 times = the 4-wavefront
// minimum for 512 bytes.
#pragma once
#include "cz_igemm.cuh"

namespace igemm {

constexpr int kMaxStages3 = 9;
constexpr int kSmemLimit3 = 232448;                            // opt-in dynamic shared memory per CTA on sm_100
constexpr int kChunkCols3 = 16;                                // columns per epilogue chunk
constexpr int kHBytes3 = 32 * kChunkCols3 * 2;                 // fp16 output tile of one chunk: 32 rows x 32 B (SWIZZLE_32B)
constexpr int kNH3 = 2;                                        // H ring
constexpr int kMaxNF3 = 4;                                     // F ring (skip tiles in flight / fp32 output tiles being stored)

struct Args3 {
  Args a;
  int stages;            // smem ring depth (<= kMaxStages3), chosen on the host from what fits beside the epilogue tiles
  int skip_mode;         // 0 none, 1 fp16 (tmSkip = fp16 map), 2 fp32 (tmSkip = fp32 map)
  int out32;             // also store the fp32 copy (tmOut32)
  int fbytes;            // bytes of one F tile: 2048 (fp32 skip and/or fp32 output), 1024 (fp16 skip only), 0 (neither)
  int nf;                // F tiles per epilogue warp (3 .. kMaxNF3): skip loads run nf - 2 chunks ahead
  int split_producer;    // 1: warp 0 issues the A (im2col) loads, warp 3 the B (weight) loads — two TMA issue streams per CTA
  int n_split;           // N tiles of N_TILE columns (1 = the tile is the whole width).  Small batches (a single game's leaves)
                         // run as pairs x n_split work items so that a launch of a few M-tiles still spreads over many SMs and
                         // each CTA's exposed epilogue is N_TILE / 16 chunks instead of C / 16; the K order per output is unchanged,
                         // so the numbers are bit for bit those of the full-width tile
};
__host__ __device__ constexpr int epi3_warp_bytes(int fbytes, int nf) { return nf * fbytes + kNH3 * kHBytes3; }

// MT = M-tiles per CTA that share one weight stage (cta_group::2: the pair computes 2*MT tiles per pass).  MT = 2 needs
// 2 (double buffer) x 2 x N_TILE TMEM columns <= 512, i.e. N_TILE <= 128 — exactly the shapes whose weight tile is too narrow
// to amortise its fill: per k-block the pipe then moves 2 x 16 KB of pixels + 8 KB of weights for TWICE the MMA work
// (40 KB per 128x128x64 instead of 48 KB), the weight TMA count halves, and one stage feeds 512 instead of 256 cycles of MMA,
// so the ring covers a TMA round trip with 5 stages where the single-tile form starved with 8.
template <int N_TILE, int MT = 1>
struct Cfg3 {
  static constexpr int kBHalfBytes = (N_TILE / 2) * 128;
  static constexpr int kStageBytes = MT * kAStageBytes + kBHalfBytes;
  static constexpr int kTmemCols = (2 * MT * N_TILE <= 32) ? 32 : (2 * MT * N_TILE <= 64) ? 64 : (2 * MT * N_TILE <= 128) ? 128
                                   : (2 * MT * N_TILE <= 256) ? 256 : 512;
  static_assert(2 * MT * N_TILE <= 512, "accumulators (double-buffered) must fit TMEM");
  static constexpr int smem_bytes(int stages, int fbytes, int nf) { return stages * kStageBytes + kEpiWarps2 * epi3_warp_bytes(fbytes, nf) + 512 + 1024; }
  static int max_stages(int fbytes, int nf) {
    int s = (kSmemLimit3 - 512 - 1024 - kEpiWarps2 * epi3_warp_bytes(fbytes, nf)) / kStageBytes;
    return s > kMaxStages3 ? kMaxStages3 : s;
  }
  static_assert(N_TILE % 64 == 0 && N_TILE <= 256, "column halves must be multiples of 32");
};

// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* t, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(t)), "r"(umma::smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// PAIRS = CTA pairs per cluster.  PAIRS = 2 (experiment, CZ_CLUSTER4=1): the two pairs of a 4-CTA cluster walk their tiles in lockstep
// and share every weight stage — pair 0's CTAs load their halves of the weight tile with TMA multicast into both pairs' shared
// memory — so a CTA pulls 16 + 8 KB per k-block instead of 16 + 16.  Every CTA's `empty` barrier then waits for BOTH pairs' MMA
// commits (a stage is refilled only when neither pair reads it any more); everything else stays per pair.
template <int N_TILE, int MT = 1, int PAIRS = 1>
__global__ void __cluster_dims__(2 * PAIRS, 1, 1) __launch_bounds__(kThreads2, 1)
k_igemm3(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut16,
         const __grid_constant__ CUtensorMap tmSkip, const __grid_constant__ CUtensorMap tmOut32, const Args3 p) {
  using C = Cfg3<N_TILE, MT>;
  const Args& a = p.a;
  const int n_stages = p.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* epi = smem + n_stages * C::kStageBytes;                 // 1024-byte aligned: stage sizes are multiples of 1024
  const int warp_bytes = epi3_warp_bytes(p.fbytes, p.nf);          // multiple of 1024
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi + kEpiWarps2 * warp_bytes);
  uint64_t* full = bars;                       // [kMaxStages3]  (used in the leader only)
  uint64_t* empty = bars + kMaxStages3;        // [kMaxStages3]  per CTA, signalled by multicast commit
  uint64_t* tfull = bars + 2 * kMaxStages3;    // [2]
  uint64_t* tempty = tfull + 2;                // [2]  leader: 512 arrivals (8 epilogue warps of both CTAs)
  uint64_t* skipbar = tempty + 2;              // [8][kMaxNF3] per epilogue warp: skip tile landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(skipbar + kMaxNF3 * kEpiWarps2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = umma::cluster_ctarank();          // rank in the cluster
  const uint32_t rank = crank & 1u;                        // rank in the CTA pair
  const uint32_t pp = crank >> 1;                          // pair in the cluster
  const uint32_t lrank = crank & ~1u;                      // cluster rank of this pair's leader
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    umma::prefetch_tmap(&tmA);
    umma::prefetch_tmap(&tmB);
    umma::prefetch_tmap(&tmOut16);
    if (p.skip_mode) umma::prefetch_tmap(&tmSkip);
    if (p.out32) umma::prefetch_tmap(&tmOut32);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < n_stages; ++s) { umma::mbar_init(&full[s], 1); umma::mbar_init(&empty[s], PAIRS); }
    for (int i = 0; i < 2; ++i) { umma::mbar_init(&tfull[i], 1); umma::mbar_init(&tempty[i], 512); }
    for (int i = 0; i < kMaxNF3 * kEpiWarps2; ++i) umma::mbar_init(&skipbar[i], 1);
    umma::fence_barrier_init();
    umma::fence_proxy_async();
  }
  if (warp == 2) umma::tmem_alloc2<C::kTmemCols>(tmem_slot);
  umma::griddep_launch_dependents();     // (PDL launches only) the next conv may set up its barriers / TMEM while this one runs
  umma::tc_fence_before();
  __syncthreads();
  umma::cluster_sync_all();
  umma::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  umma::griddep_wait();                  // (PDL launches only) the producer of this conv's input has completed; nothing above
                                         // touched global memory other than the tensor maps in the parameter space

  const int n_kb = a.n_taps * a.k_chunks;
  const int rows = args_rows(a);
  const int m_tiles = a.n_dev ? (rows + kTileM - 1) / kTileM : a.m_tiles;
  const int pairs = (m_tiles + 2 * MT - 1) / (2 * MT);     // one pass of a CTA pair = 2 * MT consecutive M-tiles
  const int ns = p.n_split > 1 ? p.n_split : 1;            // work item = (pair, N tile): item / ns, item % ns
  const int items = pairs * ns;                            // (PAIRS = 2 runs with ns = 1)
  const int items_c = (items + PAIRS - 1) / PAIRS;         // per cluster step every pair takes one item; the last may lie past the
                                                           // end: an all-out-of-bounds tile (zero-filled loads, clipped stores)
  const int n_clusters = gridDim.x / (2 * PAIRS), cluster_id = blockIdx.x / (2 * PAIRS);

  if (warp == 0 || (warp == 3 && p.split_producer)) {
    // ------------------------------------------------------------ TMA producer(s): one thread per CTA issues both operand
    // loads of a stage, or — split_producer — warp 0 the A tiles and warp 3 the B tiles (each waits for the stage to be free
    // on its own; the transaction count covers both)
    if (lane == 0) {
      const bool do_a = warp == 0, do_b = warp == 3 || !p.split_producer;
      uint32_t s = 0, ph = 0;
      for (int it = cluster_id; it < items_c; it += n_clusters) {
        const int item = it * PAIRS + (int)pp;
        const int pair = item / ns, n0 = (item % ns) * N_TILE;
        const int m_tile = MT * (2 * pair + (int)rank);     // this CTA's first tile of the pass
        for (int tap = 0; tap < a.n_taps; ++tap) {
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          for (int kc = 0; kc < a.k_chunks; ++kc) {
            umma::mbar_wait(&empty[s], ph ^ 1);
            uint8_t* sA = smem + s * C::kStageBytes;
            uint8_t* sB = sA + MT * kAStageBytes;
            if (leader && do_a) umma::mbar_expect_tx(&full[s], 2u * ((uint32_t)MT * a.a_bytes + (uint32_t)C::kBHalfBytes));
            if (do_a) {
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                const int pix0 = (m_tile + mt) * kTileM, img0 = pix0 / 90, row0 = (pix0 % 90) / 9, col0 = pix0 % 9;
                umma::tma2_load_im2col_4d_r(sA + mt * kAStageBytes, &tmA, &full[s], kc * kBlockK, col0 - 1, row0 - 1, img0, (uint16_t)(dx + 1), (uint16_t)(dy + 1), lrank);
              }
            }
            if (do_b) {
              if (PAIRS == 1) umma::tma2_load_2d(sB, &tmB, &full[s], kc * kBlockK, tap * a.n_total + n0 + (int)rank * (N_TILE / 2));
              else if (pp == 0)                             // this half of the weight tile -> the CTAs of this rank in both pairs
                umma::tma2_load_2d_mc(sB, &tmB, &full[s], kc * kBlockK, tap * a.n_total + n0 + (int)rank * (N_TILE / 2),
                                      (uint16_t)((1u << rank) | (1u << (rank + 2))));
            }
            if (++s == (uint32_t)n_stages) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma::idesc_f16(256, N_TILE);
      uint32_t s = 0, ph = 0, tcount = 0;
      for (int it = cluster_id; it < items_c; it += n_clusters, ++tcount) {
        const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
        umma::mbar_wait(&tempty[acc], aph ^ 1);
        umma::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (MT * N_TILE);
        for (int kb = 0; kb < n_kb; ++kb) {
          umma::mbar_wait(&full[s], ph);
          umma::tc_fence_after();
          const uint32_t sA = umma::smem_u32(smem + s * C::kStageBytes);
          const uint64_t db = umma::smem_desc_sw128(sA + MT * kAStageBytes);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {                 // the same weight stage against MT pixel tiles / accumulators
            const uint64_t da = umma::smem_desc_sw128(sA + mt * kAStageBytes);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma::mma2_f16_ss(d_tmem + mt * N_TILE, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          }
          umma::mma2_commit_mask(&empty[s], PAIRS == 1 ? (uint16_t)3 : (uint16_t)0xF);
          if (++s == (uint32_t)n_stages) { s = 0; ph ^= 1; }
        }
        umma::mma2_commit_mask(&tfull[acc], (uint16_t)(3u << (2 * pp)));
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue: TMEM -> registers -> swizzled smem tiles -> TMA
    // Chunk = 16 columns.  Per warp: F ring of nf tiles (fp32 32x16 = 64-byte rows, SWIZZLE_64B; or fp16 skip 32-byte rows,
    // SWIZZLE_32B) and an H ring of 2 fp16 tiles (32-byte rows, SWIZZLE_32B).  The skip tile of chunk i + nf - 2 is requested at
    // the top of chunk i — its F slot was last read by the store of chunk i - 2, which bulk_wait_read<1> has seen finish —
    // so a TMA round trip is hidden behind nf - 2 chunks of work instead of being paid once per chunk.
    const int ew = warp - 4;
    const int q = warp & 3;                                   // TMEM lane quarter (rows q*32 .. q*32+31 of the CTA's M-tile)
    const int half = ew >> 2;                                 // column half
    constexpr int kChunks = (N_TILE / 2) / kChunkCols3;       // chunks per half
    const int cbeg = half * (N_TILE / 2);
    const int fb = p.fbytes, nf = p.nf;
    uint8_t* F = epi + ew * warp_bytes;                       // F[nf]: fb bytes each
    uint8_t* H = F + nf * fb;                                 // H[2]: 1024 B each
    uint64_t* sbar = skipbar + kMaxNF3 * ew;
    const uint32_t tempty_remote[2] = {umma::mapa_shared(&tempty[0], lrank), umma::mapa_shared(&tempty[1], lrank)};
    const int r3 = (lane >> 1) & 3;                           // SWIZZLE_64B key of this thread's 64-byte row
    const int r1 = (lane >> 2) & 1;                           // SWIZZLE_32B key of this thread's 32-byte row
    const bool has_skip = p.skip_mode != 0, skip32 = p.skip_mode == 2;
    const uint32_t skip_bytes = skip32 ? 2048u : 1024u;
    // next tile's skip block -> L2, a whole tile ahead (one warp): the per-chunk TMA loads below then hit L2
    auto prefetch_skip = [&](int item) {
      if (ew != 0 || item >= items || !has_skip || ns > 1) return;   // (split launches are a few tiles: nothing to run ahead of)
      const int pr = item;
      const long long row0 = (long long)MT * (2 * pr + (int)rank) * kTileM;
      const long long nrow = rows - row0 < MT * kTileM ? rows - row0 : MT * kTileM;
      if (nrow <= 0) return;
      const size_t esz = skip32 ? 4 : 2;
      const char* base = (skip32 ? reinterpret_cast<const char*>(a.residual32) : reinterpret_cast<const char*>(a.residual)) + (size_t)row0 * a.ldo * esz;
      const size_t total = (size_t)nrow * a.ldo * esz;
      for (size_t off = (size_t)lane * 16384; off < total; off += 32 * 16384)
        umma::l2_prefetch_bulk(base + off, (uint32_t)(total - off < 16384 ? total - off : 16384));
    };
    // global chunk counter of this warp: chunk index g -> (tile = g / kChunks, chunk in tile = g % kChunks); F slot g % nf
    int my_tiles = 0;
    for (int it = cluster_id; it < items_c; it += n_clusters) ++my_tiles;
    const uint32_t total_chunks = (uint32_t)my_tiles * (MT * kChunks);
    auto request_skip = [&](uint32_t g) {                     // lane 0 only
      if (!has_skip || g >= total_chunks) return;
      const int t = (int)(g / (MT * kChunks)), rem = (int)(g % (MT * kChunks)), mt = rem / kChunks, ch = rem % kChunks;
      const int item = (cluster_id + t * n_clusters) * PAIRS + (int)pp;
      const int row = (MT * (2 * (item / ns) + (int)rank) + mt) * kTileM + q * 32;
      const uint32_t slot = g % (uint32_t)nf;
      umma::mbar_expect_tx(&sbar[slot], skip_bytes);
      umma::tma_load_2d(F + slot * fb, &tmSkip, &sbar[slot], (item % ns) * N_TILE + cbeg + ch * kChunkCols3, row);
    };
    uint32_t tcount = 0, g = 0;
    prefetch_skip(cluster_id * PAIRS + (int)pp);
    if (lane == 0)
      for (int k = 0; k < nf - 2; ++k) request_skip((uint32_t)k);   // prime the ring: chunks 0 .. nf-3
    for (int it = cluster_id; it < items_c; it += n_clusters, ++tcount) {
      const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
      const int item = it * PAIRS + (int)pp;
      const int pair = item / ns, n0 = (item % ns) * N_TILE;
      prefetch_skip(item + n_clusters * PAIRS);
      umma::mbar_wait(&tfull[acc], aph);
      umma::tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
      const int m_tile = MT * (2 * pair + (int)rank) + mt;
      const int rbase = m_tile * kTileM + q * 32;             // first global pixel row of this warp
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (acc * MT + mt) * N_TILE + cbeg;
#pragma unroll 1
      for (int ch = 0; ch < kChunks; ++ch, ++g) {
        const uint32_t slot = g % (uint32_t)nf, sph = (g / (uint32_t)nf) & 1, hb = g & 1;
        const int c0 = n0 + cbeg + ch * kChunkCols3;
        if (lane == 0) {
          // At most the store group of chunk g-1 may still be reading its tiles; chunk g-2 and older are done, so H[g & 1]
          // and the F slot of chunk g-2 — which is the slot of chunk g + nf - 2 — are free.
          bulk_wait_read<1>();
          request_skip(g + (uint32_t)nf - 2);
        }
        uint32_t v[16];
        tmem_ld_32x16(t_row + ch * kChunkCols3, v);
        __syncwarp();                                         // H[hb] is free for everybody (lane 0 waited above)
        if (has_skip) umma::mbar_wait(&sbar[slot], sph);      // skip tile of this chunk has landed in F[slot]
        uint8_t* frow = F + slot * fb + lane * 64;
        uint8_t* hrow = H + hb * kHBytes3 + lane * 32;
        const float4* bp = reinterpret_cast<const float4*>(a.bias + c0);
#pragma unroll
        for (int gq = 0; gq < 4; gq += 2) {
          float4 x0, x1;
          {
            const float4 b0 = __ldg(bp + gq), b1 = __ldg(bp + gq + 1);
            x0 = make_float4(__uint_as_float(v[4 * gq]) + b0.x, __uint_as_float(v[4 * gq + 1]) + b0.y,
                             __uint_as_float(v[4 * gq + 2]) + b0.z, __uint_as_float(v[4 * gq + 3]) + b0.w);
            x1 = make_float4(__uint_as_float(v[4 * gq + 4]) + b1.x, __uint_as_float(v[4 * gq + 5]) + b1.y,
                             __uint_as_float(v[4 * gq + 6]) + b1.z, __uint_as_float(v[4 * gq + 7]) + b1.w);
          }
          if (skip32) {
            const float4 s0 = *reinterpret_cast<const float4*>(frow + ((gq ^ r3) << 4));
            const float4 s1 = *reinterpret_cast<const float4*>(frow + (((gq + 1) ^ r3) << 4));
            x0.x += s0.x; x0.y += s0.y; x0.z += s0.z; x0.w += s0.w;
            x1.x += s1.x; x1.y += s1.y; x1.z += s1.z; x1.w += s1.w;
          } else if (has_skip) {                              // fp16 skip tile: 32-byte rows at the start of F[slot]
            const uint4 sv = *reinterpret_cast<const uint4*>(F + slot * fb + lane * 32 + (((gq >> 1) ^ r1) << 4));
            const __half2* h = reinterpret_cast<const __half2*>(&sv);
            const float2 f0 = __half22float2(h[0]), f1 = __half22float2(h[1]), f2 = __half22float2(h[2]), f3 = __half22float2(h[3]);
            x0.x += f0.x; x0.y += f0.y; x0.z += f1.x; x0.w += f1.y;
            x1.x += f2.x; x1.y += f2.y; x1.z += f3.x; x1.w += f3.y;
          }
          if (a.relu) {
            x0.x = fmaxf(x0.x, 0.f); x0.y = fmaxf(x0.y, 0.f); x0.z = fmaxf(x0.z, 0.f); x0.w = fmaxf(x0.w, 0.f);
            x1.x = fmaxf(x1.x, 0.f); x1.y = fmaxf(x1.y, 0.f); x1.z = fmaxf(x1.z, 0.f); x1.w = fmaxf(x1.w, 0.f);
          }
          if (p.out32) {                                      // in place over the skip tile
            *reinterpret_cast<float4*>(frow + ((gq ^ r3) << 4)) = x0;
            *reinterpret_cast<float4*>(frow + (((gq + 1) ^ r3) << 4)) = x1;
          }
          uint4 ov;
          __half2* oh = reinterpret_cast<__half2*>(&ov);
          oh[0] = __floats2half2_rn(x0.x, x0.y); oh[1] = __floats2half2_rn(x0.z, x0.w);
          oh[2] = __floats2half2_rn(x1.x, x1.y); oh[3] = __floats2half2_rn(x1.z, x1.w);
          *reinterpret_cast<uint4*>(hrow + (((gq >> 1) ^ r1) << 4)) = ov;
        }
        umma::fence_proxy_async();                            // generic-proxy writes above -> visible to the TMA engine
        __syncwarp();
        if (lane == 0) {
          if (p.out32) tma_store_2d(&tmOut32, F + slot * fb, c0, rbase);
          tma_store_2d(&tmOut16, H + hb * kHBytes3, c0, rbase);
          bulk_commit();
        }
      }
      }
      umma::tc_fence_before();
      umma::mbar_arrive_cluster(tempty_remote[acc]);
    }
    if (lane == 0) bulk_wait_all<0>();                        // every store of this warp has been written before the CTA exits
  }

  umma::tc_fence_before();
  __syncthreads();
  umma::cluster_sync_all();
  if (warp == 2) {
    umma::tc_fence_after();
    umma::tmem_dealloc2<C::kTmemCols>(tmem_base);
  }
}

}  // namespace igemm
