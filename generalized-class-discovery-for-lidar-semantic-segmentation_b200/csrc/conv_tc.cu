// conv_tc.cu — bf16 tcgen05/TMEM sparse convolution for sm_100a (forward and dgrad share one
// kernel; wgrad is the second kernel).
//
// Forward / dgrad: persistent, warp-specialised, output-stationary implicit GEMM.
//   tile      = 128 output rows x N (<= 256) output channels, fp32 accumulator in TMEM
//               (two accumulator buffers so the epilogue of tile i overlaps the MMAs of tile i+1)
//   iteration = (kernel offset k with at least one hit in the tile) x (64-channel slice of Cin)
//   warps 0-7 : gather producers, each OWNING ring slots (warp w fills the stages of iterations g == w mod stages;
//               a warp pair per slot when the stages are few and fat).  The owner waits for its slot, posts the
//               pre-packed weight slice (arrive.expect_tx + one cp.async.bulk on the TMA engine) and gathers the
//               128 input rows x 64 bf16 itself with 16-byte cp.async (zero fill for missing neighbours) straight into
//               the 128B-swizzled K-major operand image; cp.async.mbarrier.arrive.noinc publishes the stage when the
//               copies have landed (the warp never waits for its own copies).
//   warp 13   : table warp.  Stages the [kv][128] neighbour-table slice of tile t + 1 into the second of two shared
//               buffers while tile t streams, derives the mask of offsets with a hit and the tile's iteration count
//               (table_ready / table_free mbarriers; no CTA-wide barrier at tile boundaries).
//   warp 12   : MMA issuer.  Uniform loop (warp-vote exit) so that stage index, phase and both smem descriptors live in
//               uniform registers; the elected lane issues tcgen05.mma (M=128, N, K=16) per 16 channels by predicate and
//               commits to the stage's "empty" barrier; after the tile's last iteration it commits to "tmem_full".
//   warps 8-11: epilogue.  tcgen05.ld the accumulator (lane == output row), add bias, convert,
//               store the row; then release the accumulator buffer.
// History and measurements (what each of these choices bought, and what did not work): DESIGN.md section 4.1,
// profiles/README.md, tools/ubench/README.md.
#include <stdlib.h>
#include <mutex>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace gcd {
namespace {
using namespace ptx;

constexpr int kTileM = 128;
constexpr int kChunkK = 64;                 // bf16 per 128-byte operand row
constexpr int kRowBytes = 128;
constexpr int kABytes = kTileM * kRowBytes; // 16 KB
constexpr int kProducerWarps = 8;
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kEpilogueThreads = 128;
constexpr int kEpilogueWarp0 = kProducerWarps;        // 8: (8 + i) % 4 == i, the TMEM lane quarter rule
constexpr int kMmaWarp = kProducerWarps + 4;          // 12
constexpr int kTcThreads = kProducerThreads + kEpilogueThreads + 64;
constexpr int kMaxKV = 27;
constexpr int kMaxStages = 8;
constexpr int kTileRing = 32;               // published iteration counts (the MMA warp lags the table warp by < kTableSlots + kMaxStages tiles)
constexpr int kSliceBytes = kTileM * 4;     // one offset's [128] slice of the neighbour table
constexpr int kRingSlices = 56;             // shared-memory ring of table slices: two whole 3x3x3 tiles, ~6 tiles of a sorted table
constexpr int kTableSlots = 8;              // tiles the table warp may run ahead (barriers / masks per slot)
constexpr int kDynAhead = 4;                // ... under a dynamic tile schedule
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;             // TMEM columns between the two accumulator buffers
constexpr int kSmemBudget = 226 * 1024;

struct FwdParams {
  const __nv_bfloat16* in; int64_t ld_in;
  const int32_t* nbr; int kv; int64_t n_out;
  int c_in, c_out;                 // c_in % 16 == 0, c_out % 16 == 0
  int n_tile_cols;                 // columns per N tile (<= 256, % 16 == 0)
  int n_tiles_n;
  const uint8_t* w_packed;         // [kv][nq][c_out rows][128 B] swizzled K-major images
  const float* bias;
  void* out; int64_t ld_out; int out_is_bf16;
  int stages;
  int group;                       // producer warps that share one stage (1, 2, 4, 8)
  const int32_t* out_rows;         // kPerm kernels: output row of table column i (tile-sorted table, tilesort.cu)
  const uint32_t* tile_masks;      // optional: offsets with a hit per 128-column tile of nbr (tilesort.cu); the table warp then stages only those slices
  int32_t* sched;                  // optional: {next tile, finished CTAs}, zero on entry and on exit: tiles are claimed dynamically (see below)
  int dyn_ascending;               // dynamic schedule: 1 = tickets map to tiles in table order, 0 = last (heaviest, in a sorted table) tile first
  int dyn_ahead;                   // dynamic schedule: tiles the table warp may claim ahead of the gather warps (1 .. kTableSlots)
#ifdef GCD_TC_PROFILE
  long long* dbg;
  int ablate;                      // profile build only: 1 = no MMA issue, 2 = no gather copies, 4 = weight copies of 16 B
#endif
};

struct SmemLayout {
  uint32_t a_off, b_off, nbr_off, bar_off, total;
  uint32_t b_bytes;
};
__host__ __device__ inline SmemLayout make_layout(int stages, int n_tile_cols) {
  SmemLayout L;
  L.b_bytes = (uint32_t)n_tile_cols * kRowBytes;
  L.a_off = 0;
  L.b_off = L.a_off + (uint32_t)stages * kABytes;
  L.nbr_off = L.b_off + (uint32_t)stages * L.b_bytes;
  L.bar_off = L.nbr_off + kRingSlices * kSliceBytes;   // ring of table slices
  L.total = L.bar_off + 1024;
  return L;
}

// kNQ = number of 64-channel slices of Cin when known at compile time (1, 2, 3, 4, 6), 0 = runtime loop.
// With kNQ known the slice loop is unrolled and every copy uses an immediate offset from a per-offset base
// pointer, which is what keeps the producer loop at a few dozen instructions per stage.
// kPerm (tile-sorted tables, tilesort.cu): table column i is output row p.out_rows[i]; only the epilogue's store address changes.
// kPW = gather warps (8, or 16: two warps per ring slot; the gather is issue-bound per warp at ~90 cycles per warp-level copy,
// the LSU takes one every ~8): warps [0, kPW) gather, [kPW, kPW + 4) epilogue, kPW + 4 MMA, kPW + 5 table.
// kDyn: dynamic tile schedule (p.sched).  A template parameter, not a run-time flag: with the choice made at run time the MMA
// warp's loop state (stage index, phase, both descriptors) fell out of the uniform registers -- 31 instead of 13 R2UR in the
// SASS, every MMA preceded by five of them -- and the statically scheduled kernel lost 5 % (r2 calls 9-19).
template <int kNQ, bool kPerm = false, int kPW = kProducerWarps, bool kDyn = false>
__global__ void __launch_bounds__((kPW + 6) * 32, 1) conv_fwd_tc_kernel(const FwdParams p) {
  constexpr int kEpi0 = kPW, kMma = kPW + 4, kTab = kPW + 5;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem base is at least 16-byte aligned; the swizzle pattern needs 1024.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const SmemLayout L = make_layout(p.stages, p.n_tile_cols);
  int32_t* s_nbr0 = reinterpret_cast<int32_t*>(smem + L.nbr_off);           // [kRingSlices][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* full_bar = bars;                    // [kMaxStages]
  uint64_t* empty_bar = bars + kMaxStages;      // [kMaxStages]
  uint64_t* tmem_full = bars + 2 * kMaxStages;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint64_t* table_ready = tmem_empty + 2;       // [kTableSlots] the slices of the slot's tile have landed (table warp -> producers)
  uint64_t* table_free = table_ready + kTableSlots;   // [kTableSlots] every producer warp is done with the slot's tile
  uint32_t* s_tmem_base = reinterpret_cast<uint32_t*>(table_free + kTableSlots);
  uint32_t* s_mask = s_tmem_base + 1;           // [kTableSlots] offsets with at least one hit in the slot's tile
  uint32_t* s_base = s_mask + kTableSlots;      // [kTableSlots] ring position of the tile's first slice, bit 31: slices are compacted (mask order)
  int32_t* s_iters = reinterpret_cast<int32_t*>(s_base + kTableSlots);  // [kTileRing]
  int32_t* s_work = s_iters + kTileRing;        // [kTileRing] dynamic schedule: the tile claimed for each sequence number, -1 = no more work
  uint32_t* s_pub = reinterpret_cast<uint32_t*>(s_work + kTileRing);    // dynamic schedule: number of entries of s_work published so far

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();                   // the next kernel of the stream may be scheduled; it waits for this grid before it reads or writes
  const int S = p.stages;
  const int nq = kNQ ? kNQ : (p.c_in + kChunkK - 1) / kChunkK;
  const int last_width = p.c_in - (nq - 1) * kChunkK;      // channels of the last slice (multiple of 16)
  const int64_t tiles_m = (p.n_out + kTileM - 1) / kTileM;
  const int64_t n_work = tiles_m * p.n_tiles_n;
  // Tile schedule.  Static (p.sched == NULL): CTA b takes tiles b, b + grid, ...  Dynamic: the table warp claims the next tile
  // from a global counter and hands its index to the other roles through s_work, heaviest tiles first (a tile-sorted table
  // keeps the rows with many neighbours at its end).  A CTA that becomes resident late -- another kernel (NCCL, a weight
  // gradient on a second stream) holds its SM -- then finds little or nothing left instead of running its fixed share as a
  // second wave, and tiles of very different cost (1 .. 27 offsets in a sorted table) balance themselves.
  constexpr bool dyn = kDyn;
  auto static_work = [&](uint32_t seq) -> int64_t {
    const int64_t w = (int64_t)blockIdx.x + (int64_t)seq * gridDim.x;
    return w < n_work ? w : -1;
  };
  // other roles: wait until the table warp has published entry `seq` (warp-uniform result)
  auto published_work = [&](uint32_t seq) -> int64_t {
    const uint32_t addr = smem_u32(s_pub);
    uint32_t pub;
    do { pub = ld_acquire_shared_u32(addr); } while (!__all_sync(0xffffffffu, pub > seq));
    return (int64_t)s_work[seq & (kTileRing - 1)];
  };

  if (threadIdx.x == 0) {
    // full: one completion-triggered arrival per lane of the owning producer warp + its lane 0's arrive.expect_tx (weights)
    *s_pub = 0;
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 32 * p.group + 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], kEpilogueThreads / 32); }
    // ready: one completion-triggered arrival per lane of the table warp (its cp.async copies) + lane 0's plain arrive
    for (int b = 0; b < kTableSlots; ++b) { mbar_init(&table_ready[b], 33); mbar_init(&table_free[b], kPW); }
    fence_mbar_init();
  }
  if (warp == kMma) { tmem_alloc<kTmemCols>(s_tmem_base); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem_base;
  pdl_wait();                      // everything above overlapped the previous kernel's tail; from here on global memory is touched

  // order in which a tile's active offsets are visited (shared by the gather warps and the weight warp)
  auto rotate = [&](uint32_t mask, uint32_t rot, uint32_t& hi, uint32_t& lo) {
    if (mask == 0) mask = 1;                   // degenerate tile: run offset 0 with all-zero rows
    lo = mask & ((1u << rot) - 1u);
    hi = mask & ~((1u << rot) - 1u);
    if (hi == 0) { hi = lo; lo = 0; }
  };

  if (warp < kPW) {
    // ===================================================================== gather producers (stage owners)
    // Producer warp (or warp pair) w OWNS the stages of iterations g == w (mod PA), PA = min(8 / G, stages): it waits for the slot, posts
    // the weight slice (one bulk copy on the TMA engine) and gathers all 128 rows x 64 channels itself (32 cp.async per
    // lane), then hands the stage over with completion-triggered arrivals.  Up to PA stages are being filled at once
    // and the per-stage fixed costs (barrier wait, index loads, arrive, loop control) are paid by one warp instead of
    // by all eight in lock-step -- measured (tools/ubench/ldgsts_gather.cu): 270 -> ~130 cycles per stage.
    const int chunk = lane & 7;
    const int rsub = lane >> 3;                // this lane's rows are rsub + 4 j, j = 0..31
    const uint32_t d_even = rsub * kRowBytes + ((chunk ^ rsub) << 4);                 // j even: row & 7 == rsub
    const uint32_t d_odd = (rsub + 4) * kRowBytes + ((chunk ^ (rsub + 4)) << 4);      // j odd:  row & 7 == rsub + 4
    const char* col_base = reinterpret_cast<const char*>(p.in + chunk * 8);
    const uint32_t ld_bytes = (uint32_t)(p.ld_in * 2);   // row pitch < 4 GB: one 32 x 32 -> 64 bit multiply-add per source address
    const bool last_active = chunk * 8 < last_width;
    // A last slice of 32 channels (Cin = 32, 96, ...) fills only half of each 128-byte operand row: with the lane mapping above
    // half of the lanes of every copy instruction would idle.  Such a slice is gathered eight rows per instruction instead
    // (lane = 4 chunks x 8 rows): 16 instructions per stage instead of 32, same swizzled image.
    const bool narrow = last_width == 32;
    const int chunk4 = lane & 3, rsub8 = lane >> 2;          // this lane's rows of a narrow slice are rsub8 + 8 j, j = 0..15
    const uint32_t d_narrow = rsub8 * kRowBytes + ((chunk4 ^ rsub8) << 4);
    const char* col_base4 = reinterpret_cast<const char*>(p.in + chunk4 * 8);
    const uint32_t a_base = smem_u32(smem + L.a_off);
    const uint32_t b_base = smem_u32(smem + L.b_off);
    const uint32_t b_bytes = L.b_bytes;
    const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
    // Few, fat stages (wide N tiles: S <= 4): two warps share a stage (16 rows per lane each) so that all eight warps
    // stay busy and a stage is issued in half the time; otherwise one warp per stage.
    const int G = p.group;                     // warps per stage: 1, 2, 4 or 8
    const int grp = warp / G, sub = warp % G;
    const int PA = (kPW / G) < S ? (kPW / G) : S;   // owner groups (<= stages: a waiter may be one phase behind at most)
    const uint32_t lane0 = (lane == 0 && sub == 0) ? 1u : 0u;

    uint32_t st = grp, ph = 0;                 // stage / parity of this group's next owned iteration
    int own_skip = grp < PA ? grp : 0x7fffffff;     // iterations until the next owned one
    uint32_t tile_seq = 0;
#ifdef GCD_TC_PROFILE
    long long prof_wait = 0, prof_iters = 0, prof_table = 0; const long long prof_t0 = clock64();
#endif
    for (;; ++tile_seq) {
#ifdef GCD_TC_PROFILE
      const long long ct0 = clock64();
#endif
      // the table warp stages the slice one tile ahead: no producer-wide synchronisation at tile boundaries, a warp
      // that has issued its last stage of tile t goes straight on to its first stage of tile t + 1
      const uint32_t tb = tile_seq % kTableSlots;
      int64_t work;
      if (dyn) {
        mbar_wait(&table_ready[tb], (tile_seq / kTableSlots) & 1);      // also orders the read of s_work after its publication
        work = s_work[tile_seq & (kTileRing - 1)];
      } else {
        work = static_work(tile_seq);
      }
      if (work < 0) break;
      if (!dyn) mbar_wait(&table_ready[tb], (tile_seq / kTableSlots) & 1);
      const uint32_t s_nbr_addr = smem_u32(s_nbr0);
      const uint32_t tile_mask = s_mask[tb], tile_base = s_base[tb] & 0x7fffffffu;
      const bool compact = (s_base[tb] >> 31) != 0;
      uint32_t mask = tile_mask, lo_mask;
      // per-tile rotation of the offset order (spreads the CTAs' weight reads over the slices; a function of the tile, not of the
      // CTA, so that a tile's summation order does not depend on which CTA claims it)
      rotate(mask, (uint32_t)((work / p.n_tiles_n) * 11) % (uint32_t)p.kv, mask, lo_mask);
      const uint8_t* w_tile = p.w_packed + (int64_t)(work % p.n_tiles_n) * p.n_tile_cols * kRowBytes;
#ifdef GCD_TC_PROFILE
      prof_table += clock64() - ct0;
#endif

      while (mask) {
        const int k = __ffs(mask) - 1;
        mask &= mask - 1;
        if (mask == 0) { mask = lo_mask; lo_mask = 0; }
#pragma unroll
        for (int q = 0; q < (kNQ ? kNQ : 16); ++q) {
          if (!kNQ && q >= nq) break;
          if (own_skip != 0) { --own_skip; continue; }
          own_skip = PA - 1;
#ifdef GCD_TC_PROFILE
          const long long cw0 = clock64();
#endif
          mbar_wait_addr(empty0 + st * 8, ph ^ 1);
#ifdef GCD_TC_PROFILE
          prof_wait += clock64() - cw0; ++prof_iters;
          const uint32_t wb = (p.ablate & 4) ? 16u : b_bytes;
#else
          const uint32_t wb = b_bytes;
#endif
          // weight slice of (k, q): one bulk copy, its bytes join the stage's transaction count
          mbar_arrive_expect_tx_pred(&full_bar[st], wb, lane0);
          bulk_g2s_pred(b_base + st * b_bytes, w_tile + ((int64_t)k * nq + q) * p.c_out * kRowBytes, wb, &full_bar[st], lane0);
#ifdef GCD_TC_PROFILE
          if (!(p.ablate & 2))
#endif
          if (q + 1 == nq && narrow) {
            uint32_t pos = tile_base + (compact ? (uint32_t)__popc(tile_mask & ((1u << k) - 1u)) : (uint32_t)k);
            if (pos >= (uint32_t)kRingSlices) pos -= kRingSlices;
            const uint32_t nb = s_nbr_addr + (pos * kTileM + (uint32_t)rsub8) * 4u;
            const uint32_t a_stage = a_base + st * kABytes + d_narrow;
            const char* src_q = col_base4 + q * (kChunkK * 2);
#pragma unroll
            for (int jb = 0; jb < 16; jb += 8) {
              if (G > 1 && ((jb >> 3) * G) >> 1 != sub) continue;     // two warps per stage: eight of the sixteen copies each
              int r[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) r[j] = lds_s32(nb + (uint32_t)(jb + j) * 32u);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                cp_async_16(a_stage + (uint32_t)(jb + j) * 1024u, src_q + (uint64_t)(uint32_t)max(r[j], 0) * ld_bytes, r[j] >= 0 ? 16u : 0u);
            }
          } else if (q + 1 < nq || last_active) {
            // ring position of offset k's slice: compacted tiles hold their slices in mask order
            uint32_t pos = tile_base + (compact ? (uint32_t)__popc(tile_mask & ((1u << k) - 1u)) : (uint32_t)k);
            if (pos >= (uint32_t)kRingSlices) pos -= kRingSlices;
            const uint32_t nb = s_nbr_addr + (pos * kTileM + (uint32_t)rsub) * 4u;
            const uint32_t a_stage = a_base + st * kABytes;
            const char* src_q = col_base + q * (kChunkK * 2);
#pragma unroll
            for (int jb = 0; jb < 32; jb += 8) {
              if (((jb >> 3) * G) >> 2 != sub) continue;     // the warp's share of the rows: 32 / G per lane
              int r[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) r[j] = lds_s32(nb + (uint32_t)(jb + j) * 16u);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int jj = jb + j;
                cp_async_16(a_stage + (jj >> 1) * 1024 + ((jj & 1) ? d_odd : d_even), src_q + (uint64_t)(uint32_t)max(r[j], 0) * ld_bytes, r[j] >= 0 ? 16u : 0u);
              }
            }
          }
          cp_async_mbar_arrive_noinc_addr(full0 + st * 8);   // arrives when this lane's copies have landed
          st += PA;
          if (st >= (uint32_t)S) { st -= S; ph ^= 1; }
        }
      }
      mbar_arrive_pred(&table_free[tb], lane == 0 ? 1u : 0u);     // this warp has read everything it needs from buffer tb
    }
    cp_async_wait_all();                                 // nothing may still be writing smem at exit
#ifdef GCD_TC_PROFILE
    if (p.dbg && threadIdx.x == 0) { long long* d = p.dbg + (int64_t)blockIdx.x * 16; d[0] = clock64() - prof_t0; d[1] = prof_table; d[2] = prof_wait; d[3] = prof_iters; d[7] = tile_seq; }
#endif
  } else if (warp == kTab) {
    // ===================================================================== table warp
    // Streams neighbour-table slices ([128] int32 per kernel offset) into a shared-memory ring, as many tiles ahead of the
    // gather warps as the ring holds (kRingSlices slices, kTableSlots tiles).  With per-tile offset masks (tile_masks) only
    // the slices a tile uses are fetched, by cp.async straight into the ring: the warp never waits for a load, and the
    // tile is published by completion-triggered mbarrier arrivals.  Without masks the whole [kv][128] slice goes through
    // registers (the mask of offsets with a hit is derived on the way).  It also publishes each tile's iteration count
    // for the MMA warp.
    uint32_t tile_seq = 0, head = 0, used = 0, tail_seq = 0;
    unsigned long long counts = 0;                         // slices held by the tile of each slot, 6 bits per slot
    const uint32_t ring_addr = smem_u32(s_nbr0);
    const bool compact = p.tile_masks != nullptr;
    const bool aligned = (p.n_out & 3) == 0 && (reinterpret_cast<uintptr_t>(p.nbr) & 15) == 0;
    uint32_t next_mask = 0;
    // dynamic schedule: ticket t of the global counter is tile (tiles_m - 1 - t / n_tiles_n, t % n_tiles_n); one ticket is
    // held ahead so that the tile's offset mask is in flight while the previous tile is staged
    // position L of the schedule -> tile.  A CTA's first tile is position blockIdx.x (no atomic in front of the launch's first
    // loads); ticket t of the counter is position gridDim.x + t.
    auto tile_at = [&](int64_t pos) -> int64_t {
      if (pos >= n_work) return -1;
      const int64_t tm = pos / p.n_tiles_n;
      return (p.dyn_ascending ? tm : tiles_m - 1 - tm) * p.n_tiles_n + pos % p.n_tiles_n;
    };
    // The atomic of the tile after next is issued one iteration before its result is looked at: its round trip (and then the
    // mask load that depends on it) overlaps the staging of a whole tile instead of standing in front of every tile.
    int raw_ticket = 0;
    bool have_ticket = false;
    int64_t next_work = dyn ? tile_at(blockIdx.x) : static_work(0);
    if (dyn && next_work >= 0) {
      if (lane == 0) raw_ticket = atomicAdd(p.sched, 1);
      have_ticket = true;
    }
    if (compact && next_work >= 0) next_mask = __ldg(&p.tile_masks[next_work / p.n_tiles_n]);
    const uint32_t max_ahead = dyn ? (uint32_t)p.dyn_ahead : (uint32_t)kTableSlots;
#ifdef GCD_TC_PROFILE
    long long prof_twait = 0; const long long prof_t0 = clock64();
#endif
    for (;; ++tile_seq) {
      const int64_t work = next_work;
      if (work < 0 && !dyn) break;
      const uint32_t tb = tile_seq % kTableSlots;
      if (work < 0) {
        // end of the dynamic schedule: publish the sentinel through the slot protocol (the gather warps wait on table_ready)
        while (tile_seq - tail_seq >= (uint32_t)kTableSlots) {
          mbar_wait(&table_free[tail_seq % kTableSlots], (tail_seq / kTableSlots) & 1);
          ++tail_seq;
        }
        if (lane == 0) s_work[tile_seq & (kTileRing - 1)] = -1;
        __syncwarp();
        cp_async_mbar_arrive_noinc(&table_ready[tb]);
        mbar_arrive_pred(&table_ready[tb], lane == 0 ? 1u : 0u);
        if (lane == 0) st_release_shared_u32(smem_u32(s_pub), tile_seq + 1);
        break;
      }
      const int64_t row0 = (work / p.n_tiles_n) * kTileM;
      uint32_t mask = next_mask;
      if (dyn) {
        next_work = have_ticket ? tile_at((int64_t)gridDim.x + __shfl_sync(0xffffffffu, raw_ticket, 0)) : -1;
        have_ticket = false;
        if (next_work >= 0) {
          if (lane == 0) raw_ticket = atomicAdd(p.sched, 1);
          have_ticket = true;
        }
      } else {
        next_work = static_work(tile_seq + 1);
      }
      if (compact && next_work >= 0) next_mask = __ldg(&p.tile_masks[next_work / p.n_tiles_n]);   // one tile ahead
      const uint32_t load_mask = mask ? mask : 1u;         // degenerate tile: offset 0 runs with all-zero rows
      const uint32_t n_sl = compact ? (uint32_t)__popc(load_mask) : (uint32_t)p.kv;
#ifdef GCD_TC_PROFILE
      const long long ctw0 = clock64();
#endif
      // room in the ring and a free slot: release the oldest tiles the gather warps are done with (a dynamic schedule runs
      // fewer tiles ahead: what is claimed early is work another CTA cannot take at the end of the launch)
      while (used + n_sl > (uint32_t)kRingSlices || tile_seq - tail_seq >= max_ahead) {
        const uint32_t ts = tail_seq % kTableSlots;
        mbar_wait(&table_free[ts], (tail_seq / kTableSlots) & 1);
        used -= (uint32_t)(counts >> (6 * ts)) & 63u;
        ++tail_seq;
      }
#ifdef GCD_TC_PROFILE
      prof_twait += clock64() - ctw0;
#endif
      const uint32_t base = head;
      if (compact) {
        const bool vec = aligned && row0 + kTileM <= p.n_out;
        uint32_t todo = load_mask, pos = base;
        while (todo) {
          const int k = __ffs(todo) - 1;
          todo &= todo - 1;
          const int32_t* src = p.nbr + (int64_t)k * p.n_out + row0;
          const uint32_t dst = ring_addr + pos * (uint32_t)kSliceBytes;
          if (vec) {
            cp_async_16(dst + lane * 16, src + lane * 4, 16u);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int r = lane + 32 * j;
              if (row0 + r < p.n_out) cp_async_4(dst + r * 4, src + r);
              else s_nbr0[pos * kTileM + r] = -1;
            }
          }
          if (++pos == (uint32_t)kRingSlices) pos = 0;
        }
      } else {
        mask = 0;
        for (int kb = 0; kb < p.kv; kb += 14) {             // 56 loads in flight per lane (a load per k would cost a memory latency each)
          int v[14][4];
#pragma unroll
          for (int kk = 0; kk < 14; ++kk) {
            const int k = kb + kk;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int64_t r = row0 + lane + 32 * j;
              v[kk][j] = -1;
              if (k < p.kv && r < p.n_out) v[kk][j] = p.nbr ? __ldg(&p.nbr[(int64_t)k * p.n_out + r]) : (int)r;
            }
          }
#pragma unroll
          for (int kk = 0; kk < 14; ++kk) {
            const int k = kb + kk;
            if (k < p.kv) {
              uint32_t pos = base + (uint32_t)k;
              if (pos >= (uint32_t)kRingSlices) pos -= kRingSlices;
#pragma unroll
              for (int j = 0; j < 4; ++j) s_nbr0[pos * kTileM + lane + 32 * j] = v[kk][j];
              if (__any_sync(0xffffffffu, (v[kk][0] & v[kk][1] & v[kk][2] & v[kk][3]) >= 0)) mask |= 1u << k;     // some entry is not -1
            }
          }
        }
      }
      if (lane == 0) {
        s_mask[tb] = mask;
        s_base[tb] = base | (compact ? 0x80000000u : 0u);
        s_iters[tile_seq & (kTileRing - 1)] = __popc(mask ? mask : 1u) * nq;
        s_work[tile_seq & (kTileRing - 1)] = (int32_t)work;
      }
      __syncwarp();
      cp_async_mbar_arrive_noinc(&table_ready[tb]);        // each lane: arrives once its copies of this tile have landed
      mbar_arrive_pred(&table_ready[tb], lane == 0 ? 1u : 0u);
      if (dyn && lane == 0) st_release_shared_u32(smem_u32(s_pub), tile_seq + 1);   // MMA / epilogue warps: entry tile_seq is readable
      head = base + n_sl;
      if (head >= (uint32_t)kRingSlices) head -= kRingSlices;
      used += n_sl;
      counts = (counts & ~(63ull << (6 * tb))) | ((unsigned long long)n_sl << (6 * tb));
    }
    cp_async_wait_all();
    if (dyn && lane == 0) {
      // every CTA takes exactly one ticket beyond the work: when the last one has, nobody claims again and the counters go
      // back to zero for the next launch on this stream
      __threadfence();
      if (atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) { p.sched[0] = 0; p.sched[1] = 0; __threadfence(); }
    }
#ifdef GCD_TC_PROFILE
    if (p.dbg && lane == 0) { long long* d = p.dbg + (int64_t)blockIdx.x * 16; d[8] = clock64() - prof_t0; d[9] = prof_twait; }
#endif
  } else if (warp == kMma) {
    // ===================================================================== MMA issuer
    // The whole warp runs the loop with warp-uniform values (so descriptors live in uniform registers and there is
    // no divergence bookkeeping around every instruction); one elected lane issues the tcgen05 instructions.
    const uint32_t idesc = make_idesc_bf16((uint32_t)p.n_tile_cols, 0, 0);
    const uint64_t da0 = make_smem_desc_sw128(smem_u32(smem + L.a_off), 16, 1024);
    const uint64_t db0 = make_smem_desc_sw128(smem_u32(smem + L.b_off), 16, 1024);
    const uint32_t da_step = kABytes >> 4, db_step = L.b_bytes >> 4;      // descriptor address units are 16 bytes
    const uint32_t desc_hi = (uint32_t)(da0 >> 32);                        // identical for both operands (same layout / LBO / SBO)
    const uint32_t da0_lo = (uint32_t)da0, db0_lo = (uint32_t)db0;
    const int last_ksteps = last_width / 16;
    const bool leader = elect_one();
    uint32_t st = 0, ph = 0, tile_seq = 0;
    uint32_t da = da0_lo, db = db0_lo;
#ifdef GCD_TC_PROFILE
    long long prof_full = 0, prof_acc = 0; const long long prof_t0 = clock64();
#endif
    for (;; ++tile_seq) {
      if constexpr (kDyn) {
        const bool done = published_work(tile_seq) < 0;
        if (__all_sync(0xffffffffu, done)) break;
      } else {
        if (static_work(tile_seq) < 0) break;
      }
      const uint32_t buf = tile_seq & 1;
#ifdef GCD_TC_PROFILE
      const long long ca0 = clock64();
#endif
      mbar_wait(&tmem_empty[buf], ((tile_seq >> 1) & 1) ^ 1);
#ifdef GCD_TC_PROFILE
      prof_acc += clock64() - ca0;
#endif
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + buf * kAccStride;
      int n_iters = 1, q = 0;
      uint32_t accumulate = 0;
      for (int it = 0;;) {
#ifdef GCD_TC_PROFILE
        const long long cf0 = clock64();
#endif
        mbar_wait(&full_bar[st], ph);
#ifdef GCD_TC_PROFILE
        prof_full += clock64() - cf0;
#endif
        tc_fence_after();
        if (it == 0) n_iters = s_iters[tile_seq & (kTileRing - 1)];
        const int ksteps = (q == nq - 1) ? last_ksteps : kChunkK / 16;
        // predicated issue slots (no branch: divergence + reconvergence around an elected lane costs ~230 cycles per stage)
#pragma unroll
        for (int ks = 0; ks < kChunkK / 16; ++ks)
#ifdef GCD_TC_PROFILE
          mma_bf16_ss_pred_lo(tmem_d, da + (uint32_t)(ks * 2), db + (uint32_t)(ks * 2), desc_hi, idesc, accumulate | (uint32_t)ks, (leader && ks < ksteps && !(p.ablate & 1)) ? 1u : 0u);
#else
          mma_bf16_ss_pred_lo(tmem_d, da + (uint32_t)(ks * 2), db + (uint32_t)(ks * 2), desc_hi, idesc, accumulate | (uint32_t)ks, (leader && ks < ksteps) ? 1u : 0u);
#endif
        mma_commit_pred(&empty_bar[st], leader ? 1u : 0u);
        accumulate = 1;
        if (++q == nq) q = 0;
        if (++st == (uint32_t)S) { st = 0; ph ^= 1; da = da0_lo; db = db0_lo; } else { da += da_step; db += db_step; }
        ++it;
        if (__all_sync(0xffffffffu, it >= n_iters)) break;      // warp vote: the exit is a uniform branch, loop state stays in uniform registers
      }
      mma_commit_pred(&tmem_full[buf], leader ? 1u : 0u);
    }
#ifdef GCD_TC_PROFILE
    if (p.dbg && lane == 0) { long long* d = p.dbg + (int64_t)blockIdx.x * 16; d[4] = clock64() - prof_t0; d[5] = prof_full; d[6] = prof_acc; }
#endif
  } else if (warp < kMma) {
    // ===================================================================== epilogue
    const int ew = warp - kEpi0;      // == warp % 4: the TMEM lane quarter this warp may read
    const bool wide_store = p.out_is_bf16 && (reinterpret_cast<uintptr_t>(p.out) & 31) == 0 && (p.ld_out & 15) == 0;   // every 16-column chunk 32-byte aligned
    uint32_t tile_seq = 0;
#ifdef GCD_TC_PROFILE
    long long prof_ewait = 0; const long long prof_t0 = clock64();
#endif
    for (;; ++tile_seq) {
      int64_t work;
      if constexpr (kDyn) work = published_work(tile_seq); else work = static_work(tile_seq);
      if (work < 0) break;
      const int64_t tm = work / p.n_tiles_n;
      const int tn = (int)(work - tm * p.n_tiles_n);
      const uint32_t buf = tile_seq & 1;
      const int64_t row = tm * kTileM + ew * 32 + lane;
      int64_t orow = row;                        // where this accumulator lane's row goes
      if constexpr (kPerm) { if (row < p.n_out) orow = __ldg(&p.out_rows[row]); }     // issued before the wait: its latency hides behind the tile's MMAs
#ifdef GCD_TC_PROFILE
      const long long cew0 = clock64();
#endif
      mbar_wait_warp(&tmem_full[buf], (tile_seq >> 1) & 1, 128);
#ifdef GCD_TC_PROFILE
      prof_ewait += clock64() - cew0;
#endif
      tc_fence_after();
      const int col0 = tn * p.n_tile_cols;
      const uint32_t taddr = tmem_base + buf * kAccStride + ((uint32_t)(ew * 32) << 16);
      for (int c = 0; c < p.n_tile_cols; c += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(taddr + c, v);
        tmem_ld_wait();
        if (row < p.n_out) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + (p.bias ? p.bias[col0 + c + j] : 0.f);
          if (p.out_is_bf16) {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ld_out + col0 + c;
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
              w[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            if (wide_store) st_global_v8(o, w);      // one 32-byte store per lane: half the LSU work of two 16-byte ones (the LSU is the gather's resource)
            else {
              *reinterpret_cast<uint4*>(o) = make_uint4(w[0], w[1], w[2], w[3]);
              *reinterpret_cast<uint4*>(o + 8) = make_uint4(w[4], w[5], w[6], w[7]);
            }
          } else {
            float* o = reinterpret_cast<float*>(p.out) + orow * p.ld_out + col0 + c;
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(o + 4 * j) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
    }
#ifdef GCD_TC_PROFILE
    if (p.dbg && ew == 0 && lane == 0) { long long* d = p.dbg + (int64_t)blockIdx.x * 16; d[10] = clock64() - prof_t0; d[11] = prof_ewait; }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMma) { tc_fence_after(); tmem_dealloc<kTmemCols>(tmem_base); }
}

// ------------------------------------------------------------------------ weight packing
// Image of (offset k, 64-channel slice q): rows = output channel n (c_out of them), 128 B per row,
// element (n, kk) at n*128 + (((kk>>3) ^ (n&7))<<4) + (kk&7)*2; value = B_k[q*64+kk][n].
__device__ __forceinline__ void pack_one(const float* __restrict__ w, int kv, int c_in, int c_out, int transpose, int mirror,
                                         __nv_bfloat16* __restrict__ packed, int64_t t) {
  // logical operand: K' x N' where (K', N') = (c_in, c_out) for forward, (c_out, c_in) for dgrad
  const int kdim = transpose ? c_out : c_in, ndim = transpose ? c_in : c_out;
  const int nq = (kdim + kChunkK - 1) / kChunkK;
  const int64_t total = (int64_t)kv * nq * ndim * kChunkK;
  if (t >= total) return;
  const int kk = (int)(t % kChunkK);
  const int n = (int)((t / kChunkK) % ndim);
  const int q = (int)((t / ((int64_t)kChunkK * ndim)) % nq);
  const int k = (int)(t / ((int64_t)kChunkK * ndim * nq));
  const int c = q * kChunkK + kk;
  float v = 0.f;
  if (c < kdim) {
    const int ksrc = (transpose && mirror) ? kv - 1 - k : k;
    v = transpose ? w[((int64_t)ksrc * c_in + n) * c_out + c] : w[((int64_t)ksrc * c_in + c) * c_out + n];
  }
  const int64_t img = ((int64_t)k * nq + q) * ndim * kChunkK;  // elements
  packed[img + (int64_t)n * kChunkK + ((((kk >> 3) ^ (n & 7)) << 3) + (kk & 7))] = __float2bfloat16_rn(v);
}

__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ w, int kv, int c_in, int c_out, int transpose,
                                                            int mirror, __nv_bfloat16* __restrict__ packed) {
  pack_one(w, kv, c_in, c_out, transpose, mirror, packed, (int64_t)blockIdx.x * blockDim.x + threadIdx.x);
}

// All kernels of a model in one launch: block b works on descriptor d with block_start[d] <= b < block_start[d+1].
struct PackDesc {
  const float* w;
  __nv_bfloat16* dst;
  int32_t kv, c_in, c_out, transpose, mirror;
  int32_t block_start;
};
__global__ void __launch_bounds__(256) pack_weights_batched_kernel(const PackDesc* __restrict__ descs, int n_descs) {
  int lo = 0, hi = n_descs - 1;
  while (lo < hi) {                       // last descriptor whose block_start <= blockIdx.x
    const int mid = (lo + hi + 1) >> 1;
    if (descs[mid].block_start <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const PackDesc d = descs[lo];
  pack_one(d.w, d.kv, d.c_in, d.c_out, d.transpose, d.mirror, d.dst, (int64_t)(blockIdx.x - d.block_start) * blockDim.x + threadIdx.x);
}

// ======================================================================================= wgrad
// dW[k] (Cin x Cout, fp32) += A_k^T G_k over the pairs of offset k.  Work item = (offset k, chunk of
// pairs, 128-channel tile of Cin).  Per stage the producers gather 64 pairs: the input rows
// (their 128-channel slice) and the output-gradient rows (all Cout channels) land as 128-byte
// swizzled rows; with the *pair* index as the MMA K dimension both images are MN-major operands:
//   D[128 (Cin slice) x Cout] += A^T[128 x 16 pairs] * G[16 pairs x Cout]   (4 MMAs per stage)
// Rows of A^T beyond Cin read stale shared memory; they only produce accumulator lanes that the
// epilogue never reads.  The epilogue adds the accumulator into dW with fp32 atomics.
constexpr int kWgPairs = 64;                         // pairs (MMA K) per stage
constexpr int kSlabBytes = kWgPairs * kRowBytes;     // 8 KB: 64 rows x 64 channels
constexpr int kWgMaxOffsets = 128;

struct WgParams {
  const __nv_bfloat16* in; int64_t ld_in;
  const __nv_bfloat16* gout; int64_t ld_g;
  const int32_t* pair_in; const int32_t* pair_out; const int32_t* pair_off;
  int64_t n_rows_identity;
  int kv, c_in, c_out;
  int m_tiles;          // ceil(c_in / 128)
  int g_slabs;          // ceil(c_out / 64)
  int chunk;            // pairs per work item (multiple of 64)
  float* dw;
  int stages;
  int32_t* sched;       // optional {next item, finished CTAs}: dynamic schedule as in the forward kernel (warp 13 is the scheduler)
};
constexpr int kWgRing = 16;          // published work items (ring)
constexpr int kWgAhead = 6;          // items the scheduler may run ahead of the epilogue

template <int kGS>   // kGS = number of 64-channel slabs of the output gradient (ceil(c_out / 64), 1..4)
__global__ void __launch_bounds__(kTcThreads, 1) conv_wgrad_tc_kernel(const WgParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.stages;
  const uint32_t stage_bytes = (uint32_t)(2 + kGS) * kSlabBytes;         // A: 2 slabs, G: kGS slabs
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (uint32_t)S * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tmem_full = bars + 2 * kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* s_tmem_base = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  int32_t* s_cum = reinterpret_cast<int32_t*>(s_tmem_base + 4);        // [kv + 1] cumulative chunk counts
  int32_t* s_work = s_cum + kWgMaxOffsets + 1;                         // [kWgRing] dynamic schedule: claimed items, -1 = no more work
  uint32_t* s_pub = reinterpret_cast<uint32_t*>(s_work + kWgRing);     // entries of s_work published so far
  uint32_t* s_done = s_pub + 1;                                        // items the epilogue has finished

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();

  if (threadIdx.x == 0) {
    *s_pub = 0; *s_done = 0;
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 32); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], kEpilogueThreads / 32); }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) { tmem_alloc<kTmemCols>(s_tmem_base); tmem_relinquish(); }
  pdl_wait();                      // barrier / TMEM set-up above overlaps the previous kernel's tail; global memory from here on
  // chunk counts per offset: one parallel round of loads, then a prefix sum in shared memory (a serial loop of kv dependent
  // global loads used to open every launch)
  if ((int)threadIdx.x < p.kv) {
    const int64_t nk = p.pair_off ? (int64_t)__ldg(&p.pair_off[threadIdx.x + 1]) - __ldg(&p.pair_off[threadIdx.x]) : p.n_rows_identity;
    s_cum[threadIdx.x + 1] = (int)((nk + p.chunk - 1) / p.chunk);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    s_cum[0] = 0;
    for (int k = 0; k < p.kv; ++k) s_cum[k + 1] += s_cum[k];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem_base;
  const int64_t n_work = (int64_t)s_cum[p.kv] * p.m_tiles;
  const bool dyn = p.sched != nullptr;
  // item `seq` of this CTA: static striding, or whatever the scheduler warp claimed (see conv_fwd_tc_kernel)
  auto get_work = [&](uint32_t seq) -> int64_t {
    if (!dyn) {
      const int64_t w = (int64_t)blockIdx.x + (int64_t)seq * gridDim.x;
      return w < n_work ? w : -1;
    }
    const uint32_t addr = smem_u32(s_pub);
    uint32_t pub;
    do { pub = ld_acquire_shared_u32(addr); } while (!__all_sync(0xffffffffu, pub > seq));
    return (int64_t)s_work[seq & (kWgRing - 1)];
  };

  // work -> (k, first pair, last pair, m tile)
  auto decode = [&](int64_t work, int& k, int64_t& p_begin, int64_t& p_end, int& mt) {
    mt = (int)(work % p.m_tiles);
    const int item = (int)(work / p.m_tiles);
    k = 0;
    while (s_cum[k + 1] <= item) ++k;
    const int64_t base = p.pair_off ? (int64_t)p.pair_off[k] : 0;
    const int64_t end_k = p.pair_off ? (int64_t)p.pair_off[k + 1] : p.n_rows_identity;
    p_begin = base + (int64_t)(item - s_cum[k]) * p.chunk;
    p_end = min(p_begin + (int64_t)p.chunk, end_k);
  };

  if (warp < kProducerWarps) {
    // ===================================================================== producers (stage owners, see conv_fwd_tc_kernel)
    // Warp w owns the ring slots of stages g == w (mod PA): it gathers the 64 pairs of the stage itself (input rows: up to two
    // 64-channel slabs, gradient rows: kGS slabs; 16 rows x (2 + kGS) 16-byte copies per lane).  Its next stage is PA stages
    // away, so the pair indices of that stage are fetched from global memory while this one is being issued: the index
    // latency, which paced the lock-step version (one stage of lookahead), is hidden behind PA - 1 other stages.
    const int chunk16 = lane & 7, rsub = lane >> 3;                          // this lane's pairs of a stage: rsub + 4 j, j = 0..15
    const uint32_t d_even = rsub * kRowBytes + ((chunk16 ^ rsub) << 4);
    const uint32_t d_odd = (rsub + 4) * kRowBytes + ((chunk16 ^ (rsub + 4)) << 4);
    const uint32_t lda = (uint32_t)(p.ld_in * 2), ldg = (uint32_t)(p.ld_g * 2);   // row pitches in bytes (< 4 GB, checked by the host)
    const uint32_t stage0 = smem_u32(smem);
    const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
    uint32_t g_on = 0;                                                       // bit s: this thread's chunk exists in gradient slab s
#pragma unroll
    for (int s = 0; s < kGS; ++s) g_on |= (s * 64 + chunk16 * 8 < p.c_out ? 1u : 0u) << s;
    const char* g_col = reinterpret_cast<const char*>(p.gout + chunk16 * 8);
    const int PA = S < kProducerWarps ? S : kProducerWarps;
    uint32_t st = warp, ph = 0;
    int64_t own_skip = warp < PA ? warp : (int64_t)1 << 60;                  // stages until this warp's next owned one
    for (uint32_t seq = 0;; ++seq) {
      const int64_t work = get_work(seq);
      if (work < 0) break;
      int k, mt; int64_t p_begin, p_end;
      decode(work, k, p_begin, p_end, mt);
      const int64_t n_stages = (p_end - p_begin + kWgPairs - 1) / kWgPairs;
      if (own_skip >= n_stages) { own_skip -= n_stages; continue; }          // nothing of this item belongs to the warp
      const int c_base = mt * 128;
      const char* a_col = reinterpret_cast<const char*>(p.in + c_base + chunk16 * 8);
      const bool a_on0 = c_base + chunk16 * 8 < p.c_in, a_on1 = c_base + 64 + chunk16 * 8 < p.c_in;
      int ri[16], ro[16];
      auto fetch = [&](int64_t p0, int (&fi)[16], int (&fo)[16]) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int64_t pa = p0 + rsub + 4 * j;
          fi[j] = fo[j] = -1;
          if (pa < p_end) { fi[j] = p.pair_in ? __ldg(&p.pair_in[pa]) : (int)pa; fo[j] = p.pair_out ? __ldg(&p.pair_out[pa]) : (int)pa; }
        }
      };
      int64_t it = own_skip;
      fetch(p_begin + it * kWgPairs, ri, ro);
      for (; it < n_stages; it += PA) {
        int ni[16], no[16];
        const bool more = it + PA < n_stages;
        if (more) fetch(p_begin + (it + PA) * kWgPairs, ni, no);
        mbar_wait_addr(empty0 + st * 8, ph ^ 1);
        const uint32_t a_stage = stage0 + st * stage_bytes;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t off = (j >> 1) * 1024 + ((j & 1) ? d_odd : d_even);
          const uint32_t n = ri[j] >= 0 ? 16u : 0u;
          const char* sa = a_col + (uint64_t)(uint32_t)max(ri[j], 0) * lda;
          const char* sg = g_col + (uint64_t)(uint32_t)max(ro[j], 0) * ldg;
          if (a_on0) cp_async_16(a_stage + off, sa, n);
          if (a_on1) cp_async_16(a_stage + kSlabBytes + off, sa + 128, n);
#pragma unroll
          for (int s = 0; s < kGS; ++s)
            if (g_on & (1u << s)) cp_async_16(a_stage + (2 + s) * kSlabBytes + off, sg + s * 128, n);
        }
        cp_async_mbar_arrive_noinc_addr(full0 + st * 8);
        st += PA;
        if (st >= (uint32_t)S) { st -= S; ph ^= 1; }
        if (more) {
#pragma unroll
          for (int j = 0; j < 16; ++j) { ri[j] = ni[j]; ro[j] = no[j]; }
        }
      }
      own_skip = it - n_stages;
    }
    cp_async_wait_all();
  } else if (warp == kMmaWarp) {
    // ===================================================================== MMA issuer (uniform loop, elected lane issues)
    {
      const uint32_t idesc = make_idesc_bf16((uint32_t)p.c_out, 1, 1);   // both operands MN-major
      const uint64_t da0 = make_smem_desc_sw128(smem_u32(smem), kSlabBytes, 1024);
      const uint64_t db0 = make_smem_desc_sw128(smem_u32(smem) + 2 * kSlabBytes, kSlabBytes, 1024);
      const uint32_t desc_hi = (uint32_t)(da0 >> 32), da0_lo = (uint32_t)da0, db0_lo = (uint32_t)db0;
      const uint32_t stage_step = stage_bytes >> 4, k_step = (16 * kRowBytes) >> 4;
      const bool leader = elect_one();
      uint32_t st = 0, ph = 0, seq = 0;
      uint32_t da = da0_lo, db = db0_lo;
      for (;; ++seq) {
        const int64_t work = get_work(seq);
        if (work < 0) break;
        int k, mt; int64_t p_begin, p_end;
        decode(work, k, p_begin, p_end, mt);
        const uint32_t buf = seq & 1;
        mbar_wait(&tmem_empty[buf], ((seq >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * kAccStride;
        const int n_stages = (int)((p_end - p_begin + kWgPairs - 1) / kWgPairs);
        uint32_t accumulate = 0;
        for (int it = 0;;) {
          mbar_wait(&full_bar[st], ph);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < kWgPairs / 16; ++ks)
            mma_bf16_ss_pred_lo(tmem_d, da + ks * k_step, db + ks * k_step, desc_hi, idesc, accumulate | (uint32_t)ks, leader ? 1u : 0u);
          mma_commit_pred(&empty_bar[st], leader ? 1u : 0u);
          accumulate = 1;
          if (++st == (uint32_t)S) { st = 0; ph ^= 1; da = da0_lo; db = db0_lo; } else { da += stage_step; db += stage_step; }
          ++it;
          if (__all_sync(0xffffffffu, it >= n_stages)) break;     // warp vote: uniform exit keeps the loop state in uniform registers
        }
        mma_commit_pred(&tmem_full[buf], leader ? 1u : 0u);
      }
    }
  } else if (warp < kMmaWarp) {
    // ===================================================================== epilogue (warps 8-11; warp 13 idles here)
    const int ew = warp - kEpilogueWarp0;
    uint32_t seq = 0;
    for (;; ++seq) {
      const int64_t work = get_work(seq);
      if (work < 0) break;
      int k, mt; int64_t p_begin, p_end;
      decode(work, k, p_begin, p_end, mt);
      const uint32_t buf = seq & 1;
      mbar_wait_warp(&tmem_full[buf], (seq >> 1) & 1, 128);
      tc_fence_after();
      const int c = mt * 128 + ew * 32 + lane;          // input channel = accumulator lane
      const uint32_t taddr = tmem_base + buf * kAccStride + ((uint32_t)(ew * 32) << 16);
      float* dst = p.dw + ((int64_t)k * p.c_in + c) * p.c_out;
      for (int n0 = 0; n0 < p.c_out; n0 += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(taddr + n0, v);
        tmem_ld_wait();
        if (c < p.c_in) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)      // dst + n0 is 16-byte aligned (c_out % 16 == 0): one vector reduction per 4 columns
            red_add_v4_f32(dst + n0 + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
      if (dyn && ew == 0 && lane == 0) st_release_shared_u32(smem_u32(s_done), seq + 1);     // back-pressure for the scheduler
    }
  } else if (dyn) {
    // ===================================================================== scheduler (warp 13, dynamic schedule only)
    for (uint32_t seq = 0;; ++seq) {
      uint32_t done;
      do { done = ld_acquire_shared_u32(smem_u32(s_done)); } while (!__all_sync(0xffffffffu, seq - done < (uint32_t)kWgAhead));
      int64_t pos = blockIdx.x;                 // first item: static (no atomic in front of the launch's first loads)
      if (seq > 0) {
        int t = 0;
        if (lane == 0) t = atomicAdd(p.sched, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        pos = (int64_t)gridDim.x + t;
      }
      const int32_t work = pos < n_work ? (int32_t)pos : -1;
      if (lane == 0) {
        s_work[seq & (kWgRing - 1)] = work;
        st_release_shared_u32(smem_u32(s_pub), seq + 1);
      }
      if (work < 0) break;
    }
    if (lane == 0) {
      __threadfence();
      if (atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) { p.sched[0] = 0; p.sched[1] = 0; __threadfence(); }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) { tc_fence_after(); tmem_dealloc<kTmemCols>(tmem_base); }
}
}  // namespace

// Opt every tcgen05 kernel instantiation into the large dynamic shared memory, once per process (std::call_once: the
// autograd thread and a prefetch thread may both make the first call).
static cudaError_t tc_kernels_opt_in();

bool conv_wgrad_tc_supported(const gcd_wgrad_args* a) {
  return a->in_dtype == GCD_BF16 && a->gout_dtype == GCD_BF16 && a->c_in % 16 == 0 && a->c_out % 16 == 0 && a->c_in >= 16 &&
         a->c_out >= 16 && a->c_out <= 256 && a->kv < kWgMaxOffsets && a->ld_in % 8 == 0 && a->ld_gout % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(a->in) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->gout) & 15) == 0;
}

int32_t conv_wgrad_tc(const gcd_wgrad_args* a, cudaStream_t st) {
  WgParams p;
  p.in = (const __nv_bfloat16*)a->in; p.ld_in = a->ld_in; p.gout = (const __nv_bfloat16*)a->gout; p.ld_g = a->ld_gout;
  p.pair_in = a->pair_in; p.pair_out = a->pair_out; p.pair_off = a->pair_off; p.n_rows_identity = a->n_pairs;
  p.kv = a->kv; p.c_in = a->c_in; p.c_out = a->c_out; p.dw = a->dw;
  p.sched = a->sched;
  GCD_REQUIRE(a->ld_in > 0 && a->ld_in < (int64_t(1) << 31) && a->ld_gout > 0 && a->ld_gout < (int64_t(1) << 31), "conv_wgrad_tc: row pitch out of range");
  p.m_tiles = (a->c_in + 127) / 128;
  p.g_slabs = (a->c_out + 63) / 64;
  // n_pairs is an upper bound when pair lists are used (the exact count lives on the device);
  // real kernel maps fill about a quarter of the table.
  const int64_t expect = a->pair_in ? std::max<int64_t>(a->n_pairs / 4, 1) : a->n_pairs;
  int64_t chunk = ceil_div(expect, (int64_t)kNumSMs * 2);
  // wide outputs: each work item ends with a 128 x Cout reduction into dW, fewer and longer items keep that traffic down
  // (measured, tools/diag_tc.py: 128->128 at stride 8 23.5 -> 19.5 us, 384->256 58 -> 48 us; narrow layers prefer 512)
  int64_t chunk_min = a->c_out >= 128 ? 1024 : 512;
  if (const int o = option(GCD_OPT_WG_CHUNK_MIN); o > 0) chunk_min = std::max(64, o);      // tuning aid
  chunk = std::max<int64_t>(chunk_min, std::min<int64_t>(8192, ceil_div(chunk, kWgPairs) * kWgPairs));
  p.chunk = (int)chunk;
  const int stage_bytes = (2 + p.g_slabs) * kSlabBytes;
  int stages = std::min(kMaxStages, (kSmemBudget - 1024 - 1024) / stage_bytes);
  if (stages < 2) { set_error("conv_wgrad_tc: not enough shared memory stages"); return GCD_ERR_UNSUPPORTED; }
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 1024;
  using Kernel = void (*)(const WgParams);
  static const Kernel kernels[4] = {conv_wgrad_tc_kernel<1>, conv_wgrad_tc_kernel<2>, conv_wgrad_tc_kernel<3>, conv_wgrad_tc_kernel<4>};
  if (const cudaError_t e = tc_kernels_opt_in(); e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tcgen05 kernels)");
  const int64_t work_bound = (ceil_div(a->n_pairs, chunk) + a->kv) * p.m_tiles;
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(work_bound, kNumSMs));
  if (const cudaError_t e = launch_pdl(kernels[p.g_slabs - 1], dim3(grid), dim3(kTcThreads), smem, st, p); e != cudaSuccess) return cuda_fail(e, "gcd_conv_wgrad(tcgen05)");
  return GCD_OK;
}

#ifdef GCD_TC_PROFILE
static long long* g_debug_buffer = nullptr;
#endif

bool conv_forward_tc_supported(const gcd_conv_args* a) {
  return a->in_dtype == GCD_BF16 && a->c_in % 16 == 0 && a->c_out % 16 == 0 && a->c_in >= 16 && a->c_out >= 16 &&
         a->c_out <= 1024 && a->c_in <= 1024 && a->kv <= kMaxKV && a->w_packed != nullptr && a->ld_in % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(a->in) & 15) == 0 &&
         (a->out_dtype == GCD_BF16 ? (a->ld_out % 8 == 0) : (a->ld_out % 4 == 0)) && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0;
}

using FwdKernel = void (*)(const FwdParams);
template <bool kPerm, int kPW = kProducerWarps, bool kDyn = false>
FwdKernel pick_fwd_kernel(int nq) {
  switch (nq) {
    case 1: return conv_fwd_tc_kernel<1, kPerm, kPW, kDyn>;
    case 2: return conv_fwd_tc_kernel<2, kPerm, kPW, kDyn>;
    case 3: return conv_fwd_tc_kernel<3, kPerm, kPW, kDyn>;
    case 4: return conv_fwd_tc_kernel<4, kPerm, kPW, kDyn>;
    case 6: return conv_fwd_tc_kernel<6, kPerm, kPW, kDyn>;
    default: return conv_fwd_tc_kernel<0, kPerm, kPW, kDyn>;
  }
}

static cudaError_t tc_kernels_opt_in() {
  static std::once_flag once;
  static cudaError_t result = cudaSuccess;
  std::call_once(once, [] {
    auto opt_in = [](const void* k) {
      const cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget + 1024);
      if (e != cudaSuccess && result == cudaSuccess) result = e;
    };
    for (int nq : {0, 1, 2, 3, 4, 6}) {
      opt_in(reinterpret_cast<const void*>(pick_fwd_kernel<false>(nq)));
      opt_in(reinterpret_cast<const void*>(pick_fwd_kernel<true>(nq)));
      opt_in(reinterpret_cast<const void*>(pick_fwd_kernel<false, 16>(nq)));
      opt_in(reinterpret_cast<const void*>(pick_fwd_kernel<true, 16>(nq)));
      opt_in(reinterpret_cast<const void*>(pick_fwd_kernel<false, kProducerWarps, true>(nq)));
      opt_in(reinterpret_cast<const void*>(pick_fwd_kernel<true, kProducerWarps, true>(nq)));
    }
    opt_in(reinterpret_cast<const void*>(conv_wgrad_tc_kernel<1>));
    opt_in(reinterpret_cast<const void*>(conv_wgrad_tc_kernel<2>));
    opt_in(reinterpret_cast<const void*>(conv_wgrad_tc_kernel<3>));
    opt_in(reinterpret_cast<const void*>(conv_wgrad_tc_kernel<4>));
  });
  return result;
}

int32_t conv_forward_tc(const gcd_conv_args* a, cudaStream_t st) {
  FwdParams p;
  p.in = (const __nv_bfloat16*)a->in; p.ld_in = a->ld_in; p.nbr = a->nbr; p.kv = a->kv; p.n_out = a->n_out;
  p.c_in = a->c_in; p.c_out = a->c_out;
  p.n_tiles_n = (a->c_out + 255) / 256;
  p.n_tile_cols = a->c_out / p.n_tiles_n;
  if (p.n_tile_cols % 16 != 0 || p.n_tile_cols * p.n_tiles_n != a->c_out) { set_error("conv_forward_tc: cannot split %d output channels into equal tiles", a->c_out); return GCD_ERR_UNSUPPORTED; }
  // Deep levels have fewer row tiles than SMs and the kernel is bound by each SM's shared-memory bandwidth
  // (per stage 2 x (16 KB of rows + n_tile_cols x 128 B of weights) through a 128 B/clk port), so idle SMs are put to work
  // by splitting the output channels further: every CTA then streams a narrower weight slice.
  {
    const int64_t tiles_m = ceil_div(a->n_out, kTileM);
    while (tiles_m * p.n_tiles_n * 2 <= kNumSMs && p.n_tile_cols % 32 == 0 && p.n_tile_cols / 2 >= 32) { p.n_tiles_n *= 2; p.n_tile_cols /= 2; }
  }
  p.w_packed = (const uint8_t*)a->w_packed;  // offset mirroring (dgrad of stride-1 maps) is baked into the packed image
  GCD_REQUIRE(a->ld_in > 0 && a->ld_in < (int64_t(1) << 31), "conv_forward_tc: input row pitch out of range");
  p.bias = a->bias; p.out = a->out; p.ld_out = a->ld_out; p.out_is_bf16 = a->out_dtype == GCD_BF16;
  p.out_rows = a->out_rows;
  p.tile_masks = a->tile_masks;
  p.sched = a->sched;
  p.dyn_ascending = option(GCD_OPT_DYN_TILES) == 2;
  p.dyn_ahead = std::max(1, std::min(kTableSlots, option(GCD_OPT_DYN_AHEAD) > 0 ? option(GCD_OPT_DYN_AHEAD) : kDynAhead));
  GCD_REQUIRE(a->out_rows == nullptr || a->nbr != nullptr, "conv_forward_tc: out_rows needs a neighbour table");
  GCD_REQUIRE(a->tile_masks == nullptr || a->nbr != nullptr, "conv_forward_tc: tile_masks needs a neighbour table");
  const int stage_bytes = kABytes + p.n_tile_cols * kRowBytes;
  int stages = (kSmemBudget - 1024 - kRingSlices * kSliceBytes - 1024) / stage_bytes;
  stages = std::min(stages, kMaxStages);
  if (const int o = option(GCD_OPT_TC_STAGES); o > 0) stages = std::max(2, std::min(stages, o));   // tuning aid
  if (stages < 2) { set_error("conv_forward_tc: not enough shared memory stages"); return GCD_ERR_UNSUPPORTED; }
  p.stages = stages;
  const int pw = option(GCD_OPT_TC_WARPS) == 16 ? 16 : 8;
  p.group = stages <= 4 ? 2 : 1;      // measured: one warp per stage is best with >= 5 stages, a warp pair when the stages are few and fat
  if (pw == 16) p.group *= 2;         // sixteen gather warps: a warp pair (a quad when the stages are few and fat) per slot
  if (const int g = option(GCD_OPT_TC_GROUP); g == 1 || g == 2 || g == 4 || g == 8) p.group = g;   // tuning aid
#ifdef GCD_TC_PROFILE
  p.dbg = g_debug_buffer;
  p.ablate = getenv("GCD_TC_ABLATE") ? atoi(getenv("GCD_TC_ABLATE")) : 0;
#endif
  const SmemLayout L = make_layout(stages, p.n_tile_cols);
  const size_t smem = L.total + 1024;
  const int nq_sel = (a->c_in + kChunkK - 1) / kChunkK;
  if (pw == 16) p.sched = nullptr;      // (the sixteen-warp tuning variant exists with the static schedule only)
  const FwdKernel kernel = pw == 16 ? (p.out_rows ? pick_fwd_kernel<true, 16>(nq_sel) : pick_fwd_kernel<false, 16>(nq_sel))
                           : p.sched ? (p.out_rows ? pick_fwd_kernel<true, kProducerWarps, true>(nq_sel) : pick_fwd_kernel<false, kProducerWarps, true>(nq_sel))
                                     : (p.out_rows ? pick_fwd_kernel<true>(nq_sel) : pick_fwd_kernel<false>(nq_sel));
  if (const cudaError_t e = tc_kernels_opt_in(); e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tcgen05 kernels)");
  if (a->c_in > 16 * kChunkK) { set_error("conv_forward_tc: more than 1024 input channels unsupported"); return GCD_ERR_UNSUPPORTED; }
  const int64_t n_work = ceil_div(a->n_out, kTileM) * p.n_tiles_n;
  const unsigned grid = (unsigned)std::min<int64_t>(n_work, kNumSMs);
  if (const cudaError_t e = launch_pdl(kernel, dim3(grid), dim3((pw + 6) * 32), smem, st, p); e != cudaSuccess) return cuda_fail(e, "gcd_conv_forward(tcgen05)");
  return GCD_OK;
}
}  // namespace gcd

using namespace gcd;

extern "C" size_t gcd_conv_packed_weight_bytes(int32_t kv, int32_t c_in, int32_t c_out) {
  // large enough for either orientation
  const int64_t a = (int64_t)kv * ((c_in + 63) / 64) * c_out * 128;
  const int64_t b = (int64_t)kv * ((c_out + 63) / 64) * c_in * 128;
  return (size_t)std::max(a, b);
}

extern "C" int32_t gcd_conv_pack_weights(const float* w, int32_t kv, int32_t c_in, int32_t c_out, int32_t transpose, int32_t mirror,
                                         void* packed, void* stream) {
  GCD_REQUIRE(w && packed && kv >= 1 && c_in >= 1 && c_out >= 1, "gcd_conv_pack_weights: bad arguments");
  const int kdim = transpose ? c_out : c_in, ndim = transpose ? c_in : c_out;
  const int64_t total = (int64_t)kv * ((kdim + 63) / 64) * ndim * 64;
  pack_weights_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(w, kv, c_in, c_out, transpose, mirror, (__nv_bfloat16*)packed);
  GCD_LAUNCH_CHECK("gcd_conv_pack_weights");
  return GCD_OK;
}

extern "C" int32_t gcd_conv_pack_weights_batched(const void* descs, int32_t n_descs, int32_t total_blocks, void* stream) {
  GCD_REQUIRE(descs && n_descs >= 1 && total_blocks >= 1, "gcd_conv_pack_weights_batched: bad arguments");
  static_assert(sizeof(PackDesc) == 40, "PackDesc layout is part of the C ABI (gcd_pack_desc)");
  pack_weights_batched_kernel<<<(unsigned)total_blocks, 256, 0, as_stream(stream)>>>((const PackDesc*)descs, n_descs);
  GCD_LAUNCH_CHECK("gcd_conv_pack_weights_batched");
  return GCD_OK;
}

#ifdef GCD_TC_PROFILE
// tuning builds only (make PROFILE=1): per-CTA cycle counters of the forward kernel's producer / MMA warps
extern "C" int32_t gcd_debug_set_buffer(void* device_buffer) { gcd::g_debug_buffer = static_cast<long long*>(device_buffer); return 0; }
#endif

extern "C" int32_t gcd_conv_tc_supported(int32_t c_in, int32_t c_out, int32_t kv) {
  return c_in % 16 == 0 && c_out % 16 == 0 && c_in >= 16 && c_out >= 16 && c_out <= 1024 && c_in <= 1024 && kv <= 27;
}
