// MakeTransfers (qrad3.c:185-293) and BounceLight (qrad3.c:383-473) on the GPU.
//
// Layout ("slot order", plan.hpp): patches are grouped by the PVS cluster of their origin, every cluster padded to a
// multiple of 32 slots, so one 32-bit word of the visibility matrix = one source patch x the 32 receivers of one slice.
// The matrix is stored slice-major: word(source row, slice s) = bits[wbase(cluster of the row, word) + own row], i.e.
// the words of 32 consecutive source rows towards one slice are 32 consecutive words.  PVS pruning is structural:
// invisible pairs have no word.
//
//   pass 1  k_visibility   one warp per run = (row group of <= 128 sources, chunk of <= 32 words = 1024 receivers): the
//                          same receivers seen from neighbouring sources.  Whole-word interval proofs, form factor
//                          sign tests, blocker-guess walks, full TestLine for the rest; writes the accept bit
//                          trans > 0  (qrad3.c:216-257).
//   pass 2  k_row_totals32 one warp per 32 source rows, lane = row: walks the receivers in ascending patch index (the
//                          reference's j order) run by run, reading its words coalesced; every lane forms the SERIAL
//                          float sum `total` (qrad3.c:254) and numtransfers of its own row.
//   pass 3  k_column_counts / k_columns
//                          one warp per slice (32 receivers): walks the sources that can see it in ascending patch
//                          index, streaming ONE word per source from the slice's column, and emits the quantised
//                          weights (int)(trans * 0x10000 / total) (qrad3.c:283) as "dense slice" lists -- the
//                          transposed (gather) form ShootLight's scatter needs (finding F7), already in the source
//                          order the reference adds them.
//   bounce  k_bounce       persistent warps, one slice at a time, lane = receiver, 128-bit loads of 8 weights per
//                          lane, radiosity vectors broadcast from shared memory, packed multiplies, serial
//                          accumulation in source order, CollectLight fused.
//
// Several GPUs (a context that joined a communicator, qrad_cuda_comm_init): source rows are owned in row groups, receiver
// slices in contiguous runs (plan.hpp).  Pass 1 and pass 2 are local; before pass 3 every rank sends each peer the
// contiguous piece of its row buffer that describes the peer's slices (grouped ncclSend / ncclRecv), row totals are
// summed over ranks, and every bounce ends with the owners broadcasting their part of the radiosity vector.
//
// Everything is float arithmetic in the reference's operation order; this TU is compiled with -fmad=false.
#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdlib>
#include <cstring>

#include "bsp_flat.hpp"
#include "comm.hpp"
#include "qrad_device.cuh"

namespace qrad {
namespace {

constexpr int kThreads = 256;
constexpr int kWarpsPerBlock = kThreads / 32;

// form factor of source (o, n) towards receiver (po = origin+area, pn): qrad3.c:232-249.
// Returns trans (0 when rejected); *wants_ray = the reference would call TestLine_r here.
__device__ __forceinline__ float form_factor_pre(const float4 so, const float4 sn, const float4 po, const float4 pn,
                                                 float &dist_out, bool &wants_ray) {
  float dx = po.x - so.x, dy = po.y - so.y, dz = po.z - so.z;
  const float dist = vnormalize(dx, dy, dz);
  wants_ray = false;
  dist_out = dist;
  if (dist == 0) return 0.f;
  float scale = dx * sn.x + dy * sn.y + dz * sn.z;
  scale *= -(dx * pn.x + dy * pn.y + dz * pn.z);
  if (!(scale > 0)) return 0.f;
  wants_ray = true;
  float trans = scale * po.w / (dist * dist);
  if (trans < 0) trans = 0;
  return trans;
}

// The part of the pair test that decides whether the reference calls TestLine_r (qrad3.c:232-244), and whether the pair
// can carry a transfer at all (trans > 0).  sn.w / pn.w hold +-(axis + 1) for exactly axial plane normals: the normalised
// dot product then is ONE coordinate of delta times 1 / dist -- a positive factor that cannot underflow for world
// coordinates -- so the sign of scale = (delta . n_i) * -(delta . n_j) follows from two coordinate differences, without
// the square root and the division.  For such pairs trans = scale * area / dist^2 > 0 exactly when scale > 0 (area >= 1,
// scale >= 1e-17: nothing underflows).  Other normals take the reference's arithmetic.
__device__ __forceinline__ bool pair_wants_ray(const float4 so, const float4 sn, const float4 po, const float4 pn,
                                               bool &positive) {
  const float dx = po.x - so.x, dy = po.y - so.y, dz = po.z - so.z;
  const int cs = (int) sn.w, cr = (int) pn.w;
  if (cs != 0 && cr != 0) {
    const float a = axis_pick(abs(cs) - 1, dx, dy, dz), b = axis_pick(abs(cr) - 1, dx, dy, dz);
    // a' = +-a / dist, b' = +-b / dist;  scale = a' * -b' > 0
    const bool apos = cs > 0 ? a > 0 : a < 0, aneg = cs > 0 ? a < 0 : a > 0;
    const bool bpos = cr > 0 ? b > 0 : b < 0, bneg = cr > 0 ? b < 0 : b > 0;
    positive = (apos && bneg) || (aneg && bpos);
    return positive;
  }
  float dist;
  bool wants;
  const float trans = form_factor_pre(so, sn, po, pn, dist, wants);
  positive = trans > 0;
  return wants;
}

// What every kernel needs to find a word of the matrix (device copies of the plan, plan.hpp).
struct Geo {
  const int32_t *cl_slot_start, *cl_count;   // [nc + 1], [nc]
  const int32_t *slot_patch, *slot_cluster;  // [M]
  const uint8_t *slot_flags;                 // [M] bit 0: holds a patch, bit 1: may shoot (its origin lies in a valid cluster)
  const int32_t *slice_count;                // [M / 32] valid receivers of a slice
  const int64_t *wb_first;                   // [nc + 1] first entry of a source cluster in the word tables
  const int32_t *wslice;                     // receiver slice of word w of a row of the cluster
  const int64_t *wbase;                      // address of that word for own row 0 of the cluster in this rank's buffer
  const int32_t *nwords;                     // [nc]
  int nc;
};

struct VisArgs {
  Tree tree;
  const float4 *wbox_min, *wbox_max;  // [M/32] bounding box of the patch origins of each slice
  const float4 *s_origin_area, *s_normal;
  Geo g;
  const int4 *run_list;           // this rank's runs = (row group, chunk, first row, rows), the expensive ones first
  const int32_t *grp_cluster, *grp_slot0, *grp_rows, *grp_ownbase;
  uint64_t num_runs;
  uint32_t *bits;
  unsigned long long *counters;   // [0] rays, [1] node visits, [2] candidate pairs
  const float4 *cell_lo, *cell_hi;   // [numleafs] leaf cells (shaft_blocked)
  float margin;                   // how deep inside a solid cell a segment must pass to count as proven blocked
  int mode;                       // bit 0: slice x word cell proofs, 1: source x word cell proofs, 2: per-ray cell proofs,
                                  // 5: walks start at the node all rays of (source slice, word) reach unclipped
};

// One warp lights one run = the rows of one row group (<= 128 consecutive source patches of one cluster) against one
// chunk of <= 32 words (1024 receivers), row after row: the SAME receivers seen from neighbouring source patches, so
// what blocked a receiver last time is an excellent guess for this time (measured with the CPU oracle on C5: 95 % of
// the rays are blocked, 94-99 % of those by the leaf that blocked the same receiver from the previous source).
// No ray is traced that a cheap exact argument settles.  Per row ("task"):
//   (1) lane w, for word w: shaft_blocked() with the leaf that blocked somebody in the word last -- if every segment from
//       the source (or from all 32 sources of its slice: kept for the following rows) to the word's bounding box passes
//       deep through that leaf's solid cell, all 32 receivers are blocked (73 % of the rays on C5, no tree walk);
//   (2) which pairs would the reference hand to TestLine_r (qrad3.c:232-244)?  They are counted as rays either way.
//       Word level first: with axial normals and the word's box on one side of the source the sign test has one
//       outcome for all 32 receivers; only the other words are tested receiver by receiver.  Pairs of unproven words
//       are compacted into a queue;
//   (3) the queue is drained 32 rays at a time: segment_through_cell() with the receiver's last blocker and with the
//       word's (another 19 %); the rest collects in a second queue and takes the full TestLine walk in full batches
//       (9 % of the rays: the 4.6 % that are clear, and blocked ones whose guess was wrong or only grazed), which
//       refreshes the guesses;
//   (4) accept bits are OR-ed into 32 result words in shared memory and stored into the 32 slice columns.
// Measured alternatives (profiles/README.md): interval walks along the root -> leaf path of the guess
// (round 1: 37 % of the rays per word, and a walk per remaining ray) cost 2.5 x the time of the cell proofs; one word per
// warp without compaction leaves 46 % of the lanes idle; refilling finished lanes from the queue decorrelates the lanes.
constexpr int kQueue = 256;   // pairs waiting for stage (3); drained whenever another word might not fit
template <int STACK, bool COUNT>
__global__ void __launch_bounds__(kThreads, 4) k_visibility(const VisArgs a) {
  __shared__ uint16_t s_queue[kWarpsPerBlock][kQueue];
  __shared__ uint16_t s_rguess[kWarpsPerBlock][kChunkWords * 32];  // per receiver of the chunk: last blocker leaf
  __shared__ uint16_t s_wguess[kWarpsPerBlock][kChunkWords];       // per word: last blocker leaf seen in it
  __shared__ uint32_t s_res[kWarpsPerBlock][kChunkWords];
  __shared__ int32_t s_wbase[kWarpsPerBlock][kChunkWords];         // first receiver slot of the word
  __shared__ int32_t s_wcnt[kWarpsPerBlock][kChunkWords];          // valid receivers in the word
  __shared__ int64_t s_waddr[kWarpsPerBlock][kChunkWords];         // where the word of own row 0 of the cluster is stored
  __shared__ uint16_t s_fb[kWarpsPerBlock][64];                    // rays the cell proofs left open (receiver index)
  __shared__ uint16_t s_wstart[kWarpsPerBlock][kChunkWords];       // where the walks of (current source slice, word) may start
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1;
  unsigned long long rays = 0, pairs = 0, visits_total = 0, boxed = 0, shafted = 0;
  unsigned long long grazed = 0, uncelled = 0, wrong_guess = 0, grazed_other = 0, retested = 0, walked = 0;
  long long tk0 = 0, tk_box = 0, tk_pairs = 0, tk_guess = 0, tk_full = 0, tk_misc = 0;   // COUNT: warp cycles per phase
  for (;;) {
    // runs are drawn from a counter: their cost varies by an order of magnitude (a run whose words are all proven costs
    // a few microseconds, one inside a room with thousands of clear rays a millisecond), and a shard has only ~20 runs
    // per resident warp -- with a fixed stride the last warps set the kernel time (8 ranks: 31 % more warp time than 1)
    unsigned long long run = 0;
    if (lane == 0) run = atomicAdd(a.counters + 16, 1ull);
    run = __shfl_sync(0xffffffffu, run, 0);
    if (run >= a.num_runs) break;
    const int4 rc = __ldg(a.run_list + run);
    const int grp = rc.x, chunk = rc.y;
    const int c = a.grp_cluster[grp];
    const int nw = a.g.nwords[c];
    const int w0 = chunk * kChunkWords;
    const int nwc = min(kChunkWords, nw - w0);
    const int slot0 = a.grp_slot0[grp], row_lo = rc.z, row_hi = rc.z + rc.w, ownbase = a.grp_ownbase[grp];
    __syncwarp();
    if (lane < nwc) {
      const int64_t e = a.g.wb_first[c] + w0 + lane;
      const int sl = a.g.wslice[e];
      s_wbase[wid][lane] = sl * 32;
      s_wcnt[wid][lane] = a.g.slice_count[sl];
      s_waddr[wid][lane] = a.g.wbase[e];
    }
    for (int k = lane; k < kChunkWords * 32; k += 32) s_rguess[wid][k] = 0xffffu;   // new receivers: forget the guesses
    s_wguess[wid][lane] = 0xffffu;
    unsigned slice_proven = 0;      // words proven blocked for ALL sources of the current source slice
    __syncwarp();
  for (int r = row_lo; r < row_hi; r++) {
    const int sslot = slot0 + r;
    const unsigned flags = a.g.slot_flags[sslot];
    if (!(flags & 2u)) {   // padding slot at the end of a cluster, or (only under -nopvs) a source in solid, which never
                           // shoots (qrad3.c:208): zero words, so that every word of the row buffer is defined
      if (lane < nwc) a.bits[s_waddr[wid][lane] + ownbase + r] = 0u;
      continue;
    }
    const float4 so = __ldg(a.s_origin_area + sslot);
    const float4 sn = __ldg(a.s_normal + sslot);
    s_res[wid][lane] = 0;
    if (COUNT) { const long long t = clock64(); if (tk0) tk_misc += t - tk0; tk0 = t; }
    // ---- (1) per word: proofs that settle all 32 receivers at once; (2a) the word-level sign test ----
    // every 32 rows the sources move on to the next slice: its two-sided proofs start afresh
    if ((r & 31) == 0 || r == row_lo) {
      slice_proven = 0;
      if (a.mode & 32) {   // the node every ray of (this source slice, word) reaches unclipped: their walks start there
        s_wstart[wid][lane] = 0;
        if (lane < nwc)
          s_wstart[wid][lane] = (uint16_t) common_start_node(a.tree, __ldg(a.wbox_min + (sslot >> 5)), __ldg(a.wbox_max + (sslot >> 5)),
                                                             __ldg(a.wbox_min + (s_wbase[wid][lane] >> 5)),
                                                             __ldg(a.wbox_max + (s_wbase[wid][lane] >> 5)));
        __syncwarp();
      }
    }
    bool proven = (slice_proven >> lane) & 1u;
    bool proven_now2 = false;
    int wclass = 0;   // 0: no receiver passes the sign test, 1: all do, 2: receiver by receiver
    if (lane < nwc) {
      const int base = s_wbase[wid][lane];
      const float4 bmn = __ldg(a.wbox_min + (base >> 5)), bmx = __ldg(a.wbox_max + (base >> 5));
      const int g = s_wguess[wid][lane];
      if (!proven && g != 0xffff) {
        const float4 clo = __ldg(a.cell_lo + g), chi = __ldg(a.cell_hi + g);
        // (a) all 32 sources of the slice x all 32 receivers of the word
        if ((a.mode & 1) && shaft_blocked(clo, chi, a.margin, __ldg(a.wbox_min + (sslot >> 5)), __ldg(a.wbox_max + (sslot >> 5)), bmn, bmx))
          proven = proven_now2 = true;
        // (b) this source x all 32 receivers
        else if ((a.mode & 2) && shaft_blocked(clo, chi, a.margin, so, so, bmn, bmx))
          proven = true;
      }
      // When source and receivers have axial normals and the word's bounding box lies strictly on one side of the
      // source along both normals' axes, the sign test has the same outcome for all 32 receivers (the coordinate
      // differences are monotone in the receiver's coordinate): such a word costs one test, not 32 loads.
      wclass = 2;
      const int cs = (int) sn.w, cr = (int) bmn.w;
      if (cs != 0 && cr != 0) {
        const int ks = abs(cs) - 1, kr = abs(cr) - 1;
        const float os = axis_pick(ks, so.x, so.y, so.z), orr = axis_pick(kr, so.x, so.y, so.z);
        const float alo = axis_pick(ks, bmn.x, bmn.y, bmn.z) - os, ahi = axis_pick(ks, bmx.x, bmx.y, bmx.z) - os;
        const float blo = axis_pick(kr, bmn.x, bmn.y, bmn.z) - orr, bhi = axis_pick(kr, bmx.x, bmx.y, bmx.z) - orr;
        const int sa = alo > 0 ? 1 : (ahi < 0 ? -1 : 0), sb = blo > 0 ? 1 : (bhi < 0 ? -1 : 0);
        if (sa != 0 && sb != 0) {
          const bool apos = (cs > 0) == (sa > 0), bpos = (cr > 0) == (sb > 0);
          wclass = (apos != bpos) ? 1 : 0;
          const int cnt = s_wcnt[wid][lane];
          pairs += cnt;
          if (wclass == 1) {
            rays += cnt;
            if (COUNT && proven) boxed += cnt;
          }
        }
      }
    }
    slice_proven |= __ballot_sync(0xffffffffu, proven_now2);
    const unsigned proven_mask = __ballot_sync(0xffffffffu, proven);
    const unsigned all_mask = __ballot_sync(0xffffffffu, wclass == 1) & ~proven_mask;
    const unsigned each_mask = __ballot_sync(0xffffffffu, wclass == 2);
    __syncwarp();
    if (COUNT) { const long long t = clock64(); tk_box += t - tk0; tk0 = t; }
    // ---- (3) cell proofs per ray, full walks for the rest ----
    int n = 0, fbn = 0;
    auto full_batch = [&](const int cnt) {  // the first `cnt` entries of the fallback queue
      long long tf0 = 0;
      if (COUNT) tf0 = clock64();
      if (lane < cnt) {
        const int cand = s_fb[wid][lane];
        const float4 po = __ldcg(a.s_origin_area + s_wbase[wid][cand >> 5] + (cand & 31));
        unsigned visits = 0;
        int leaf = -1;
        int blocked;
        // the word's guess may have been refreshed by an earlier batch of this very row (a group of neighbouring
        // receivers usually changes its blocker together): one more cell test before the walk
        const int gw = s_wguess[wid][cand >> 5];
        if ((a.mode & 4) && gw != 0xffff && gw != s_rguess[wid][cand] &&
            segment_through_cell(__ldg(a.cell_lo + gw), __ldg(a.cell_hi + gw), a.margin, so, po)) {
          blocked = 1;
          leaf = gw;
          if (COUNT) retested++;
        } else {
          blocked = test_line<STACK, COUNT>(a.tree, so.x, so.y, so.z, po.x, po.y, po.z, visits, &leaf,
                                            (a.mode & 32) ? (int) s_wstart[wid][cand >> 5] : 0);
          if (COUNT) walked++;
        }
        if (COUNT) visits_total += visits;
        if (!blocked) atomicOr(&s_res[wid][cand >> 5], 1u << (cand & 31));
        else if (leaf < 0xffff) {
          if (COUNT) {   // why did the proofs miss this blocked ray: the guess was right but the ray only grazes the cell ...
            if (leaf == s_rguess[wid][cand] || leaf == s_wguess[wid][cand >> 5]) grazed++;
            else if (__ldg(a.cell_lo + leaf).w == 0.f) uncelled++;   // ... or the blocker has no box cell ...
            else if (segment_through_cell(__ldg(a.cell_lo + leaf), __ldg(a.cell_hi + leaf), a.margin, so, po)) wrong_guess++;
            else grazed_other++;                                     // ... or it was another leaf (provable / not)
          }
          s_rguess[wid][cand] = (uint16_t) leaf;
          s_wguess[wid][cand >> 5] = (uint16_t) leaf;
        }
      }
      __syncwarp();
      if (COUNT) tk_full += clock64() - tf0;
    };
    auto drain = [&]() {   // stage (3) for the n pairs in the queue
      long long td0 = 0;
      if (COUNT) td0 = clock64();
      for (int base = 0; base < n; base += 32) {
        const int k = base + lane;
        bool miss = false;
        int cand = 0;
        if (k < n) {
          cand = s_queue[wid][k];
          const int g = s_rguess[wid][cand], g2 = s_wguess[wid][cand >> 5];
          miss = true;
          if ((a.mode & 4) && (g != 0xffff || g2 != 0xffff)) {
            const float4 po = __ldcg(a.s_origin_area + s_wbase[wid][cand >> 5] + (cand & 31));
            // (c) this ray alone: does it pass deep through the cell of the leaf that blocked this receiver last time,
            // or of the leaf that blocked a neighbour of it most recently?
            if ((g != 0xffff && segment_through_cell(__ldg(a.cell_lo + g), __ldg(a.cell_hi + g), a.margin, so, po)) ||
                (g2 != g && g2 != 0xffff &&
                 segment_through_cell(__ldg(a.cell_lo + g2), __ldg(a.cell_hi + g2), a.margin, so, po))) {
              miss = false;
              if (COUNT) shafted++;
            }
          }
        }
        const unsigned m = __ballot_sync(0xffffffffu, miss);
        if (miss) s_fb[wid][fbn + __popc(m & lt)] = (uint16_t) cand;
        fbn += __popc(m);
        __syncwarp();
        if (fbn >= 32) {
          full_batch(32);
          uint16_t t = 0;
          if (lane < fbn - 32) t = s_fb[wid][32 + lane];
          __syncwarp();
          if (lane < fbn - 32) s_fb[wid][lane] = t;
          fbn -= 32;
          __syncwarp();
        }
      }
      n = 0;
      if (COUNT) tk_guess += clock64() - td0;
    };
    // ---- (2b) receiver-by-receiver sign tests where needed, compaction ----
    for (unsigned todo = all_mask | each_mask; todo;) {
      const int wi = __ffs(todo) - 1;
      todo &= todo - 1;
      bool want;
      if ((each_mask >> wi) & 1u) {
        const int rslot = s_wbase[wid][wi] + lane;
        want = false;
        if (lane < s_wcnt[wid][wi] && rslot != sslot) {
          // receivers stream past L1, which holds the tree and the cells
          const float4 po = __ldcg(a.s_origin_area + rslot), pn = __ldcg(a.s_normal + rslot);
          bool positive;
          const bool wants_ray = pair_wants_ray(so, sn, po, pn, positive);
          pairs++;
          if (wants_ray) {
            rays++;
            want = positive && !((proven_mask >> wi) & 1u);
            if (COUNT && positive && !want) boxed++;
          }
        }
      } else
        want = lane < s_wcnt[wid][wi];
      const unsigned m = __ballot_sync(0xffffffffu, want);
      if (want) s_queue[wid][n + __popc(m & lt)] = (uint16_t) (wi * 32 + lane);
      n += __popc(m);
      __syncwarp();
      if (n > kQueue - 32) drain();
    }
    if (COUNT) { const long long t = clock64(); tk_pairs += t - tk0; tk0 = t; }
    drain();
    if (fbn > 0) full_batch(fbn);
    __syncwarp();
    if (COUNT) { const long long t = clock64(); tk0 = t; }
    // consecutive rows of the run fill consecutive words of each column: L2 merges them into whole sectors
    if (lane < nwc) a.bits[s_waddr[wid][lane] + ownbase + r] = s_res[wid][lane];
    __syncwarp();
  }
  }
  for (int o = 16; o; o >>= 1) {
    rays += __shfl_xor_sync(0xffffffffu, rays, o);
    pairs += __shfl_xor_sync(0xffffffffu, pairs, o);
    if (COUNT) {
      visits_total += __shfl_xor_sync(0xffffffffu, visits_total, o);
      boxed += __shfl_xor_sync(0xffffffffu, boxed, o);
      shafted += __shfl_xor_sync(0xffffffffu, shafted, o);
      grazed += __shfl_xor_sync(0xffffffffu, grazed, o);
      uncelled += __shfl_xor_sync(0xffffffffu, uncelled, o);
      wrong_guess += __shfl_xor_sync(0xffffffffu, wrong_guess, o);
      grazed_other += __shfl_xor_sync(0xffffffffu, grazed_other, o);
      retested += __shfl_xor_sync(0xffffffffu, retested, o);
      walked += __shfl_xor_sync(0xffffffffu, walked, o);
    }
  }
  if (lane == 0) {
    atomicAdd(a.counters + 0, rays);
    atomicAdd(a.counters + 2, pairs);
    if (COUNT) {
      atomicAdd(a.counters + 1, visits_total);
      atomicAdd(a.counters + 3, walked);   // rays that took the full TestLine walk
      atomicAdd(a.counters + 4, boxed);
      atomicAdd(a.counters + 5, shafted);
      atomicAdd(a.counters + 6, grazed);
      atomicAdd(a.counters + 7, uncelled);
      atomicAdd(a.counters + 13, wrong_guess);
      atomicAdd(a.counters + 14, grazed_other);
      atomicAdd(a.counters + 15, retested);
      atomicAdd(a.counters + 8, (unsigned long long) tk_box);
      atomicAdd(a.counters + 9, (unsigned long long) (tk_pairs - tk_guess));
      atomicAdd(a.counters + 10, (unsigned long long) (tk_guess - tk_full));
      atomicAdd(a.counters + 11, (unsigned long long) tk_full);
      atomicAdd(a.counters + 12, (unsigned long long) tk_misc);
    }
  }
}

// ---- word tables ---------------------------------------------------------------------------------------------------
// For every source cluster a and every word w of its rows: the receiver slice the word is about and the address of the
// word of own row 0 (plan.hpp).  One warp per visible pair (a, d).
__global__ void __launch_bounds__(kThreads) k_word_tables(int nvis, const int32_t *__restrict__ vis_src,
                                                          const int32_t *__restrict__ vis_cluster,
                                                          const int32_t *__restrict__ vis_woff, const int32_t *__restrict__ vis_seer,
                                                          const int32_t *__restrict__ cl_slot_start, const int64_t *__restrict__ wb_first,
                                                          const int64_t *__restrict__ sbase, const int32_t *__restrict__ coff,
                                                          int32_t *__restrict__ wslice, int64_t *__restrict__ wbase) {
  const int v = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (v >= nvis) return;
  const int a = vis_src[v], d = vis_cluster[v];
  const int s0 = cl_slot_start[d] >> 5, ns = (cl_slot_start[d + 1] >> 5) - s0;
  const int64_t e0 = wb_first[a] + vis_woff[v];
  const int co = coff[vis_seer[v]];
  for (int j = threadIdx.x & 31; j < ns; j += 32) {
    wslice[e0 + j] = s0 + j;
    wbase[e0 + j] = sbase[s0 + j] + co;
  }
}

// ---- ordered run lists ---------------------------------------------------------------------------------------------
// The receivers of a row must be visited in ascending PATCH index (the reference's j loop), the sources of a column
// likewise (the order ShootLight adds them).  Slots are cluster-major, but consecutive patches mostly share a cluster and
// then have consecutive slots, so the patch order is kept as a list of RUNS = (first slot, length <= 32) of consecutive
// patches inside one slice.  Per source cluster c the runs whose cluster c can see form its row list; per receiver
// cluster d the runs whose cluster can see d form its column list.  Both are an ordered compaction of the one global
// run list: one block per cluster.
struct RunListArgs {
  int nruns;
  const int2 *runs;                 // {first slot, length}
  const int32_t *slot_cluster, *cl_slot_start;
  const int32_t *vis_ptr, *vis_cluster, *vis_woff, *vis_seer;
  const int32_t *clusters;          // clusters handled by this launch (blockIdx.x indexes it)
  const int64_t *list_ptr;          // [nc + 1] where the list of a cluster starts (fill mode)
  int32_t *count;                   // [nc] count mode output
  // row lists
  const int64_t *wb_first, *wbase;
  const int32_t *wslice;
  int4 *vl;                         // {addr lo, addr hi, mask, slice}
  // column lists
  const int32_t *grp_first, *grp_owner, *grp_ownbase, *grp_slot0;
  const int32_t *coff;              // [world][E]
  int64_t E;
  int4 *vr;                         // {first slot, offset in the owner's column, length, owner}
};

__device__ __forceinline__ int find_visible(const RunListArgs &a, const int src_cluster, const int dst_cluster) {
  int lo = a.vis_ptr[src_cluster], hi = a.vis_ptr[src_cluster + 1];
  while (lo < hi) {   // lower bound of dst_cluster in the (ascending) visible list of src_cluster
    const int mid = (lo + hi) >> 1;
    if (a.vis_cluster[mid] < dst_cluster) lo = mid + 1; else hi = mid;
  }
  return (lo < a.vis_ptr[src_cluster + 1] && a.vis_cluster[lo] == dst_cluster) ? lo : -1;
}

template <bool COLUMNS, bool FILL>
__global__ void __launch_bounds__(1024) k_run_lists(const RunListArgs a) {
  const int c = a.clusters[blockIdx.x];
  __shared__ int warp_sums[32];
  __shared__ int64_t base_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = FILL ? a.list_ptr[c] : 0;
  __syncthreads();
  for (int k0 = 0; k0 < a.nruns; k0 += blockDim.x) {
    const int k = k0 + threadIdx.x;
    int v = -1;
    int2 run = make_int2(0, 0);
    int rc = 0;
    if (k < a.nruns) {
      run = a.runs[k];
      rc = a.slot_cluster[run.x];
      v = COLUMNS ? find_visible(a, rc, c) : find_visible(a, c, rc);
    }
    const bool flag = v >= 0;
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) warp_sums[wid] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
    for (int j = 0; j < (int) (blockDim.x >> 5); j++) {
      const int n = warp_sums[j];
      if (j < wid) before += n;
      total += n;
    }
    if (FILL && flag) {
      const int64_t o = base_s + before + __popc(m & ((1u << lane) - 1));
      const int rel = run.x - a.cl_slot_start[rc];
      if (!COLUMNS) {
        const int64_t e = a.wb_first[c] + a.vis_woff[v] + (rel >> 5);
        const int64_t addr = a.wbase[e];
        const unsigned mask = (run.y >= 32 ? 0xffffffffu : ((1u << run.y) - 1u)) << (rel & 31);
        a.vl[o] = make_int4((int) (uint32_t) (addr & 0xffffffffll), (int) (addr >> 32), (int) mask, a.wslice[e]);
      } else {
        const int grp = a.grp_first[rc] + rel / kGroupRows;
        const int q = a.grp_owner[grp];
        const int off = a.coff[(int64_t) q * a.E + a.vis_seer[v]] + a.grp_ownbase[grp] + (run.x - a.grp_slot0[grp]);
        a.vr[o] = make_int4(run.x, off, run.y, q);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) base_s += total;
    __syncthreads();
  }
  if (!FILL && threadIdx.x == 0) a.count[c] = (int) base_s;
}

__device__ __forceinline__ int64_t vl_addr(const int4 e) { return ((int64_t) e.y << 32) | (uint32_t) e.x; }

// ---- pass 2: serial row sums in ascending receiver order ----------------------------------------------------------
struct RowArgs {
  const float4 *s_origin_area, *s_normal;
  Geo g;
  const int64_t *vl_ptr;           // [nc + 1]
  const int4 *vl;
  const uint32_t *bits;
  const int32_t *rowslice_slot0, *rowslice_row0;   // the 32-row slices this launch handles: first slot, first own row
  int n_rowslices;
  float *row_total;                // [M]
  int32_t *row_count;              // [M] numtransfers of the source in slot s
  // download of selected rows (reference row format): warp k handles source slot sel_slots[k] / own row sel_rows[k]
  const int32_t *sel_slots, *sel_rows;
  int nsel;
  const int64_t *row_ptr;          // [nsel + 1]
  uint32_t *out_patch;
  uint16_t *out_transfer;
  long long *sel_rays;             // [nsel] root TestLine_r calls of the row (k_row_rays)
};

// lane = row: one warp sums 32 consecutive source rows of one cluster at once.  The rows share the cluster's row list,
// and neighbouring sources accept nearly the same receivers, so the warp walks the list once; per run (<= 32 receivers of
// one slice) every lane reads its own word of that slice -- 32 consecutive words -- and a run no row accepts is skipped.
// Otherwise the receivers some row accepts are staged in shared memory and every accepting lane adds its own form factor
// to its own SERIAL sum (qrad3.c:254): `total` stays bit-identical while 32 sums advance together.
__global__ void __launch_bounds__(kThreads, 5) k_row_totals32(const RowArgs a) {
  __shared__ float4 s_po[kWarpsPerBlock][32], s_pn[kWarpsPerBlock][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int rs = blockIdx.x * kWarpsPerBlock + wid;
  if (rs >= a.n_rowslices) return;
  const int sslot0 = a.rowslice_slot0[rs];
  const int sslot = sslot0 + lane;
  const int64_t row = (int64_t) a.rowslice_row0[rs] + lane;
  const int c = a.g.slot_cluster[sslot0];              // a 32-slot slice lies in one cluster
  const bool live_row = (a.g.slot_flags[sslot] & 2u) != 0;
  const float4 so = a.s_origin_area[sslot];
  const float4 sn = a.s_normal[sslot];
  const int64_t e0 = a.vl_ptr[c], e1 = a.vl_ptr[c + 1];
  float total = 0.f;
  int count = 0;
  // The list is taken 32 runs at a time (lane j holds run j: one coalesced load), and the words of four runs are in
  // flight while the previous four are summed: the walk is bound by the latency of the word loads otherwise (one
  // dependent HBM round trip per run, profiles/r02b_launches_c5.txt: 68 ms).
  for (int64_t eb = e0; eb < e1; eb += 32) {
    const int nb = (int) min((int64_t) 32, e1 - eb);
    const int4 mine = lane < nb ? __ldg(a.vl + eb + lane) : make_int4(0, 0, 0, 0);
    const int64_t myaddr = vl_addr(mine);
    auto fetch = [&](const int j) -> uint32_t {   // this lane's word of run j (0 past the end of the batch)
      const int64_t ad = __shfl_sync(0xffffffffu, myaddr, j & 31);
      return j < nb ? __ldcs(a.bits + ad + row) : 0u;
    };
    uint32_t cur[4], nxt[4];
#pragma unroll
    for (int k = 0; k < 4; k++) cur[k] = fetch(k);
    for (int j0 = 0; j0 < nb; j0 += 4) {
#pragma unroll
      for (int k = 0; k < 4; k++) nxt[k] = fetch(j0 + 4 + k);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int j = j0 + k;
        if (j >= nb) break;
        const uint32_t mask = (uint32_t) __shfl_sync(0xffffffffu, mine.z, j);
        const int slice = __shfl_sync(0xffffffffu, mine.w, j);
        const uint32_t acc = live_row ? (cur[k] & mask) : 0u;
        unsigned hot = __reduce_or_sync(0xffffffffu, acc);
        if (!hot) continue;
        __syncwarp();
        if ((hot >> lane) & 1u) {
          s_po[wid][lane] = a.s_origin_area[slice * 32 + lane];
          s_pn[wid][lane] = a.s_normal[slice * 32 + lane];
        }
        __syncwarp();
        while (hot) {
          const int kk = __ffs(hot) - 1;
          hot &= hot - 1;
          if ((acc >> kk) & 1u) {
            float dist;
            bool wr;
            const float trans = form_factor_pre(so, sn, s_po[wid][kk], s_pn[wid][kk], dist, wr);
            total += trans;
            count++;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; k++) cur[k] = nxt[k];
    }
  }
  if (a.g.slot_flags[sslot] & 1u) {
    a.row_total[sslot] = total;
    a.row_count[sslot] = count;
  }
}

// The same sums with four times the warps: a warp takes 8 of the 32 rows of a slice, lane = (row, one of four receivers).
// The four lanes of a row compute the form factors of four accepted-or-not receivers at once, then every one of them adds
// the four values in receiver order (a rejected receiver contributes +0.0f, which leaves a float sum as it is): the sum
// is the same serial sum, the square roots and divisions run four abreast.  One warp per 32 rows needs 20 ms for the
// longest row list of C5 all by itself -- the whole pass on an eighth of the map should take 8.
__global__ void __launch_bounds__(kThreads) k_row_totals8(const RowArgs a) {
  __shared__ float4 s_po[kWarpsPerBlock][32], s_pn[kWarpsPerBlock][32];
  __shared__ uint8_t s_hot[kWarpsPerBlock][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1;
  const int q = blockIdx.x * kWarpsPerBlock + wid;
  const int rs = q >> 2, part = q & 3;
  if (rs >= a.n_rowslices) return;
  const int r = lane >> 2, sub = lane & 3;
  const int sslot = a.rowslice_slot0[rs] + part * 8 + r;
  const int64_t row = (int64_t) a.rowslice_row0[rs] + part * 8 + r;
  const int c = a.g.slot_cluster[a.rowslice_slot0[rs]];
  const bool live_row = (a.g.slot_flags[sslot] & 2u) != 0;
  const float4 so = a.s_origin_area[sslot];
  const float4 sn = a.s_normal[sslot];
  const int64_t e0 = a.vl_ptr[c], e1 = a.vl_ptr[c + 1];
  float total = 0.f;
  int count = 0;
  for (int64_t eb = e0; eb < e1; eb += 32) {
    const int nb = (int) min((int64_t) 32, e1 - eb);
    const int4 mine = lane < nb ? __ldg(a.vl + eb + lane) : make_int4(0, 0, 0, 0);
    const int64_t myaddr = vl_addr(mine);
    auto fetch = [&](const int j) -> uint32_t {   // this lane's row's word of run j (0 past the end of the batch)
      const int64_t ad = __shfl_sync(0xffffffffu, myaddr, j & 31);
      return j < nb ? __ldcs(a.bits + ad + row) : 0u;
    };
    uint32_t cur[4], nxt[4];
#pragma unroll
    for (int k = 0; k < 4; k++) cur[k] = fetch(k);
    for (int j0 = 0; j0 < nb; j0 += 4) {
#pragma unroll
      for (int k = 0; k < 4; k++) nxt[k] = fetch(j0 + 4 + k);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int j = j0 + k;
        if (j >= nb) break;
        const uint32_t mask = (uint32_t) __shfl_sync(0xffffffffu, mine.z, j);
        const int slice = __shfl_sync(0xffffffffu, mine.w, j);
        const uint32_t acc = live_row ? (cur[k] & mask) : 0u;
        const unsigned hot = __reduce_or_sync(0xffffffffu, acc);
        if (!hot) continue;
        __syncwarp();
        if ((hot >> lane) & 1u) {
          s_po[wid][lane] = a.s_origin_area[slice * 32 + lane];
          s_pn[wid][lane] = a.s_normal[slice * 32 + lane];
          s_hot[wid][__popc(hot & lt)] = (uint8_t) lane;
        }
        __syncwarp();
        count += __popc(acc);
        const int nh = __popc(hot);
        for (int g = 0; g < nh; g += 4) {
          float t = 0.f;
          if (g + sub < nh) {
            const int kk = s_hot[wid][g + sub];
            if ((acc >> kk) & 1u) {
              float dist;
              bool wr;
              t = form_factor_pre(so, sn, s_po[wid][kk], s_pn[wid][kk], dist, wr);
            }
          }
#pragma unroll
          for (int i = 0; i < 4; i++) total += __shfl_sync(0xffffffffu, t, (lane & ~3) + i);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; k++) cur[k] = nxt[k];
    }
  }
  if (sub == 0 && (a.g.slot_flags[sslot] & 1u)) {
    a.row_total[sslot] = total;
    a.row_count[sslot] = count;
  }
}

// Reference row format of selected rows (parity tests): one warp per row; the receivers of a run are tested by 32 lanes
// and written in ascending patch index: {patch, (int)(trans * 0x10000 / total)} (qrad3.c:283-289).
__global__ void __launch_bounds__(kThreads) k_rows_emit(const RowArgs a) {
  const int lane = threadIdx.x & 31;
  const int widx = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (widx >= a.nsel) return;
  const int sslot = a.sel_slots[widx];
  if (sslot < 0 || !(a.g.slot_flags[sslot] & 1u)) return;
  const int64_t row = a.sel_rows[widx];
  const int c = a.g.slot_cluster[sslot];
  const float4 so = a.s_origin_area[sslot];
  const float4 sn = a.s_normal[sslot];
  const float rtotal = a.row_total[sslot];
  int64_t outp = a.row_ptr[widx];
  for (int64_t e = a.vl_ptr[c]; e < a.vl_ptr[c + 1]; e++) {
    const int4 ent = __ldg(a.vl + e);
    const uint32_t word = a.bits[vl_addr(ent) + row] & (uint32_t) ent.z;   // the same word for every lane
    if (!word) continue;
    if ((word >> lane) & 1u) {
      const int rslot = ent.w * 32 + lane;
      float dist;
      bool wr;
      const float trans = form_factor_pre(so, sn, a.s_origin_area[rslot], a.s_normal[rslot], dist, wr);
      const int64_t o = outp + __popc(word & ((1u << lane) - 1));
      a.out_patch[o] = (uint32_t) a.g.slot_patch[rslot];
      a.out_transfer[o] = (uint16_t) (int) (trans * 65536.0f / rtotal);
    }
    outp += __popc(word);
  }
}

// Root TestLine_r calls of selected rows: the candidates of the row (PVS-visible receivers other than the source) that
// pass the sign test of qrad3.c:238-241 -- the pairs the reference hands to TestLine_r at qrad3.c:244.
__global__ void __launch_bounds__(kThreads) k_row_rays(const RowArgs a) {
  const int lane = threadIdx.x & 31;
  const int widx = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (widx >= a.nsel) return;
  const int sslot = a.sel_slots[widx];
  long long n = 0;
  if (sslot >= 0 && (a.g.slot_flags[sslot] & 2u)) {
    const int c = a.g.slot_cluster[sslot];
    const float4 so = a.s_origin_area[sslot];
    const float4 sn = a.s_normal[sslot];
    for (int64_t e = a.vl_ptr[c]; e < a.vl_ptr[c + 1]; e++) {
      const int4 ent = __ldg(a.vl + e);
      const int rslot = ent.w * 32 + lane;
      if (!(((uint32_t) ent.z >> lane) & 1u) || rslot == sslot) continue;
      float dist;
      bool wr;
      form_factor_pre(so, sn, a.s_origin_area[rslot], a.s_normal[rslot], dist, wr);
      n += wr ? 1 : 0;
    }
  }
  for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if (lane == 0) a.sel_rays[widx] = n;
}

// ---- pass 3: columns -----------------------------------------------------------------------------------------------
// pass 3 + bounce share the gather-form layout ("dense slices"): for every slice of 32 receivers the list of
// sources that reach at least one of them (ascending patch index = the order ShootLight adds them), one u32 source
// slot per list entry and 32 u16 weights per entry (0 = no transfer to that receiver; adding send * 0 leaves the
// float sum unchanged, and the add is skipped anyway).  Entries are grouped by 8: a lane's 8 weights are 16
// contiguous bytes (one 128-bit load), a warp reads 512 contiguous bytes per group.  Measured on C5: 18.8 of the
// 32 receivers of a listed source are hit on average, so this stores 3.6 B per transfer instead of 6 and turns
// the per-transfer gather of the radiosity vector into one broadcast load per 18.8 transfers (the per-column ELL
// of the first version was bound by L1 sector lookups at 43 % of HBM peak, profiles/r01c_full_c5_k_bounce.txt).
constexpr int kMaxWorld = 16;
struct ColArgs {
  const float4 *s_origin_area, *s_normal;
  Geo g;
  const int64_t *vr_ptr;           // [nc + 1]
  const int4 *vr;                  // {first source slot, offset in the owner's column, sources (<= 32), owner rank}
  const uint32_t *col[kMaxWorld];  // per owner rank: its words about this rank's slices, biased so that the index is
                                   //   sbase[owner][slice] + offset  (the own rank points into the row buffer itself)
  const int64_t *sbase;            // [world][nslices + 1]
  int nslices;
  const float *row_total;
  int slice_lo, slice_hi;          // this rank's slices
  const int32_t *col_order;        // fill mode: this rank's slices, longest list first (one block each)
  int32_t *slice_len;              // count mode output: list entries of the slice
  const int64_t *slice_base;       // fill mode: first entry of the slice (multiple of 32)
  uint32_t *g_src;                 // [entries]      source slot
  uint4 *g_w;                      // [entries / 8][32] 8 x u16 weights per lane
};

// pass 3a: list lengths = non-zero words of the slice's columns (one per owner rank), each a contiguous piece of memory
// streamed once with 128-bit loads.  One warp per slice.  (Words of padding rows are zero: k_visibility writes them.)
__global__ void __launch_bounds__(kThreads) k_column_counts(const ColArgs a, const int world) {
  const int lane = threadIdx.x & 31;
  const int slice = a.slice_lo + blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (slice >= a.slice_hi) return;
  int cnt = 0;
  for (int q = 0; q < world; q++) {
    const int64_t *sb = a.sbase + (int64_t) q * (a.nslices + 1);
    const int64_t first = sb[slice], n = sb[slice + 1] - first;   // multiples of 32 words
    const uint4 *p = reinterpret_cast<const uint4 *>(a.col[q] + first);
    const int64_t n4 = n >> 2;
    for (int64_t i = lane; i < n4; i += 128) {   // four loads in flight per lane
      uint4 w[4];
#pragma unroll
      for (int k = 0; k < 4; k++) w[k] = i + 32 * k < n4 ? __ldcs(p + i + 32 * k) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int k = 0; k < 4; k++) cnt += (w[k].x != 0) + (w[k].y != 0) + (w[k].z != 0) + (w[k].w != 0);
    }
  }
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) a.slice_len[slice] = cnt;
}

// pass 3b: the lists, lane = receiver.  The slice's sources come in runs of <= 32 consecutive slots whose words are
// consecutive in the slice's column: lane k fetches the word of source k of a run (the words of four runs are in flight
// while the previous four are emitted); when some word is non-zero the sources that own one are staged in shared memory
// (origin, normal, row total: one gather per run instead of a dependent L2 round trip per list entry), and the warp
// writes the entries in source order.
// SPLIT = false: one warp per slice, slices longest list first -- for many waves of slices (one GPU: 95 ms on C5).
// SPLIT = true:  one BLOCK per slice.  A shard of an eighth of C5 is less than one wave of warps, and took as long as its
//                longest list (20 of 33 ms).  The eight warps first count the non-zero words -- the list entries -- of the
//                run list in cells of 32 runs, a prefix over the cells cuts the list into eight parts of (nearly) equal
//                entries at cell boundaries, and every warp emits one part from the entry the prefix names.
constexpr int kColCells = 1024;
template <bool SPLIT>
__global__ void __launch_bounds__(kThreads, 4) k_columns(const ColArgs a, const int nslices_own) {
  __shared__ float4 s_so[kWarpsPerBlock][32], s_sn[kWarpsPerBlock][32];
  __shared__ float s_rt[kWarpsPerBlock][32];
  __shared__ int s_cell[SPLIT ? kColCells + 1 : 1];
  __shared__ int s_wsum[kWarpsPerBlock];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int oi = SPLIT ? (int) blockIdx.x : (int) blockIdx.x * kWarpsPerBlock + wid;
  if (oi >= nslices_own) return;   // (SPLIT: the whole block)
  const int slice = a.col_order[oi];
  const int rslot = slice * 32 + lane;
  const int d = a.g.slot_cluster[slice * 32];
  const bool valid = (a.g.slot_flags[rslot] & 1u) != 0;
  float4 po = make_float4(0, 0, 0, 0), pn = po;
  if (valid) {
    po = a.s_origin_area[rslot];
    pn = a.s_normal[rslot];
  }
  const int64_t l0 = a.vr_ptr[d], nruns = a.vr_ptr[d + 1] - l0;
  auto run_ptr = [&](const int4 run) -> const uint32_t * {
    return a.col[run.w & (kMaxWorld - 1)] + a.sbase[(int64_t) run.w * (a.nslices + 1) + slice] + run.y;
  };
  int64_t e0 = l0, e1 = l0 + nruns;   // this warp's part of the run list
  int n = 0;                          // list entries before this warp's next one (uniform)
  if (SPLIT) {
    const int64_t nbatches = (nruns + 31) >> 5;
    const int per_cell = (int) ((nbatches + kColCells - 1) / kColCells);   // batches of 32 runs per cell (1 unless the list is huge)
    const int ncells = per_cell ? (int) ((nbatches + per_cell - 1) / per_cell) : 0;
    for (int cidx = wid; cidx < ncells; cidx += kWarpsPerBlock) {
      int cnt = 0;
      const int64_t b0 = l0 + (int64_t) cidx * per_cell * 32, b1 = min(l0 + nruns, b0 + (int64_t) per_cell * 32);
      for (int64_t eb = b0; eb < b1; eb += 32) {
        const int nb = (int) min((int64_t) 32, b1 - eb);
        const int4 mine = lane < nb ? __ldg(a.vr + eb + lane) : make_int4(0, 0, 0, 0);
        const uint32_t *myptr = run_ptr(mine);
        for (int j = 0; j < nb; j += 4) {
          uint32_t w[4];
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const unsigned long long pj = __shfl_sync(0xffffffffu, (unsigned long long) myptr, (j + k) & 31);
            const int len = __shfl_sync(0xffffffffu, mine.z, (j + k) & 31);
            w[k] = (j + k < nb && lane < len) ? __ldg(reinterpret_cast<const uint32_t *>(pj) + lane) : 0u;
          }
#pragma unroll
          for (int k = 0; k < 4; k++) cnt += w[k] != 0;
        }
      }
      for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      if (lane == 0) s_cell[cidx] = cnt;
    }
    __syncthreads();
    // exclusive prefix over the cells (kColCells = 4 per thread)
    int v[4], tsum = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int cidx = threadIdx.x * 4 + k;
      v[k] = cidx < ncells ? s_cell[cidx] : 0;
      tsum += v[k];
    }
    int incl = tsum;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_wsum[wid] = incl;
    __syncthreads();
    int before = incl - tsum;
    for (int w = 0; w < wid; w++) before += s_wsum[w];
    int total = 0;
    for (int w = 0; w < kWarpsPerBlock; w++) total += s_wsum[w];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int cidx = threadIdx.x * 4 + k;
      if (cidx <= ncells && cidx <= kColCells) s_cell[cidx] = before;
      before += v[k];
    }
    __syncthreads();
    // warp w takes the cells from the first whose prefix reaches total * w / 8
    auto first_cell = [&](const int w) -> int {
      if (w >= kWarpsPerBlock) return ncells;
      const int target = (int) ((int64_t) total * w / kWarpsPerBlock);
      int lo = 0, hi = ncells;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_cell[mid] < target) lo = mid + 1;
        else hi = mid;
      }
      return lo;
    };
    const int c0 = first_cell(wid), c1 = first_cell(wid + 1);
    e0 = min(l0 + nruns, l0 + (int64_t) c0 * per_cell * 32);
    e1 = min(l0 + nruns, l0 + (int64_t) c1 * per_cell * 32);
    n = c0 < ncells ? s_cell[c0] : total;
  }
  const int head_end = (n + 7) & ~7;                   // entries before this one share their group of 8 with the previous warp
  const int64_t base = a.slice_base[slice];
  uint16_t *const w16 = reinterpret_cast<uint16_t *>(a.g_w);
  uint32_t wq[4] = {0, 0, 0, 0};                       // this lane's weights of the current group of 8 entries
  // 32 runs at a time (lane j holds run j), the words of four runs in flight while the previous four are emitted
  for (int64_t eb = e0; eb < e1; eb += 32) {
    const int nb = (int) min((int64_t) 32, e1 - eb);
    const int4 mine = lane < nb ? __ldg(a.vr + eb + lane) : make_int4(0, 0, 0, 0);
    const uint32_t *myptr = run_ptr(mine);
    auto fetch = [&](const int j) -> uint32_t {   // the word of source `lane` of run j (0 past the run / the batch)
      const unsigned long long pj = __shfl_sync(0xffffffffu, (unsigned long long) myptr, j & 31);
      const int len = __shfl_sync(0xffffffffu, mine.z, j & 31);
      return (j < nb && lane < len) ? __ldcs(reinterpret_cast<const uint32_t *>(pj) + lane) : 0u;
    };
    uint32_t cur[4], nxt[4];
#pragma unroll
    for (int k = 0; k < 4; k++) cur[k] = fetch(k);
    for (int j0 = 0; j0 < nb; j0 += 4) {
#pragma unroll
      for (int k = 0; k < 4; k++) nxt[k] = fetch(j0 + 4 + k);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int j = j0 + k;
        if (j >= nb) break;
        const uint32_t word = cur[k];
        unsigned nz = __ballot_sync(0xffffffffu, word != 0);
        if (nz == 0) continue;
        const uint32_t slot0 = (uint32_t) __shfl_sync(0xffffffffu, mine.x, j);
        const uint32_t sslot = slot0 + lane;
        __syncwarp();
        if (word != 0) {
          s_so[wid][lane] = a.s_origin_area[sslot];
          s_sn[wid][lane] = a.s_normal[sslot];
          s_rt[wid][lane] = a.row_total[sslot];
        }
        __syncwarp();
        while (nz) {
          const int src = __ffs(nz) - 1;
          nz &= nz - 1;
          const uint32_t w = __shfl_sync(0xffffffffu, word, src);
          uint32_t itrans = 0;
          if ((w >> lane) & 1u) {
            float dist;
            bool wr;
            const float trans = form_factor_pre(s_so[wid][src], s_sn[wid][src], po, pn, dist, wr);
            itrans = (uint32_t) (int) (trans * 65536.0f / s_rt[wid][src]) & 0xffffu;   // u16 store wraps (qrad3.c:283-285)
          }
          const int q = n & 7;
          if (SPLIT && n < head_end) w16[(((base + n) >> 3) * 32 + lane) * 8 + q] = (uint16_t) itrans;   // group shared with the warp before
          else wq[q >> 1] |= itrans << ((q & 1) * 16);
          if (lane == 0) a.g_src[base + n] = slot0 + src;
          n++;
          if ((n & 7) == 0 && n > head_end) {
            a.g_w[((base + n) / 8 - 1) * 32 + lane] = make_uint4(wq[0], wq[1], wq[2], wq[3]);
            wq[0] = wq[1] = wq[2] = wq[3] = 0;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; k++) cur[k] = nxt[k];
    }
  }
  if ((n & 7) && n > head_end) {
    // last, partial group: followed by the next warp's entries, or by zeros up to the end of the 32-entry batch (source
    // slot 0, zero weights: the buffers are cleared)
    if (SPLIT) {
      for (int q = 0; q < (n & 7); q++) w16[(((base + n) >> 3) * 32 + lane) * 8 + q] = (uint16_t) (wq[q >> 1] >> ((q & 1) * 16));
    } else
      a.g_w[((base + n) / 8) * 32 + lane] = make_uint4(wq[0], wq[1], wq[2], wq[3]);
  }
}

// ---- bounce --------------------------------------------------------------------------------
struct BounceArgs {
  const int32_t *slice_len;
  const int64_t *slice_base;
  const uint32_t *g_src;
  const uint4 *g_w;
  const float4 *send;            // radiosity / 0x10000 per source slot
  float4 *send_next;
  float4 *total;                 // totallight per slot
  const float4 *refl_area;       // xyz reflectivity, w area; w < 0 marks a sky patch
  float *rad_sum;                // [M] r0 + r1 + r2 of the new radiosity (for `added`), may be null
  int M;
  int slice_lo, slice_hi;        // this rank's slices
  const int32_t *order;          // this rank's slices, longest list first
  int32_t *next;                 // [2] work counters: launch `parity` draws from next[parity] and clears the other
  int parity;
  int slice_cut;                 // tools/bounce_shard_probe.py: slices >= slice_cut are skipped (INT_MAX otherwise)
};

// (float) lo16(W), (float) hi16(W) without the quarter-rate I2F: 0x4B000000 | w is the float 2^23 + w, exactly, so one
// byte permute per half and ONE packed subtraction give both weights of a 32-bit word.
__device__ __forceinline__ float2 u16x2_to_float2(const uint32_t w) {
  const float2 m = make_float2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7410)),
                               __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7432)));
  return __fadd2_rn(m, make_float2(-8388608.0f, -8388608.0f));
}

// Streaming (evict-first) / read-only loads of k_bounce's prefetch stage.
__device__ __forceinline__ uint4 ld_stream_u4(const uint4 *p) {
  uint4 v;
  asm volatile("ld.global.cs.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_vec_f4(const float4 *p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// ---- bulk-copy ring of k_bounce (sm_90+ PTX: cp.async.bulk + mbarrier complete_tx) ----
constexpr int kBounceStages = 4;                    // batches resident per warp: one in use, three in flight
constexpr int kBounceStageBytes = 2048 + 128;       // 32 entries: 32 lanes x 4 x 16 B of weights, 32 source slots
constexpr int kBounceWarpBytes = 9600;              // stages + parked vectors (2 x 96 floats) + mbarriers, 128-aligned
constexpr int kBounceVecOfs = kBounceStages * kBounceStageBytes;
constexpr int kBounceBarOfs = kBounceVecOfs + 2 * 96 * 4;
static_assert(kBounceBarOfs + kBounceStages * 8 <= kBounceWarpBytes, "k_bounce shared memory layout");

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(const uint32_t bar, const uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(const uint32_t bar, const uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(const uint32_t dst, const void *src, const uint32_t bytes, const uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(const uint32_t bar, const uint32_t parity) {
  uint32_t ok = 0;
  for (unsigned spins = 0; !ok; spins++) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (spins > (1u << 28)) __trap();   // a copy that never lands is a bug: fail loudly instead of hanging the GPU
  }
}

// ShootLight (as a gather) + CollectLight.  One warp per slice, lane = receiver; the list is walked in order, so
// every receiver accumulates its transfers in ascending source order like the reference's serial loop:
// illumination[j][l] += send[l] * trans->transfer (qrad3.c:439), mul and add rounded separately.  A zero weight adds
// +-0 and leaves the sum unchanged (the sum starts at +0 and never becomes -0).
// Data movement: the list is consumed in batches of 32 entries = 2 KB of weights + 128 B of source slots, contiguous
// in HBM.  Each warp owns a ring of four batch buffers (kBounceStages) in shared memory; lane 0 refills a buffer with two
// bulk async copies (cp.async.bulk, completion counted on an mbarrier) as soon as the warp has consumed it, so three batches
// are always in flight per warp without holding a register, and the radiosity vectors of the next batches are gathered while
// the current one is summed.  This matters when a GPU has few slices (8-GPU shard: 3 900 slices on 148 SMs): the
// kernel then lasts as long as its longest list, i.e. as fast as ONE warp can stream (the register-prefetch version
// took 2.0 ms for 2.5 GB per GPU at 8 GPUs because ptxas sinks the loads next to their use).
// Arithmetic (the r01g kernel issued 12 instructions per list entry -- 3 shuffles, 3 unpack/convert, 3 FMUL, 3 FADD --
// and sat at 58 % of the issue rate with HBM at 54 %, profiles/r01g_full_c5_k_bounce.txt): the 32 vectors of a batch
// are parked in shared memory channel-major (x[32] y[32] z[32], conflict-free stores), so FOUR entries cost three
// broadcast 128-bit loads, four byte permutes + two packed subtracts for their weights, six packed multiplies (sm_100
// FMUL2: two entries of a channel at once) and twelve scalar adds, which keep the serial order: 6.75 instructions per
// entry.  (Packed adds are not used: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under
// -fmad=false, which would change the rounding.)
// Slices are handed out longest-first from a work counter (persistent warps): with one block per 8 consecutive
// slices only 16 of 24 resident warps were alive on average (profiles/r01i_full_c5_k_bounce.txt).
__device__ __forceinline__ void bounce_ring_role(const BounceArgs &a, unsigned char *s_bounce, const int first_block) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned char *const wbase = s_bounce + wid * kBounceWarpBytes;
  float *const s_vec = reinterpret_cast<float *>(wbase + kBounceVecOfs);
  const uint32_t stage0 = smem_u32(wbase), bar0 = stage0 + kBounceBarOfs;
  if (lane == 0) {
    for (int s = 0; s < kBounceStages; s++) mbar_init(bar0 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if ((int) blockIdx.x == first_block && threadIdx.x == 0) a.next[a.parity ^ 1] = 0;
  const int nslices = a.slice_hi - a.slice_lo;
  uint32_t gb = 0;   // batches this warp has consumed so far: batch gb + i lives in buffer (gb + i) % stages
  for (;;) {
    int k = 0;
    if (lane == 0) k = atomicAdd(a.next + a.parity, 1);
    k = __shfl_sync(0xffffffffu, k, 0);     // also: every lane is done with the previous slice's shared memory
    if (k >= nslices) break;
    const int slice = a.order[k];
    if (slice >= a.slice_cut) continue;
    const int slot = slice * 32 + lane;
    const int64_t base = a.slice_base[slice];
    const int batches = (a.slice_len[slice] + 31) >> 5;
    const uint32_t *const ps = a.g_src + base;
    const uint4 *const pw = a.g_w + (base >> 3) * 32;
    auto issue = [&](const int i) {   // lane 0: start the copies of batch i of this slice
      const uint32_t s = (gb + (uint32_t) i) % kBounceStages;
      const uint32_t dst = stage0 + s * kBounceStageBytes, bar = bar0 + 8 * s;
      mbar_expect_tx(bar, kBounceStageBytes);
      bulk_g2s(dst, pw + (int64_t) i * 128, 2048, bar);
      bulk_g2s(dst + 2048, ps + (int64_t) i * 32, 128, bar);
    };
    auto arrived = [&](const int i) {   // wait for batch i and return this lane's source slot of it
      const uint32_t u = gb + (uint32_t) i, s = u % kBounceStages;
      mbar_wait(bar0 + 8 * s, (u / kBounceStages) & 1u);
      return *reinterpret_cast<const uint32_t *>(wbase + s * kBounceStageBytes + 2048 + lane * 4);
    };
    if (lane == 0)
      for (int i = 0; i < kBounceStages && i < batches; i++) issue(i);
    float ix = 0.f, iy = 0.f, iz = 0.f;
    // radiosity vectors of entry `lane` of the next two batches: gathered two batches ahead, so that a warp that has
    // an SM to itself does not wait for an L2 round trip per batch
    float4 nx = make_float4(0, 0, 0, 0), nx2 = nx;
    if (batches > 0) nx = __ldg(a.send + arrived(0));
    if (batches > 1) nx2 = __ldg(a.send + arrived(1));
// entries 4Q .. 4Q+3 of the batch: W0 holds the weights of the first two, W1 of the last two
#define QRAD_QUAD(Q, W0, W1)                                                         \
  do {                                                                               \
    const float4 vx = *reinterpret_cast<const float4 *>(sv + (Q) * 4);               \
    const float4 vy = *reinterpret_cast<const float4 *>(sv + 32 + (Q) * 4);          \
    const float4 vz = *reinterpret_cast<const float4 *>(sv + 64 + (Q) * 4);          \
    const float2 f0 = u16x2_to_float2(W0), f1 = u16x2_to_float2(W1);                 \
    const float2 px0 = __fmul2_rn(make_float2(vx.x, vx.y), f0);                      \
    const float2 py0 = __fmul2_rn(make_float2(vy.x, vy.y), f0);                      \
    const float2 pz0 = __fmul2_rn(make_float2(vz.x, vz.y), f0);                      \
    const float2 px1 = __fmul2_rn(make_float2(vx.z, vx.w), f1);                      \
    const float2 py1 = __fmul2_rn(make_float2(vy.z, vy.w), f1);                      \
    const float2 pz1 = __fmul2_rn(make_float2(vz.z, vz.w), f1);                      \
    ix = ix + px0.x;                                                                 \
    iy = iy + py0.x;                                                                 \
    iz = iz + pz0.x;                                                                 \
    ix = ix + px0.y;                                                                 \
    iy = iy + py0.y;                                                                 \
    iz = iz + pz0.y;                                                                 \
    ix = ix + px1.x;                                                                 \
    iy = iy + py1.x;                                                                 \
    iz = iz + pz1.x;                                                                 \
    ix = ix + px1.y;                                                                 \
    iy = iy + py1.y;                                                                 \
    iz = iz + pz1.y;                                                                 \
  } while (0)
// batch B: park its vectors V (double buffered: the other half may still be read by slower lanes), gather those of
// batch B + 2 into the same registers, consume the weights, hand the buffer back.  Two register sets alternate, so no
// instruction touches a gathered vector before it is parked (a register copy would wait for the load).
#define QRAD_BATCH(B, V)                                                                                              \
  do {                                                                                                                \
    float *const sv = s_vec + ((B) & 1) * 96;                                                                         \
    sv[lane] = V.x;                                                                                                   \
    sv[32 + lane] = V.y;                                                                                              \
    sv[64 + lane] = V.z;                                                                                              \
    __syncwarp();                                                                                                     \
    if ((B) + 2 < batches) V = __ldg(a.send + arrived((B) + 2));                                                      \
    const uint4 *const cw =                                                                                           \
        reinterpret_cast<const uint4 *>(wbase + ((gb + (uint32_t) (B)) % kBounceStages) * kBounceStageBytes) + lane;  \
    const uint4 c0 = cw[0], c1 = cw[32], c2 = cw[64], c3 = cw[96];                                                    \
    QRAD_QUAD(0, c0.x, c0.y);                                                                                         \
    QRAD_QUAD(1, c0.z, c0.w);                                                                                         \
    QRAD_QUAD(2, c1.x, c1.y);                                                                                         \
    QRAD_QUAD(3, c1.z, c1.w);                                                                                         \
    QRAD_QUAD(4, c2.x, c2.y);                                                                                         \
    QRAD_QUAD(5, c2.z, c2.w);                                                                                         \
    QRAD_QUAD(6, c3.x, c3.y);                                                                                         \
    QRAD_QUAD(7, c3.z, c3.w);                                                                                         \
    __syncwarp(); /* every lane holds its weights of batch B in registers: the buffer may be refilled */              \
    if (lane == 0 && (B) + kBounceStages < batches) issue((B) + kBounceStages);                                       \
  } while (0)
    for (int b = 0; b < batches; b += 2) {
      QRAD_BATCH(b, nx);
      if (b + 1 < batches) QRAD_BATCH(b + 1, nx2);
    }
#undef QRAD_BATCH
#undef QRAD_QUAD
    gb += (uint32_t) batches;
    // CollectLight, qrad3.c:383-409
    const float4 ra = a.refl_area[slot];
    float4 t = a.total[slot];
    float rx = 0.f, ry = 0.f, rz = 0.f;
    if (ra.w > 0) {  // not sky (and not a padding slot)
      t.x += ix / ra.w;
      t.y += iy / ra.w;
      t.z += iz / ra.w;
      rx = ix * ra.x;
      ry = iy * ra.y;
      rz = iz * ra.z;
      a.total[slot] = t;
    }
    if (a.rad_sum) a.rad_sum[slot] = rx + ry + rz;
    a.send_next[slot] = make_float4(rx / 65536.0f, ry / 65536.0f, rz / 65536.0f, 0.f);
  }
}

__global__ void __launch_bounds__(kThreads) k_bounce(const BounceArgs a) {
  extern __shared__ __align__(128) unsigned char s_bounce[];
  bounce_ring_role(a, s_bounce, 0);
}

// k_bounce with register prefetch instead of the bulk-copy ring: same arithmetic, same order; the weights of the next
// batch are ordinary 128-bit streaming loads.  ptxas places them at the end of the batch and keeps ONE set of weight
// registers (48 registers, 40 resident warps per SM), so it is the other warps that cover the HBM latency: the
// faster kernel when a GPU has many slices per resident warp (one B200, C5: 3.2 ms per bounce = 6.26 TB/s of stored
// matrix against 3.5 ms with the ring), the slower one when it has few (2 GPUs: 2.6 ms against 2.2 ms; 8 GPUs: 2.0 ms
// for an eighth of the matrix).  bounce_step() picks by slices per resident warp.
// The prefetch is unconditional: g_src / g_w are padded by two batches / one batch and hold valid slots.
__global__ void __launch_bounds__(kThreads) k_bounce_regs(const BounceArgs a) {
  __shared__ __align__(16) float s_vec[kWarpsPerBlock][2][3 * 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (blockIdx.x == 0 && threadIdx.x == 0) a.next[a.parity ^ 1] = 0;
  const int nslices = a.slice_hi - a.slice_lo;
  for (;;) {
    int k = 0;
    if (lane == 0) k = atomicAdd(a.next + a.parity, 1);
    k = __shfl_sync(0xffffffffu, k, 0);     // also: every lane is done with the previous slice's shared memory
    if (k >= nslices) break;
    const int slice = a.order[k];
    if (slice >= a.slice_cut) continue;
    const int slot = slice * 32 + lane;
    const int64_t base = a.slice_base[slice];
    const int batches = (a.slice_len[slice] + 31) >> 5;
    const uint32_t *ps = a.g_src + base + lane;
    const uint4 *pw = a.g_w + (base >> 3) * 32 + lane;
    float ix = 0.f, iy = 0.f, iz = 0.f;
    uint32_t nidx = 0;                        // source slot of entry `lane` of the batch after next
    float4 nx = make_float4(0, 0, 0, 0);      // radiosity vector of entry `lane` of the next batch
    // this lane's weights of the current / next batch: two register sets that swap roles every batch (no copies)
    uint4 wa0 = make_uint4(0, 0, 0, 0), wa1 = wa0, wa2 = wa0, wa3 = wa0, wb0 = wa0, wb1 = wa0, wb2 = wa0, wb3 = wa0;
    if (batches > 0) {
      nx = __ldg(a.send + __ldcs(ps));
      wa0 = __ldcs(pw);
      wa1 = __ldcs(pw + 32);
      wa2 = __ldcs(pw + 64);
      wa3 = __ldcs(pw + 96);
      nidx = __ldcs(ps + 32);
    }
// entries 4Q .. 4Q+3 of the batch: W0 holds the weights of the first two, W1 of the last two
#define QRAD_QUAD(Q, W0, W1)                                                         \
  do {                                                                               \
    const float4 vx = *reinterpret_cast<const float4 *>(sv + (Q) * 4);               \
    const float4 vy = *reinterpret_cast<const float4 *>(sv + 32 + (Q) * 4);          \
    const float4 vz = *reinterpret_cast<const float4 *>(sv + 64 + (Q) * 4);          \
    const float2 f0 = u16x2_to_float2(W0), f1 = u16x2_to_float2(W1);                 \
    const float2 px0 = __fmul2_rn(make_float2(vx.x, vx.y), f0);                      \
    const float2 py0 = __fmul2_rn(make_float2(vy.x, vy.y), f0);                      \
    const float2 pz0 = __fmul2_rn(make_float2(vz.x, vz.y), f0);                      \
    const float2 px1 = __fmul2_rn(make_float2(vx.z, vx.w), f1);                      \
    const float2 py1 = __fmul2_rn(make_float2(vy.z, vy.w), f1);                      \
    const float2 pz1 = __fmul2_rn(make_float2(vz.z, vz.w), f1);                      \
    ix = ix + px0.x;                                                                 \
    iy = iy + py0.x;                                                                 \
    iz = iz + pz0.x;                                                                 \
    ix = ix + px0.y;                                                                 \
    iy = iy + py0.y;                                                                 \
    iz = iz + pz0.y;                                                                 \
    ix = ix + px1.x;                                                                 \
    iy = iy + py1.x;                                                                 \
    iz = iz + pz1.x;                                                                 \
    ix = ix + px1.y;                                                                 \
    iy = iy + py1.y;                                                                 \
    iz = iz + pz1.y;                                                                 \
  } while (0)
// batch B: park the vectors (double buffered: the other half may still be read by slower lanes), start the loads of
// batch B + 1 into the N registers, then consume the C registers
#define QRAD_BATCH(B, C0, C1, C2, C3, N0, N1, N2, N3)            \
  do {                                                           \
    float *sv = s_vec[wid][(B) & 1];                             \
    sv[lane] = nx.x;                                             \
    sv[32 + lane] = nx.y;                                        \
    sv[64 + lane] = nx.z;                                        \
    __syncwarp();                                                \
    /* unconditional: the buffers are padded by two batches and hold valid slots */ \
    nx = ld_vec_f4(a.send + nidx);                               \
    {                                                            \
      const uint4 *q = pw + 128 * ((B) + 1);                     \
      N0 = ld_stream_u4(q);                                      \
      N1 = ld_stream_u4(q + 32);                                 \
      N2 = ld_stream_u4(q + 64);                                 \
      N3 = ld_stream_u4(q + 96);                                 \
      nidx = ld_stream_u32(ps + 32 * ((B) + 2));                 \
    }                                                            \
    QRAD_QUAD(0, C0.x, C0.y);                                    \
    QRAD_QUAD(1, C0.z, C0.w);                                    \
    QRAD_QUAD(2, C1.x, C1.y);                                    \
    QRAD_QUAD(3, C1.z, C1.w);                                    \
    QRAD_QUAD(4, C2.x, C2.y);                                    \
    QRAD_QUAD(5, C2.z, C2.w);                                    \
    QRAD_QUAD(6, C3.x, C3.y);                                    \
    QRAD_QUAD(7, C3.z, C3.w);                                    \
  } while (0)
    for (int b = 0; b < batches; b += 2) {
      QRAD_BATCH(b, wa0, wa1, wa2, wa3, wb0, wb1, wb2, wb3);
      if (b + 1 < batches) QRAD_BATCH(b + 1, wb0, wb1, wb2, wb3, wa0, wa1, wa2, wa3);
    }
#undef QRAD_BATCH
#undef QRAD_QUAD
    // CollectLight, qrad3.c:383-409
    const float4 ra = a.refl_area[slot];
    float4 t = a.total[slot];
    float rx = 0.f, ry = 0.f, rz = 0.f;
    if (ra.w > 0) {  // not sky (and not a padding slot)
      t.x += ix / ra.w;
      t.y += iy / ra.w;
      t.z += iz / ra.w;
      rx = ix * ra.x;
      ry = iy * ra.y;
      rz = iz * ra.z;
      a.total[slot] = t;
    }
    if (a.rad_sum) a.rad_sum[slot] = rx + ry + rz;
    a.send_next[slot] = make_float4(rx / 65536.0f, ry / 65536.0f, rz / 65536.0f, 0.f);
  }
}

// ---- block-cooperative role for the few very long lists -------------------------------------------------------------
// The sum of a receiver is a serial fold in source order, so a list cannot be split across warps -- but everything
// EXCEPT the adds can: 7 producer warps take the batches of 32 entries in turn (batch b -> producer b mod 7), gather the
// radiosity vectors, convert the weights and form the products send * weight exactly as k_bounce does, and park them in
// shared memory; the consumer warp only waits for the next parked batch, pulls it into registers (24 LDS.128) and adds
// it to the three running sums of its receiver in list order.  One warp on its own needs about 0.5 us per batch
// (latency of the gathers and of the weight stream); the consumer's chain of dependent adds needs a fraction of that, so
// the launch of a shard no longer lasts as long as its longest list (C5: 66 637 entries against a mean of 9 096;
// DESIGN.md section 7).  Hand-over: one mbarrier pair (full / empty) per producer, one elected arrival each; a producer
// parks one batch (four groups of 8 entries) and refills it as soon as the consumer has taken it, six batches of the
// other producers later.  Two batches are in flight per producer (two register sets that swap roles), their source slots
// are fetched three batches ahead.  Results are bit-identical to k_bounce: the same products (packed FMUL, rounded once),
// added in the same order.
constexpr int kCoopProducers = kWarpsPerBlock - 1;
constexpr int kCoopGroupBytes = 6 * 32 * 16;                    // 8 entries x 3 channels x 32 lanes, as 6 float4 per lane
constexpr int kCoopBarOfs = kCoopProducers * 4 * kCoopGroupBytes;                 // full[7], empty[7]
constexpr int kCoopVecOfs = kCoopBarOfs + 128;                                    // parked vectors: [7][2][96] floats
constexpr int kCoopSmemUsed = kCoopVecOfs + kCoopProducers * 2 * 96 * 4;
constexpr int kShardSmemBytes = kCoopSmemUsed > kWarpsPerBlock * kBounceWarpBytes ? kCoopSmemUsed : kWarpsPerBlock * kBounceWarpBytes;

__device__ __forceinline__ void mbar_arrive(const uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void bounce_coop_role(const BounceArgs &a, unsigned char *s_coop, const int32_t *__restrict__ coop_slices,
                                                 const int ncoop, int32_t *__restrict__ coop_next) {
  __shared__ int s_pick;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t sbase = smem_u32(s_coop);
  const uint32_t bar_full = sbase + kCoopBarOfs, bar_empty = bar_full + kCoopProducers * 8;
  if (threadIdx.x == 0) {
    for (int k = 0; k < kCoopProducers; k++) {
      mbar_init(bar_full + 8 * k, 1);
      mbar_init(bar_empty + 8 * k, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) coop_next[a.parity ^ 1] = 0;
  int uses = 0;   // consumer: lane p counts the batches taken from producer p; producer: the batches it has parked
  constexpr int P = kCoopProducers;
  for (;;) {
    __syncthreads();   // everybody is done with the previous slice
    if (threadIdx.x == 0) s_pick = atomicAdd(coop_next + a.parity, 1);
    __syncthreads();
    const int pick = s_pick;
    if (pick >= ncoop) break;
    const int slice = coop_slices[pick];
    const int64_t base = a.slice_base[slice];
    const int batches = (a.slice_len[slice] + 31) >> 5;
    if (wid == 0) {
      // ---- consumer: the serial adds, in list order ----
      float ix = 0.f, iy = 0.f, iz = 0.f;
      for (int b = 0; b < batches; b++) {
        const int p = b % P;
        const uint32_t par = (uint32_t) __shfl_sync(0xffffffffu, uses, p) & 1u;
        if (lane == p) uses++;
        mbar_wait(bar_full + 8 * p, par);
        // the whole batch into registers first, so that the producer can refill at once and no add waits for shared memory
        float4 v[4][6];
#pragma unroll
        for (int g = 0; g < 4; g++) {
          const float4 *grp = reinterpret_cast<const float4 *>(s_coop + (p * 4 + g) * kCoopGroupBytes) + lane;
#pragma unroll
          for (int k = 0; k < 6; k++) v[g][k] = grp[32 * k];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * p);
#pragma unroll
        for (int g = 0; g < 4; g++) {
          const float4 x0 = v[g][0], x1 = v[g][1], y0 = v[g][2], y1 = v[g][3], z0 = v[g][4], z1 = v[g][5];
          ix = ix + x0.x; iy = iy + y0.x; iz = iz + z0.x;
          ix = ix + x0.y; iy = iy + y0.y; iz = iz + z0.y;
          ix = ix + x0.z; iy = iy + y0.z; iz = iz + z0.z;
          ix = ix + x0.w; iy = iy + y0.w; iz = iz + z0.w;
          ix = ix + x1.x; iy = iy + y1.x; iz = iz + z1.x;
          ix = ix + x1.y; iy = iy + y1.y; iz = iz + z1.y;
          ix = ix + x1.z; iy = iy + y1.z; iz = iz + z1.z;
          ix = ix + x1.w; iy = iy + y1.w; iz = iz + z1.w;
        }
      }
      // CollectLight, qrad3.c:383-409
      const int slot = slice * 32 + lane;
      const float4 ra = a.refl_area[slot];
      float4 t = a.total[slot];
      float rx = 0.f, ry = 0.f, rz = 0.f;
      if (ra.w > 0) {  // not sky (and not a padding slot)
        t.x += ix / ra.w;
        t.y += iy / ra.w;
        t.z += iz / ra.w;
        rx = ix * ra.x;
        ry = iy * ra.y;
        rz = iz * ra.z;
        a.total[slot] = t;
      }
      if (a.rad_sum) a.rad_sum[slot] = rx + ry + rz;
      a.send_next[slot] = make_float4(rx / 65536.0f, ry / 65536.0f, rz / 65536.0f, 0.f);
    } else {
      // ---- producers: batch b = p, p + 7, ... ----
      const int p = wid - 1;
      float *const s_vec = reinterpret_cast<float *>(s_coop + kCoopVecOfs) + p * 2 * 96;
      const uint32_t *ps = a.g_src + base + lane;
      const uint4 *pw = a.g_w + (base >> 3) * 32 + lane;
      uint4 wa0 = make_uint4(0, 0, 0, 0), wa1 = wa0, wa2 = wa0, wa3 = wa0, wb0 = wa0, wb1 = wa0, wb2 = wa0, wb3 = wa0;
      float4 nxa = make_float4(0, 0, 0, 0), nxb = nxa;
      uint32_t idx2 = 0;   // source slot of entry `lane` of the batch after the next two
      if (p < batches) {
        nxa = __ldg(a.send + __ldcs(ps + 32 * p));
        const uint4 *q = pw + 128 * p;
        wa0 = __ldcs(q); wa1 = __ldcs(q + 32); wa2 = __ldcs(q + 64); wa3 = __ldcs(q + 96);
      }
      if (p + P < batches) {
        nxb = __ldg(a.send + __ldcs(ps + 32 * (p + P)));
        const uint4 *q = pw + 128 * (p + P);
        wb0 = __ldcs(q); wb1 = __ldcs(q + 32); wb2 = __ldcs(q + 64); wb3 = __ldcs(q + 96);
      }
      if (p + 2 * P < batches) idx2 = __ldcs(ps + 32 * (p + 2 * P));
// entries 8G .. 8G+7 of the batch: weights WA (first four) and WB (last four), two 32-bit words each
#define QRAD_COOP_GROUP(G, WA0, WA1, WB0, WB1)                                                                  \
  do {                                                                                                           \
    const float4 vxa = *reinterpret_cast<const float4 *>(sv + (G) * 8), vxb = *reinterpret_cast<const float4 *>(sv + (G) * 8 + 4); \
    const float4 vya = *reinterpret_cast<const float4 *>(sv + 32 + (G) * 8), vyb = *reinterpret_cast<const float4 *>(sv + 32 + (G) * 8 + 4); \
    const float4 vza = *reinterpret_cast<const float4 *>(sv + 64 + (G) * 8), vzb = *reinterpret_cast<const float4 *>(sv + 64 + (G) * 8 + 4); \
    const float2 fa0 = u16x2_to_float2(WA0), fa1 = u16x2_to_float2(WA1), fb0 = u16x2_to_float2(WB0), fb1 = u16x2_to_float2(WB1); \
    const float2 xa0 = __fmul2_rn(make_float2(vxa.x, vxa.y), fa0), xa1 = __fmul2_rn(make_float2(vxa.z, vxa.w), fa1);  \
    const float2 xb0 = __fmul2_rn(make_float2(vxb.x, vxb.y), fb0), xb1 = __fmul2_rn(make_float2(vxb.z, vxb.w), fb1);  \
    const float2 ya0 = __fmul2_rn(make_float2(vya.x, vya.y), fa0), ya1 = __fmul2_rn(make_float2(vya.z, vya.w), fa1);  \
    const float2 yb0 = __fmul2_rn(make_float2(vyb.x, vyb.y), fb0), yb1 = __fmul2_rn(make_float2(vyb.z, vyb.w), fb1);  \
    const float2 za0 = __fmul2_rn(make_float2(vza.x, vza.y), fa0), za1 = __fmul2_rn(make_float2(vza.z, vza.w), fa1);  \
    const float2 zb0 = __fmul2_rn(make_float2(vzb.x, vzb.y), fb0), zb1 = __fmul2_rn(make_float2(vzb.z, vzb.w), fb1);  \
    float4 *grp = reinterpret_cast<float4 *>(s_coop + (p * 4 + (G)) * kCoopGroupBytes) + lane;                   \
    grp[0] = make_float4(xa0.x, xa0.y, xa1.x, xa1.y);                                                            \
    grp[32] = make_float4(xb0.x, xb0.y, xb1.x, xb1.y);                                                           \
    grp[64] = make_float4(ya0.x, ya0.y, ya1.x, ya1.y);                                                           \
    grp[96] = make_float4(yb0.x, yb0.y, yb1.x, yb1.y);                                                           \
    grp[128] = make_float4(za0.x, za0.y, za1.x, za1.y);                                                          \
    grp[160] = make_float4(zb0.x, zb0.y, zb1.x, zb1.y);                                                          \
  } while (0)
// one batch: park the vectors V of batch B, refill V / W0..3 with batch B + 2P (its source slots are in idx2), fetch the
// source slots of batch B + 3P, then form and park the products of the copies c0..3
#define QRAD_COOP_BATCH(B, V, W0, W1, W2, W3)                                                                    \
  do {                                                                                                           \
    const int use = uses;                                                                                        \
    float *const sv = s_vec + (use & 1) * 96;                                                                    \
    sv[lane] = V.x;                                                                                              \
    sv[32 + lane] = V.y;                                                                                         \
    sv[64 + lane] = V.z;                                                                                         \
    __syncwarp();                                                                                                \
    const uint4 c0 = W0, c1 = W1, c2 = W2, c3 = W3;                                                              \
    if ((B) + 2 * P < batches) {                                                                                 \
      V = ld_vec_f4(a.send + idx2);                                                                              \
      const uint4 *q = pw + 128 * ((B) + 2 * P);                                                                 \
      W0 = ld_stream_u4(q); W1 = ld_stream_u4(q + 32); W2 = ld_stream_u4(q + 64); W3 = ld_stream_u4(q + 96);     \
    }                                                                                                            \
    if ((B) + 3 * P < batches) idx2 = ld_stream_u32(ps + 32 * ((B) + 3 * P));                                    \
    /* the consumer must have taken this producer's previous batch */                                           \
    if (use > 0) mbar_wait(bar_empty + 8 * p, (uint32_t) (use + 1) & 1u);                                        \
    QRAD_COOP_GROUP(0, c0.x, c0.y, c0.z, c0.w);                                                                  \
    QRAD_COOP_GROUP(1, c1.x, c1.y, c1.z, c1.w);                                                                  \
    QRAD_COOP_GROUP(2, c2.x, c2.y, c2.z, c2.w);                                                                  \
    QRAD_COOP_GROUP(3, c3.x, c3.y, c3.z, c3.w);                                                                  \
    __syncwarp();                                                                                                \
    if (lane == 0) mbar_arrive(bar_full + 8 * p);                                                                \
    uses++;                                                                                                      \
  } while (0)
      for (int b = p; b < batches; b += 2 * P) {
        QRAD_COOP_BATCH(b, nxa, wa0, wa1, wa2, wa3);
        if (b + P < batches) QRAD_COOP_BATCH(b + P, nxb, wb0, wb1, wb2, wb3);
      }
#undef QRAD_COOP_BATCH
#undef QRAD_COOP_GROUP
    }
  }
}

// One launch for a shard with very long lists: the first `coop_blocks` blocks take them block-cooperatively, the others
// run the warp-per-slice ring on the rest.  One grid, so that the long lists start at once (a kernel on a second stream
// only starts when the persistent blocks of the first leave room).
__global__ void __launch_bounds__(kThreads, 2) k_bounce_shard(const BounceArgs a, const int32_t *__restrict__ coop_slices,
                                                               const int ncoop, int32_t *__restrict__ coop_next, const int coop_blocks) {
  extern __shared__ __align__(128) unsigned char s_shard[];
  if ((int) blockIdx.x < coop_blocks) bounce_coop_role(a, s_shard, coop_slices, ncoop, coop_next);
  else bounce_ring_role(a, s_shard, coop_blocks);
}

// serial float sum of rad_sum over PATCH order (CollectLight's `total`), one warp
__global__ void k_ordered_sum(const float *__restrict__ v, const int32_t *__restrict__ patch_slot, int n,
                              float *__restrict__ out) {
  const int lane = threadIdx.x;
  float total = 0.f;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    float x = 0.f;
    bool has = false;
    if (i < n) {
      const int s = patch_slot[i];
      if (s >= 0) {
        x = v[s];
        has = true;
      }
    }
    unsigned m = __ballot_sync(0xffffffffu, has);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      total += __shfl_sync(0xffffffffu, x, src);
    }
  }
  if (lane == 0) *out = total;
}

__global__ void k_init_bounce(int M, const int32_t *__restrict__ slot_patch, const float *__restrict__ samplelight,
                              const float *__restrict__ totallight, const float4 *__restrict__ p_refl,
                              const float4 *__restrict__ p_origin_area, float4 *__restrict__ send,
                              float4 *__restrict__ total, float4 *__restrict__ refl_area) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= M) return;
  const int p = slot_patch[s];
  if (p < 0) {
    send[s] = make_float4(0, 0, 0, 0);
    total[s] = make_float4(0, 0, 0, 0);
    refl_area[s] = make_float4(0, 0, 0, -1.f);
    return;
  }
  const float4 r = p_refl[p];
  const float area = p_origin_area[p].w;
  // radiosity = samplelight * reflectivity * area   (qrad3.c:455-459), then send = radiosity / 0x10000
  const float rx = samplelight[p * 3] * r.x * area;
  const float ry = samplelight[p * 3 + 1] * r.y * area;
  const float rz = samplelight[p * 3 + 2] * r.z * area;
  send[s] = make_float4(rx / 65536.0f, ry / 65536.0f, rz / 65536.0f, 0.f);
  total[s] = make_float4(totallight[p * 3], totallight[p * 3 + 1], totallight[p * 3 + 2], 0.f);
  refl_area[s] = make_float4(r.x, r.y, r.z, r.w != 0.f ? -area : area);
}

__global__ void k_finish_bounce(int N, const int32_t *__restrict__ patch_slot, const float4 *__restrict__ total,
                                const float4 *__restrict__ send, float *__restrict__ totallight,
                                float *__restrict__ radiosity, int32_t *__restrict__ negative) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  const int s = patch_slot[p];
  if (s < 0) return;
  const float4 t = total[s];
  if (t.x < 0 || t.y < 0 || t.z < 0) *negative = 1;   // CheckPatches, qrad3.c:476-486
  totallight[p * 3] = t.x;
  totallight[p * 3 + 1] = t.y;
  totallight[p * 3 + 2] = t.z;
  const float4 r = send[s];
  radiosity[p * 3] = r.x * 65536.0f;
  radiosity[p * 3 + 1] = r.y * 65536.0f;
  radiosity[p * 3 + 2] = r.z * 65536.0f;
}


// slot-order copies of the patch geometry, and the bounding box of every slice's patch origins (shaft_blocked)
// The run list of k_visibility: per own row group (one block) its (chunk, part) runs, the ones that hold words of the group's
// own cluster in the front section of the list ("heavy": their rays are mostly clear and get walked), the rest behind.
__global__ void __launch_bounds__(128) k_run_list(const RunGroup *groups, int run_rows, int4 *list) {
  const RunGroup r = groups[blockIdx.x];
  const int parts = (r.rows + run_rows - 1) / run_rows, nself = r.self_hi - r.self_lo + 1;
  for (int k = threadIdx.x; k < r.nchunks * parts; k += blockDim.x) {
    const int ch = k / parts, part = k - ch * parts, r0 = part * run_rows;
    const int4 e = make_int4(r.grp, ch, r0, min(run_rows, r.rows - r0));
    if (ch >= r.self_lo && ch <= r.self_hi) list[r.heavy_at + (int64_t) (ch - r.self_lo) * parts + part] = e;
    else list[r.light_at + (int64_t) (ch < r.self_lo ? ch : ch - nself) * parts + part] = e;
  }
}

__global__ void k_slot_geometry(int M, const int32_t *__restrict__ slot_patch, const float4 *__restrict__ p_origin_area,
                                const float4 *__restrict__ p_normal, float4 *__restrict__ s_origin_area,
                                float4 *__restrict__ s_normal, float4 *__restrict__ box_min, float4 *__restrict__ box_max) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;   // blockDim.x is a multiple of 32 and so is M
  if (s >= M) return;
  const int p = slot_patch[s];
  float4 o = make_float4(0, 0, 0, 0), n = o;
  if (p >= 0) {
    o = p_origin_area[p];
    n = p_normal[p];
  }
  s_origin_area[s] = o;
  s_normal[s] = n;
  float lo[3] = {p >= 0 ? o.x : 3e38f, p >= 0 ? o.y : 3e38f, p >= 0 ? o.z : 3e38f};
  float hi[3] = {p >= 0 ? o.x : -3e38f, p >= 0 ? o.y : -3e38f, p >= 0 ? o.z : -3e38f};
  for (int d = 16; d; d >>= 1)
    for (int k = 0; k < 3; k++) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], d));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], d));
    }
  // the axis code all patches of the slice share (0: none, mixed, or a normal that is not axial), for the word-level
  // sign test of k_visibility
  const int code = p >= 0 ? (int) n.w : 0;
  const unsigned have = __ballot_sync(0xffffffffu, p >= 0);
  const int first = __shfl_sync(0xffffffffu, code, have ? __ffs(have) - 1 : 0);
  const bool same = __all_sync(0xffffffffu, p < 0 || code == first);
  if ((threadIdx.x & 31) == 0) {
    box_min[s >> 5] = make_float4(lo[0], lo[1], lo[2], (have && same) ? (float) first : 0.f);
    box_max[s >> 5] = make_float4(hi[0], hi[1], hi[2], 0);
  }
}

void nccl_ok(ncclResult_t r, const char *what) {
  if (r != ncclSuccess) dfail("%s failed: %s", what, nccl().GetErrorString(r));
}

Geo geo_of(qrad_cuda_ctx *c) {
  Geo g{};
  g.cl_slot_start = c->d_cl_slot_start.p;
  g.cl_count = c->d_cl_count.p;
  g.slot_patch = c->slot_patch.p;
  g.slot_cluster = c->slot_cluster.p;
  g.slot_flags = c->slot_flags.p;
  g.slice_count = c->slice_count.p;
  g.wb_first = c->d_wb_first.p;
  g.wslice = c->wslice.p;
  g.wbase = c->wbase.p;
  g.nwords = c->d_nwords.p;
  g.nc = c->plan.nc;
  return g;
}

RowArgs row_args(qrad_cuda_ctx *c) {
  RowArgs ra{};
  ra.s_origin_area = c->s_origin_area.p;
  ra.s_normal = c->s_normal.p;
  ra.g = geo_of(c);
  ra.vl_ptr = c->d_vl_ptr.p;
  ra.vl = c->vl.p;
  ra.bits = c->bits.p;
  ra.row_total = c->row_total.p;
  ra.row_count = c->row_count.p;
  return ra;
}

// Every owner sends its part of a per-slot array to everybody (in place).  The parts differ in length (receiver slices are
// dealt by work, not by count), so there is no all-gather that fits the array as it lies; a grouped broadcast per rank
// does, at 0.45 ms per call on eight GPUs -- and BounceLight calls this once per bounce.  16-byte records therefore go
// through a staging buffer of `world` equal parts: pack the own part, ONE in-place all-gather, unpack the others.
struct ShareArgs {
  int world, rank, maxlen;
  int first[kMaxWorld], len[kMaxWorld];   // slots
};
__global__ void __launch_bounds__(256) k_share_pack(const ShareArgs a, const float4 *__restrict__ buf, float4 *__restrict__ staging) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.len[a.rank]) staging[(size_t) a.rank * a.maxlen + i] = buf[a.first[a.rank] + i];
}
__global__ void __launch_bounds__(256) k_share_unpack(const ShareArgs a, const float4 *__restrict__ staging, float4 *__restrict__ buf) {
  const int q = blockIdx.y;
  if (q == a.rank) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.len[q]) buf[a.first[q] + i] = staging[(size_t) q * a.maxlen + i];
}

void share_slices(qrad_cuda_ctx *c, void *buf, size_t bytes_per_slot, cudaStream_t st) {
  const TransferPlan &P = c->plan;
  const NcclApi &n = nccl();
  if (bytes_per_slot == sizeof(float4) && P.world <= kMaxWorld && !getenv("QRAD_SHARE_BROADCAST")) {
    ShareArgs a{};
    a.world = P.world;
    a.rank = P.rank;
    for (int q = 0; q < P.world; q++) {
      a.first[q] = P.slice_lo[q] * 32;
      a.len[q] = (P.slice_hi[q] - P.slice_lo[q]) * 32;
      a.maxlen = std::max(a.maxlen, a.len[q]);
    }
    if (a.maxlen == 0) return;
    c->share_staging.alloc((size_t) a.maxlen * P.world);
    const int blocks = (a.maxlen + 255) / 256;
    k_share_pack<<<blocks, 256, 0, st>>>(a, (const float4 *) buf, c->share_staging.p);
    c->launch_check("k_share_pack");
    nccl_ok(n.AllGather(c->share_staging.p + (size_t) a.rank * a.maxlen, c->share_staging.p, (size_t) a.maxlen * 4, ncclFloat, c->comm, st),
            "ncclAllGather");
    k_share_unpack<<<dim3(blocks, P.world), 256, 0, st>>>(a, c->share_staging.p, (float4 *) buf);
    c->launch_check("k_share_unpack");
    return;
  }
  nccl_ok(n.GroupStart(), "ncclGroupStart");
  for (int q = 0; q < P.world; q++) {
    const size_t first = (size_t) P.slice_lo[q] * 32 * bytes_per_slot, count = (size_t) (P.slice_hi[q] - P.slice_lo[q]) * 32 * bytes_per_slot;
    if (!count) continue;
    char *p = (char *) buf + first;
    nccl_ok(n.Broadcast(p, p, count, ncclChar, q, c->comm, st), "ncclBroadcast");
  }
  nccl_ok(n.GroupEnd(), "ncclGroupEnd");
}

float elapsed(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0;
  QCUDA(cudaEventElapsedTime(&ms, a, b));
  return ms;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
static void transfers_lists(qrad_cuda_ctx *c);

// phase 0: the plan, slot order, word tables, pass 1 (this rank's row groups)
void transfers_trace(qrad_cuda_ctx *c, int nopvs) {
  if (!c->numnodes || !c->N) dfail("make_transfers: BSP and patches must be uploaded first");
  if (!c->have_vis) dfail("make_transfers: map has no visibility data (the reference forces numbounce = 0)");
  c->t_transfers.start(c->stream);
  const auto wall0 = std::chrono::steady_clock::now();
  c->transfers_done = false;
  const int N = c->N;
  cudaStream_t st = c->stream;
  // With -nopvs every patch receives (even those in solid); sources still need a valid cluster
  // (PvsForOrigin fails for them, qrad3.c:208-209).  That case is one pseudo cluster holding everybody.
  const int nc = nopvs ? 1 : c->numclusters;
  {
    bool have = false;
    if (c->plan_future.valid()) {   // started by upload_patches
      try {
        std::shared_ptr<TransferPlan> p = c->plan_future.get();
        if (!nopvs && p->N == N && p->world == c->world && p->rank == c->rank && p->nc == nc) {
          c->plan = std::move(*p);
          have = true;
        }
      } catch (const std::exception &e) {
        dfail("%s", e.what());
      }
    }
    if (!have) {
      std::vector<int32_t> eff((size_t) N);
      for (int i = 0; i < N; i++) {
        const int cl = c->h_cluster[i];
        if (cl >= c->numclusters) dfail("patch %d has cluster %d >= numclusters %d", i, cl, c->numclusters);
        eff[i] = nopvs ? 0 : cl;
      }
      try {
        build_plan(c->plan, N, eff.data(), nc, nopvs ? nullptr : c->h_pvs.data(), c->pvs_words, c->world, c->rank, 0);
      } catch (const std::exception &e) {
        dfail("%s", e.what());
      }
    }
  }
  const TransferPlan &P = c->plan;
  if (P.world > kMaxWorld) dfail("make_transfers: at most %d ranks", kMaxWorld);
  const int M = P.M;
  c->M = M;
  {
    std::vector<uint8_t> flags((size_t) M, 0);
    for (int s = 0; s < M; s++) {
      const int p = P.slot_patch[s];
      if (p >= 0) flags[s] = (uint8_t) (1 | (c->h_cluster[p] >= 0 ? 2 : 0));
    }
    c->slot_flags.upload(flags, st);
  }
  c->slot_patch.upload(P.slot_patch, st);
  c->patch_slot.upload(P.patch_slot, st);
  c->slot_cluster.upload(P.slot_cluster, st);
  c->slice_count.upload(P.slice_count, st);
  c->d_cl_slot_start.upload(P.cl_slot_start, st);
  c->d_cl_count.upload(P.cl_count, st);
  c->d_nwords.upload(P.nwords, st);
  c->d_wb_first.upload(P.wb_first, st);
  c->d_vis_ptr.upload(P.vis_ptr, st);
  c->d_vis_cluster.upload(P.vis_cluster, st);
  c->d_vis_woff.upload(P.vis_woff, st);
  c->d_vis_seer.upload(P.vis_seer, st);
  c->d_coff.upload(P.coff.data(), P.coff.size(), st);
  c->d_sbase.upload(P.sbase.data(), P.sbase.size(), st);
  c->d_grp_first.upload(P.grp_first, st);
  c->d_grp_cluster.upload(P.grp_cluster, st);
  c->d_grp_slot0.upload(P.grp_slot0, st);
  c->d_grp_rows.upload(P.grp_rows, st);
  c->d_grp_owner.upload(P.grp_owner, st);
  c->d_grp_ownbase.upload(P.grp_ownbase, st);
  {
    // The runs of this rank, the expensive ones first: a run costs what its clear rays cost (they are walked), and those
    // are the pairs inside one room -- the chunks that hold the source cluster's own words.  Warps draw runs from a
    // counter, so starting with the long ones keeps the tail of the kernel short (a run is 128 rows x 1 024 receivers:
    // milliseconds; a shard has only ~20 of them per resident warp).
    // A run covers all rows of its group when there are plenty of runs per resident warp; a shard has few, and then the
    // runs that end last set the kernel time, so they are cut in halves / quarters (each part starts with cold guesses:
    // 32-row runs cost 5.6 % more work on one GPU, and save 20 % of the kernel time on an eighth of C5).
    int run_rows = kGroupRows;
    {
      const int64_t warps = (int64_t) c->sm_count * 4 * kWarpsPerBlock;
      while (run_rows > 32 && P.own_runbase.back() * (kGroupRows / run_rows) < 100 * warps) run_rows /= 2;
    }
    if (const char *e = getenv("QRAD_RUN_ROWS")) run_rows = std::max(32, atoi(e) & ~31);
    // The list has a million entries on one GPU: the host only says where each group's entries go, a kernel writes them
    std::vector<RunGroup> rg;
    rg.reserve(P.own_groups.size());
    int64_t nheavy = 0, nlight = 0;
    for (int g : P.own_groups) {
      const int cl = P.grp_cluster[g];
      RunGroup r{};
      r.grp = g;
      r.nchunks = (P.nwords[cl] + kChunkWords - 1) / kChunkWords;
      r.rows = P.grp_rows[g];
      r.self_lo = 0, r.self_hi = -1;   // chunks that hold words of the group's own cluster
      for (int v = P.vis_ptr[cl]; v < P.vis_ptr[cl + 1]; v++)
        if (P.vis_cluster[v] == cl) {
          const int w0 = P.vis_woff[v], w1 = w0 + ((P.cl_slot_start[cl + 1] - P.cl_slot_start[cl]) >> 5);
          r.self_lo = w0 / kChunkWords;
          r.self_hi = (w1 - 1) / kChunkWords;
        }
      const int parts = (r.rows + run_rows - 1) / run_rows, nself = r.self_hi - r.self_lo + 1;
      r.heavy_at = nheavy;
      r.light_at = nlight;
      nheavy += (int64_t) nself * parts;
      nlight += (int64_t) (r.nchunks - nself) * parts;
      rg.push_back(r);
    }
    for (RunGroup &r : rg) r.light_at += nheavy;
    c->num_runs = nheavy + nlight;
    c->d_run_list.alloc((size_t) c->num_runs + 1);
    if (!rg.empty()) {
      c->d_run_groups.upload(rg, st);
      k_run_list<<<(int) rg.size(), 128, 0, st>>>(c->d_run_groups.p, run_rows, c->d_run_list.p);
      c->launch_check("k_run_list");
    }
  }
  const int nvis = (int) P.vis_cluster.size();
  {
    std::vector<int32_t> vis_src((size_t) std::max(nvis, 1));
    for (int a = 0; a < nc; a++)
      for (int v = P.vis_ptr[a]; v < P.vis_ptr[a + 1]; v++) vis_src[v] = a;
    c->d_vis_src.upload(vis_src, st);
  }
  c->s_origin_area.alloc((size_t) M);
  c->s_normal.alloc((size_t) M);
  c->wbox_min.alloc((size_t) M / 32);
  c->wbox_max.alloc((size_t) M / 32);
  k_slot_geometry<<<(M + 255) / 256, 256, 0, st>>>(M, c->slot_patch.p, c->p_origin_area.p, c->p_normal.p, c->s_origin_area.p,
                                                   c->s_normal.p, c->wbox_min.p, c->wbox_max.p);
  c->launch_check("k_slot_geometry");
  const int64_t nwords_all = P.wb_first[nc];
  c->wslice.alloc((size_t) nwords_all + 1);
  c->wbase.alloc((size_t) nwords_all + 1);
  const size_t E = P.seer_cluster.size();
  if (nvis > 0) {
    k_word_tables<<<(nvis + kWarpsPerBlock - 1) / kWarpsPerBlock, kThreads, 0, st>>>(
        nvis, c->d_vis_src.p, c->d_vis_cluster.p, c->d_vis_woff.p, c->d_vis_seer.p, c->d_cl_slot_start.p, c->d_wb_first.p,
        c->d_sbase.p + (size_t) P.rank * (P.nslices + 1), c->d_coff.p + (size_t) P.rank * E, c->wslice.p, c->wbase.p);
    c->launch_check("k_word_tables");
  }
  const int64_t W = P.rowbuf_words();
  c->bits.alloc((size_t) W + 1);
  c->counters.alloc(24);   // [0..15] statistics of k_visibility, [16] its run counter
  c->counters.zero(st);
  c->sync("make_transfers set-up");
  c->stats.ms_tr_setup = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
  // ---- pass 1 ----
  QCUDA(cudaEventRecord(c->ev[0], st));
  const uint64_t runs = (uint64_t) c->num_runs;
  if (runs) {
    VisArgs a{};
    a.tree = c->tree();
    a.wbox_min = c->wbox_min.p;
    a.wbox_max = c->wbox_max.p;
    a.s_origin_area = c->s_origin_area.p;
    a.s_normal = c->s_normal.p;
    a.g = geo_of(c);
    a.run_list = c->d_run_list.p;
    a.grp_cluster = c->d_grp_cluster.p;
    a.grp_slot0 = c->d_grp_slot0.p;
    a.grp_rows = c->d_grp_rows.p;
    a.grp_ownbase = c->d_grp_ownbase.p;
    a.num_runs = runs;
    a.bits = c->bits.p;
    a.counters = c->counters.p;
    a.cell_lo = c->leaf_cell_lo.p;
    a.cell_hi = c->leaf_cell_hi.p;
    a.margin = proof_margin(std::max(c->bsp_extent, c->patch_extent));
    a.mode = 2 | 4 | 32;
    if (const char *e = getenv("QRAD_VIS_MODE")) a.mode = atoi(e);
    const uint64_t want_blocks = (runs + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const int blocks = (int) std::max<uint64_t>(1, std::min<uint64_t>(want_blocks, (uint64_t) c->sm_count * 8));
    const bool big = c->tree_depth > kTraceStack;
    if (c->count_nodes) {
      if (big) k_visibility<kTraceStackBig, true><<<blocks, kThreads, 0, st>>>(a);
      else k_visibility<kTraceStack, true><<<blocks, kThreads, 0, st>>>(a);
    } else {
      if (big) k_visibility<kTraceStackBig, false><<<blocks, kThreads, 0, st>>>(a);
      else k_visibility<kTraceStack, false><<<blocks, kThreads, 0, st>>>(a);
    }
    c->launch_check("k_visibility");
  }
  QCUDA(cudaEventRecord(c->ev[1], st));
  // several GPUs: while this rank goes on with its lists and pass 2, every peer receives the contiguous piece of the
  // row buffer that describes its receiver slices (second stream; pass 3 waits for it)
  c->stats.ms_x_matrix = 0;
  if (P.world > 1 && c->comm) {   // (no communicator: qrad_cuda_trace_shard times a hypothetical rank's pass 1 on its own)
    const NcclApi &n = nccl();
    int64_t recv_total = 0;
    c->piece_first.assign(P.world, 0);
    for (int q = 0; q < P.world; q++) {
      c->piece_first[q] = recv_total;
      if (q != P.rank) recv_total += P.words_for(q, P.rank);
    }
    c->colbuf.alloc((size_t) recv_total + 1);
    QCUDA(cudaStreamWaitEvent(c->xstream, c->ev[1], 0));
    QCUDA(cudaEventRecord(c->ev[2], c->xstream));
    nccl_ok(n.GroupStart(), "ncclGroupStart");
    for (int q = 0; q < P.world; q++) {
      if (q == P.rank) continue;
      const int64_t sfirst = P.sbase[(size_t) P.rank * (P.nslices + 1) + P.slice_lo[q]], scount = P.words_for(P.rank, q);
      if (scount) nccl_ok(n.Send(c->bits.p + sfirst, (size_t) scount, ncclUint32, q, c->comm, c->xstream), "ncclSend");
      const int64_t rcount = P.words_for(q, P.rank);
      if (rcount) nccl_ok(n.Recv(c->colbuf.p + c->piece_first[q], (size_t) rcount, ncclUint32, q, c->comm, c->xstream), "ncclRecv");
    }
    nccl_ok(n.GroupEnd(), "ncclGroupEnd");
    QCUDA(cudaEventRecord(c->ev[3], c->xstream));
  }
  if (!(P.world > 1 && !c->comm)) transfers_lists(c);   // (not for qrad_cuda_trace_shard)
  c->sync("make_transfers pass 1");
  c->stats.ms_transfers_trace = elapsed(c->ev[0], c->ev[1]);
}

// phase 1: ordered run lists, pass 2 (row totals of this rank's rows)
// The row and column lists of passes 2 and 3 follow from the plan alone, not from pass 1: they are prepared on a stream
// of their own while k_visibility runs (the host part overlaps it entirely, the kernels fill the SMs its last runs leave).
static void transfers_lists(qrad_cuda_ctx *c) {
  cudaStream_t st = c->lstream;
  const TransferPlan &P = c->plan;
  const int N = c->N, M = P.M, nc = P.nc;
  // ---- the patch order as runs of consecutive slots inside one slice ----
  std::vector<int2> runs;
  runs.reserve((size_t) M / 8);
  for (int i = 0; i < N; i++) {
    const int s = P.patch_slot[i];
    if (s < 0) continue;
    if (!runs.empty()) {
      int2 &r = runs.back();
      if (s == r.x + r.y && (s >> 5) == (r.x >> 5)) {   // the next slot of the same slice: the run goes on
        r.y++;
        continue;
      }
    }
    runs.push_back(make_int2(s, 1));
  }
  c->d_runs.upload(runs, st);
  std::vector<int32_t> row_clusters, col_clusters;
  for (int k = 0; k < nc; k++) {
    if (P.Lown_ptr[k + 1] > P.Lown_ptr[k]) row_clusters.push_back(k);
    if (P.Rown_ptr[k + 1] > P.Rown_ptr[k]) col_clusters.push_back(k);
  }
  RunListArgs la{};
  la.nruns = (int) runs.size();
  la.runs = c->d_runs.p;
  la.slot_cluster = c->slot_cluster.p;
  la.cl_slot_start = c->d_cl_slot_start.p;
  la.vis_ptr = c->d_vis_ptr.p;
  la.vis_cluster = c->d_vis_cluster.p;
  la.vis_woff = c->d_vis_woff.p;
  la.vis_seer = c->d_vis_seer.p;
  la.wb_first = c->d_wb_first.p;
  la.wbase = c->wbase.p;
  la.wslice = c->wslice.p;
  la.grp_first = c->d_grp_first.p;
  la.grp_owner = c->d_grp_owner.p;
  la.grp_ownbase = c->d_grp_ownbase.p;
  la.grp_slot0 = c->d_grp_slot0.p;
  la.coff = c->d_coff.p;
  la.E = (int64_t) P.seer_cluster.size();
  DBuf<int32_t> &d_rowcl = c->lists_i32[0], &d_colcl = c->lists_i32[1], &d_cnt_row = c->lists_i32[2], &d_cnt_col = c->lists_i32[3];
  d_rowcl.upload(row_clusters, st);
  d_colcl.upload(col_clusters, st);
  d_cnt_row.alloc((size_t) nc);
  d_cnt_col.alloc((size_t) nc);
  d_cnt_row.zero(st);
  d_cnt_col.zero(st);
  if (!row_clusters.empty()) {
    la.clusters = d_rowcl.p;
    la.count = d_cnt_row.p;
    k_run_lists<false, false><<<(int) row_clusters.size(), 1024, 0, st>>>(la);
    c->launch_check("k_run_lists(rows, count)");
  }
  if (!col_clusters.empty()) {
    la.clusters = d_colcl.p;
    la.count = d_cnt_col.p;
    k_run_lists<true, false><<<(int) col_clusters.size(), 1024, 0, st>>>(la);
    c->launch_check("k_run_lists(columns, count)");
  }
  std::vector<int32_t> cnt_row((size_t) nc), cnt_col((size_t) nc);
  d_cnt_row.download(cnt_row.data(), (size_t) nc, st);
  d_cnt_col.download(cnt_col.data(), (size_t) nc, st);
  std::vector<int64_t> vl_ptr((size_t) nc + 1, 0), vr_ptr((size_t) nc + 1, 0);
  for (int k = 0; k < nc; k++) {
    vl_ptr[k + 1] = vl_ptr[k] + cnt_row[k];
    vr_ptr[k + 1] = vr_ptr[k] + cnt_col[k];
  }
  c->d_vl_ptr.upload(vl_ptr, st);
  c->d_vr_ptr.upload(vr_ptr, st);
  c->vl.alloc((size_t) vl_ptr[nc] + 1);
  c->vr.alloc((size_t) vr_ptr[nc] + 1);
  la.vl = c->vl.p;
  la.vr = c->vr.p;
  if (!row_clusters.empty()) {
    la.clusters = d_rowcl.p;
    la.list_ptr = c->d_vl_ptr.p;
    k_run_lists<false, true><<<(int) row_clusters.size(), 1024, 0, st>>>(la);
    c->launch_check("k_run_lists(rows)");
  }
  if (!col_clusters.empty()) {
    la.clusters = d_colcl.p;
    la.list_ptr = c->d_vr_ptr.p;
    k_run_lists<true, true><<<(int) col_clusters.size(), 1024, 0, st>>>(la);
    c->launch_check("k_run_lists(columns)");
  }
  QCUDA(cudaEventRecord(c->ev_lists, st));
}

void transfers_rows(qrad_cuda_ctx *c) {
  cudaStream_t st = c->stream;
  auto wall0 = std::chrono::steady_clock::now();
  const TransferPlan &P = c->plan;
  const int M = P.M;
  QCUDA(cudaStreamWaitEvent(st, c->ev_lists, 0));
  c->sync("make_transfers lists");   // what is left of them once pass 1 has ended
  c->stats.ms_tr_lists = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
  wall0 = std::chrono::steady_clock::now();
  // ---- pass 2 ----
  c->row_total.alloc((size_t) M);
  c->row_count.alloc((size_t) M);
  c->row_total.zero(st);
  c->row_count.zero(st);
  const int nrs = (int) P.own_rowslice_slot0.size();
  if (nrs > 0) {
    // the slices with the longest row lists first: a warp is serial in its list, the last wave should hold the short ones
    std::vector<int32_t> ord((size_t) nrs), slot0((size_t) nrs), row0((size_t) nrs);
    for (int i = 0; i < nrs; i++) ord[i] = i;
    std::stable_sort(ord.begin(), ord.end(), [&](int32_t x, int32_t y) {
      const int cx = P.slot_cluster[P.own_rowslice_slot0[x]], cy = P.slot_cluster[P.own_rowslice_slot0[y]];
      return P.Lptr[cx + 1] - P.Lptr[cx] > P.Lptr[cy + 1] - P.Lptr[cy];
    });
    for (int i = 0; i < nrs; i++) {
      slot0[i] = P.own_rowslice_slot0[ord[i]];
      row0[i] = P.own_rowslice_row0[ord[i]];
    }
    c->d_rowslice_slot0.upload(slot0, st);
    c->d_rowslice_row0.upload(row0, st);
    RowArgs ra = row_args(c);
    ra.rowslice_slot0 = c->d_rowslice_slot0.p;
    ra.rowslice_row0 = c->d_rowslice_row0.p;
    ra.n_rowslices = nrs;
    // about one wave of 32-row warps or less (a shard): the longest row list would set the time of the launch; the
    // quarter-slice kernel does 1.7 times the work (every warp scans the whole row list) in a quarter of the time per slice
    bool quarters = (int64_t) nrs * 4 < (int64_t) 5 * c->sm_count * 32;
    if (const char *e = getenv("QRAD_ROWS_KERNEL")) quarters = atoi(e) == 8;   // tests run both
    if (quarters) k_row_totals8<<<(4 * nrs + kWarpsPerBlock - 1) / kWarpsPerBlock, kThreads, 0, st>>>(ra);
    else k_row_totals32<<<(nrs + kWarpsPerBlock - 1) / kWarpsPerBlock, kThreads, 0, st>>>(ra);
    c->launch_check("k_row_totals");
  }
  if (P.world > 1) {   // every row was summed by exactly one rank
    // one communicator: its collectives are kept in one order by making this stream wait for the matrix exchange
    QCUDA(cudaStreamWaitEvent(st, c->ev[3], 0));
    const NcclApi &n = nccl();
    nccl_ok(n.AllReduce(c->row_total.p, c->row_total.p, (size_t) M, ncclFloat, ncclSum, c->comm, st), "ncclAllReduce");
    nccl_ok(n.AllReduce(c->row_count.p, c->row_count.p, (size_t) M, ncclInt32, ncclSum, c->comm, st), "ncclAllReduce");
  }
  c->sync("make_transfers pass 2");
  c->stats.ms_tr_rows = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
}

// phase 2: pass 3 (gather lists of this rank's slices).  Needs the row totals of every source and, on several GPUs, the
// pieces of the other ranks' row buffers.
void transfers_columns(qrad_cuda_ctx *c) {
  cudaStream_t st = c->stream;
  const auto wall0 = std::chrono::steady_clock::now();
  const TransferPlan &P = c->plan;
  const int M = P.M, nslices = P.nslices;
  const int lo = P.slice_lo[P.rank], hi = P.slice_hi[P.rank];
  c->slice_len.alloc((size_t) nslices);
  c->slice_len.zero(st);
  ColArgs ca{};
  ca.s_origin_area = c->s_origin_area.p;
  ca.s_normal = c->s_normal.p;
  ca.g = geo_of(c);
  ca.vr_ptr = c->d_vr_ptr.p;
  ca.vr = c->vr.p;
  for (int q = 0; q < P.world; q++) {
    if (q == P.rank) ca.col[q] = c->bits.p;
    else ca.col[q] = c->colbuf.p + c->piece_first[q] - P.sbase[(size_t) q * (nslices + 1) + lo];
  }
  ca.sbase = c->d_sbase.p;
  ca.nslices = nslices;
  ca.row_total = c->row_total.p;
  ca.slice_lo = lo;
  ca.slice_hi = hi;
  ca.slice_len = c->slice_len.p;
  const int cblocks = (hi - lo + kWarpsPerBlock - 1) / kWarpsPerBlock;
  if (cblocks > 0) {
    k_column_counts<<<cblocks, kThreads, 0, st>>>(ca, P.world);
    c->launch_check("k_column_counts");
  }
  std::vector<int32_t> h_len((size_t) nslices);
  c->slice_len.download(h_len.data(), (size_t) nslices, st);
  std::vector<int64_t> h_slice_base((size_t) nslices + 1, 0);
  int64_t nzw = 0;
  for (int s = 0; s < nslices; s++) {
    nzw += h_len[s];
    h_slice_base[s + 1] = h_slice_base[s] + ((h_len[s] + 31) & ~31);
  }
  c->sell_entries = h_slice_base[nslices];
  if (getenv("QRAD_DEBUG_SLICES")) {   // list length statistics for tools/bounce_shard_probe.py
    int mx = 0, mx8 = 0;
    for (int s = 0; s < nslices; s++) {
      mx = std::max(mx, h_len[s]);
      if (s < nslices / 8) mx8 = std::max(mx8, h_len[s]);
    }
    fprintf(stderr, "slices %d, list entries: mean %.0f, longest %d, longest in the first eighth %d\n", nslices,
            nslices ? (double) nzw / nslices : 0.0, mx, mx8);
    for (int thr : {12000, 16000, 24000, 32000, 48000}) {
      int cnt = 0;
      int64_t ent = 0;
      for (int s = 0; s < nslices; s++)
        if (h_len[s] > thr) {
          cnt++;
          ent += h_len[s];
        }
      fprintf(stderr, "  lists longer than %d entries: %d slices, %.2f %% of all entries\n", thr, cnt, nzw ? 100.0 * ent / nzw : 0.0);
    }
  }
  c->slice_base.upload(h_slice_base, st);
  {  // k_bounce hands this rank's slices to its warps longest list first
    std::vector<int32_t> order;
    int hi_work = hi;
    if (const char *f = getenv("QRAD_DEBUG_BOUNCE_SHARD"))   // tools/bounce_shard_probe.py: the launches an n-GPU shard
      hi_work = lo + std::max(1, (int) ((hi - lo) * atof(f)));   // would see, timed on one GPU (results are not lighting)
    for (int s = lo; s < hi_work; s++) order.push_back(s);
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return h_len[x] > h_len[y]; });
    // A launch cannot be shorter than ONE warp needs for the longest list (about 0.5 us per 32 entries).  Lists that
    // would take a lone warp more than 0.7 of what the whole shard takes at full bandwidth go to block-cooperative
    // blocks instead (none on one GPU with C5; the longest 2 % of the slices on an eighth of it).
    int64_t own_entries = 0;
    for (int s = lo; s < hi_work; s++) own_entries += (h_len[s] + 31) & ~31;
    const double est_s = (double) own_entries * 68 / 6.0e12;
    double factor = 0.7;
    if (const char *e = getenv("QRAD_BOUNCE_COOP_FACTOR")) factor = atof(e);   // tools/bounce_shard_probe.py sweeps it
    int64_t threshold = std::max<int64_t>(4096, (int64_t) (factor * est_s / 0.5e-6 * 32));
    if (const char *e = getenv("QRAD_BOUNCE_COOP")) threshold = atoll(e);   // tests force it (0 = every slice)
    std::vector<int32_t> coop, regular;
    for (int32_t s : order) (h_len[s] > threshold ? coop : regular).push_back(s);
    {
      std::vector<int32_t> all;
      for (int s = lo; s < hi; s++) all.push_back(s);
      std::stable_sort(all.begin(), all.end(), [&](int32_t x, int32_t y) { return h_len[x] > h_len[y]; });
      if (all.empty()) all.push_back(0);
      c->col_order.upload(all, st);
    }
    c->n_coop = (int) coop.size();
    c->n_regular = (int) regular.size();
    if (coop.empty()) coop.push_back(0);
    if (regular.empty()) regular.push_back(0);
    c->slice_order.upload(regular, st);
    c->coop_slices.upload(coop, st);
  }
  // k_bounce prefetches without bounds checks: slots two batches, weights one batch past the end of a list
  c->g_src.alloc((size_t) c->sell_entries + 96);
  c->g_w.alloc((size_t) (c->sell_entries / 8 + 8) * 32);
  c->g_src.zero(st);
  c->g_w.zero(st);
  ca.slice_base = c->slice_base.p;
  ca.g_src = c->g_src.p;
  ca.g_w = c->g_w.p;
  ca.col_order = c->col_order.p;
  if (hi > lo) {
    // fewer than three waves of warps: the longest lists would set the time of a warp-per-slice launch
    bool split = (int64_t) (hi - lo) < (int64_t) 3 * c->sm_count * 32;
    if (const char *e = getenv("QRAD_COLUMNS_SPLIT")) split = atoi(e) != 0;   // tests run both
    if (split) k_columns<true><<<hi - lo, kThreads, 0, st>>>(ca, hi - lo);
    else k_columns<false><<<(hi - lo + kWarpsPerBlock - 1) / kWarpsPerBlock, kThreads, 0, st>>>(ca, hi - lo);
    c->launch_check("k_columns");
  }
  std::vector<int32_t> h_rc((size_t) M);
  c->row_count.download(h_rc.data(), (size_t) M, st);
  int64_t nnz = 0;
  for (int s = 0; s < M; s++) nnz += h_rc[s];
  c->sync("make_transfers");
  if (P.world > 1) c->stats.ms_x_matrix = elapsed(c->ev[2], c->ev[3]);

  unsigned long long hc[16];
  c->counters.download(hc, 16, st);
  if (c->count_nodes) {   // warp cycles per phase of k_visibility (tools/gpu_raycounts.py)
    fprintf(stderr, "k_visibility: %llu rays settled by per-ray cell proofs; blocked full walks: %llu grazed the guessed cell, "
            "%llu hit another leaf that a proof would have caught, %llu another leaf only grazed, %llu a leaf without a box cell; "
            "%llu fallback rays settled by a refreshed guess\n",
            hc[5], hc[6], hc[13], hc[14], hc[7], hc[15]);
    const double tot = (double) (hc[8] + hc[9] + hc[10] + hc[11] + hc[12]);
    if (tot > 0)
      fprintf(stderr, "k_visibility warp cycles: word proofs %.1f %%, sign tests %.1f %%, per-ray proofs %.1f %%, full walks %.1f %%, "
              "task set-up / stores %.1f %%\n", 100 * hc[8] / tot, 100 * hc[9] / tot, 100 * hc[10] / tot, 100 * hc[11] / tot, 100 * hc[12] / tot);
  }
  c->stats.rays_transfers = (int64_t) hc[0];
  c->stats.node_visits += (int64_t) hc[1];
  c->stats.candidate_pairs = (int64_t) hc[2];
  c->stats.full_walks = (int64_t) hc[3];
  c->stats.box_pruned_rays = (int64_t) hc[4];
  c->stats.nonzero_words = nzw;
  c->total_transfers = nnz;
  c->stats.total_transfers = nnz;
  c->stats.padded_entries = c->sell_entries;
  c->stats.transfer_bytes = c->sell_entries * 68;   // stored: u32 source + 32 x u16 weights per list entry
  c->transfers_done = true;
  c->stats.ms_tr_columns = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
  c->stats.ms_transfers = c->t_transfers.stop();
}

void make_transfers_impl(qrad_cuda_ctx *c, int nopvs) {
  transfers_trace(c, nopvs);
  transfers_rows(c);
  transfers_columns(c);
}

// reference row format for parity tests: the rows of selected source patches, receivers ascending, and the root
// TestLine_r calls of each row
void download_rows_impl(qrad_cuda_ctx *c, int nrows, const int32_t *rows, int64_t *row_ptr, uint32_t *patch,
                        uint16_t *transfer, int64_t capacity, int64_t *rays) {
  if (c->world > 1) dfail("download of transfer rows: not available on a context that is part of a communicator");
  if (nrows <= 0) {
    row_ptr[0] = 0;
    return;
  }
  const TransferPlan &P = c->plan;
  cudaStream_t st = c->stream;
  std::vector<int32_t> rc((size_t) P.M);
  c->row_count.download(rc.data(), rc.size(), st);
  std::vector<int32_t> sel((size_t) nrows), selrow((size_t) nrows, 0);
  row_ptr[0] = 0;
  for (int k = 0; k < nrows; k++) {
    if (rows[k] < 0 || rows[k] >= c->N) dfail("download_rows: patch %d out of range", rows[k]);
    sel[k] = P.patch_slot[rows[k]];
    if (sel[k] >= 0) {
      const int g = P.group_of_slot(sel[k]);
      selrow[k] = P.grp_ownbase[g] + (sel[k] - P.grp_slot0[g]);
    }
    row_ptr[k + 1] = row_ptr[k] + (sel[k] >= 0 ? rc[sel[k]] : 0);
  }
  const int64_t nnz = row_ptr[nrows];
  if (nnz > capacity) dfail("download_rows: %lld records do not fit the caller's %lld", (long long) nnz, (long long) capacity);
  DBuf<int64_t> d_row_ptr;
  d_row_ptr.upload(row_ptr, (size_t) nrows + 1, st);
  DBuf<int32_t> d_sel, d_selrow;
  d_sel.upload(sel, st);
  d_selrow.upload(selrow, st);
  DBuf<uint32_t> d_patch;
  DBuf<uint16_t> d_tr;
  DBuf<long long> d_rays;
  d_patch.alloc((size_t) nnz + 1);
  d_tr.alloc((size_t) nnz + 1);
  d_rays.alloc((size_t) nrows);
  RowArgs ra = row_args(c);
  ra.row_ptr = d_row_ptr.p;
  ra.out_patch = d_patch.p;
  ra.out_transfer = d_tr.p;
  ra.sel_slots = d_sel.p;
  ra.sel_rows = d_selrow.p;
  ra.nsel = nrows;
  ra.sel_rays = d_rays.p;
  const int blocks = (nrows + kWarpsPerBlock - 1) / kWarpsPerBlock;
  if (patch && transfer) {
    k_rows_emit<<<blocks, kThreads, 0, st>>>(ra);
    c->launch_check("k_rows_emit");
    d_patch.download(patch, (size_t) nnz, st);
    d_tr.download(transfer, (size_t) nnz, st);
  }
  if (rays) {
    k_row_rays<<<blocks, kThreads, 0, st>>>(ra);
    c->launch_check("k_row_rays");
    std::vector<long long> hr((size_t) nrows);
    d_rays.download(hr.data(), hr.size(), st);
    for (int k = 0; k < nrows; k++) rays[k] = hr[k];
  }
}

void download_transfers_impl(qrad_cuda_ctx *c, int64_t *row_ptr, uint32_t *patch, uint16_t *transfer) {
  std::vector<int32_t> all((size_t) c->N);
  for (int i = 0; i < c->N; i++) all[i] = i;
  download_rows_impl(c, c->N, all.data(), row_ptr, patch, transfer, INT64_MAX, nullptr);
}

// BounceLight, qrad3.c:448-473: begin (radiosity = samplelight * reflectivity * area), step (one ShootLight +
// CollectLight over this rank's slices; on several GPUs the owners then broadcast their part of the new radiosity), end
// (totallight back to patch order).
void bounce_begin(qrad_cuda_ctx *c, bool want_sum) {
  if (!c->transfers_done) dfail("bounce: make_transfers has not run");
  c->t_bounce.start(c->stream);
  cudaStream_t st = c->stream;
  const int M = c->M;
  c->rad_a.alloc((size_t) M);
  c->rad_b.alloc((size_t) M);
  c->s_total.alloc((size_t) M);
  c->s_refl_area.alloc((size_t) M);
  c->rad_a.zero(st);
  c->rad_b.zero(st);
  c->s_total.zero(st);
  k_init_bounce<<<(M + 255) / 256, 256, 0, st>>>(M, c->slot_patch.p, c->p_samplelight.p, c->p_totallight.p, c->p_refl.p,
                                                 c->p_origin_area.p, c->rad_a.p, c->s_total.p, c->s_refl_area.p);
  c->launch_check("k_init_bounce");
  if (want_sum) c->rad_sum.alloc((size_t) M);
  else c->rad_sum.release();
  c->rad_cur = c->rad_a.p;
  c->rad_nxt = c->rad_b.p;
  c->bounce_steps = 0;
  c->bounce_next.alloc(4);   // work counters of the warp-per-slice [0..1] and block-cooperative [2..3] roles, two parities each
  c->bounce_next.zero(st);
  c->negative_light.alloc(1);
  c->negative_light.zero(st);
  if (c->bounce_blocks_per_sm == 0) {
    int per_sm = 0;
    QCUDA(cudaFuncSetAttribute(k_bounce, cudaFuncAttributeMaxDynamicSharedMemorySize, kWarpsPerBlock * kBounceWarpBytes));
    QCUDA(cudaFuncSetAttribute(k_bounce_shard, cudaFuncAttributeMaxDynamicSharedMemorySize, kShardSmemBytes));
    QCUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bounce_shard, kThreads, kShardSmemBytes));
    c->shard_blocks_per_sm = std::max(1, per_sm);
    QCUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bounce, kThreads, kWarpsPerBlock * kBounceWarpBytes));
    c->bounce_blocks_per_sm = std::max(1, per_sm);
    QCUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bounce_regs, kThreads, 0));
    c->bounce_regs_blocks_per_sm = std::max(1, per_sm);
  }
}

void bounce_step(qrad_cuda_ctx *c) {
  cudaStream_t st = c->stream;
  const TransferPlan &P = c->plan;
  BounceArgs a{};
  a.slice_len = c->slice_len.p;
  a.slice_base = c->slice_base.p;
  a.g_src = c->g_src.p;
  a.g_w = c->g_w.p;
  a.total = c->s_total.p;
  a.refl_area = c->s_refl_area.p;
  a.rad_sum = c->rad_sum.p;
  a.M = c->M;
  a.slice_lo = P.slice_lo[P.rank];
  a.slice_hi = P.slice_hi[P.rank];
  a.send = c->rad_cur;
  a.send_next = c->rad_nxt;
  a.order = c->slice_order.p;
  a.next = c->bounce_next.p;
  a.parity = c->bounce_steps & 1;
  a.slice_cut = INT_MAX;
  const int nsl = c->n_regular + c->n_coop;
  a.slice_hi = a.slice_lo + c->n_regular;   // the work list of the warp-per-slice role holds the slices without a very long list
  // persistent warps: as many blocks as fit on the chip, slices are drawn from the work counter
  const int want = std::max(1, (a.slice_hi - a.slice_lo + kWarpsPerBlock - 1) / kWarpsPerBlock);
  // CUDA events around every launch, read once at the end: no host synchronisation between bounces
  while ((int) c->bounce_ev.size() < 2 * (c->bounce_steps + 1)) {
    cudaEvent_t e;
    QCUDA(cudaEventCreate(&e));
    c->bounce_ev.push_back(e);
  }
  QCUDA(cudaEventRecord(c->bounce_ev[2 * c->bounce_steps], st));
  // few slices per resident warp: the launch lasts as long as ONE warp needs for the longest list -> bulk-copy ring
  bool ring = (int64_t) nsl < (int64_t) 4 * c->sm_count * 40;
  if (const char *force = getenv("QRAD_BOUNCE_KERNEL")) ring = strcmp(force, "regs") != 0;   // tests cover both
  if (c->n_coop > 0) {
    // a shard with very long lists: they go to block-cooperative blocks of the same launch
    const int coop_blocks = std::min(c->n_coop, c->sm_count);
    const int ring_blocks = c->n_regular > 0 ? std::max(0, std::min(want, c->sm_count * c->shard_blocks_per_sm - coop_blocks)) : 0;
    k_bounce_shard<<<coop_blocks + ring_blocks, kThreads, kShardSmemBytes, st>>>(a, c->coop_slices.p, c->n_coop,
                                                                               c->bounce_next.p + 2, coop_blocks);
    c->launch_check("k_bounce_shard");
  } else if (c->n_regular > 0) {
    if (ring) k_bounce<<<std::min(want, c->sm_count * c->bounce_blocks_per_sm), kThreads, kWarpsPerBlock * kBounceWarpBytes, st>>>(a);
    else k_bounce_regs<<<std::min(want, c->sm_count * c->bounce_regs_blocks_per_sm), kThreads, 0, st>>>(a);
    c->launch_check("k_bounce");
  }
  QCUDA(cudaEventRecord(c->bounce_ev[2 * c->bounce_steps + 1], st));
  c->bounce_steps++;
  std::swap(c->rad_cur, c->rad_nxt);   // rad_cur now holds the new radiosity (this rank's slots)
  if (P.world > 1) {
    share_slices(c, c->rad_cur, sizeof(float4), st);
    if (c->rad_sum.p) share_slices(c, c->rad_sum.p, sizeof(float), st);
  }
}

void bounce_end(qrad_cuda_ctx *c) {
  cudaStream_t st = c->stream;
  if (c->plan.world > 1) share_slices(c, c->s_total.p, sizeof(float4), st);
  k_finish_bounce<<<(c->N + 255) / 256, 256, 0, st>>>(c->N, c->patch_slot.p, c->s_total.p, c->rad_cur, c->p_totallight.p,
                                                      c->d_radiosity.p, c->negative_light.p);
  c->launch_check("k_finish_bounce");
  c->sync("bounce");
  float kernel_ms = 0;
  for (int b = 0; b < c->bounce_steps; b++) kernel_ms += elapsed(c->bounce_ev[2 * b], c->bounce_ev[2 * b + 1]);
  c->stats.ms_bounce_kernel = c->bounce_steps > 0 ? kernel_ms / c->bounce_steps : 0;
  if (getenv("QRAD_DEBUG_SLICES") && c->bounce_steps > 0)
    fprintf(stderr, "bounce: %d slices warp-per-slice, %d slices block-cooperative, last launch %.3f ms\n", c->n_regular, c->n_coop,
            elapsed(c->bounce_ev[2 * (c->bounce_steps - 1)], c->bounce_ev[2 * (c->bounce_steps - 1) + 1]));
  c->stats.ms_bounce = c->t_bounce.stop();
  c->stats.ms_x_radiosity = c->plan.world > 1 ? std::max(0.f, c->stats.ms_bounce - kernel_ms) : 0;
}

void bounce_impl(qrad_cuda_ctx *c, int numbounce, float *added_out) {
  bounce_begin(c, added_out != nullptr);
  DBuf<float> d_added;
  if (added_out) d_added.alloc((size_t) std::max(numbounce, 1));
  for (int b = 0; b < numbounce; b++) {
    bounce_step(c);
    if (added_out) {
      k_ordered_sum<<<1, 32, 0, c->stream>>>(c->rad_sum.p, c->patch_slot.p, c->N, d_added.p + b);
      c->launch_check("k_ordered_sum");
    }
  }
  if (added_out && numbounce > 0) d_added.download(added_out, (size_t) numbounce, c->stream);
  bounce_end(c);
  // CheckPatches, qrad3.c:476-486
  if (c->negative_light.p) {
    int neg = 0;
    c->negative_light.download(&neg, 1, c->stream);
    if (neg) dfail("negative patch totallight\n");
  }
}

}  // namespace qrad
