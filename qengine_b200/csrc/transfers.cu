// MakeTransfers (qrad3.c:185-293) and BounceLight (qrad3.c:383-473) on the GPU.
//
// Layout ("slot order"): patches are grouped by the PVS cluster of their origin, every
// cluster padded to a multiple of 32 slots, so one 32-bit word of the visibility matrix
// = one source patch x 32 consecutive receivers of one cluster.
//
//   pass 1  k_visibility   one warp per task (1 source x <= 1024 receivers, PVS pruning is structural):
//                          whole-word interval proofs, form factor sign tests, blocker-guess walks, full
//                          TestLine for the rest; writes the accept bit  trans > 0  (qrad3.c:216-257).
//   pass 2  k_row_totals32 one warp per 32 source rows of a cluster, lane = row: walks the shared candidate list in
//                          ascending receiver index (the reference's j order); every lane forms the SERIAL float
//                          sum `total` (qrad3.c:254) and numtransfers of its own row.
//   pass 3  k_column_counts / k_columns
//                          one warp per 32 receivers (slice): walks the potential sources in ascending
//                          patch index, reads ONE word per source, and emits the quantised weights
//                          (int)(trans * 0x10000 / total)  (qrad3.c:283) as "dense slice" lists -- the
//                          transposed (gather) form ShootLight's scatter needs (finding F7), already in
//                          the source order the reference adds them.
//   bounce  k_bounce       persistent warps, one slice at a time, lane = receiver, 128-bit loads of 8 weights per
//                          lane, radiosity vectors broadcast from shared memory, packed multiplies, serial
//                          accumulation in source order, CollectLight fused.
//
// Multi-GPU: every phase takes this rank's share (task granules / slices); the host exchanges in between
// (qrad_cuda_make_transfers_phase, qrad_cuda_bounce_phase, qrad_cuda_exchange_buffer).
//
// Everything is float arithmetic in the reference's operation order; this TU is compiled
// with -fmad=false.
#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdlib>
#include <cstring>

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

constexpr int kChunkWords = 32;   // one task = one source patch x up to 32 words (1024 receivers)
constexpr int kTaskRun = 128;       // consecutive tasks (same receivers, neighbouring sources) handled by one warp
constexpr int kTaskGranule = 512;   // multi-GPU: granule g of tasks belongs to rank g % world

struct VisArgs {
  Tree tree;
  const int4 *leaf_path;          // root -> leaf ancestor records (reaches_leaf)
  const int32_t *leaf_path_first;
  const float4 *wbox_min, *wbox_max;  // [M/32] bounding box of the patch origins of each 32-slot word
  const float4 *s_origin_area, *s_normal;
  const int64_t *bits_base;      // [nc+1] word offset of the first row of each source cluster
  const int64_t *task_base;      // [nc+1] task offset of each source cluster (rows x chunks per row)
  const int32_t *nwords;         // [nc]   words per row
  const int32_t *cl_slot_start;  // [nc]
  const int32_t *cl_count;       // [nc]
  const int32_t *vis_ptr;        // [nc+1] CSR of visible receiver clusters per source cluster
  const int32_t *vis_cluster;    // receiver cluster
  const int32_t *vis_woff;       // word offset of that cluster inside a row
  int nc;
  uint64_t num_tasks;             // all tasks
  uint64_t my_tasks;              // upper bound of this rank's local task index
  int rank, world;                // tasks are dealt to ranks in granules of kTaskGranule
  uint32_t *bits;
  unsigned long long *counters;   // [0] rays, [1] node visits, [2] candidate pairs
};

// One warp lights one task = one source patch x one chunk of <= 32 words (1024 receivers).  Tasks are numbered
// chunk-major inside a source cluster, so the consecutive tasks a warp takes are the SAME receivers seen from
// neighbouring source patches; what blocked a receiver last time is an excellent guess for this time
// (measured with the CPU oracle on C5: 95 % of the rays are blocked, 94-99 % of those by the leaf that blocked the
// same receiver from the previous source, and a single leaf blocks a whole word for about half of the words).
//   (1) lane w looks up word w of the chunk and runs the interval walk box_reaches_leaf() for the bounding box of its
//       32 receivers with the word's last blocker: a proven word needs no ray at all;
//   (2) form-factor sign tests for all receivers (they define which pairs the reference would hand to TestLine_r,
//       qrad3.c:244, and are counted as rays either way); pairs of unproven words are compacted into a queue;
//   (3) the queue is consumed 32 rays at a time: reaches_leaf() with the receiver's last blocker (one root -> leaf
//       path, no stack); rays that are not settled collect in a second queue and take the full TestLine walk in full
//       batches, which refreshes the guesses;
//   (4) accept bits are OR-ed into 32 result words in shared memory and stored with one coalesced write.
// Measured alternatives (profiles/README.md): one word per warp without compaction leaves 46 % of the lanes idle;
// refilling finished lanes from the queue decorrelates the lanes and is slower.
template <int STACK, bool COUNT>
__global__ void __launch_bounds__(kThreads, 4) k_visibility(const VisArgs a) {
  __shared__ uint16_t s_queue[kWarpsPerBlock][kChunkWords * 32];
  __shared__ uint16_t s_rguess[kWarpsPerBlock][kChunkWords * 32];  // per receiver of the chunk: last blocker leaf
  __shared__ uint16_t s_wguess[kWarpsPerBlock][kChunkWords];       // per word: last blocker leaf seen in it
  __shared__ uint32_t s_res[kWarpsPerBlock][kChunkWords];
  __shared__ int32_t s_wbase[kWarpsPerBlock][kChunkWords];         // first receiver slot of the word
  __shared__ int32_t s_wcnt[kWarpsPerBlock][kChunkWords];          // valid receivers in the word
  __shared__ uint16_t s_fb[kWarpsPerBlock][64];                    // rays whose guess failed (receiver index)
  // what the word's box walk learned when it did not prove the word (see box_reaches_leaf): the leaf it was run for,
  // the first ancestor not passed by all rays (0x80 | .. = all rays leave the path there) and the splits before it
  __shared__ uint16_t s_bleaf[kWarpsPerBlock][kChunkWords];
  __shared__ uint8_t s_bfirst[kWarpsPerBlock][kChunkWords];
  __shared__ unsigned long long s_bsplits[kWarpsPerBlock][kChunkWords];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1;
  const uint64_t warp0 = (uint64_t) blockIdx.x * kWarpsPerBlock + wid;
  const uint64_t nwarps = (uint64_t) gridDim.x * kWarpsPerBlock;
  unsigned long long rays = 0, pairs = 0, visits_total = 0, fallbacks = 0, boxed = 0;
  int guess_c = -1, guess_w0 = -1;   // which (source cluster, chunk) the guesses belong to
  long long tk0 = 0, tk_box = 0, tk_pairs = 0, tk_guess = 0, tk_full = 0, tk_misc = 0;   // COUNT: warp cycles per phase
  for (uint64_t lrun = warp0 * kTaskRun; lrun < a.my_tasks; lrun += nwarps * kTaskRun)
  for (uint64_t ltask = lrun; ltask < lrun + kTaskRun && ltask < a.my_tasks; ltask++) {
    const uint64_t task = ((ltask / kTaskGranule) * a.world + a.rank) * kTaskGranule + ltask % kTaskGranule;
    if (task >= a.num_tasks) continue;
    // ---- decode: source cluster, chunk, row (chunk-major) ----
    int lo = 0, hi = a.nc;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if ((uint64_t) a.task_base[mid] <= task) lo = mid; else hi = mid;
    }
    const int c = lo;
    const int nw = a.nwords[c];
    const int rows = a.cl_count[c];
    const uint64_t rel = task - (uint64_t) a.task_base[c];
    const int chunk = (int) (rel / (uint64_t) rows);
    const int row = (int) (rel - (uint64_t) chunk * rows);
    const int w0 = chunk * kChunkWords;
    const int nwc = min(kChunkWords, nw - w0);
    const int sslot = a.cl_slot_start[c] + row;
    const float4 so = __ldg(a.s_origin_area + sslot);
    const float4 sn = __ldg(a.s_normal + sslot);
    if (c != guess_c || w0 != guess_w0) {   // new receivers: forget the guesses
      guess_c = c;
      guess_w0 = w0;
      for (int k = lane; k < kChunkWords * 32; k += 32) s_rguess[wid][k] = 0xffffu;
      s_wguess[wid][lane] = 0xffffu;
    }
    s_res[wid][lane] = 0;
    if (COUNT) { const long long t = clock64(); if (tk0) tk_misc += t - tk0; tk0 = t; }
    // ---- (1) per word: receiver slots, box test ----
    bool proven = false;
    if (lane < nwc) {
      const int w = w0 + lane;
      int v = a.vis_ptr[c], vhi = a.vis_ptr[c + 1];
      while (vhi - v > 1) {  // last entry of c's visible list with woff <= w
        const int mid = (v + vhi) >> 1;
        if (a.vis_woff[mid] <= w) v = mid; else vhi = mid;
      }
      const int d = a.vis_cluster[v];
      const int off = (w - a.vis_woff[v]) * 32;
      const int base = a.cl_slot_start[d] + off;
      s_wbase[wid][lane] = base;
      s_wcnt[wid][lane] = min(32, a.cl_count[d] - off);
      const int g = s_wguess[wid][lane];
      s_bleaf[wid][lane] = 0xffffu;
      if (g != 0xffff) {
        const int p0 = __ldg(a.leaf_path_first + g), p1 = __ldg(a.leaf_path_first + g + 1);
        int first;
        unsigned long long splits;
        const int r = box_reaches_leaf(a.leaf_path + p0, p1 - p0, so, __ldg(a.wbox_min + (base >> 5)),
                                       __ldg(a.wbox_max + (base >> 5)), first, splits);
        proven = r == kBoxProven;
        if (!proven) {
          s_bleaf[wid][lane] = (uint16_t) g;
          s_bfirst[wid][lane] = (uint8_t) (first | (r == kBoxElsewhere ? 0x80 : 0));
          s_bsplits[wid][lane] = splits;
        }
      }
    }
    const unsigned proven_mask = __ballot_sync(0xffffffffu, proven);
    __syncwarp();
    if (COUNT) { const long long t = clock64(); tk_box += t - tk0; tk0 = t; }
    // ---- (2) cheap tests, compaction ----
    int n = 0;
    // receivers stream through (32 KB per task and warp): they bypass L1, which holds the tree and the paths, and the
    // next word's 32 receivers are requested before this word's are tested (an L2 round trip per word otherwise)
    float4 po_n = make_float4(0, 0, 0, 0), pn_n = po_n;
    if (nwc > 0) {
      po_n = __ldcg(a.s_origin_area + s_wbase[wid][0] + lane);
      pn_n = __ldcg(a.s_normal + s_wbase[wid][0] + lane);
    }
    for (int wi = 0; wi < nwc; wi++) {
      const int rslot = s_wbase[wid][wi] + lane;
      const float4 po = po_n, pn = pn_n;
      if (wi + 1 < nwc) {   // slots are padded to whole words: every lane has a valid address
        po_n = __ldcg(a.s_origin_area + s_wbase[wid][wi + 1] + lane);
        pn_n = __ldcg(a.s_normal + s_wbase[wid][wi + 1] + lane);
      }
      bool want = false;
      if (lane < s_wcnt[wid][wi] && rslot != sslot) {
        float dist;
        bool wants_ray;
        const float trans = form_factor_pre(so, sn, po, pn, dist, wants_ray);
        pairs++;
        if (wants_ray) {
          rays++;
          want = trans > 0 && !((proven_mask >> wi) & 1u);
          if (COUNT && trans > 0 && !want) boxed++;
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, want);
      if (want) s_queue[wid][n + __popc(m & lt)] = (uint16_t) (wi * 32 + lane);
      n += __popc(m);
    }
    __syncwarp();
    if (COUNT) { const long long t = clock64(); tk_pairs += t - tk0; tk0 = t; }
    // ---- (3) guess walks, full walks for the rest ----
    int fbn = 0;
    auto full_batch = [&](const int cnt) {  // full TestLine for the first `cnt` entries of the fallback queue
      long long tf0 = 0;
      if (COUNT) tf0 = clock64();
      if (lane < cnt) {
        const int cand = s_fb[wid][lane];
        const float4 po = __ldcg(a.s_origin_area + s_wbase[wid][cand >> 5] + (cand & 31));
        unsigned visits = 0;
        int leaf = -1;
        const int blocked = test_line<STACK, COUNT>(a.tree, so.x, so.y, so.z, po.x, po.y, po.z, visits, &leaf);
        if (COUNT) visits_total += visits;
        if (!blocked) atomicOr(&s_res[wid][cand >> 5], 1u << (cand & 31));
        else if (leaf < 0xffff) {
          s_rguess[wid][cand] = (uint16_t) leaf;
          s_wguess[wid][cand >> 5] = (uint16_t) leaf;
        }
      }
      __syncwarp();
      if (COUNT) tk_full += clock64() - tf0;
    };
    for (int base = 0; base < n; base += 32) {
      const int k = base + lane;
      bool miss = false;
      int cand = 0;
      if (k < n) {
        cand = s_queue[wid][k];
        const int g = s_rguess[wid][cand];
        miss = true;
        int first = 0;
        unsigned long long splits = 0;
        if (g != 0xffff && g == s_bleaf[wid][cand >> 5]) {   // the word's box walk was about this very leaf
          first = s_bfirst[wid][cand >> 5];
          splits = s_bsplits[wid][cand >> 5];
        }
        if (g != 0xffff && !(first & 0x80)) {   // 0x80: no ray into this word reaches g; straight to the full walk
          const float4 po = __ldcg(a.s_origin_area + s_wbase[wid][cand >> 5] + (cand & 31));
          const int p0 = __ldg(a.leaf_path_first + g), p1 = __ldg(a.leaf_path_first + g + 1);
          unsigned visits = 0;
          miss = !reaches_leaf<COUNT>(a.tree, a.leaf_path + p0, p1 - p0, so.x, so.y, so.z, po.x, po.y, po.z, visits,
                                      first, splits);
          if (COUNT) visits_total += visits;
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, miss);
      if (COUNT && miss) fallbacks++;
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
    if (fbn > 0) full_batch(fbn);
    __syncwarp();
    if (COUNT) { const long long t = clock64(); tk_guess += t - tk0; tk0 = t; }
    if (lane < nwc) a.bits[a.bits_base[c] + (int64_t) row * nw + w0 + lane] = s_res[wid][lane];
    __syncwarp();
  }
  for (int o = 16; o; o >>= 1) {
    rays += __shfl_xor_sync(0xffffffffu, rays, o);
    pairs += __shfl_xor_sync(0xffffffffu, pairs, o);
    if (COUNT) visits_total += __shfl_xor_sync(0xffffffffu, visits_total, o);
    if (COUNT) fallbacks += __shfl_xor_sync(0xffffffffu, fallbacks, o);
    if (COUNT) boxed += __shfl_xor_sync(0xffffffffu, boxed, o);
  }
  if (lane == 0) {
    atomicAdd(a.counters + 0, rays);
    atomicAdd(a.counters + 2, pairs);
    if (COUNT) atomicAdd(a.counters + 1, visits_total);
    if (COUNT) atomicAdd(a.counters + 3, fallbacks);
    if (COUNT) atomicAdd(a.counters + 4, boxed);
    if (COUNT) {
      atomicAdd(a.counters + 8, (unsigned long long) tk_box);
      atomicAdd(a.counters + 9, (unsigned long long) tk_pairs);
      atomicAdd(a.counters + 10, (unsigned long long) (tk_guess - tk_full));
      atomicAdd(a.counters + 11, (unsigned long long) tk_full);
      atomicAdd(a.counters + 12, (unsigned long long) tk_misc);
    }
  }
}

// Ordered list builder: for each cluster c (one block), all slots j (walking PATCH index ascending)
// whose cluster is related to c by table[c][cluster(j)] >= 0 (row lists) or table[cluster(j)][c] >= 0
// (column lists).  out_a = slot of j, out_b = bit position / word address material.
struct ListArgs {
  int N, nc;
  const int32_t *patch_slot;     // [N]
  const int32_t *slot_cluster;   // [M]
  const int32_t *woff_table;     // [nc][nc] word offset of receiver cluster d in a row of source cluster c, -1 = not visible
  const int32_t *cl_slot_start;
  const int64_t *list_ptr;       // [nc+1]
  uint32_t *out_slot;
  uint32_t *out_aux;
  int c_first;                   // first cluster handled by this launch (blockIdx.x is relative to it)
  int transpose;                 // 0: row lists (aux = bit position in the row); 1: column lists (addr = word address)
  const int64_t *bits_base;      // column lists: geometry of the bit matrix
  const int32_t *nwords;
  int64_t *out_addr;             // column lists: word of (source row, first slice of the receiver cluster)
};

__global__ void __launch_bounds__(1024) k_build_lists(const ListArgs a) {
  const int c = a.c_first + blockIdx.x;
  __shared__ int warp_sums[32];
  __shared__ int64_t base_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = a.list_ptr[c];
  __syncthreads();
  for (int j0 = 0; j0 < a.N; j0 += blockDim.x) {
    const int j = j0 + threadIdx.x;
    bool flag = false;
    int slot = -1, woff = -1, d = -1;
    if (j < a.N) {
      slot = a.patch_slot[j];
      if (slot >= 0) {
        d = a.slot_cluster[slot];
        woff = a.transpose ? a.woff_table[(size_t) d * a.nc + c] : a.woff_table[(size_t) c * a.nc + d];
        flag = woff >= 0;
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) warp_sums[wid] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
    for (int k = 0; k < (int) (blockDim.x >> 5); k++) {
      const int v = warp_sums[k];
      if (k < wid) before += v;
      total += v;
    }
    if (flag) {
      const int64_t o = base_s + before + __popc(m & ((1u << lane) - 1));
      a.out_slot[o] = (uint32_t) slot;
      if (!a.transpose) a.out_aux[o] = (uint32_t) woff * 32u + (uint32_t) (slot - a.cl_slot_start[d]);
      else a.out_addr[o] = a.bits_base[d] + (int64_t) (slot - a.cl_slot_start[d]) * a.nwords[d] + woff;
    }
    __syncthreads();
    if (threadIdx.x == 0) base_s += total;
    __syncthreads();
  }
}

// Summaries of the row lists: which row words a 32-entry chunk can touch, as up to four {word, mask} pairs in two
// uint4 (entries that are neighbours in patch order are neighbours in slot order as long as they lie in one cluster, so
// a chunk usually lies inside one or two words, three or four where the list moves on to another cluster).
// out[2 * chunk].x == 0xffffffff marks a chunk with more than four words; an unused pair has word 0xffffffff.
__global__ void __launch_bounds__(kThreads) k_list_summaries(int nc, const int64_t *__restrict__ Lptr,
                                                             const int64_t *__restrict__ chunk_base,
                                                             const uint32_t *__restrict__ L_pos, uint4 *__restrict__ out,
                                                             const int64_t chunk_lo, const int64_t chunk_hi) {
  const int lane = threadIdx.x & 31;
  const int64_t chunk = chunk_lo + (int64_t) blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (chunk >= chunk_hi) return;
  int lo = 0, hi = nc;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (chunk_base[mid] <= chunk) lo = mid; else hi = mid;
  }
  const int64_t k = Lptr[lo] + ((chunk - chunk_base[lo]) << 5) + lane;
  const bool have = k < Lptr[lo + 1];
  const uint32_t pos = have ? L_pos[k] : 0;
  const uint32_t word = pos >> 5, bit = have ? 1u << (pos & 31) : 0u;
  uint32_t w[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu}, m[4] = {0, 0, 0, 0};
  bool left = have;   // lane 0 always has an entry
#pragma unroll
  for (int t = 0; t < 4; t++) {
    const unsigned rest = __ballot_sync(0xffffffffu, left);
    if (rest) {
      w[t] = __shfl_sync(0xffffffffu, word, __ffs(rest) - 1);
      m[t] = __reduce_or_sync(0xffffffffu, (left && word == w[t]) ? bit : 0u);
      left = left && word != w[t];
    }
  }
  const bool complex_chunk = __ballot_sync(0xffffffffu, left) != 0;
  if (lane == 0) {
    out[2 * chunk] = complex_chunk ? make_uint4(0xffffffffu, 0, 0, 0) : make_uint4(w[0], m[0], w[1], m[1]);
    out[2 * chunk + 1] = make_uint4(w[2], m[2], w[3], m[3]);
  }
}

// pass 2: serial row sums in ascending receiver order.
struct RowArgs {
  const float4 *s_origin_area, *s_normal;
  const int32_t *slot_patch, *slot_cluster;
  const int64_t *bits_base, *Lptr;
  const int32_t *nwords, *cl_slot_start;
  const uint32_t *L_slot, *L_pos;
  const uint4 *L_sum;            // per 32-entry chunk of a list: {word0, mask0, word1, mask1} it can touch
  const int64_t *chunk_base;     // [nc+1] first chunk of each cluster's list
  const uint32_t *bits;
  int M;
  int slot_lo, slot_hi;          // this rank's rows
  float *row_total;
  int32_t *row_count;
  // download mode (reference row format): when out_patch != nullptr, rows are written at row_ptr[patch]
  const int64_t *row_ptr;
  uint32_t *out_patch;
  uint16_t *out_transfer;
  // download of selected rows: warp k handles source slot sel_slots[k] and writes at row_ptr[k]
  const int32_t *sel_slots;
  int nsel;
  long long *sel_rays;             // [nsel] root TestLine_r calls of the row (k_row_rays)
};

template <bool EMIT>
__global__ void __launch_bounds__(kThreads) k_row_totals(const RowArgs a) {
  const int lane = threadIdx.x & 31;
  const int widx = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  int sslot = a.slot_lo + widx;
  if (a.sel_slots) {
    if (widx >= a.nsel) return;
    sslot = a.sel_slots[widx];
    if (sslot < 0) return;
  } else if (sslot >= a.slot_hi)
    return;
  const int spatch = a.slot_patch[sslot];
  if (spatch < 0) return;
  const int c = a.slot_cluster[sslot];
  const uint32_t *row = a.bits + a.bits_base[c] + (int64_t) (sslot - a.cl_slot_start[c]) * a.nwords[c];
  const float4 so = a.s_origin_area[sslot];
  const float4 sn = a.s_normal[sslot];
  const int64_t l0 = a.Lptr[c], l1 = a.Lptr[c + 1];
  float total = 0.f;
  int count = 0;
  float rtotal = 0.f;
  int64_t outp = 0;
  if (EMIT) {
    rtotal = a.row_total[sslot];
    outp = a.sel_slots ? a.row_ptr[widx] : a.row_ptr[spatch];
  }
  // The list has ~40 candidates per accepted receiver.  Each 32-entry chunk has a precomputed summary of the (at
  // most two) row words it can touch, so 32 chunks are ruled out per step with one masked test per lane and only
  // chunks holding an accept bit take the ordered walk below (the order of the float sum is untouched).
  const int64_t nchunks = (l1 - l0 + 31) >> 5;
  const uint4 *sum = a.L_sum + 2 * a.chunk_base[c];
  for (int64_t g0 = 0; g0 < nchunks; g0 += 32) {
    bool live = false;
    if (g0 + lane < nchunks) {
      const uint4 sm = __ldg(sum + 2 * (g0 + lane));
      if (sm.x == 0xffffffffu) live = true;  // touches more than four words: always walk it
      else {
        live = (row[sm.x] & sm.y) != 0;
        if (!live && sm.z != 0xffffffffu) {
          live = (row[sm.z] & sm.w) != 0;
          const uint4 sn2 = __ldg(sum + 2 * (g0 + lane) + 1);
          if (!live && sn2.x != 0xffffffffu) live = (row[sn2.x] & sn2.y) != 0;
          if (!live && sn2.z != 0xffffffffu) live = (row[sn2.z] & sn2.w) != 0;
        }
      }
    }
    unsigned lm = __ballot_sync(0xffffffffu, live);
    while (lm) {
      const int ci = __ffs(lm) - 1;
      lm &= lm - 1;
      const int64_t k = l0 + ((g0 + ci) << 5) + lane;
      bool set = false;
      float trans = 0.f;
      uint32_t rslot = 0;
      if (k < l1) {
        rslot = a.L_slot[k];
        const uint32_t pos = a.L_pos[k];
        set = (row[pos >> 5] >> (pos & 31)) & 1u;
        if (set) {
          float dist;
          bool wr;
          trans = form_factor_pre(so, sn, a.s_origin_area[rslot], a.s_normal[rslot], dist, wr);
        }
      }
      unsigned m = __ballot_sync(0xffffffffu, set);
      if (EMIT) {
        if (set) {
          const int64_t o = outp + __popc(m & ((1u << lane) - 1));
          a.out_patch[o] = (uint32_t) a.slot_patch[rslot];
          a.out_transfer[o] = (uint16_t) (int) (trans * 65536.0f / rtotal);
        }
        outp += __popc(m);
      } else {
        count += __popc(m);
        while (m) {  // serial float sum in receiver order (qrad3.c:254); uniform across the warp
          const int src = __ffs(m) - 1;
          m &= m - 1;
          total += __shfl_sync(0xffffffffu, trans, src);
        }
      }
    }
  }
  if (!EMIT && lane == 0) {
    a.row_total[sslot] = total;
    a.row_count[sslot] = count;
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
  if (sslot >= 0 && a.slot_patch[sslot] >= 0) {
    const int c = a.slot_cluster[sslot];
    const float4 so = a.s_origin_area[sslot];
    const float4 sn = a.s_normal[sslot];
    for (int64_t k = a.Lptr[c] + lane; k < a.Lptr[c + 1]; k += 32) {
      const uint32_t rslot = a.L_slot[k];
      if ((int) rslot == sslot) continue;
      float dist;
      bool wr;
      form_factor_pre(so, sn, a.s_origin_area[rslot], a.s_normal[rslot], dist, wr);
      n += wr ? 1 : 0;
    }
  }
  for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if (lane == 0) a.sel_rays[widx] = n;
}

// pass 2, lane = row: one warp sums 32 consecutive source rows of one cluster at once.  The rows share the cluster's
// candidate list, and neighbouring sources accept nearly the same receivers, so the warp walks the list once, skips a
// 32-candidate chunk when none of its 32 rows has an accept bit in it, and otherwise visits only the candidates some
// row accepts: their receiver is staged in shared memory and every lane that accepts it adds its own form factor to
// its own serial sum (qrad3.c:254) -- the sums of 32 rows advance together instead of one shuffle + add per transfer
// with 31 lanes looking on (k_row_totals, kept for the download path: 206 ms on C5 against ~40 ms here).
__global__ void __launch_bounds__(kThreads) k_row_totals32(const RowArgs a) {
  __shared__ float4 s_po[kWarpsPerBlock][32], s_pn[kWarpsPerBlock][32];
  __shared__ uint32_t s_pos[kWarpsPerBlock][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int sslot0 = a.slot_lo + (blockIdx.x * kWarpsPerBlock + wid) * 32;
  if (sslot0 >= a.slot_hi) return;
  const int sslot = sslot0 + lane;
  const int c = a.slot_cluster[sslot0];              // a 32-slot slice lies in one cluster
  const bool live_row = a.slot_patch[sslot] >= 0;
  const uint32_t *row = a.bits + a.bits_base[c] + (int64_t) (sslot - a.cl_slot_start[c]) * a.nwords[c];
  const float4 so = a.s_origin_area[sslot];
  const float4 sn = a.s_normal[sslot];
  const int64_t l0 = a.Lptr[c], l1 = a.Lptr[c + 1];
  const int64_t nchunks = (l1 - l0 + 31) >> 5;
  const uint4 *sum = a.L_sum + 2 * a.chunk_base[c];
  float total = 0.f;
  int count = 0;
  uint4 sa_n = make_uint4(0, 0, 0, 0), sb_n = sa_n;
  if (nchunks > 0) {
    sa_n = __ldg(sum);
    sb_n = __ldg(sum + 1);
  }
  for (int64_t g = 0; g < nchunks; g++) {
    const uint4 sa = sa_n, sb = sb_n;
    if (g + 1 < nchunks) {
      sa_n = __ldg(sum + 2 * (g + 1));
      sb_n = __ldg(sum + 2 * (g + 1) + 1);
    }
    const bool cx = sa.x == 0xffffffffu;             // chunk touches more than four row words
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;         // this row's accept bits in the chunk's words
    if (!cx) {
      if (live_row) {
        a0 = row[sa.x] & sa.y;
        if (sa.z != 0xffffffffu) a1 = row[sa.z] & sa.w;
        if (sb.x != 0xffffffffu) a2 = row[sb.x] & sb.y;
        if (sb.z != 0xffffffffu) a3 = row[sb.z] & sb.w;
      }
      if (!__any_sync(0xffffffffu, (a0 | a1 | a2 | a3) != 0)) continue;
    }
    // stage the chunk's candidates: lane k holds candidate k; s_pos = which of the four words << 5 | bit
    const int64_t k = l0 + (g << 5) + lane;
    const bool have = k < l1;
    const uint32_t rslot = have ? a.L_slot[k] : 0u;
    const uint32_t pos = have ? a.L_pos[k] : 0u;
    const uint32_t w = pos >> 5;
    const uint32_t t = w == sa.x ? 0u : (w == sa.z ? 1u : (w == sb.x ? 2u : 3u));
    __syncwarp();
    s_pos[wid][lane] = cx ? pos : (t << 5 | (pos & 31));
    s_po[wid][lane] = a.s_origin_area[rslot];
    s_pn[wid][lane] = a.s_normal[rslot];
    __syncwarp();
    // candidates accepted by at least one of the 32 rows
    unsigned hot;
    if (!cx) {
      const uint32_t u0 = __reduce_or_sync(0xffffffffu, a0), u1 = __reduce_or_sync(0xffffffffu, a1);
      const uint32_t u2 = __reduce_or_sync(0xffffffffu, a2), u3 = __reduce_or_sync(0xffffffffu, a3);
      const uint32_t u = t == 0 ? u0 : (t == 1 ? u1 : (t == 2 ? u2 : u3));
      hot = __ballot_sync(0xffffffffu, have && ((u >> (pos & 31)) & 1u));
    } else
      hot = __ballot_sync(0xffffffffu, have);
    while (hot) {
      const int kk = __ffs(hot) - 1;
      hot &= hot - 1;
      const uint32_t pk = s_pos[wid][kk];   // the same for every lane
      uint32_t word;
      if (!cx) {
        const uint32_t tk = pk >> 5;
        word = tk == 0 ? a0 : (tk == 1 ? a1 : (tk == 2 ? a2 : a3));
      } else
        word = live_row ? row[pk >> 5] : 0u;
      if ((word >> (pk & 31)) & 1u) {
        float dist;
        bool wr;
        const float trans = form_factor_pre(so, sn, s_po[wid][kk], s_pn[wid][kk], dist, wr);
        total += trans;
        count++;
      }
    }
  }
  if (live_row) {
    a.row_total[sslot] = total;
    a.row_count[sslot] = count;
  }
}

// pass 3: columns.  One warp = 32 receivers (a slice), lane = receiver.
// pass 3 + bounce share the gather-form layout ("dense slices"): for every slice of 32 receivers the list of
// sources that reach at least one of them (ascending patch index = the order ShootLight adds them), one u32 source
// slot per list entry and 32 u16 weights per entry (0 = no transfer to that receiver; adding send * 0 leaves the
// float sum unchanged, and the add is skipped anyway).  Entries are grouped by 8: a lane's 8 weights are 16
// contiguous bytes (one 128-bit load), a warp reads 512 contiguous bytes per group.  Measured on C5: 18.8 of the
// 32 receivers of a listed source are hit on average, so this stores 3.6 B per transfer instead of 6 and turns
// the per-transfer gather of the radiosity vector into one broadcast load per 18.8 transfers (the per-column ELL
// of the first version was bound by L1 sector lookups at 43 % of HBM peak, profiles/r01c_full_c5_k_bounce.txt).
struct ColArgs {
  const float4 *s_origin_area, *s_normal;
  const int32_t *slot_patch, *slot_cluster;
  const int64_t *bits_base, *Rptr;
  const int32_t *nwords, *cl_slot_start;
  const int32_t *woff_table;
  int nc;
  const uint32_t *R_slot;
  const int64_t *R_addr;         // word address of (source row, first slice of this receiver cluster)
  const uint32_t *bits;
  const float *row_total;
  int M;
  const int32_t *blk_slice0;     // k_columns, per block: first slice (up to 8 consecutive slices of ONE cluster)
  const int32_t *blk_n;          // per block: slices (1..8)
  int32_t *slice_len;            // count mode output: list entries of the slice
  const int64_t *slice_base;     // fill mode: first entry of the slice (multiple of 32)
  uint32_t *g_src;               // [entries]      source slot
  uint4 *g_w;                    // [entries / 8][32] 8 x u16 weights per lane
};

// pass 3a: list lengths.  One block per receiver cluster, lane = SOURCE: per batch of 32 sources a lane reads the
// words of its source row for up to 32 slices of the cluster back to back (consecutive sectors of one row, used
// completely) and the warp counts the non-zero words per slice with one ballot each; lane j keeps the count of slice
// j; the 8 warps of the block take every 8th batch.  Measured on C5 (profiles/r01k_full_c5_columns.txt): counting with
// one word per lane and slice reads the 26.7 GB matrix ten times over from HBM.
__global__ void __launch_bounds__(kThreads) k_column_counts(const ColArgs a, const int c_first, const int slice_lo,
                                                            const int slice_hi) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int d = c_first + blockIdx.x;
  const int cs0 = a.cl_slot_start[d] >> 5, cs1 = a.cl_slot_start[d + 1] >> 5;   // the cluster's slices
  const int s_lo = max(cs0, slice_lo), s_hi = min(cs1, slice_hi);               // ... that belong to this rank
  const int64_t r0 = a.Rptr[d], r1 = a.Rptr[d + 1];
  constexpr int64_t kStep = 32 * kWarpsPerBlock;
  for (int p0 = s_lo; p0 < s_hi; p0 += 32) {             // passes of 32 slices (one for all but huge clusters)
    const int np = min(32, s_hi - p0);
    const int woff = p0 - cs0;
    int mycnt = 0;
    const int64_t first = r0 + wid * 32 + lane;
    int64_t addr_n = first < r1 ? a.R_addr[first] : -1;
    for (int64_t k = first; k - lane < r1; k += kStep) {   // k - lane is warp-uniform
      const int64_t addr = addr_n;
      addr_n = k + kStep < r1 ? a.R_addr[k + kStep] : -1;
      for (int j0 = 0; j0 < np; j0 += 8) {
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; j++) w[j] = (addr >= 0 && j0 + j < np) ? __ldcs(a.bits + addr + woff + j0 + j) : 0u;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const int cnt = __popc(__ballot_sync(0xffffffffu, w[j] != 0));
          if (lane == j0 + j) mycnt += cnt;
        }
      }
    }
    if (lane < np && mycnt) atomicAdd(a.slice_len + p0 + lane, mycnt);
  }
}

// pass 3b: the lists.  One warp per slice, lane = receiver; the 8 warps of a block own 8 neighbouring slices of one
// cluster and walk the cluster's source list together, re-aligned by a barrier every 16 batches of 32 sources, so
// that the words they fetch for one source (one 32-byte sector) come from HBM once and are then served by L1 -- a
// warp per slice on its own reads the matrix 25 times over (profiles/r01k_full_c5_columns.txt: 661 GB, HBM-bound),
// and a barrier per batch makes every warp wait for the one with the most entries.
// Sources are taken 32 at a time: lane k fetches the word of source k (addresses and words are requested two / one
// batch ahead); when some word is non-zero the sources that own one are staged in shared memory (origin, normal, row
// total: one gather per batch instead of a dependent L2 round trip per list entry), and the warp then emits the
// entries in source order.
__global__ void __launch_bounds__(kThreads) k_columns(const ColArgs a) {
  __shared__ float4 s_so[kWarpsPerBlock][32], s_sn[kWarpsPerBlock][32];
  __shared__ float s_rt[kWarpsPerBlock][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int slice0 = a.blk_slice0[blockIdx.x];
  const bool active = wid < a.blk_n[blockIdx.x];
  const int slice = slice0 + (active ? wid : 0);
  const int rslot = slice * 32 + lane;
  const int d = a.slot_cluster[slice0 * 32];           // all slots of the block share the cluster
  const int wslice = (slice * 32 - a.cl_slot_start[d]) >> 5;
  const bool valid = active && a.slot_patch[rslot] >= 0;
  float4 po = make_float4(0, 0, 0, 0), pn = po;
  if (valid) {
    po = a.s_origin_area[rslot];
    pn = a.s_normal[rslot];
  }
  const int64_t r0 = a.Rptr[d], r1 = a.Rptr[d + 1];
  int n = 0;                                           // list entries so far (uniform)
  const int64_t base = a.slice_base[slice];
  uint32_t wq[4] = {0, 0, 0, 0};                       // this lane's weights of the current group of 8 entries
  // pipeline: addresses two batches ahead, words one batch ahead
  int64_t addr_n = (active && r0 + lane < r1) ? a.R_addr[r0 + lane] : -1;
  uint32_t word_n = addr_n >= 0 ? a.bits[addr_n + wslice] : 0u;
  addr_n = (active && r0 + 32 + lane < r1) ? a.R_addr[r0 + 32 + lane] : -1;
  int batch = 0;
  for (int64_t k0 = r0; k0 < r1; k0 += 32, batch++) {
    if ((batch & 15) == 0) __syncthreads();
    const uint32_t word = word_n;
    word_n = addr_n >= 0 ? a.bits[addr_n + wslice] : 0u;
    addr_n = (active && k0 + 64 + lane < r1) ? a.R_addr[k0 + 64 + lane] : -1;
    unsigned nz = __ballot_sync(0xffffffffu, word != 0);
    if (nz == 0) continue;
    uint32_t sslot = 0;
    __syncwarp();
    if (word != 0) {
      sslot = a.R_slot[k0 + lane];
      s_so[wid][lane] = a.s_origin_area[sslot];
      s_sn[wid][lane] = a.s_normal[sslot];
      s_rt[wid][lane] = a.row_total[sslot];
    }
    __syncwarp();
    while (nz) {
      const int src = __ffs(nz) - 1;
      nz &= nz - 1;
      const uint32_t w = __shfl_sync(0xffffffffu, word, src);
      const uint32_t ss = __shfl_sync(0xffffffffu, sslot, src);
      uint32_t itrans = 0;
      if ((w >> lane) & 1u) {
        float dist;
        bool wr;
        const float trans = form_factor_pre(s_so[wid][src], s_sn[wid][src], po, pn, dist, wr);
        itrans = (uint32_t) (int) (trans * 65536.0f / s_rt[wid][src]) & 0xffffu;   // u16 store wraps (qrad3.c:283-285)
      }
      const int q = n & 7;
      wq[q >> 1] |= itrans << ((q & 1) * 16);
      if (lane == 0) a.g_src[base + n] = ss;
      n++;
      if ((n & 7) == 0) {
        a.g_w[((base + n) / 8 - 1) * 32 + lane] = make_uint4(wq[0], wq[1], wq[2], wq[3]);
        wq[0] = wq[1] = wq[2] = wq[3] = 0;
      }
    }
  }
  if (active && (n & 7)) {
    // last, partial group; the rest of the 32-entry batch stays zero (source slot 0, zero weights: buffers are cleared)
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
// in HBM.  Each warp owns a ring of three batch buffers in shared memory; lane 0 refills a buffer with two bulk async
// copies (cp.async.bulk, completion counted on an mbarrier) as soon as the warp has consumed it, so two batches are
// always in flight per warp without holding a register, and the radiosity vectors of the next batch are gathered while
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
__global__ void __launch_bounds__(kThreads) k_bounce(const BounceArgs a) {
  extern __shared__ __align__(128) unsigned char s_bounce[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned char *const wbase = s_bounce + wid * kBounceWarpBytes;
  float *const s_vec = reinterpret_cast<float *>(wbase + kBounceVecOfs);
  const uint32_t stage0 = smem_u32(wbase), bar0 = stage0 + kBounceBarOfs;
  if (lane == 0) {
    for (int s = 0; s < kBounceStages; s++) mbar_init(bar0 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (blockIdx.x == 0 && threadIdx.x == 0) a.next[a.parity ^ 1] = 0;
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
                                float *__restrict__ radiosity) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  const int s = patch_slot[p];
  if (s < 0) return;
  const float4 t = total[s];
  totallight[p * 3] = t.x;
  totallight[p * 3 + 1] = t.y;
  totallight[p * 3 + 2] = t.z;
  const float4 r = send[s];
  radiosity[p * 3] = r.x * 65536.0f;
  radiosity[p * 3 + 1] = r.y * 65536.0f;
  radiosity[p * 3 + 2] = r.z * 65536.0f;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// phase 0: slot order, row geometry, pass 1 (this rank's granules of tasks)
void transfers_trace(qrad_cuda_ctx *c, int nopvs) {
  if (!c->numnodes || !c->N) dfail("make_transfers: BSP and patches must be uploaded first");
  if (!c->have_vis) dfail("make_transfers: map has no visibility data (the reference forces numbounce = 0)");
  c->t_transfers.start(c->stream);
  const auto wall0 = std::chrono::steady_clock::now();
  c->transfers_done = false;
  const int N = c->N;
  // ---- host: slot order ----
  // With -nopvs every patch receives (even those in solid); sources still need a valid cluster
  // (PvsForOrigin fails for them, qrad3.c:208-209).  That case is one pseudo cluster holding everybody.
  const int nc = nopvs ? 1 : c->numclusters;
  if (nc <= 0) dfail("make_transfers: no clusters");
  if ((int64_t) nc * nc > (int64_t) 1 << 28) dfail("make_transfers: %d clusters exceed the dense PVS table limit", nc);
  std::vector<int32_t> eff(N);
  for (int i = 0; i < N; i++) {
    int cl = c->h_cluster[i];
    if (cl >= c->numclusters) dfail("patch %d has cluster %d >= numclusters %d", i, cl, c->numclusters);
    eff[i] = nopvs ? 0 : cl;
  }
  c->nc_eff = nc;
  c->h_cl_count.assign(nc, 0);
  for (int i = 0; i < N; i++)
    if (eff[i] >= 0) c->h_cl_count[eff[i]]++;
  c->h_cl_slot_start.assign(nc + 1, 0);
  for (int k = 0; k < nc; k++) c->h_cl_slot_start[k + 1] = c->h_cl_slot_start[k] + ((c->h_cl_count[k] + 31) & ~31);
  const int M = c->h_cl_slot_start[nc];
  c->M = M;
  {  // shards of slots for passes 2, 3 and the bounces: equal runs of whole slices
    const int nsl = M / 32, per = (nsl + c->world - 1) / c->world;
    c->slice_lo = std::min(nsl, per * c->rank);
    c->slice_hi = std::min(nsl, per * (c->rank + 1));
    c->shard_slots = per * 32;
    c->Mpad = c->shard_slots * c->world;
  }
  if (M == 0) dfail("make_transfers: no patch lies in a valid cluster");
  c->h_slot_patch.assign(M, -1);
  c->h_patch_slot.assign(N, -1);
  c->h_slot_cluster.assign(M, 0);
  {
    std::vector<int32_t> fill(c->h_cl_slot_start.begin(), c->h_cl_slot_start.end() - 1);
    for (int i = 0; i < N; i++)
      if (eff[i] >= 0) {
        int s = fill[eff[i]]++;
        c->h_slot_patch[s] = i;
        c->h_patch_slot[i] = s;
      }
    for (int k = 0; k < nc; k++)
      for (int s = c->h_cl_slot_start[k]; s < c->h_cl_slot_start[k + 1]; s++) c->h_slot_cluster[s] = k;
  }
  // sources in solid never shoot under -nopvs: remember them so their rows are skipped
  std::vector<uint8_t> source_ok(M, 0);
  for (int s = 0; s < M; s++) {
    int p = c->h_slot_patch[s];
    source_ok[s] = p >= 0 && c->h_cluster[p] >= 0;
  }
  // ---- host: visible cluster lists and row geometry ----
  std::vector<int32_t> vis_ptr(nc + 1, 0), vis_cluster, vis_woff, woff_table((size_t) nc * nc, -1);
  c->h_nwords.assign(nc, 0);
  c->h_bits_base.assign(nc + 1, 0);
  c->h_Lptr.assign(nc + 1, 0);
  std::vector<int64_t> task_base(nc + 1, 0);
  std::vector<int64_t> Rptr(nc + 1, 0);
  for (int a = 0; a < nc; a++) {
    int w = 0;
    int64_t cand = 0;
    for (int d = 0; d < nc; d++) {
      bool vis = nopvs ? true : ((c->h_pvs[(size_t) a * c->pvs_words + (d >> 5)] >> (d & 31)) & 1u);
      if (!vis || c->h_cl_count[d] == 0) continue;
      vis_cluster.push_back(d);
      vis_woff.push_back(w);
      woff_table[(size_t) a * nc + d] = w;
      w += (c->h_cl_slot_start[d + 1] - c->h_cl_slot_start[d]) >> 5;
      cand += c->h_cl_count[d];
      Rptr[d + 1] += c->h_cl_count[a];
    }
    vis_ptr[a + 1] = (int32_t) vis_cluster.size();
    c->h_nwords[a] = w;
    c->h_bits_base[a + 1] = c->h_bits_base[a] + (int64_t) c->h_cl_count[a] * w;
    task_base[a + 1] = task_base[a] + (int64_t) c->h_cl_count[a] * ((w + kChunkWords - 1) / kChunkWords);
    c->h_Lptr[a + 1] = c->h_Lptr[a] + cand;
  }
  for (int d = 0; d < nc; d++) Rptr[d + 1] += Rptr[d];
  // a cluster that sees nothing would leave a row of zero words; bits_base must stay strictly usable
  const uint64_t W = (uint64_t) c->h_bits_base[nc];
  c->bits_words = W;
  if (c->h_Lptr[nc] >= ((int64_t) 1 << 40)) dfail("make_transfers: candidate lists too large");

  cudaStream_t st = c->stream;
  c->slot_patch.upload(c->h_slot_patch, st);
  c->patch_slot.upload(c->h_patch_slot, st);
  c->slot_cluster.upload(c->h_slot_cluster, st);
  {
    std::vector<float4> so(M, make_float4(0, 0, 0, 0)), sn(M, make_float4(0, 0, 0, 0));
    const std::vector<float4> &po = c->h_origin_area, &pn = c->h_normal;   // host copies kept by upload_patches
    for (int s = 0; s < M; s++) {
      int p = c->h_slot_patch[s];
      if (p >= 0) {
        so[s] = po[p];
        sn[s] = pn[p];
      }
    }
    c->s_origin_area.upload(so, st);
    c->s_normal.upload(sn, st);
    std::vector<float4> bmn((size_t) M / 32, make_float4(3e38f, 3e38f, 3e38f, 0)), bmx((size_t) M / 32, make_float4(-3e38f, -3e38f, -3e38f, 0));
    for (int s = 0; s < M; s++) {
      if (c->h_slot_patch[s] < 0) continue;
      float4 &lo4 = bmn[s >> 5], &hi4 = bmx[s >> 5];
      lo4.x = std::min(lo4.x, so[s].x); lo4.y = std::min(lo4.y, so[s].y); lo4.z = std::min(lo4.z, so[s].z);
      hi4.x = std::max(hi4.x, so[s].x); hi4.y = std::max(hi4.y, so[s].y); hi4.z = std::max(hi4.z, so[s].z);
    }
    c->wbox_min.upload(bmn, st);
    c->wbox_max.upload(bmx, st);
  }
  c->d_vis_ptr.upload(vis_ptr, st);
  c->d_vis_cluster.upload(vis_cluster, st);
  c->d_vis_woff.upload(vis_woff, st);
  c->d_woff_table.upload(woff_table, st);
  c->d_bits_base.upload(c->h_bits_base, st);
  c->d_Lptr.upload(c->h_Lptr, st);
  c->d_nwords.upload(c->h_nwords, st);
  c->d_cl_slot_start.upload(c->h_cl_slot_start, st);
  c->d_cl_count.upload(c->h_cl_count, st);
  c->h_Rptr = Rptr;
  c->d_Rptr.upload(Rptr, st);
  c->d_task_base.upload(task_base, st);
  c->bits.alloc((size_t) W + 1);
  c->counters.alloc(16);
  c->counters.zero(st);

  c->sync("make_transfers uploads");
  c->stats.ms_tr_setup = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
  // ---- pass 1 ----
  Timer ttrace;
  ttrace.start(st);
  if (W) {
    // rows of patches that sit in solid (only possible under -nopvs) are cleared afterwards
    VisArgs a;
    a.tree = c->tree();
    a.leaf_path = c->leaf_path.p;
    a.leaf_path_first = c->leaf_path_first.p;
    a.wbox_min = c->wbox_min.p;
    a.wbox_max = c->wbox_max.p;
    a.s_origin_area = c->s_origin_area.p;
    a.s_normal = c->s_normal.p;
    a.bits_base = c->d_bits_base.p;
    a.task_base = c->d_task_base.p;
    a.nwords = c->d_nwords.p;
    a.cl_slot_start = c->d_cl_slot_start.p;
    a.cl_count = c->d_cl_count.p;
    a.vis_ptr = c->d_vis_ptr.p;
    a.vis_cluster = c->d_vis_cluster.p;
    a.vis_woff = c->d_vis_woff.p;
    a.nc = nc;
    // multi-GPU: tasks are dealt to the ranks in granules; words of other ranks stay zero so that the exchange
    // is one sum all-reduce of the matrix (each word is written by exactly one rank)
    const uint64_t T = (uint64_t) task_base[nc];
    const uint64_t granules = (T + kTaskGranule - 1) / kTaskGranule;
    const uint64_t mine_g = granules / c->world + ((granules % c->world) > (uint64_t) c->rank ? 1 : 0);
    a.num_tasks = T;
    a.my_tasks = mine_g * kTaskGranule;
    a.rank = c->rank;
    a.world = c->world;
    if (c->world > 1) c->bits.zero(st);
    a.bits = c->bits.p;
    a.counters = c->counters.p;
    const uint64_t mine = a.my_tasks;
    const uint64_t want_blocks = (mine / kTaskRun + kWarpsPerBlock) / kWarpsPerBlock;
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
  c->stats.ms_transfers_trace = ttrace.stop();
  if (nopvs) {
    // clear rows of sources in solid: they never shoot (qrad3.c:208)
    for (int s = 0; s < M; s++)
      if (c->h_slot_patch[s] >= 0 && !source_ok[s])
        QCUDA(cudaMemsetAsync(c->bits.p + c->h_bits_base[0] + (int64_t) s * c->h_nwords[0], 0, (size_t) c->h_nwords[0] * 4, st));
  }

  c->sync("make_transfers pass 1");
}

// phase 1: ordered lists + pass 2 (row totals of this rank's slots).  Needs the complete bit matrix.
void transfers_rows(qrad_cuda_ctx *c) {
  cudaStream_t st = c->stream;
  auto wall0 = std::chrono::steady_clock::now();
  const int N = c->N, M = c->M, nc = c->nc_eff;
  // ---- ordered lists ----
  c->L_slot.alloc((size_t) c->h_Lptr[nc] + 1);
  c->L_pos.alloc((size_t) c->h_Lptr[nc] + 1);
  c->R_slot.alloc((size_t) c->h_Rptr[nc] + 1);
  {
    ListArgs la;
    la.N = N;
    la.nc = nc;
    la.patch_slot = c->patch_slot.p;
    la.slot_cluster = c->slot_cluster.p;
    la.woff_table = c->d_woff_table.p;
    la.cl_slot_start = c->d_cl_slot_start.p;
    la.list_ptr = c->d_Lptr.p;
    la.out_slot = c->L_slot.p;
    la.out_aux = c->L_pos.p;
    la.transpose = 0;
    la.bits_base = nullptr;
    la.nwords = nullptr;
    la.out_addr = nullptr;
    // only the clusters this rank's slots belong to (all of them on one GPU)
    const int c_lo = c->slice_hi > c->slice_lo ? c->h_slot_cluster[(size_t) c->slice_lo * 32] : 0;
    const int c_hi = c->slice_hi > c->slice_lo ? c->h_slot_cluster[(size_t) c->slice_hi * 32 - 1] : -1;
    const int c_n = c_hi - c_lo + 1;
    la.c_first = c_lo;
    if (c_n > 0) k_build_lists<<<c_n, 1024, 0, st>>>(la);
    c->launch_check("k_build_lists(rows)");
    la.list_ptr = c->d_Rptr.p;
    la.out_slot = c->R_slot.p;
    la.out_aux = nullptr;
    la.transpose = 1;
    la.bits_base = c->d_bits_base.p;
    la.nwords = c->d_nwords.p;
    c->R_addr.alloc((size_t) c->h_Rptr[nc] + 1);
    la.out_addr = c->R_addr.p;
    if (c_n > 0) k_build_lists<<<c_n, 1024, 0, st>>>(la);
    c->launch_check("k_build_lists(cols)");
  }

  {
    std::vector<int64_t> chunk_base((size_t) nc + 1, 0);
    for (int k = 0; k < nc; k++) chunk_base[k + 1] = chunk_base[k] + ((c->h_Lptr[k + 1] - c->h_Lptr[k] + 31) >> 5);
    c->d_chunk_base.upload(chunk_base, st);
    c->L_sum.alloc((size_t) 2 * chunk_base[nc] + 2);
    const int c_lo = c->slice_hi > c->slice_lo ? c->h_slot_cluster[(size_t) c->slice_lo * 32] : 0;
    const int c_hi = c->slice_hi > c->slice_lo ? c->h_slot_cluster[(size_t) c->slice_hi * 32 - 1] : -1;
    const int64_t ch_lo = chunk_base[c_lo], ch_hi = c_hi >= c_lo ? chunk_base[c_hi + 1] : chunk_base[c_lo];
    if (ch_hi > ch_lo) {
      k_list_summaries<<<(unsigned) ((ch_hi - ch_lo + kWarpsPerBlock - 1) / kWarpsPerBlock), kThreads, 0, st>>>(
          nc, c->d_Lptr.p, c->d_chunk_base.p, c->L_pos.p, c->L_sum.p, ch_lo, ch_hi);
      c->launch_check("k_list_summaries");
    }
  }
  c->sync("make_transfers lists");
  c->stats.ms_tr_lists = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
  wall0 = std::chrono::steady_clock::now();
  // ---- pass 2 ----
  c->row_total.alloc((size_t) c->Mpad);
  c->row_count.alloc((size_t) c->Mpad);
  c->row_total.zero(st);
  c->row_count.zero(st);
  {
    RowArgs ra{};
    ra.s_origin_area = c->s_origin_area.p;
    ra.s_normal = c->s_normal.p;
    ra.slot_patch = c->slot_patch.p;
    ra.slot_cluster = c->slot_cluster.p;
    ra.bits_base = c->d_bits_base.p;
    ra.Lptr = c->d_Lptr.p;
    ra.nwords = c->d_nwords.p;
    ra.cl_slot_start = c->d_cl_slot_start.p;
    ra.L_slot = c->L_slot.p;
    ra.L_pos = c->L_pos.p;
    ra.L_sum = c->L_sum.p;
    ra.chunk_base = c->d_chunk_base.p;
    ra.bits = c->bits.p;
    ra.M = M;
    ra.slot_lo = c->slice_lo * 32;
    ra.slot_hi = c->slice_hi * 32;
    ra.row_total = c->row_total.p;
    ra.row_count = c->row_count.p;
    const int rows = ra.slot_hi - ra.slot_lo;   // whole 32-slot slices
    if (rows > 0) {
      k_row_totals32<<<(rows / 32 + kWarpsPerBlock - 1) / kWarpsPerBlock, kThreads, 0, st>>>(ra);
      c->launch_check("k_row_totals32");
    }
  }
  c->sync("make_transfers pass 2");
  c->stats.ms_tr_rows = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
}

// phase 2: pass 3 (gather lists of this rank's slices).  Needs the row totals of every source.
void transfers_columns(qrad_cuda_ctx *c) {
  cudaStream_t st = c->stream;
  const auto wall0 = std::chrono::steady_clock::now();
  const int M = c->M, nc = c->nc_eff;
  const int nslices = M / 32;
  c->slice_len.alloc((size_t) nslices);
  c->slice_len.zero(st);
  ColArgs ca{};
  ca.s_origin_area = c->s_origin_area.p;
  ca.s_normal = c->s_normal.p;
  ca.slot_patch = c->slot_patch.p;
  ca.slot_cluster = c->slot_cluster.p;
  ca.bits_base = c->d_bits_base.p;
  ca.Rptr = c->d_Rptr.p;
  ca.nwords = c->d_nwords.p;
  ca.cl_slot_start = c->d_cl_slot_start.p;
  ca.woff_table = c->d_woff_table.p;
  ca.nc = nc;
  ca.R_slot = c->R_slot.p;
  ca.R_addr = c->R_addr.p;
  ca.bits = c->bits.p;
  ca.row_total = c->row_total.p;
  ca.M = M;
  // blocks: runs of up to 8 consecutive slices that lie in one cluster and in this rank's range
  std::vector<int32_t> blk_slice0, blk_n;
  for (int s = c->slice_lo; s < c->slice_hi;) {
    const int cl = c->h_slot_cluster[(size_t) s * 32];
    const int cl_end = c->h_cl_slot_start[cl + 1] / 32;
    const int e = std::min(std::min(s + kWarpsPerBlock, cl_end), c->slice_hi);
    blk_slice0.push_back(s);
    blk_n.push_back(e - s);
    s = e;
  }
  const int cblocks = (int) blk_slice0.size();
  c->col_blk_slice0.upload(blk_slice0, st);
  c->col_blk_n.upload(blk_n, st);
  ca.blk_slice0 = c->col_blk_slice0.p;
  ca.blk_n = c->col_blk_n.p;
  ca.slice_len = c->slice_len.p;
  if (c->slice_hi > c->slice_lo) {   // slice_len was cleared above
    const int c_lo = c->h_slot_cluster[(size_t) c->slice_lo * 32], c_hi = c->h_slot_cluster[(size_t) c->slice_hi * 32 - 1];
    k_column_counts<<<c_hi - c_lo + 1, kThreads, 0, st>>>(ca, c_lo, c->slice_lo, c->slice_hi);
  }
  c->launch_check("k_column_counts");
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
  }
  c->slice_base.upload(h_slice_base, st);
  {  // k_bounce hands this rank's slices to its warps longest list first
    std::vector<int32_t> order;
    for (int s = c->slice_lo; s < c->slice_hi; s++) order.push_back(s);
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return h_len[x] > h_len[y]; });
    if (order.empty()) order.push_back(0);
    c->slice_order.upload(order, st);
  }
  // k_bounce prefetches without bounds checks: slots two batches, weights one batch past the end of a list
  c->g_src.alloc((size_t) c->sell_entries + 96);
  c->g_w.alloc((size_t) (c->sell_entries / 8 + 8) * 32);
  c->g_src.zero(st);
  c->g_w.zero(st);
  ca.slice_base = c->slice_base.p;
  ca.g_src = c->g_src.p;
  ca.g_w = c->g_w.p;
  if (cblocks > 0) k_columns<<<cblocks, kThreads, 0, st>>>(ca);
  c->launch_check("k_columns");
  std::vector<int32_t> h_rc((size_t) M);
  c->row_count.download(h_rc.data(), (size_t) M, st);
  int64_t nnz = 0;
  for (int s = 0; s < M; s++) nnz += h_rc[s];
  c->sync("make_transfers");

  unsigned long long hc[16];
  c->counters.download(hc, 16, st);
  if (c->count_nodes) {   // warp cycles per phase of k_visibility (tools/gpu_raycounts.py)
    const double tot = (double) (hc[8] + hc[9] + hc[10] + hc[11] + hc[12]);
    if (tot > 0)
      fprintf(stderr, "k_visibility warp cycles: box walks %.1f %%, pair tests %.1f %%, guess walks %.1f %%, full walks %.1f %%, "
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
  if (c->world > 1) dfail("make_transfers: this context is a shard (%d/%d); use qrad_cuda_make_transfers_phase and exchange", c->rank, c->world);
  transfers_trace(c, nopvs);
  transfers_rows(c);
  transfers_columns(c);
}

// reference row format for parity tests: rows by source patch, receivers ascending
void download_transfers_impl(qrad_cuda_ctx *c, int64_t *row_ptr, uint32_t *patch, uint16_t *transfer) {
  if (c->world > 1) dfail("download_transfers: not available on a shard (row lists exist only for this rank's clusters)");
  const int N = c->N, M = c->M;
  cudaStream_t st = c->stream;
  std::vector<int32_t> rc((size_t) M);
  c->row_count.download(rc.data(), rc.size(), st);
  row_ptr[0] = 0;
  for (int i = 0; i < N; i++) row_ptr[i + 1] = row_ptr[i] + (c->h_patch_slot[i] >= 0 ? rc[c->h_patch_slot[i]] : 0);
  const int64_t nnz = row_ptr[N];
  DBuf<int64_t> d_row_ptr;
  d_row_ptr.upload(row_ptr, (size_t) N + 1, st);
  DBuf<uint32_t> d_patch;
  DBuf<uint16_t> d_tr;
  d_patch.alloc((size_t) nnz + 1);
  d_tr.alloc((size_t) nnz + 1);
  RowArgs ra{};
  ra.s_origin_area = c->s_origin_area.p;
  ra.s_normal = c->s_normal.p;
  ra.slot_patch = c->slot_patch.p;
  ra.slot_cluster = c->slot_cluster.p;
  ra.bits_base = c->d_bits_base.p;
  ra.Lptr = c->d_Lptr.p;
  ra.nwords = c->d_nwords.p;
  ra.cl_slot_start = c->d_cl_slot_start.p;
  ra.L_slot = c->L_slot.p;
  ra.L_pos = c->L_pos.p;
  ra.L_sum = c->L_sum.p;
  ra.chunk_base = c->d_chunk_base.p;
  ra.bits = c->bits.p;
  ra.M = M;
  ra.slot_lo = 0;
  ra.slot_hi = M;
  ra.row_total = c->row_total.p;
  ra.row_count = c->row_count.p;
  ra.row_ptr = d_row_ptr.p;
  ra.out_patch = d_patch.p;
  ra.out_transfer = d_tr.p;
  k_row_totals<true><<<(M + kWarpsPerBlock - 1) / kWarpsPerBlock, kThreads, 0, st>>>(ra);
  c->launch_check("k_row_totals<emit>");
  d_patch.download(patch, (size_t) nnz, st);
  d_tr.download(transfer, (size_t) nnz, st);
}

// the same for selected source patches only (row-sample parity checks on maps whose whole matrix is tens of GB)
void download_rows_impl(qrad_cuda_ctx *c, int nrows, const int32_t *rows, int64_t *row_ptr, uint32_t *patch,
                        uint16_t *transfer, int64_t capacity, int64_t *rays) {
  if (c->world > 1) dfail("download_rows: not available on a shard");
  if (nrows <= 0) return;
  cudaStream_t st = c->stream;
  std::vector<int32_t> rc((size_t) c->M);
  c->row_count.download(rc.data(), rc.size(), st);
  std::vector<int32_t> sel((size_t) nrows);
  row_ptr[0] = 0;
  for (int k = 0; k < nrows; k++) {
    if (rows[k] < 0 || rows[k] >= c->N) dfail("download_rows: patch %d out of range", rows[k]);
    sel[k] = c->h_patch_slot[rows[k]];
    row_ptr[k + 1] = row_ptr[k] + (sel[k] >= 0 ? rc[sel[k]] : 0);
  }
  const int64_t nnz = row_ptr[nrows];
  if (nnz > capacity) dfail("download_rows: %lld records do not fit the caller's %lld", (long long) nnz, (long long) capacity);
  DBuf<int64_t> d_row_ptr;
  d_row_ptr.upload(row_ptr, (size_t) nrows + 1, st);
  DBuf<int32_t> d_sel;
  d_sel.upload(sel, st);
  DBuf<uint32_t> d_patch;
  DBuf<uint16_t> d_tr;
  DBuf<long long> d_rays;
  d_patch.alloc((size_t) nnz + 1);
  d_tr.alloc((size_t) nnz + 1);
  d_rays.alloc((size_t) nrows);
  RowArgs ra{};
  ra.s_origin_area = c->s_origin_area.p;
  ra.s_normal = c->s_normal.p;
  ra.slot_patch = c->slot_patch.p;
  ra.slot_cluster = c->slot_cluster.p;
  ra.bits_base = c->d_bits_base.p;
  ra.Lptr = c->d_Lptr.p;
  ra.nwords = c->d_nwords.p;
  ra.cl_slot_start = c->d_cl_slot_start.p;
  ra.L_slot = c->L_slot.p;
  ra.L_pos = c->L_pos.p;
  ra.L_sum = c->L_sum.p;
  ra.chunk_base = c->d_chunk_base.p;
  ra.bits = c->bits.p;
  ra.M = c->M;
  ra.slot_lo = 0;
  ra.slot_hi = c->M;
  ra.row_total = c->row_total.p;
  ra.row_count = c->row_count.p;
  ra.row_ptr = d_row_ptr.p;
  ra.out_patch = d_patch.p;
  ra.out_transfer = d_tr.p;
  ra.sel_slots = d_sel.p;
  ra.nsel = nrows;
  ra.sel_rays = d_rays.p;
  const int blocks = (nrows + kWarpsPerBlock - 1) / kWarpsPerBlock;
  k_row_totals<true><<<blocks, kThreads, 0, st>>>(ra);
  c->launch_check("k_row_totals<emit, selected>");
  if (rays) {
    k_row_rays<<<blocks, kThreads, 0, st>>>(ra);
    c->launch_check("k_row_rays");
  }
  d_patch.download(patch, (size_t) nnz, st);
  d_tr.download(transfer, (size_t) nnz, st);
  if (rays) {
    std::vector<long long> hr((size_t) nrows);
    d_rays.download(hr.data(), hr.size(), st);
    for (int k = 0; k < nrows; k++) rays[k] = hr[k];
  }
}

// BounceLight, qrad3.c:448-473, in three parts so that a multi-GPU host can all-gather the radiosity vector
// between bounces: begin (radiosity = samplelight * reflectivity * area), step (one ShootLight + CollectLight
// over this rank's slices), end (totallight back to patch order).
void bounce_begin(qrad_cuda_ctx *c, bool want_sum) {
  if (!c->transfers_done) dfail("bounce: make_transfers has not run");
  c->t_bounce.start(c->stream);
  cudaStream_t st = c->stream;
  const int M = c->M;
  c->rad_a.alloc((size_t) c->Mpad);
  c->rad_b.alloc((size_t) c->Mpad);
  c->s_total.alloc((size_t) c->Mpad);
  c->s_refl_area.alloc((size_t) c->Mpad);
  c->rad_a.zero(st);
  c->rad_b.zero(st);
  c->s_total.zero(st);
  k_init_bounce<<<(M + 255) / 256, 256, 0, st>>>(M, c->slot_patch.p, c->p_samplelight.p, c->p_totallight.p, c->p_refl.p,
                                                 c->p_origin_area.p, c->rad_a.p, c->s_total.p, c->s_refl_area.p);
  c->launch_check("k_init_bounce");
  if (want_sum) c->rad_sum.alloc((size_t) c->Mpad);
  else c->rad_sum.release();
  c->rad_cur = c->rad_a.p;
  c->rad_nxt = c->rad_b.p;
  c->bounce_kernel_ms = 0;
  c->bounce_steps = 0;
  c->bounce_next.alloc(2);
  c->bounce_next.zero(st);
  if (c->bounce_blocks_per_sm == 0) {
    int per_sm = 0;
    QCUDA(cudaFuncSetAttribute(k_bounce, cudaFuncAttributeMaxDynamicSharedMemorySize, kWarpsPerBlock * kBounceWarpBytes));
    QCUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bounce, kThreads, kWarpsPerBlock * kBounceWarpBytes));
    c->bounce_blocks_per_sm = std::max(1, per_sm);
    QCUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bounce_regs, kThreads, 0));
    c->bounce_regs_blocks_per_sm = std::max(1, per_sm);
  }
  c->sync("bounce_begin");
}

void bounce_step(qrad_cuda_ctx *c) {
  cudaStream_t st = c->stream;
  BounceArgs a{};
  a.slice_len = c->slice_len.p;
  a.slice_base = c->slice_base.p;
  a.g_src = c->g_src.p;
  a.g_w = c->g_w.p;
  a.total = c->s_total.p;
  a.refl_area = c->s_refl_area.p;
  a.rad_sum = c->rad_sum.p;
  a.M = c->M;
  a.slice_lo = c->slice_lo;
  a.slice_hi = c->slice_hi;
  a.send = c->rad_cur;
  a.send_next = c->rad_nxt;
  a.order = c->slice_order.p;
  a.next = c->bounce_next.p;
  a.parity = c->bounce_steps & 1;
  a.slice_cut = INT_MAX;
  int nsl = c->slice_hi - c->slice_lo;
  if (const char *f = getenv("QRAD_DEBUG_BOUNCE_SHARD")) {   // time the launch an n-GPU shard would see, on one GPU
    nsl = std::max(1, (int) (nsl * atof(f)));
    a.slice_cut = c->slice_lo + nsl;
  }
  // persistent warps: as many blocks as fit on the chip, slices are drawn from the work counter
  const int want = std::max(1, (c->slice_hi - c->slice_lo + kWarpsPerBlock - 1) / kWarpsPerBlock);
  Timer tk;
  tk.start(st);
  // few slices per resident warp: the launch lasts as long as ONE warp needs for the longest list -> bulk-copy ring
  bool ring = (int64_t) nsl < (int64_t) 4 * c->sm_count * 40;
  if (const char *force = getenv("QRAD_BOUNCE_KERNEL")) ring = strcmp(force, "regs") != 0;   // tests cover both
  if (ring) k_bounce<<<std::min(want, c->sm_count * c->bounce_blocks_per_sm), kThreads, kWarpsPerBlock * kBounceWarpBytes, st>>>(a);
  else k_bounce_regs<<<std::min(want, c->sm_count * c->bounce_regs_blocks_per_sm), kThreads, 0, st>>>(a);
  c->launch_check("k_bounce");
  c->bounce_kernel_ms += tk.stop();
  c->bounce_steps++;
  std::swap(c->rad_cur, c->rad_nxt);   // rad_cur now holds the new radiosity (this rank's slots)
}

void bounce_end(qrad_cuda_ctx *c) {
  cudaStream_t st = c->stream;
  k_finish_bounce<<<(c->N + 255) / 256, 256, 0, st>>>(c->N, c->patch_slot.p, c->s_total.p, c->rad_cur, c->p_totallight.p,
                                                      c->d_radiosity.p);
  c->launch_check("k_finish_bounce");
  c->sync("bounce");
  c->stats.ms_bounce_kernel = c->bounce_steps > 0 ? c->bounce_kernel_ms / c->bounce_steps : 0;
  c->stats.ms_bounce = c->t_bounce.stop();
}

void bounce_impl(qrad_cuda_ctx *c, int numbounce, float *added_out) {
  if (c->world > 1) dfail("bounce: this context is a shard (%d/%d); use qrad_cuda_bounce_phase and exchange", c->rank, c->world);
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
}

}  // namespace qrad
