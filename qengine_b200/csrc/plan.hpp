// Geometry of the transfer matrix (MakeTransfers, qrad3.c:185-293) on one or several GPUs: pure host code, no CUDA,
// replicated on every rank and identical there.  Everything the kernels of transfers.cu index with comes from here,
// and the same object drives the exchanges between GPUs, so the CPU tests (tests/test_multi_cpu.py) exercise the
// product's own addressing through the C entry points at the bottom of plan.cpp.
//
// Slots:     patches grouped by the PVS cluster of their origin, every cluster padded to a multiple of 32 slots.
//            A "slice" is 32 consecutive slots (always inside one cluster).
// Word:      accept bits of ONE source patch towards the 32 receivers of one slice.
// Row group: up to 128 consecutive slots (4 slices) of one cluster = the unit of source-row ownership.  The rows a
//            rank owns inside cluster c are numbered compactly ("own rows" 0 .. nown[q][c]).
// Storage on rank q (its "row buffer"): for every slice s, in order, one column of words holding, for every source
//            cluster c that can see s's cluster (ascending c), the own rows of c:
//                 word(source row, s) = rowbuf[ sbase[q][s] + coff[q](d, c) + ownrow ]          d = cluster of s
//            so the 32 consecutive rows a warp handles are 32 consecutive words, and everything rank q knows about
//            the slices [lo, hi) of another rank is ONE contiguous piece of its buffer: the exchange before pass 3 is
//            an all-to-all of contiguous pieces, with no packing.
// Receiver slices are owned in contiguous runs [slice_lo[r], slice_hi[r]) balanced by candidate sources; row groups
// are dealt to the ranks in many small contiguous chunks each, balanced by candidate words.
#pragma once
#include <cstdint>
#include <memory>
#include <utility>
#include <vector>

namespace qrad {

// std::vector whose resize() leaves new elements uninitialised (they are written, in parallel, right afterwards)
template <class T> struct UninitAlloc : std::allocator<T> {
  template <class U> struct rebind { using other = UninitAlloc<U>; };
  using std::allocator<T>::allocator;
  template <class U> void construct(U *p) noexcept { ::new (static_cast<void *>(p)) U; }
  template <class U, class... A> void construct(U *p, A &&...a) { ::new (static_cast<void *>(p)) U(std::forward<A>(a)...); }
};
template <class T> using UninitVec = std::vector<T, UninitAlloc<T>>;

constexpr int kGroupRows = 128;   // rows of a row group (= consecutive tasks of one warp run in k_visibility)
constexpr int kChunkWords = 32;   // one task = one source patch x up to 32 words (1024 receivers)

struct TransferPlan {
  int N = 0, nc = 0, M = 0, nslices = 0, world = 1, rank = 0;
  // ---- slots ----
  std::vector<int32_t> cl_count, cl_slot_start;          // [nc], [nc + 1]
  std::vector<int32_t> slot_patch, patch_slot, slot_cluster;   // [M] patch or -1, [N] slot or -1, [M]
  std::vector<int32_t> slice_count;                      // [nslices] valid receivers of a slice (1 .. 32)
  // ---- which clusters see which (only clusters that hold patches) ----
  std::vector<int32_t> vis_ptr, vis_cluster, vis_woff, vis_seer;   // source a -> receivers d ascending; word offset of d in
                                                                   // a row of a; index of (a, d) in the seer arrays
  std::vector<int32_t> nwords;                           // [nc] words of a row of source cluster a
  std::vector<int64_t> wb_first;                         // [nc + 1] prefix of nwords: index of word 0 of cluster a in the word tables
  std::vector<int32_t> seer_ptr, seer_cluster;           // receiver d -> sources c ascending that see d
  // ---- row groups ----
  std::vector<int32_t> grp_first;                        // [nc + 1] first group of a cluster
  std::vector<int32_t> grp_cluster, grp_slot0, grp_rows, grp_owner, grp_ownbase;
  std::vector<int32_t> nown;                             // [world][nc] own rows (multiple of 32)
  UninitVec<int32_t> coff;                               // [world][seer entries] offset of c's own rows in a column of d
  std::vector<int32_t> colh;                             // [world][nc] words of one slice column of cluster d
  UninitVec<int64_t> sbase;                              // [world][nslices + 1] first word of slice s in rank q's buffer
  // ---- receiver slices ----
  std::vector<int32_t> slice_lo, slice_hi;               // [world]
  // ---- candidate counts ----
  std::vector<int64_t> Lptr;   // [nc + 1] receivers a source of cluster c must look at (row lists), all clusters
  std::vector<int64_t> Rptr;   // [nc + 1] sources that can see cluster d (column lists), all clusters
  // ---- this rank ----
  std::vector<int32_t> own_groups;                       // groups of this rank, ascending
  std::vector<int64_t> own_runbase;                      // [own groups + 1] prefix of chunks-per-row: run r of pass 1 = (group, chunk)
  std::vector<int32_t> own_rowslice_slot0, own_rowslice_row0;   // the 32-row slices of the own groups: first slot, first own row
  std::vector<int64_t> Lown_ptr;                         // [nc + 1] row-list storage of this rank (clusters with own groups only)
  std::vector<int64_t> Rown_ptr;                         // [nc + 1] column-list storage (clusters with own receiver slices only)
  int64_t rowbuf_words() const { return sbase[(size_t) rank * (nslices + 1) + nslices]; }
  int64_t words_for(int q, int r) const {                // words rank q holds about the receiver slices of rank r
    const int64_t *sb = &sbase[(size_t) q * (nslices + 1)];
    return sb[slice_hi[r]] - sb[slice_lo[r]];
  }
  int group_of_slot(int slot) const {
    const int c = slot_cluster[slot];
    return grp_first[c] + (slot - cl_slot_start[c]) / kGroupRows;
  }
  // where rank `q`'s buffer keeps the words of source `slot` inside a column of receiver cluster d (relative to sbase)
  int64_t column_offset(int q, int d, int slot) const;
};

// eff_cluster[i]: cluster of patch i (or -1); pvs: nc rows of pvs_words 32-bit words (bit d of row a: d visible from a),
// nullptr = everything sees everything (-nopvs uses one pseudo cluster).  owner_chunks: contiguous chunks of row groups
// per rank (0 = default).
void build_plan(TransferPlan &p, int N, const int32_t *eff_cluster, int nc, const uint32_t *pvs, int pvs_words, int world,
                int rank, int owner_chunks);

}  // namespace qrad
