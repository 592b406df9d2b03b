// Host-side geometry of the transfer matrix and of its distribution over GPUs (see plan.hpp).  No CUDA in this file.
#include "plan.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>

#include "../../include/qrad_cuda.h"

namespace qrad {

namespace {
[[noreturn]] void pfail(const std::string &m) { throw std::runtime_error(m); }
}  // namespace

// DecompressVis (bspfile.c:129-154) for every cluster at once: row c, bit d = cluster d is in the PVS of cluster c.
void expand_pvs(const uint8_t *visdata, int visdatasize, int &nc, int &pvs_words, std::vector<uint32_t> &out) {
  if (visdatasize < 4) pfail("visibility lump too small");
  int32_t n;
  memcpy(&n, visdata, 4);
  if (n < 0 || n > 65536 || (int64_t) 4 + (int64_t) n * 8 > visdatasize) pfail("bad visibility header");
  nc = n;
  const int rowbytes = (nc + 7) >> 3;
  pvs_words = (nc + 31) >> 5;
  out.assign((size_t) nc * pvs_words, 0);
  std::vector<uint8_t> row((size_t) pvs_words * 4 + 8);
  for (int cl = 0; cl < nc; cl++) {
    int32_t ofs;
    memcpy(&ofs, visdata + 4 + (size_t) cl * 8, 4);  // bitofs[cl][DVIS_PVS]
    if (ofs < 0 || ofs >= visdatasize) pfail("PVS offset of cluster " + std::to_string(cl) + " out of range");
    const uint8_t *in = visdata + ofs, *end = visdata + visdatasize;
    std::fill(row.begin(), row.end(), 0);
    int o = 0;
    do {
      if (in >= end) pfail("PVS row of cluster " + std::to_string(cl) + " runs past the lump");
      if (*in) {
        row[o++] = *in++;
        continue;
      }
      if (in + 1 >= end) pfail("PVS row of cluster " + std::to_string(cl) + " runs past the lump");
      int rep = in[1];
      if (!rep) pfail("DecompressVis: 0 repeat");
      in += 2;
      while (rep && o < rowbytes) {
        row[o++] = 0;
        rep--;
      }
    } while (o < rowbytes);
    memcpy(&out[(size_t) cl * pvs_words], row.data(), (size_t) pvs_words * 4);
  }
}

int64_t TransferPlan::column_offset(int q, int d, int slot) const {
  const int c = slot_cluster[slot];
  const int32_t *b = &seer_cluster[seer_ptr[d]], *e = &seer_cluster[seer_ptr[d + 1]];
  const int32_t *it = std::lower_bound(b, e, c);
  if (it == e || *it != c) return -1;
  const int g = group_of_slot(slot);
  if (grp_owner[g] != q) return -1;
  const size_t entry = (size_t) (it - seer_cluster.data());
  return (int64_t) coff[(size_t) q * seer_cluster.size() + entry] + grp_ownbase[g] + (slot - grp_slot0[g]);
}

void build_plan(TransferPlan &p, int N, const int32_t *eff, int nc, const uint32_t *pvs, int pvs_words, int world, int rank,
                int owner_chunks) {
  if (world < 1 || world > 64 || rank < 0 || rank >= world) pfail("bad rank / world");
  if (nc <= 0) pfail("make_transfers: no clusters");
  const bool timing = getenv("QRAD_DEBUG_TIMING") != nullptr;
  auto tp = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!timing) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "  plan: %-22s %7.2f ms\n", what, std::chrono::duration<double, std::milli>(now - tp).count());
    tp = now;
  };
  p = TransferPlan();
  p.N = N;
  p.nc = nc;
  p.world = world;
  p.rank = rank;
  // Both counting sorts below (patches by cluster, visible pairs by receiver) run on T threads: thread t takes a
  // contiguous range of the input, counts into a histogram of its own, and after a prefix over (bucket, thread) writes its
  // items behind those of the threads before it -- the order inside a bucket stays the input order.
  int T = (int) std::thread::hardware_concurrency() / std::max(world, 1);
  T = std::max(1, std::min(T, 8));
  if (const char *e = getenv("QRAD_PLAN_THREADS")) T = std::max(1, std::min(64, atoi(e)));
  auto run_threads = [&](auto &&fn) {
    if (T == 1) {
      fn(0);
      return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) th.emplace_back([&fn, t] { fn(t); });
    for (auto &x : th) x.join();
  };
  // ---- slots ----
  std::vector<int32_t> hist((size_t) T * nc, 0);
  std::vector<int> bad_patch((size_t) T, -1);
  run_threads([&](int t) {
    int32_t *h = &hist[(size_t) t * nc];
    const int i0 = (int) ((int64_t) N * t / T), i1 = (int) ((int64_t) N * (t + 1) / T);
    for (int i = i0; i < i1; i++) {
      if (eff[i] >= nc) {
        if (bad_patch[t] < 0) bad_patch[t] = i;
      } else if (eff[i] >= 0)
        h[eff[i]]++;
    }
  });
  for (int t = 0; t < T; t++)
    if (bad_patch[t] >= 0)
      pfail("patch " + std::to_string(bad_patch[t]) + " has cluster " + std::to_string(eff[bad_patch[t]]) + " >= numclusters");
  p.cl_count.assign(nc, 0);
  p.cl_slot_start.assign(nc + 1, 0);
  for (int k = 0; k < nc; k++) {
    int32_t n = 0;
    for (int t = 0; t < T; t++) n += hist[(size_t) t * nc + k];
    p.cl_count[k] = n;
    p.cl_slot_start[k + 1] = p.cl_slot_start[k] + ((n + 31) & ~31);
    int32_t at = p.cl_slot_start[k];   // hist becomes: first slot of thread t in cluster k
    for (int t = 0; t < T; t++) {
      const int32_t c = hist[(size_t) t * nc + k];
      hist[(size_t) t * nc + k] = at;
      at += c;
    }
  }
  const int M = p.cl_slot_start[nc];
  p.M = M;
  p.nslices = M / 32;
  if (M == 0) pfail("make_transfers: no patch lies in a valid cluster");
  p.slot_patch.assign(M, -1);
  p.patch_slot.assign(N, -1);
  p.slot_cluster.assign(M, 0);
  p.slice_count.assign(p.nslices, 0);
  run_threads([&](int t) {
    int32_t *fill = &hist[(size_t) t * nc];
    const int i0 = (int) ((int64_t) N * t / T), i1 = (int) ((int64_t) N * (t + 1) / T);
    for (int i = i0; i < i1; i++)
      if (eff[i] >= 0) {
        const int s = fill[eff[i]]++;
        p.slot_patch[s] = i;
        p.patch_slot[i] = s;
      }
    const int k0 = (int) ((int64_t) nc * t / T), k1 = (int) ((int64_t) nc * (t + 1) / T);
    for (int k = k0; k < k1; k++) {
      for (int s = p.cl_slot_start[k]; s < p.cl_slot_start[k + 1]; s++) p.slot_cluster[s] = k;
      for (int s = p.cl_slot_start[k] >> 5; s < p.cl_slot_start[k + 1] >> 5; s++)
        p.slice_count[s] = std::min(32, p.cl_count[k] - (s * 32 - p.cl_slot_start[k]));
    }
  });
  lap("slots");
  // ---- who sees whom ----
  p.vis_ptr.assign(nc + 1, 0);
  p.nwords.assign(nc, 0);
  p.wb_first.assign(nc + 1, 0);
  p.Lptr.assign(nc + 1, 0);
  p.Rptr.assign(nc + 1, 0);
  // clusters that hold patches, as a bit row: a visible list is (PVS row & this) scanned word by word
  const int nw32 = (nc + 31) >> 5;
  std::vector<uint32_t> populated((size_t) nw32, 0);
  for (int d = 0; d < nc; d++)
    if (p.cl_count[d] > 0) populated[d >> 5] |= 1u << (d & 31);
  auto row_bits = [&](int a, int k) -> uint32_t { return (pvs ? pvs[(size_t) a * pvs_words + k] : 0xffffffffu) & populated[k]; };
  for (int a = 0; a < nc; a++) {
    int n = 0;
    if (p.cl_count[a] > 0)
      for (int k = 0; k < nw32; k++) n += __builtin_popcount(row_bits(a, k));
    p.vis_ptr[a + 1] = p.vis_ptr[a] + n;
  }
  const size_t E = (size_t) p.vis_ptr[nc];
  p.vis_cluster.resize(E);
  p.vis_woff.resize(E);
  p.seer_cluster.assign(E, 0);
  p.vis_seer.assign(E, 0);
  std::vector<int32_t> seen((size_t) T * nc, 0);      // [t][d] sources of thread t's range that see d
  std::vector<int64_t> rsum((size_t) T * nc, 0);      // [t][d] their patches
  std::vector<int64_t> cand((size_t) nc, 0);
  run_threads([&](int t) {
    const int a0 = (int) ((int64_t) nc * t / T), a1 = (int) ((int64_t) nc * (t + 1) / T);
    int32_t *sn = &seen[(size_t) t * nc];
    int64_t *rs = &rsum[(size_t) t * nc];
    for (int a = a0; a < a1; a++) {
      if (p.cl_count[a] == 0) continue;
      int w = 0, v = p.vis_ptr[a];
      int64_t cnd = 0;
      for (int k = 0; k < nw32; k++) {
        uint32_t bits = row_bits(a, k);
        while (bits) {
          const int d = k * 32 + __builtin_ctz(bits);
          bits &= bits - 1;
          p.vis_cluster[v] = d;
          p.vis_woff[v] = w;
          v++;
          w += (p.cl_slot_start[d + 1] - p.cl_slot_start[d]) >> 5;
          cnd += p.cl_count[d];
          rs[d] += p.cl_count[a];
          sn[d]++;
        }
      }
      p.nwords[a] = w;
      cand[a] = cnd;
    }
  });
  p.seer_ptr.assign(nc + 1, 0);
  for (int a = 0; a < nc; a++) {
    p.wb_first[a + 1] = p.wb_first[a] + p.nwords[a];
    p.Lptr[a + 1] = p.Lptr[a] + cand[a];
  }
  for (int d = 0; d < nc; d++) {
    int32_t at = p.seer_ptr[d];
    int64_t r = 0;
    for (int t = 0; t < T; t++) {   // seen becomes: where thread t's first seer of d goes
      const int32_t c = seen[(size_t) t * nc + d];
      seen[(size_t) t * nc + d] = at;
      at += c;
      r += rsum[(size_t) t * nc + d];
    }
    p.seer_ptr[d + 1] = at;
    p.Rptr[d + 1] = p.Rptr[d] + r;
  }
  if (p.Lptr[nc] >= ((int64_t) 1 << 40)) pfail("make_transfers: candidate lists too large");
  run_threads([&](int t) {   // a ascending inside a thread, threads in order: every seer list comes out sorted
    const int a0 = (int) ((int64_t) nc * t / T), a1 = (int) ((int64_t) nc * (t + 1) / T);
    int32_t *fill = &seen[(size_t) t * nc];
    for (int a = a0; a < a1; a++)
      for (int v = p.vis_ptr[a]; v < p.vis_ptr[a + 1]; v++) {
        const int e = fill[p.vis_cluster[v]]++;
        p.seer_cluster[e] = a;
        p.vis_seer[v] = e;
      }
  });
  lap("visibility lists");
  // ---- row groups and their owners ----
  p.grp_first.assign(nc + 1, 0);
  for (int c = 0; c < nc; c++) {
    const int padded = p.cl_slot_start[c + 1] - p.cl_slot_start[c];
    const int ng = (padded + kGroupRows - 1) / kGroupRows;
    p.grp_first[c + 1] = p.grp_first[c] + ng;
    for (int g = 0; g < ng; g++) {
      p.grp_cluster.push_back(c);
      p.grp_slot0.push_back(p.cl_slot_start[c] + g * kGroupRows);
      p.grp_rows.push_back(std::min(kGroupRows, padded - g * kGroupRows));
    }
  }
  const int G = (int) p.grp_cluster.size();
  p.grp_owner.assign(G, 0);
  if (world > 1) {
    // dealt in `chunks` contiguous chunks per rank of (nearly) equal candidate words, shuffled among the ranks.  The
    // cost of a candidate word varies over the map (it is the clear rays inside a room that are walked), so the chunks
    // are small -- a couple of row groups: measured on C5 at 8 ranks (tools/trace_balance.py) the slowest rank is 12-14 %
    // above the mean with 8-128 chunks per rank and 4.7 % with 512.  Every rank then owns rows in most clusters; its
    // row lists cover them all, which costs a few milliseconds.
    int chunks = owner_chunks > 0 ? owner_chunks : 512;
    if (const char *e = getenv("QRAD_OWNER_CHUNKS")) chunks = std::max(1, atoi(e));
    double total = 0;
    for (int g = 0; g < G; g++) total += (double) p.grp_rows[g] * p.nwords[p.grp_cluster[g]];
    const int nchunks = world * chunks;
    const double per = total > 0 ? total / nchunks : 1;
    double cum = 0;
    for (int g = 0; g < G; g++) {
      const double w = (double) p.grp_rows[g] * p.nwords[p.grp_cluster[g]];
      int ch = (int) ((cum + 0.5 * w) / per);
      ch = std::max(0, std::min(nchunks - 1, ch));
      // every `world` consecutive chunks go to the `world` ranks in a pseudo-random order of their own (a rotation would
      // alias with the 3-4 row groups a cluster has: the first group of a cluster is not like its last)
      const int block = ch / world, pos = ch % world;
      uint32_t h = (uint32_t) block * 2654435761u + 0x9e3779b9u;
      int perm[64];
      for (int k = 0; k < world; k++) perm[k] = k;
      for (int k = world - 1; k > 0; k--) {   // Fisher-Yates with an LCG
        h = h * 1664525u + 1013904223u;
        const int j = (int) ((h >> 8) % (uint32_t) (k + 1));
        std::swap(perm[k], perm[j]);
      }
      p.grp_owner[g] = perm[pos];
      cum += w;
    }
  }
  p.nown.assign((size_t) world * nc, 0);
  p.grp_ownbase.assign(G, 0);
  for (int g = 0; g < G; g++) {
    int32_t &n = p.nown[(size_t) p.grp_owner[g] * nc + p.grp_cluster[g]];
    p.grp_ownbase[g] = n;
    n += p.grp_rows[g];
  }
  p.coff.resize((size_t) world * E);
  p.sbase.resize((size_t) world * (p.nslices + 1));
  p.colh.assign((size_t) world * nc, 0);
  {
    std::vector<std::thread> th;
    std::vector<int> bad((size_t) world, 0);
    auto one = [&](int q) {   // the columns of rank q
      for (int d = 0; d < nc; d++) {
        int64_t acc = 0;
        for (int e = p.seer_ptr[d]; e < p.seer_ptr[d + 1]; e++) {
          p.coff[(size_t) q * E + e] = (int32_t) acc;
          acc += p.nown[(size_t) q * nc + p.seer_cluster[e]];
        }
        if (acc > 0x7fffffff) bad[q] = 1;
        p.colh[(size_t) q * nc + d] = (int32_t) acc;
      }
      int64_t *sb = &p.sbase[(size_t) q * (p.nslices + 1)];
      sb[0] = 0;
      for (int s = 0; s < p.nslices; s++) sb[s + 1] = sb[s] + p.colh[(size_t) q * nc + p.slot_cluster[(size_t) s * 32]];
    };
    for (int q = 1; q < world; q++) th.emplace_back(one, q);
    one(0);
    for (auto &t : th) t.join();
    for (int q = 0; q < world; q++)
      if (bad[q]) pfail("make_transfers: a slice column exceeds 2^31 words");
  }
  lap("groups, column offsets");
  // ---- receiver slices: contiguous runs of (nearly) equal candidate sources ----
  p.slice_lo.assign(world, 0);
  p.slice_hi.assign(world, 0);
  {
    double total = 0;
    for (int s = 0; s < p.nslices; s++) {
      const int d = p.slot_cluster[(size_t) s * 32];
      total += (double) (p.Rptr[d + 1] - p.Rptr[d]);
    }
    double cum = 0;
    int r = 0;
    for (int s = 0; s < p.nslices; s++) {
      const int d = p.slot_cluster[(size_t) s * 32];
      const double w = (double) (p.Rptr[d + 1] - p.Rptr[d]);
      while (r < world - 1 && cum + 0.5 * w >= total * (r + 1) / world) {
        p.slice_hi[r] = s;
        p.slice_lo[++r] = s;
      }
      cum += w;
    }
    while (r < world - 1) {
      p.slice_hi[r] = p.nslices;
      p.slice_lo[++r] = p.nslices;
    }
    p.slice_hi[world - 1] = p.nslices;
  }
  lap("slices");
  // ---- this rank ----
  p.own_runbase.assign(1, 0);
  for (int g = 0; g < G; g++) {
    if (p.grp_owner[g] != rank) continue;
    const int c = p.grp_cluster[g];
    p.own_groups.push_back(g);
    p.own_runbase.push_back(p.own_runbase.back() + (p.nwords[c] + kChunkWords - 1) / kChunkWords);
    for (int r = 0; r < p.grp_rows[g]; r += 32) {
      p.own_rowslice_slot0.push_back(p.grp_slot0[g] + r);
      p.own_rowslice_row0.push_back(p.grp_ownbase[g] + r);
    }
  }
  p.Lown_ptr.assign(nc + 1, 0);
  p.Rown_ptr.assign(nc + 1, 0);
  for (int c = 0; c < nc; c++) {
    const bool rows_here = p.nown[(size_t) rank * nc + c] > 0;
    p.Lown_ptr[c + 1] = p.Lown_ptr[c] + (rows_here ? p.Lptr[c + 1] - p.Lptr[c] : 0);
    const int s0 = p.cl_slot_start[c] >> 5, s1 = p.cl_slot_start[c + 1] >> 5;
    const bool slices_here = std::max(s0, p.slice_lo[rank]) < std::min(s1, p.slice_hi[rank]);
    p.Rown_ptr[c + 1] = p.Rown_ptr[c] + (slices_here ? p.Rptr[c + 1] - p.Rptr[c] : 0);
  }
}

}  // namespace qrad

// ---------------------------------------------------------------------------------------------------------------------
// C entry points (include/qrad_cuda.h, "planning"): the same plan the CUDA half builds, without a GPU.
struct qrad_plan {
  qrad::TransferPlan p;
};

static thread_local std::string g_plan_error;
extern "C" {

const char *qrad_plan_last_error(void) { return g_plan_error.c_str(); }

int qrad_plan_create(int32_t num_patches, const int32_t *cluster, const uint8_t *visdata, int32_t visdatasize, int nopvs,
                     int world, int rank, qrad_plan **out) {
  *out = nullptr;
  try {
    int nc = 0, pw = 0;
    std::vector<uint32_t> pvs;
    qrad::expand_pvs(visdata, visdatasize, nc, pw, pvs);
    std::vector<int32_t> eff(cluster, cluster + num_patches);
    if (nopvs) std::fill(eff.begin(), eff.end(), 0);
    qrad_plan *h = new qrad_plan;
    try {
      qrad::build_plan(h->p, num_patches, eff.data(), nopvs ? 1 : nc, nopvs ? nullptr : pvs.data(), pw, world, rank, 0);
    } catch (...) {
      delete h;
      throw;
    }
    *out = h;
    return 0;
  } catch (const std::exception &e) {
    g_plan_error = e.what();
    return 1;
  }
}

void qrad_plan_destroy(qrad_plan *h) { delete h; }

// out[0] slots, [1] slices, [2] words of this rank's row buffer, [3] row groups, [4] own groups, [5] pass-1 runs of this
// rank, [6] clusters, [7] words of the whole matrix (sum over ranks)
int qrad_plan_info(const qrad_plan *h, int64_t *out) {
  const qrad::TransferPlan &p = h->p;
  out[0] = p.M;
  out[1] = p.nslices;
  out[2] = p.rowbuf_words();
  out[3] = (int64_t) p.grp_cluster.size();
  out[4] = (int64_t) p.own_groups.size();
  out[5] = p.own_runbase.back();
  out[6] = p.nc;
  int64_t all = 0;
  for (int q = 0; q < p.world; q++) all += p.sbase[(size_t) q * (p.nslices + 1) + p.nslices];
  out[7] = all;
  return 0;
}

int qrad_plan_slices(const qrad_plan *h, int32_t *lo, int32_t *hi) {
  for (int r = 0; r < h->p.world; r++) {
    lo[r] = h->p.slice_lo[r];
    hi[r] = h->p.slice_hi[r];
  }
  return 0;
}

// per slot: owner rank of its row group; per patch: its slot (-1 = no valid cluster)
int qrad_plan_slots(const qrad_plan *h, int32_t *slot_owner /*[M]*/, int32_t *patch_slot /*[N]*/) {
  const qrad::TransferPlan &p = h->p;
  if (slot_owner)
    for (int s = 0; s < p.M; s++) slot_owner[s] = p.grp_owner[p.group_of_slot(s)];
  if (patch_slot) memcpy(patch_slot, p.patch_slot.data(), sizeof(int32_t) * (size_t) p.N);
  return 0;
}

// first word and number of words of the contiguous piece of rank q's row buffer that describes the receiver slices of rank r
int qrad_plan_piece(const qrad_plan *h, int q, int r, int64_t *first, int64_t *count) {
  const qrad::TransferPlan &p = h->p;
  if (q < 0 || q >= p.world || r < 0 || r >= p.world) return 1;
  *first = p.sbase[(size_t) q * (p.nslices + 1) + p.slice_lo[r]];
  *count = p.words_for(q, r);
  return 0;
}

// address, in the row buffer of the rank that owns source slot `source_slot`, of its word towards `slice`; -1 when the
// source's cluster cannot see the slice's cluster (no such word exists)
int64_t qrad_plan_word_address(const qrad_plan *h, int32_t source_slot, int32_t slice) {
  const qrad::TransferPlan &p = h->p;
  if (source_slot < 0 || source_slot >= p.M || slice < 0 || slice >= p.nslices) return -1;
  const int q = p.grp_owner[p.group_of_slot(source_slot)];
  const int64_t off = p.column_offset(q, p.slot_cluster[(size_t) slice * 32], source_slot);
  if (off < 0) return -1;
  return p.sbase[(size_t) q * (p.nslices + 1) + slice] + off;
}

// The column list of `slice` as pass 3 walks it: candidate sources in ascending patch index, each with the rank that
// owns its row and the address of its word for this slice inside THE PIECE that rank sends to the slice's owner.
int qrad_plan_column_list(const qrad_plan *h, int32_t slice, int32_t *slots, int32_t *owner, int64_t *addr_in_piece,
                          int64_t capacity, int64_t *count) {
  const qrad::TransferPlan &p = h->p;
  if (slice < 0 || slice >= p.nslices) return 1;
  const int d = p.slot_cluster[(size_t) slice * 32];
  int r = 0;
  while (r < p.world - 1 && slice >= p.slice_hi[r]) r++;
  int64_t n = 0;
  for (int i = 0; i < p.N; i++) {
    const int s = p.patch_slot[i];
    if (s < 0) continue;
    const int q = p.grp_owner[p.group_of_slot(s)];
    const int64_t off = p.column_offset(q, d, s);
    if (off < 0) continue;
    if (n < capacity) {
      slots[n] = s;
      owner[n] = q;
      const int64_t *sb = &p.sbase[(size_t) q * (p.nslices + 1)];
      addr_in_piece[n] = sb[slice] - sb[p.slice_lo[r]] + off;
    }
    n++;
  }
  *count = n;
  return 0;
}

}  // extern "C"
