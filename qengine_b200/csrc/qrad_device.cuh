// Device-side core of libqrad_cuda: context, buffers and the two BSP walkers every
// kernel uses.  Compiled for sm_100a only, with -fmad=false so that every float and
// double operation is a separately rounded IEEE op exactly as in the reference's
// x86-64 SSE build (finding F6: FMA contraction changes bytes).
#pragma once
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <future>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <nccl.h>

#include "../../include/qrad_cuda.h"
#include "plan.hpp"

namespace qrad {

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
[[noreturn]] inline void dfail(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw CudaError(buf);
}
#define QCUDA(expr)                                                                                 \
  do {                                                                                              \
    cudaError_t e__ = (expr);                                                                       \
    if (e__ != cudaSuccess) ::qrad::dfail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// Host <-> device traffic of this thread's contexts, counted where the copies are issued (qrad_cuda_stats.h2d_bytes /
// d2h_bytes report the totals since the context was created).
// where the runs of one own row group go in k_visibility's run list (transfers.cu: k_run_list)
struct RunGroup {
  int grp, nchunks, rows, self_lo, self_hi;
  int64_t heavy_at, light_at;
};

struct Traffic {
  int64_t h2d = 0, d2h = 0;
};
inline Traffic &traffic() {
  static thread_local Traffic t;
  return t;
}

// Device buffer.  alloc() keeps the allocation when it is already large enough: a host that lights map after map
// (or a benchmark that repeats a step) does not pay cudaMalloc/cudaFree of multi-GB buffers every time.
template <class T> struct DBuf {
  T *p = nullptr;
  size_t n = 0, cap = 0;
  DBuf() = default;
  DBuf(const DBuf &) = delete;
  DBuf &operator=(const DBuf &) = delete;
  ~DBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = cap = 0;
  }
  void alloc(size_t count) {
    if (count > cap || (cap > (size_t) 1 << 20 && count < cap / 4)) {
      release();
      if (count) {
        QCUDA(cudaMalloc(&p, count * sizeof(T)));
        cap = count;
      }
    }
    n = count;
  }
  void zero(cudaStream_t s) {
    if (n) QCUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
  }
  void upload(const T *h, size_t count, cudaStream_t s) {
    alloc(count);
    if (count) QCUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
    traffic().h2d += (int64_t) (count * sizeof(T));
  }
  void upload(const std::vector<T> &h, cudaStream_t s) { upload(h.data(), h.size(), s); }
  void download(T *h, size_t count, cudaStream_t s) const {
    if (count) QCUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
    traffic().d2h += (int64_t) (count * sizeof(T));
    QCUDA(cudaStreamSynchronize(s));
  }
};

}  // namespace qrad
#include "walkers.cuh"
namespace qrad {

#ifdef __CUDACC__
// VectorNormalize, mathlib.c:106-122.  (float)sqrt((double)x) == sqrtf(x) and (float)(1.0/(double)x) == 1.0f/x
// for every float x (53 >= 2*24+2), so the float intrinsics reproduce the reference's double round trip.
__device__ __forceinline__ float vnormalize(float &x, float &y, float &z) {
  const float len = __fsqrt_rn(x * x + y * y + z * z);
  if (len == 0) {
    x = y = z = 0;
    return 0;
  }
  const float il = __fdiv_rn(1.0f, len);
  x *= il;
  y *= il;
  z *= il;
  return len;
}
#endif  // __CUDACC__

// QRAD_DEBUG_TIMING=1: wall-clock laps of a host function, the stream drained at every lap (tools only: it serialises)
struct DebugLaps {
  bool on = getenv("QRAD_DEBUG_TIMING") != nullptr;
  const char *who;
  cudaStream_t st;
  std::chrono::steady_clock::time_point tp;
  DebugLaps(const char *who_, cudaStream_t st_) : who(who_), st(st_) {
    if (on) {
      cudaStreamSynchronize(st);
      tp = std::chrono::steady_clock::now();
    }
  }
  void lap(const char *what) {
    if (!on) return;
    cudaStreamSynchronize(st);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "%s: %-28s %7.2f ms\n", who, what, std::chrono::duration<double, std::milli>(now - tp).count());
    tp = now;
  }
};

struct Timer {
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t s = nullptr;
  void start(cudaStream_t st) {
    s = st;
    if (!a) {
      QCUDA(cudaEventCreate(&a));
      QCUDA(cudaEventCreate(&b));
    }
    QCUDA(cudaEventRecord(a, s));
  }
  float stop() {
    QCUDA(cudaEventRecord(b, s));
    QCUDA(cudaEventSynchronize(b));
    float ms = 0;
    QCUDA(cudaEventElapsedTime(&ms, a, b));
    return ms;
  }
  ~Timer() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
  }
};

}  // namespace qrad

// The context: everything libqrad_cuda keeps on one GPU between the RadWorld seams.
struct qrad_cuda_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  int rank = 0, world = 1;
  bool count_nodes = false;
  qrad_cuda_stats stats{};
  qrad::Traffic traffic0;              // thread counters when the context was created

  // ---- BSP (qrad_cuda_upload_bsp) ----
  int numnodes = 0, numleafs = 0, numclusters = 0, tree_depth = 0, pvs_words = 0;
  bool have_vis = false;
  qrad::DBuf<int4> nodes;
  qrad::DBuf<float4> node_normals;
  qrad::DBuf<float4> leaf_cell_lo, leaf_cell_hi;   // [numleafs] box cells of the leaves, see shaft_blocked()
  float bsp_extent = 0, patch_extent = 0;           // largest |coordinate| of the axial planes / patch origins
  qrad::Tree tree() const { return qrad::Tree{nodes.p, node_normals.p}; }
  qrad::DBuf<int32_t> leaf_contents, leaf_cluster;
  qrad::DBuf<uint32_t> pvs;            // [numclusters][pvs_words], bit d of row c = cluster d visible from c
  std::vector<uint32_t> h_pvs;

  // ---- patches, original order (qrad_cuda_upload_patches) ----
  int N = 0, numfaces_p = 0;
  std::vector<int32_t> h_cluster;
  std::vector<uint8_t> h_sky;
  std::vector<float4> h_origin_area, h_normal;
  qrad::DBuf<float4> p_origin_area;    // xyz = origin, w = area
  qrad::DBuf<float4> p_normal;         // xyz = plane normal
  qrad::DBuf<float4> p_refl;           // xyz = reflectivity, w = sky flag
  qrad::DBuf<float> p_baselight, p_totallight, p_samplelight, p_bounds;  // [N][3], [N][3], [N][3], [N][6]
  qrad::DBuf<int32_t> p_samples, p_face, p_next, face_head, p_cluster;

  // ---- lights ----
  int numlights = 0, numstyles_map = 1;
  std::vector<int32_t> h_style_ids;    // distinct light styles ascending, [0] == 0
  qrad::DBuf<int32_t> light_first, l_type, l_styleslot, l_cluster;
  qrad::DBuf<float4> l_origin_int;     // xyz origin, w intensity
  qrad::DBuf<float4> l_color;          // xyz color, w stopdot
  qrad::DBuf<float4> l_normal;

  // ---- faces ----
  int F = 0, num_luxels = 0;
  std::vector<uint8_t> h_lit;
  std::vector<int32_t> h_luxel_first, h_texsize, h_planelink_head, h_facelinks;
  std::vector<float> h_face_bounds, h_minlight;
  qrad::DBuf<uint8_t> f_lit;
  qrad::DBuf<float> f_normal, f_texorg, f_textoworld, f_exactmins, f_exactmaxs, f_bounds, f_minlight, f_plane_normal;
  qrad::DBuf<int32_t> f_texmins, f_texsize, f_luxel_first, luxel_face;

  // ---- facelight results (qrad_cuda_build_facelights) ----
  int fl_numsamples = 1;
  qrad::DBuf<float> fl_surfpt;         // [numsamples][L][3]; sample 0 = facelight origins
  qrad::DBuf<float> fl_samples;        // [numstyles_map][L][3]
  qrad::DBuf<uint32_t> fl_touched;     // [F][numstyles_map] flag: style table exists for the face
  std::vector<int32_t> h_fl_numstyles, h_fl_stylenums;  // [F], [F][32] (slot indices into h_style_ids)
  bool facelights_done = false;

  // ---- transfer matrix (qrad_cuda_make_transfers): geometry in plan.hpp ----
  qrad::TransferPlan plan;
  // the plan only needs the patch clusters and the PVS: upload_patches starts it on a host thread, so that it is ready
  // when make_transfers asks for it (it runs beside the other uploads and BuildFacelights)
  std::future<std::shared_ptr<qrad::TransferPlan>> plan_future;
  int M = 0;                            // number of slots (= plan.M)
  qrad::DBuf<int32_t> slot_patch, patch_slot, slot_cluster, slice_count, d_cl_slot_start, d_cl_count, d_nwords;
  qrad::DBuf<uint8_t> slot_flags;
  qrad::DBuf<int64_t> d_wb_first, d_sbase, wbase;
  qrad::DBuf<int32_t> wslice, d_vis_ptr, d_vis_cluster, d_vis_woff, d_vis_seer, d_vis_src, d_coff;
  qrad::DBuf<int32_t> d_grp_first, d_grp_cluster, d_grp_slot0, d_grp_rows, d_grp_owner, d_grp_ownbase;
  qrad::DBuf<int32_t> d_rowslice_slot0, d_rowslice_row0;
  qrad::DBuf<float4> s_origin_area, s_normal, wbox_min, wbox_max;
  qrad::DBuf<uint32_t> bits;            // this rank's rows of the visibility*formfactor>0 bit matrix, slice-major
  qrad::DBuf<uint32_t> colbuf;          // several GPUs: the other ranks' words about this rank's slices
  std::vector<int64_t> piece_first;     // [world] where rank q's piece starts in colbuf
  qrad::DBuf<int2> d_runs;              // the patch order as runs of consecutive slots
  qrad::DBuf<qrad::RunGroup> d_run_groups;
  qrad::DBuf<int4> d_run_list;          // pass 1: this rank's (row group, chunk, first row, rows) runs in the order the warps draw them
  int64_t num_runs = 0;
  qrad::DBuf<int64_t> d_vl_ptr, d_vr_ptr;
  qrad::DBuf<int4> vl, vr;              // row / column run lists (transfers.cu)
  qrad::DBuf<float> row_total;          // [M]
  qrad::DBuf<int32_t> row_count;        // [M] numtransfers of the source in slot s
  // gather-form matrix, "dense slices": see k_columns / k_bounce in transfers.cu
  int64_t sell_entries = 0;             // list entries over all slices (each slice padded to a multiple of 32)
  qrad::DBuf<int32_t> slice_len;        // [M/32] list entries of a slice
  qrad::DBuf<int64_t> slice_base;       // [M/32 + 1] first entry of a slice
  qrad::DBuf<uint32_t> g_src;           // [entries] source slot
  qrad::DBuf<uint4> g_w;                // [entries/8][32] 8 u16 weights per lane
  qrad::DBuf<int32_t> slice_order;      // this rank's slices, longest list first (k_bounce work order)
  qrad::DBuf<int32_t> col_order;        // all of this rank's slices, longest list first (k_columns: one block each)
  qrad::DBuf<int32_t> coop_slices;      // the slices with the longest lists: work list of the block-cooperative role
  int n_coop = 0, n_regular = 0;
  qrad::DBuf<int32_t> bounce_next;      // [4] work counters of the bounce kernels
  qrad::DBuf<int32_t> negative_light;   // CheckPatches flag
  int bounce_blocks_per_sm = 0, bounce_regs_blocks_per_sm = 0, shard_blocks_per_sm = 0;
  bool transfers_done = false;
  int64_t total_transfers = 0;
  qrad::DBuf<unsigned long long> counters;
  qrad::DBuf<float> rad_sum;
  qrad::Timer t_transfers, t_bounce;
  float4 *rad_cur = nullptr, *rad_nxt = nullptr;
  int bounce_steps = 0;
  std::vector<cudaEvent_t> bounce_ev;   // two per bounce: around the kernel launch
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // pass 1 begin / end, matrix exchange begin / end
  // ---- several GPUs ----
  ncclComm_t comm = nullptr;
  cudaStream_t xstream = nullptr;       // the matrix exchange runs beside the lists and pass 2
  cudaStream_t lstream = nullptr;       // the lists of passes 2 and 3 are prepared beside pass 1
  cudaEvent_t ev_lists = nullptr;
  qrad::DBuf<int32_t> lists_i32[4];
  qrad::DBuf<float4> share_staging;     // share_slices: `world` equal parts for the all-gather of a per-slot array
  int face_lo = 0, face_hi = 0;         // faces this rank lights (all of them on one GPU)

  // ---- plane groups of FinalLightFace: for every lit face the patches of all faces on its plane chain, in chain order
  // (lightmap.c:1123-1134).  Host work that depends on the uploads only: upload_faces starts it on a host thread. ----
  struct PlaneGroups {
    std::vector<int32_t> group_of_face, group_first, group_patch;
    std::string error;
  };
  std::vector<int32_t> h_next, h_face_head;        // patch->next, face_patches[] as uploaded
  std::future<std::shared_ptr<PlaneGroups>> groups_future;

  // ---- FinalLightFace scratch (per resident warp) ----
  qrad::DBuf<uint16_t> fin_matrix, fin_ep0, fin_ep1, fin_tedges, fin_stack;
  qrad::DBuf<int16_t> fin_etri;
  qrad::DBuf<float4> fin_pts, fin_eplane;
  qrad::DBuf<int32_t> fin_i32[11];
  qrad::DBuf<int64_t> fin_i64;
  qrad::DBuf<uint8_t> fin_out;
  qrad::DBuf<float> fin_lb;

  // ---- bounce state (slot order) ----
  qrad::DBuf<float4> rad_a, rad_b;      // radiosity / 0x10000 ("send", qrad3.c:430-431), double buffered
  qrad::DBuf<float4> s_total;           // totallight per slot
  qrad::DBuf<float4> s_refl_area;       // xyz reflectivity, w = area (negative area marks sky)
  qrad::DBuf<float> d_radiosity;        // [N][3] radiosity in patch order (inspection)

  void launch_check(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) qrad::dfail("kernel launch %s failed: %s", what, cudaGetErrorString(e));
    stats.kernel_launches++;
  }
  void sync(const char *what) {
    cudaError_t e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) qrad::dfail("%s failed: %s", what, cudaGetErrorString(e));
  }
};

namespace qrad {
// implemented in the phase translation units
void face_range(qrad_cuda_ctx *c);
std::shared_ptr<qrad_cuda_ctx::PlaneGroups> compute_plane_groups(const qrad_cuda_ctx *c);
void build_facelights_impl(qrad_cuda_ctx *c, int extrasamples, int numbounce);
void make_transfers_impl(qrad_cuda_ctx *c, int nopvs);
void bounce_impl(qrad_cuda_ctx *c, int numbounce, float *added_out);
void transfers_trace(qrad_cuda_ctx *c, int nopvs);
void transfers_rows(qrad_cuda_ctx *c);
void transfers_columns(qrad_cuda_ctx *c);
void bounce_begin(qrad_cuda_ctx *c, bool want_sum);
void bounce_step(qrad_cuda_ctx *c);
void bounce_end(qrad_cuda_ctx *c);
void final_light_impl(qrad_cuda_ctx *c, const qrad_final_params *p, uint8_t *out, int32_t cap, int32_t *size,
                      int32_t *lightofs, uint8_t *styles);
void download_transfers_impl(qrad_cuda_ctx *c, int64_t *row_ptr, uint32_t *patch, uint16_t *transfer);
void expand_pvs(const uint8_t *visdata, int visdatasize, int &nc, int &pvs_words, std::vector<uint32_t> &out);
void download_rows_impl(qrad_cuda_ctx *c, int nrows, const int32_t *rows, int64_t *row_ptr, uint32_t *patch,
                        uint16_t *transfer, int64_t capacity, int64_t *rays);
void test_lines_impl(qrad_cuda_ctx *c, int64_t n, const float *starts, const float *stops, uint8_t *results);
void point_in_leaf_impl(qrad_cuda_ctx *c, int64_t n, const float *pts, int32_t *out);
}  // namespace qrad
