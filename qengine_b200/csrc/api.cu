// extern "C" surface of the CUDA half (include/qrad_cuda.h): context, uploads,
// batched ray / point queries, inspection downloads.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "bsp_flat.hpp"
#include "comm.hpp"
#include "qrad_device.cuh"

using namespace qrad;

static thread_local std::string g_cuda_error;

template <class F> static int guarded(F &&f) {
  try {
    f();
    return 0;
  } catch (const std::exception &e) {
    g_cuda_error = e.what();
    return 1;
  }
}

// ---------------------------------------------------------------------------
// kernels for the query entry points
// ---------------------------------------------------------------------------
template <int STACK>
__global__ void k_test_lines(const Tree nodes, int64_t n, const float *__restrict__ starts,
                             const float *__restrict__ stops, uint8_t *__restrict__ out) {
  int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned v = 0;
  out[i] = (uint8_t) test_line<STACK, false>(nodes, starts[i * 3], starts[i * 3 + 1], starts[i * 3 + 2], stops[i * 3],
                                             stops[i * 3 + 1], stops[i * 3 + 2], v);
}

__global__ void k_point_in_leaf(const Tree nodes, int64_t n, const float *__restrict__ pts,
                                int32_t *__restrict__ out) {
  int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = point_in_leaf(nodes, pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]);
}

namespace qrad {

void test_lines_impl(qrad_cuda_ctx *c, int64_t n, const float *starts, const float *stops, uint8_t *results) {
  if (!c->numnodes) dfail("qrad_cuda_test_lines: no BSP uploaded");
  if (n <= 0) return;
  DBuf<float> ds, de;
  DBuf<uint8_t> dr;
  ds.upload(starts, (size_t) n * 3, c->stream);
  de.upload(stops, (size_t) n * 3, c->stream);
  dr.alloc((size_t) n);
  int blocks = (int) ((n + 255) / 256);
  if (c->tree_depth <= kTraceStack)
    k_test_lines<kTraceStack><<<blocks, 256, 0, c->stream>>>(c->tree(), n, ds.p, de.p, dr.p);
  else
    k_test_lines<kTraceStackBig><<<blocks, 256, 0, c->stream>>>(c->tree(), n, ds.p, de.p, dr.p);
  c->launch_check("k_test_lines");
  dr.download(results, (size_t) n, c->stream);
}

void point_in_leaf_impl(qrad_cuda_ctx *c, int64_t n, const float *pts, int32_t *out) {
  if (!c->numnodes) dfail("qrad_cuda_point_in_leaf: no BSP uploaded");
  if (n <= 0) return;
  DBuf<float> dp;
  DBuf<int32_t> dr;
  dp.upload(pts, (size_t) n * 3, c->stream);
  dr.alloc((size_t) n);
  k_point_in_leaf<<<(int) ((n + 255) / 256), 256, 0, c->stream>>>(c->tree(), n, dp.p, dr.p);
  c->launch_check("k_point_in_leaf");
  dr.download(out, (size_t) n, c->stream);
}

}  // namespace qrad

// ---------------------------------------------------------------------------
extern "C" {

int qrad_cuda_abi_version(void) { return QRAD_CUDA_ABI_VERSION; }
const char *qrad_cuda_last_error(void) { return g_cuda_error.c_str(); }

int qrad_cuda_create(int device, qrad_cuda_ctx **out) {
  *out = nullptr;
  return guarded([&] {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
      dfail("libqrad_cuda needs a CUDA device and there is no CPU fallback (%s)",
            e != cudaSuccess ? cudaGetErrorString(e) : "no devices");
    if (device < 0 || device >= count) dfail("CUDA device %d out of range (%d visible)", device, count);
    QCUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    QCUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) dfail("libqrad_cuda is built for sm_100a; device %d is sm_%d%d", device, prop.major, prop.minor);
    qrad_cuda_ctx *c = new qrad_cuda_ctx;
    c->device = device;
    c->traffic0 = traffic();
    c->sm_count = prop.multiProcessorCount;
    QCUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    QCUDA(cudaStreamCreateWithFlags(&c->xstream, cudaStreamNonBlocking));
    QCUDA(cudaStreamCreateWithFlags(&c->lstream, cudaStreamNonBlocking));
    QCUDA(cudaEventCreate(&c->ev_lists));
    for (auto &e : c->ev) QCUDA(cudaEventCreate(&e));
    *out = c;
  });
}

// Host-side derived data of a lighting pass, computed in the background while the GPU works on the phases before the one
// that needs it: the plan of the transfer matrix (make_transfers) and the plane groups (final_light).  Started by the
// uploads and again by build_facelights -- the first call of every pass -- so that a pass over data uploaded earlier
// does the same work as the first one.
static void start_plan(qrad_cuda_ctx *c) {
  if (c->plan_future.valid() || !c->have_vis || c->numclusters <= 0 || c->N <= 0) return;
  const int n = c->N, world = c->world, rank = c->rank, nc = c->numclusters, pw = c->pvs_words;
  const uint32_t *pvs = c->h_pvs.data();   // stays put until the next upload_bsp, which waits for this task
  std::vector<int32_t> eff(c->h_cluster);
  bool ok = true;
  for (int cl : eff) ok = ok && cl < nc;
  if (ok)
    c->plan_future = std::async(std::launch::async, [eff = std::move(eff), n, nc, pvs, pw, world, rank]() {
      auto p = std::make_shared<TransferPlan>();
      build_plan(*p, n, eff.data(), nc, pvs, pw, world, rank, 0);
      return p;
    });
}

static void start_plane_groups(qrad_cuda_ctx *c) {
  if (c->groups_future.valid() || c->N <= 0 || c->F <= 0 || (int) c->h_face_head.size() != c->F) return;
  c->groups_future = std::async(std::launch::async, [c] { return compute_plane_groups(c); });
}

void qrad_cuda_destroy(qrad_cuda_ctx *c) {
  if (!c) return;
  if (c->plan_future.valid()) c->plan_future.wait();
  if (c->groups_future.valid()) c->groups_future.wait();
  cudaSetDevice(c->device);
  cudaStream_t s = c->stream, x = c->xstream, l = c->lstream;
  if (s) cudaStreamSynchronize(s);
  if (x) cudaStreamSynchronize(x);
  if (l) cudaStreamSynchronize(l);
  if (c->ev_lists) cudaEventDestroy(c->ev_lists);
  if (c->comm) {
    try {
      nccl().CommDestroy(c->comm);
    } catch (...) {
    }
  }
  for (auto e : c->bounce_ev) cudaEventDestroy(e);
  for (auto e : c->ev)
    if (e) cudaEventDestroy(e);
  delete c;
  if (s) cudaStreamDestroy(s);
  if (x) cudaStreamDestroy(x);
  if (l) cudaStreamDestroy(l);
}

// ---- several GPUs: the context joins an NCCL communicator; from then on the same entry points do their share of the work
// and exchange inside the library (transfers.cu, direct.cu, final.cu) ----
int qrad_cuda_comm_unique_id(uint8_t *id128) {
  return guarded([&] {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId");
    ncclUniqueId id;
    ncclResult_t r = nccl().GetUniqueId(&id);
    if (r != ncclSuccess) dfail("ncclGetUniqueId failed: %s", nccl().GetErrorString(r));
    memcpy(id128, &id, 128);
  });
}

int qrad_cuda_comm_init(qrad_cuda_ctx *c, int rank, int world, const uint8_t *id128) {
  return guarded([&] {
    if (world < 1 || rank < 0 || rank >= world) dfail("bad rank %d of %d", rank, world);
    if (world > 16) dfail("at most 16 ranks");
    QCUDA(cudaSetDevice(c->device));
    if (c->comm) {
      nccl().CommDestroy(c->comm);
      c->comm = nullptr;
    }
    c->rank = 0;
    c->world = 1;
    if (world > 1) {
      ncclUniqueId id;
      memcpy(&id, id128, 128);
      ncclResult_t r = nccl().CommInitRank(&c->comm, world, id, rank);
      if (r != ncclSuccess) dfail("ncclCommInitRank failed: %s", nccl().GetErrorString(r));
      c->rank = rank;
      c->world = world;
    }
    c->transfers_done = c->facelights_done = false;
  });
}

int qrad_cuda_set_count_nodes(qrad_cuda_ctx *c, int enable) {
  c->count_nodes = enable != 0;
  return 0;
}

// Flatten dnodes/dplanes/dleafs into the trace tree (MakeTnodes, trace.c:47-88; breadth-first here), compute the box
// cells of the leaves (shaft_blocked, qrad_device.cuh), and expand the RLE PVS rows (DecompressVis, bspfile.c:129-154) once, so kernels
// test a cluster bit with one load.
int qrad_cuda_upload_bsp(qrad_cuda_ctx *c, const qrad_bsp_view *b) {
  return guarded([&] {
    QCUDA(cudaSetDevice(c->device));
    if (c->plan_future.valid()) c->plan_future.wait();   // a plan under construction reads the old PVS
    c->plan_future = {};
    Timer t;
    t.start(c->stream);
    FlatBsp flat;
    try {
      flatten_bsp(b, flat);
    } catch (const std::exception &e) {
      dfail("%s", e.what());
    }
    if (flat.depth > kTraceStackBig) dfail("BSP depth %d exceeds the supported trace stack (%d)", flat.depth, kTraceStackBig);
    c->numnodes = (int) flat.nodes.size();
    c->numleafs = b->numleafs;
    c->tree_depth = flat.depth;
    c->bsp_extent = flat.extent;
    c->nodes.upload(flat.nodes, c->stream);
    c->node_normals.upload(flat.normals, c->stream);
    c->leaf_contents.upload(flat.leaf_contents, c->stream);
    c->leaf_cluster.upload(flat.leaf_cluster, c->stream);
    c->leaf_cell_lo.upload(flat.cell_lo, c->stream);
    c->leaf_cell_hi.upload(flat.cell_hi, c->stream);
    c->have_vis = b->visdatasize > 0;
    c->numclusters = 0;
    c->h_pvs.clear();
    if (c->have_vis) {
      try {
        expand_pvs(b->visdata, b->visdatasize, c->numclusters, c->pvs_words, c->h_pvs);
      } catch (const std::exception &e) {
        dfail("%s", e.what());
      }
      c->pvs.upload(c->h_pvs, c->stream);
    }
    c->sync("upload_bsp");
    c->stats.tree_depth = c->tree_depth;
    c->stats.num_clusters = c->numclusters;
    c->stats.ms_upload = t.stop();
    c->transfers_done = c->facelights_done = false;
  });
}

static std::vector<float4> pack4(const float *xyz, const float *w, int n, float wdefault = 0.f) {
  std::vector<float4> out((size_t) n);
  for (int i = 0; i < n; i++) out[i] = make_float4(xyz[i * 3], xyz[i * 3 + 1], xyz[i * 3 + 2], w ? w[i] : wdefault);
  return out;
}

int qrad_cuda_upload_patches(qrad_cuda_ctx *c, const qrad_patches_view *p) {
  return guarded([&] {
    QCUDA(cudaSetDevice(c->device));
    int n = p->num_patches;
    if (n <= 0) dfail("no patches");
    c->N = n;
    c->numfaces_p = p->numfaces;
    if (c->groups_future.valid()) c->groups_future.wait();
    c->groups_future = {};
    c->h_cluster.assign(p->cluster, p->cluster + n);
    c->h_sky.assign(p->sky, p->sky + n);
    c->h_next.assign(p->next, p->next + n);
    c->h_face_head.assign(p->face_head, p->face_head + p->numfaces);
    for (int i = 0; i < n; i++) {
      if (p->face[i] < 0 || p->face[i] >= p->numfaces) dfail("patch %d has bad face %d", i, p->face[i]);
      if (p->next[i] < -1 || p->next[i] >= n) dfail("patch %d has bad next %d", i, p->next[i]);
    }
    c->h_origin_area = pack4(p->origin, p->area, n);
    c->h_normal = pack4(p->normal, nullptr, n);
    // w of the normal: +-(axis + 1) for an exactly axial normal, 0 otherwise (the sign test of MakeTransfers then
    // needs one coordinate difference instead of a normalised dot product, form_factor_sign() in transfers.cu)
    float extent = 0;
    for (int i = 0; i < n; i++) {
      float4 &nn = c->h_normal[i];
      const float v[3] = {nn.x, nn.y, nn.z};
      int code = 0;
      for (int k = 0; k < 3; k++)
        if (fabsf(v[k]) == 1.f && v[(k + 1) % 3] == 0.f && v[(k + 2) % 3] == 0.f) code = v[k] > 0 ? k + 1 : -(k + 1);
      nn.w = (float) code;
      const float4 &o = c->h_origin_area[i];
      extent = std::max(extent, std::max(fabsf(o.x), std::max(fabsf(o.y), fabsf(o.z))));
    }
    c->patch_extent = extent;
    c->p_origin_area.upload(c->h_origin_area, c->stream);
    c->p_normal.upload(c->h_normal, c->stream);
    std::vector<float> skyf(n);
    for (int i = 0; i < n; i++) skyf[i] = p->sky[i] ? 1.f : 0.f;
    c->p_refl.upload(pack4(p->reflectivity, skyf.data(), n), c->stream);
    c->p_baselight.upload(p->baselight, (size_t) n * 3, c->stream);
    c->p_totallight.upload(p->totallight, (size_t) n * 3, c->stream);
    c->p_bounds.upload(p->bounds, (size_t) n * 6, c->stream);
    c->p_face.upload(p->face, (size_t) n, c->stream);
    c->p_next.upload(p->next, (size_t) n, c->stream);
    c->p_cluster.upload(p->cluster, (size_t) n, c->stream);
    c->face_head.upload(p->face_head, (size_t) p->numfaces, c->stream);
    c->p_samplelight.alloc((size_t) n * 3);
    c->p_samplelight.zero(c->stream);
    c->p_samples.alloc((size_t) n);
    c->p_samples.zero(c->stream);
    c->d_radiosity.alloc((size_t) n * 3);
    c->d_radiosity.zero(c->stream);
    c->sync("upload_patches");
    if (c->plan_future.valid()) c->plan_future.wait();
    c->plan_future = {};
    start_plan(c);   // plan of the PVS-mode transfer matrix, in the background
    c->stats.num_patches = n;
    c->transfers_done = c->facelights_done = false;
  });
}

int qrad_cuda_upload_lights(qrad_cuda_ctx *c, const qrad_lights_view *l) {
  return guarded([&] {
    QCUDA(cudaSetDevice(c->device));
    if (l->numclusters != c->numclusters) dfail("lights are bucketed for %d clusters, BSP has %d", l->numclusters, c->numclusters);
    int n = l->numlights;
    c->numlights = n;
    // distinct styles, ascending; style 0 always has a table (lightmap.c:1010-1011)
    std::vector<int32_t> ids{0};
    for (int i = 0; i < n; i++) {
      if (l->style[i] < 0 || l->style[i] >= 256) dfail("light %d has style %d", i, l->style[i]);
      ids.push_back(l->style[i]);
    }
    std::sort(ids.begin(), ids.end());
    ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
    c->h_style_ids = ids;
    c->numstyles_map = (int) ids.size();
    std::vector<int32_t> slot(n);
    for (int i = 0; i < n; i++) slot[i] = (int32_t) (std::lower_bound(ids.begin(), ids.end(), l->style[i]) - ids.begin());
    std::vector<int32_t> first(l->light_first, l->light_first + l->numclusters + 1);
    if (first.empty()) first.push_back(0);
    c->light_first.upload(first, c->stream);
    std::vector<int32_t> lcl((size_t) n);
    for (int cl = 0; cl < l->numclusters; cl++)
      for (int k = first[cl]; k < first[cl + 1]; k++) lcl[k] = cl;
    c->l_cluster.upload(lcl, c->stream);
    c->l_type.upload(l->type, (size_t) n, c->stream);
    c->l_styleslot.upload(slot, c->stream);
    c->l_origin_int.upload(pack4(l->origin, l->intensity, n), c->stream);
    c->l_color.upload(pack4(l->color, l->stopdot, n), c->stream);
    c->l_normal.upload(pack4(l->normal, nullptr, n), c->stream);
    c->sync("upload_lights");
    c->facelights_done = false;
  });
}

int qrad_cuda_upload_faces(qrad_cuda_ctx *c, const qrad_faces_view *f) {
  return guarded([&] {
    QCUDA(cudaSetDevice(c->device));
    int nf = f->numfaces;
    if (nf <= 0) dfail("no faces");
    if (c->N && nf != c->numfaces_p) dfail("faces view has %d faces, patches view %d", nf, c->numfaces_p);
    if (c->groups_future.valid()) c->groups_future.wait();
    c->groups_future = {};
    c->F = nf;
    c->h_lit.assign(f->lit, f->lit + nf);
    c->h_texsize.assign(f->texsize, f->texsize + (size_t) nf * 2);
    c->h_planelink_head.assign(f->planelink_head, f->planelink_head + nf);
    c->h_facelinks.assign(f->facelinks, f->facelinks + nf);
    c->h_face_bounds.assign(f->face_bounds, f->face_bounds + (size_t) nf * 6);
    c->h_minlight.assign(f->minlight, f->minlight + nf);
    c->h_luxel_first.assign(nf + 1, 0);
    for (int i = 0; i < nf; i++) {
      int n = 0;
      if (f->lit[i]) {
        int w = f->texsize[i * 2] + 1, h = f->texsize[i * 2 + 1] + 1;
        if (w < 1 || h < 1 || (int64_t) (w - 1) * (h - 1) > 4096) dfail("Surface to large to map");
        n = w * h;
      }
      c->h_luxel_first[i + 1] = c->h_luxel_first[i] + n;
    }
    c->num_luxels = c->h_luxel_first[nf];
    std::vector<int32_t> lface((size_t) c->num_luxels);
    for (int i = 0; i < nf; i++)
      for (int k = c->h_luxel_first[i]; k < c->h_luxel_first[i + 1]; k++) lface[k] = i;
    c->f_lit.upload(f->lit, (size_t) nf, c->stream);
    c->f_normal.upload(f->facenormal, (size_t) nf * 3, c->stream);
    c->f_texorg.upload(f->texorg, (size_t) nf * 3, c->stream);
    c->f_textoworld.upload(f->textoworld, (size_t) nf * 6, c->stream);
    c->f_exactmins.upload(f->exactmins, (size_t) nf * 2, c->stream);
    c->f_exactmaxs.upload(f->exactmaxs, (size_t) nf * 2, c->stream);
    c->f_texmins.upload(f->texmins, (size_t) nf * 2, c->stream);
    c->f_texsize.upload(f->texsize, (size_t) nf * 2, c->stream);
    c->f_bounds.upload(f->face_bounds, (size_t) nf * 6, c->stream);
    c->f_minlight.upload(f->minlight, (size_t) nf, c->stream);
    c->f_plane_normal.upload(f->plane_normal, (size_t) nf * 3, c->stream);
    c->f_luxel_first.upload(c->h_luxel_first, c->stream);
    c->luxel_face.upload(lface, c->stream);
    c->sync("upload_faces");
    start_plane_groups(c);   // the plane groups FinalLightFace will need, in the background
    c->stats.num_luxels = c->num_luxels;
    c->facelights_done = false;
  });
}

int qrad_cuda_test_lines(qrad_cuda_ctx *c, int64_t n, const float *starts, const float *stops, uint8_t *results) {
  return guarded([&] {
    QCUDA(cudaSetDevice(c->device));
    test_lines_impl(c, n, starts, stops, results);
  });
}

int qrad_cuda_point_in_leaf(qrad_cuda_ctx *c, int64_t n, const float *points, int32_t *leafnums) {
  return guarded([&] {
    QCUDA(cudaSetDevice(c->device));
    point_in_leaf_impl(c, n, points, leafnums);
  });
}

int qrad_cuda_build_facelights(qrad_cuda_ctx *c, int extrasamples, int numbounce) {
  return guarded([&] {
    QCUDA(cudaSetDevice(c->device));
    if (numbounce > 0) {
      start_plan(c);
      start_plane_groups(c);
    }
    build_facelights_impl(c, extrasamples, numbounce);
  });
}

int qrad_cuda_make_transfers(qrad_cuda_ctx *c, int nopvs, int64_t *total_transfer) {
  return guarded([&] {
    QCUDA(cudaSetDevice(c->device));
    make_transfers_impl(c, nopvs);
    if (total_transfer) *total_transfer = c->total_transfers;
  });
}

// Pass 1 of MakeTransfers as rank `rank` of `world` would run it, alone on this GPU (no exchange follows): how the
// visibility work of a map divides over ranks can be measured without the ranks (tools/trace_balance.py).
int qrad_cuda_trace_shard(qrad_cuda_ctx *c, int nopvs, int rank, int world, float *trace_ms, int64_t *rays) {
  return guarded([&] {
    QCUDA(cudaSetDevice(c->device));
    if (c->comm) dfail("trace_shard: this context is part of a communicator");
    if (world < 1 || rank < 0 || rank >= world) dfail("bad rank %d of %d", rank, world);
    const int r0 = c->rank, w0 = c->world;
    c->rank = rank;
    c->world = world;
    try {
      transfers_trace(c, nopvs);
    } catch (...) {
      c->rank = r0;
      c->world = w0;
      throw;
    }
    c->rank = r0;
    c->world = w0;
    unsigned long long hc[16];
    c->counters.download(hc, 16, c->stream);
    if (trace_ms) *trace_ms = c->stats.ms_transfers_trace;
    if (rays) *rays = (int64_t) hc[0];
    c->transfers_done = false;
  });
}

int qrad_cuda_bounce(qrad_cuda_ctx *c, int numbounce, float *added_out) {
  return guarded([&] {
    QCUDA(cudaSetDevice(c->device));
    bounce_impl(c, numbounce, added_out);
  });
}

int qrad_cuda_final_light(qrad_cuda_ctx *c, const qrad_final_params *p, uint8_t *out, int32_t cap, int32_t *size,
                          int32_t *lightofs, uint8_t *styles) {
  return guarded([&] {
    QCUDA(cudaSetDevice(c->device));
    final_light_impl(c, p, out, cap, size, lightofs, styles);
  });
}

int qrad_cuda_get_stats(qrad_cuda_ctx *c, qrad_cuda_stats *out) {
  *out = c->stats;
  out->h2d_bytes = traffic().h2d - c->traffic0.h2d;
  out->d2h_bytes = traffic().d2h - c->traffic0.d2h;
  out->world = c->world;
  out->rank = c->rank;
  return 0;
}

int qrad_cuda_download_numtransfers(qrad_cuda_ctx *c, int32_t *out) {
  return guarded([&] {
    QCUDA(cudaSetDevice(c->device));
    if (!c->transfers_done) dfail("make_transfers has not run");
    std::vector<int32_t> rc((size_t) c->M);
    c->row_count.download(rc.data(), rc.size(), c->stream);
    for (int i = 0; i < c->N; i++) out[i] = c->plan.patch_slot[i] >= 0 ? rc[c->plan.patch_slot[i]] : 0;
  });
}

int qrad_cuda_download_transfers(qrad_cuda_ctx *c, int64_t *row_ptr, uint32_t *patch, uint16_t *transfer) {
  return guarded([&] {
    QCUDA(cudaSetDevice(c->device));
    if (!c->transfers_done) dfail("make_transfers has not run");
    download_transfers_impl(c, row_ptr, patch, transfer);
  });
}

int qrad_cuda_download_rows(qrad_cuda_ctx *c, int32_t nrows, const int32_t *rows, int64_t *row_ptr, uint32_t *patch,
                            uint16_t *transfer, int64_t capacity, int64_t *rays) {
  return guarded([&] {
    QCUDA(cudaSetDevice(c->device));
    if (!c->transfers_done) dfail("make_transfers has not run");
    download_rows_impl(c, nrows, rows, row_ptr, patch, transfer, capacity, rays);
  });
}

int qrad_cuda_download_patch_light(qrad_cuda_ctx *c, float *totallight, float *samplelight, int32_t *samples,
                                   float *radiosity) {
  return guarded([&] {
    QCUDA(cudaSetDevice(c->device));
    size_t n = (size_t) c->N;
    if (totallight) c->p_totallight.download(totallight, n * 3, c->stream);
    if (samplelight) c->p_samplelight.download(samplelight, n * 3, c->stream);
    if (samples) c->p_samples.download(samples, n, c->stream);
    if (radiosity) c->d_radiosity.download(radiosity, n * 3, c->stream);
  });
}

int qrad_cuda_upload_patch_light(qrad_cuda_ctx *c, const float *totallight, const float *samplelight,
                                 const int32_t *samples) {
  return guarded([&] {
    QCUDA(cudaSetDevice(c->device));
    size_t n = (size_t) c->N;
    if (!n) dfail("no patches uploaded");
    if (totallight) QCUDA(cudaMemcpyAsync(c->p_totallight.p, totallight, n * 12, cudaMemcpyHostToDevice, c->stream));
    if (samplelight) QCUDA(cudaMemcpyAsync(c->p_samplelight.p, samplelight, n * 12, cudaMemcpyHostToDevice, c->stream));
    if (samples) QCUDA(cudaMemcpyAsync(c->p_samples.p, samples, n * 4, cudaMemcpyHostToDevice, c->stream));
    c->sync("upload_patch_light");
  });
}

int qrad_cuda_facelight_layout(qrad_cuda_ctx *c, int32_t *luxel_first, int32_t *numstyles, int32_t *stylenums) {
  return guarded([&] {
    if (!c->facelights_done) dfail("build_facelights has not run");
    memcpy(luxel_first, c->h_luxel_first.data(), sizeof(int32_t) * (c->F + 1));
    memcpy(numstyles, c->h_fl_numstyles.data(), sizeof(int32_t) * c->F);
    // report real style ids, in the order FinalLightFace will use them
    for (int f = 0; f < c->F; f++)
      for (int s = 0; s < 32; s++) {
        int slot = c->h_fl_stylenums[(size_t) f * 32 + s];
        stylenums[(size_t) f * 32 + s] = (s < c->h_fl_numstyles[f]) ? c->h_style_ids[slot] : 0;
      }
  });
}

int qrad_cuda_style_ids(qrad_cuda_ctx *c, int32_t *ids, int32_t *count) {
  *count = (int32_t) c->h_style_ids.size();
  for (size_t i = 0; i < c->h_style_ids.size(); i++) ids[i] = c->h_style_ids[i];
  return 0;
}

// samples of the style_slot-th *map* style table (slot 0 = style 0); origins = facelight origins.
int qrad_cuda_download_facelight(qrad_cuda_ctx *c, float *origins, int32_t style_slot, float *samples) {
  return guarded([&] {
    QCUDA(cudaSetDevice(c->device));
    if (!c->facelights_done) dfail("build_facelights has not run");
    size_t L = (size_t) c->num_luxels;
    if (origins && L) QCUDA(cudaMemcpyAsync(origins, c->fl_surfpt.p, L * 12, cudaMemcpyDeviceToHost, c->stream));
    if (samples && L) {
      if (style_slot < 0 || style_slot >= c->numstyles_map) dfail("style slot %d out of range", style_slot);
      QCUDA(cudaMemcpyAsync(samples, c->fl_samples.p + (size_t) style_slot * L * 3, L * 12, cudaMemcpyDeviceToHost, c->stream));
    }
    c->sync("download_facelight");
  });
}

}  // extern "C"
