// First stage of the reference's qvis3 on the GPU (include/qvis_cuda.h): BasePortalVis and SimpleFlood, flow.c:576-652.
// Two kernels, both bit-matrix work over (portal, portal):
//   k_portal_front   portalfront[p] bit j: some point of j's winding lies in front of p's plane by more than ON_EPSILON
//                    AND some point of p's winding lies behind j's plane by more than ON_EPSILON (flow.c:623-644).
//                    A warp per portal p, lane = j: the accept word of 32 portals is one ballot.
//   k_simple_flood   portalflood[p]: the closure of "cross a fronted portal of the cluster, continue in the cluster behind
//                    it" from p's cluster (flow.c:576-598).  A warp per portal; a cluster is scanned once (a second visit
//                    would find its fronted portals flooded already), so the work list holds at most `clusters` entries.
// Float arithmetic in the reference's order (this library is compiled with -fmad=false); the distance is compared with
// the DOUBLE 0.1 of vis.h:31.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/qvis_cuda.h"

void qvis_set_error(const std::string &m);

namespace {

struct FrontArgs {
  int P, words;
  const float4 *planes;
  const int32_t *wfirst;
  const float *points;
  uint32_t *front;   // [P][words]
};

__device__ __forceinline__ float plane_dist(const float *pt, const float4 pl) {
  return (pt[0] * pl.x + pt[1] * pl.y + pt[2] * pl.z) - pl.w;
}

__global__ void __launch_bounds__(256) k_portal_front(const FrontArgs a) {
  __shared__ float s_pts[8][64 * 3];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int p = blockIdx.x * 8 + wid;
  if (p >= a.P) return;
  const float4 pl = a.planes[p];
  const int w0 = a.wfirst[p], np = a.wfirst[p + 1] - w0;
  for (int k = lane; k < np * 3; k += 32) s_pts[wid][k] = a.points[(size_t) w0 * 3 + k];
  __syncwarp();
  for (int j0 = 0; j0 < a.words * 32; j0 += 32) {
    const int j = j0 + lane;
    bool front = false;
    if (j < a.P && j != p) {
      const int v0 = a.wfirst[j], nj = a.wfirst[j + 1] - v0;
      bool any = false;
      for (int k = 0; k < nj && !any; k++) any = (double) plane_dist(a.points + (size_t) (v0 + k) * 3, pl) > 0.1;
      if (any) {
        const float4 pj = a.planes[j];
        bool back = false;
        for (int k = 0; k < np && !back; k++) back = (double) plane_dist(&s_pts[wid][k * 3], pj) < -0.1;
        front = back;
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, front);
    if (lane == 0) a.front[(size_t) p * a.words + (j0 >> 5)] = m;
  }
}

struct FloodArgs {
  int P, nc, words, leaf_words, nwarps;
  const uint32_t *front;
  const int32_t *leaf, *leaf_first, *leaf_portals;
  uint32_t *flood;      // [P][words], zeroed
  int32_t *might;       // [P]
  int32_t *queue;       // [nwarps][nc]
  uint32_t *seen;       // [nwarps][leaf_words]
};

__global__ void __launch_bounds__(256) k_simple_flood(const FloodArgs a) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= a.nwarps) return;
  int32_t *queue = a.queue + (size_t) w * a.nc;
  uint32_t *seen = a.seen + (size_t) w * a.leaf_words;
  for (int p = w; p < a.P; p += a.nwarps) {
    const uint32_t *fr = a.front + (size_t) p * a.words;
    uint32_t *fl = a.flood + (size_t) p * a.words;
    for (int k = lane; k < a.leaf_words; k += 32) seen[k] = 0;
    __syncwarp();
    int head = 0, tail = 0;
    if (lane == 0) {
      const int l = a.leaf[p];
      queue[0] = l;
      seen[l >> 5] = 1u << (l & 31);
    }
    tail = 1;
    __syncwarp();
    while (head < tail) {
      const int l = queue[head++];
      const int e0 = a.leaf_first[l], e1 = a.leaf_first[l + 1];
      for (int eb = e0; eb < e1; eb += 32) {
        const int e = eb + lane;
        int push = -1;
        if (e < e1) {
          const int q = a.leaf_portals[e];
          const uint32_t bit = 1u << (q & 31);
          if ((fr[q >> 5] & bit) && !(atomicOr(&fl[q >> 5], bit) & bit)) {
            const int nl = a.leaf[q];
            const uint32_t lb = 1u << (nl & 31);
            if (!(atomicOr(&seen[nl >> 5], lb) & lb)) push = nl;
          }
        }
        const unsigned m = __ballot_sync(0xffffffffu, push >= 0);
        if (push >= 0) queue[tail + __popc(m & ((1u << lane) - 1))] = push;
        tail += __popc(m);
        __syncwarp();
      }
    }
    int cnt = 0;
    for (int k = lane; k < a.words; k += 32) cnt += __popc(fl[k]);
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) a.might[p] = cnt;
    __syncwarp();
  }
}

struct Bufs {
  std::vector<void *> ptrs;
  ~Bufs() {
    for (void *p : ptrs) cudaFree(p);
  }
  template <class T> T *alloc(size_t n) {
    void *p = nullptr;
    if (cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) throw std::runtime_error("cudaMalloc failed");
    ptrs.push_back(p);
    return (T *) p;
  }
  template <class T> T *upload(const T *h, size_t n) {
    T *d = alloc<T>(n);
    if (n && cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) throw std::runtime_error("cudaMemcpy failed");
    return d;
  }
};

void ck(cudaError_t e, const char *what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

}  // namespace

extern "C" int qvis_cuda_base_portal_vis(int device, const qvis_portals_view *v, uint8_t *portalfront, uint8_t *portalflood,
                                         int32_t *nummightsee, float *ms_front, float *ms_flood) {
  try {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
      throw std::runtime_error("libqrad_cuda needs a CUDA device and there is no CPU fallback");
    if (device < 0 || device >= count) throw std::runtime_error("CUDA device out of range");
    ck(cudaSetDevice(device), "cudaSetDevice");
    const int P = v->numportals, nc = v->numclusters;
    if (P <= 0 || nc <= 0) throw std::runtime_error("BasePortalVis: no portals");
    for (int p = 0; p < P; p++) {
      const int np = v->winding_first[p + 1] - v->winding_first[p];
      if (np < 1 || np > 64) throw std::runtime_error("BasePortalVis: a winding has more than MAX_POINTS_ON_WINDING points");
      if ((unsigned) v->leaf[p] >= (unsigned) nc) throw std::runtime_error("BasePortalVis: portal leads into no cluster");
    }
    const int pb = qvis_portal_bytes(P), words = pb / 4, leaf_words = (nc + 31) / 32;
    Bufs b;
    FrontArgs fa{};
    fa.P = P;
    fa.words = words;
    fa.planes = (const float4 *) b.upload(v->planes, (size_t) P * 4);
    fa.wfirst = b.upload(v->winding_first, (size_t) P + 1);
    fa.points = b.upload(v->points, (size_t) v->winding_first[P] * 3);
    fa.front = b.alloc<uint32_t>((size_t) P * words);
    cudaDeviceProp prop;
    ck(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
    FloodArgs fl{};
    fl.P = P;
    fl.nc = nc;
    fl.words = words;
    fl.leaf_words = leaf_words;
    // warps in flight: every SM full, but not more work lists than 256 MB hold
    fl.nwarps = (int) std::max<int64_t>(1, std::min<int64_t>({(int64_t) P, (int64_t) prop.multiProcessorCount * 64, ((int64_t) 256 << 20) / ((int64_t) nc * 4)}));
    fl.front = fa.front;
    fl.leaf = b.upload(v->leaf, (size_t) P);
    fl.leaf_first = b.upload(v->leaf_first, (size_t) nc + 1);
    fl.leaf_portals = b.upload(v->leaf_portals, (size_t) v->leaf_first[nc]);
    fl.flood = b.alloc<uint32_t>((size_t) P * words);
    fl.might = b.alloc<int32_t>((size_t) P);
    fl.queue = b.alloc<int32_t>((size_t) fl.nwarps * nc);
    fl.seen = b.alloc<uint32_t>((size_t) fl.nwarps * leaf_words);
    ck(cudaMemset(fl.flood, 0, (size_t) P * words * 4), "cudaMemset");
    cudaEvent_t ev[3];
    for (auto &e : ev) ck(cudaEventCreate(&e), "cudaEventCreate");
    ck(cudaEventRecord(ev[0]), "cudaEventRecord");
    k_portal_front<<<(P + 7) / 8, 256>>>(fa);
    ck(cudaGetLastError(), "k_portal_front");
    ck(cudaEventRecord(ev[1]), "cudaEventRecord");
    k_simple_flood<<<(fl.nwarps + 7) / 8, 256>>>(fl);
    ck(cudaGetLastError(), "k_simple_flood");
    ck(cudaEventRecord(ev[2]), "cudaEventRecord");
    ck(cudaDeviceSynchronize(), "BasePortalVis kernels");
    float t0 = 0, t1 = 0;
    cudaEventElapsedTime(&t0, ev[0], ev[1]);
    cudaEventElapsedTime(&t1, ev[1], ev[2]);
    for (auto &e : ev) cudaEventDestroy(e);
    if (ms_front) *ms_front = t0;
    if (ms_flood) *ms_flood = t1;
    ck(cudaMemcpy(portalfront, fa.front, (size_t) P * pb, cudaMemcpyDeviceToHost), "cudaMemcpy");
    ck(cudaMemcpy(portalflood, fl.flood, (size_t) P * pb, cudaMemcpyDeviceToHost), "cudaMemcpy");
    ck(cudaMemcpy(nummightsee, fl.might, (size_t) P * 4, cudaMemcpyDeviceToHost), "cudaMemcpy");
    return 0;
  } catch (const std::exception &e) {
    qvis_set_error(e.what());
    return 1;
  }
}
