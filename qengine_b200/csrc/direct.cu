// BuildFacelights (lightmap.c:967-1056) on the GPU: CalcPoints, GatherSampleLight,
// AddSampleToPatch and the emissive-face baselight add, one kernel each.
// Float / double mix and operation order follow the reference;
// compiled with -fmad=false.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "comm.hpp"
#include "qrad_device.cuh"

namespace qrad {
namespace {

__constant__ float c_sampleofs[5][2] = {{0, 0}, {-0.25f, -0.25f}, {0.25f, -0.25f}, {0.25f, 0.25f}, {-0.25f, 0.25f}};

struct FaceArgs {
  Tree nodes;
  const int32_t *leaf_contents, *leaf_cluster;
  const uint32_t *pvs;
  int pvs_words, numclusters;
  const int32_t *luxel_face, *luxel_first;
  const float *f_normal, *f_texorg, *f_textoworld, *f_exactmins, *f_exactmaxs;
  const int32_t *f_texmins, *f_texsize;
  int L, numsamples;
  int lux_lo, lux_n;             // the luxels this launch lights (several GPUs: the luxels of this rank's faces)
  float *surfpt;                 // [numsamples][L][3]
  // lights
  const int32_t *light_first, *l_type, *l_styleslot, *l_cluster;
  int numlights;
  const float4 *l_origin_int, *l_color, *l_normal;
  int numstyles;
  float *samples;                // [numstyles][L][3]
  uint32_t *touched;             // [F][numstyles]
  unsigned long long *counters;  // [0] rays, [1] node visits
};

// CalcPoints, lightmap.c:608-677: one thread per (sample offset, luxel).
template <int STACK, bool COUNT>
__global__ void __launch_bounds__(128) k_calc_points(const FaceArgs a) {
  const int64_t gid = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
  unsigned long long rays = 0, visits_total = 0;
  if (gid < (int64_t) a.lux_n * a.numsamples) {
    const int smp = (int) (gid / a.lux_n);
    const int lux = a.lux_lo + (int) (gid - (int64_t) smp * a.lux_n);
    const int f = a.luxel_face[lux];
    const int idx = lux - a.luxel_first[f];
    const int w = a.f_texsize[f * 2] + 1;
    const int s = idx % w, t = idx / w;
    const float sofs = c_sampleofs[smp][0], tofs = c_sampleofs[smp][1];
    const float *texorg = a.f_texorg + f * 3;
    const float *tw0 = a.f_textoworld + f * 6, *tw1 = tw0 + 3;
    const float mids = (a.f_exactmaxs[f * 2] + a.f_exactmins[f * 2]) / 2;
    const float midt = (a.f_exactmaxs[f * 2 + 1] + a.f_exactmins[f * 2 + 1]) / 2;
    float facemid[3], surf[3];
    for (int j = 0; j < 3; j++) facemid[j] = texorg[j] + tw0[j] * mids + tw1[j] * midt;
    const float starts = (float) (a.f_texmins[f * 2] * 16);
    const float startt = (float) (a.f_texmins[f * 2 + 1] * 16);
    float us = starts + ((float) s + sofs) * 16.0f;
    float ut = startt + ((float) t + tofs) * 16.0f;
    for (int i = 0; i < 6; i++) {
      for (int j = 0; j < 3; j++) surf[j] = texorg[j] + tw0[j] * us + tw1[j] * ut;
      const int leaf = point_in_leaf(a.nodes, surf[0], surf[1], surf[2]);
      if (a.leaf_contents[leaf] != 1) {  // != CONTENTS_SOLID
        unsigned v = 0;
        rays++;
        const int blocked = test_line<STACK, COUNT>(a.nodes, facemid[0], facemid[1], facemid[2], surf[0], surf[1], surf[2], v);
        if (COUNT) visits_total += v;
        if (!blocked) break;
      }
      if (i & 1) {
        if (us > mids) {
          us -= 8;
          if (us < mids) us = mids;
        } else {
          us += 8;
          if (us > mids) us = mids;
        }
      } else {
        if (ut > midt) {
          ut -= 8;
          if (ut < midt) ut = midt;
        } else {
          ut += 8;
          if (ut > midt) ut = midt;
        }
      }
    }
    float *o = a.surfpt + ((size_t) smp * a.L + lux) * 3;
    o[0] = surf[0];
    o[1] = surf[1];
    o[2] = surf[2];
  }
  for (int o = 16; o; o >>= 1) {
    rays += __shfl_xor_sync(0xffffffffu, rays, o);
    if (COUNT) visits_total += __shfl_xor_sync(0xffffffffu, visits_total, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (rays) atomicAdd(a.counters + 0, rays);
    if (COUNT && visits_total) atomicAdd(a.counters + 1, visits_total);
  }
}

// GatherSampleLight, lightmap.c:847-917, for all samples of one luxel in the reference's order
// (sample j ascending, cluster ascending, lights in list order): one thread per luxel, because
// the accumulation into a style table is a serial chain of rounded double multiply-adds.
template <int STACK, bool COUNT>
__global__ void __launch_bounds__(128) k_gather(const FaceArgs a) {
  const int lux = a.lux_lo + blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long rays = 0, visits_total = 0;
  if (lux < a.lux_lo + a.lux_n) {
    const int f = a.luxel_face[lux];
    const float nx = a.f_normal[f * 3], ny = a.f_normal[f * 3 + 1], nz = a.f_normal[f * 3 + 2];
    const float lightscale = a.numsamples == 5 ? (float) (1.0 / 5) : 1.0f;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;  // style slot 0 (style 0)
    bool touched0 = false;
    for (int smp = 0; smp < a.numsamples; smp++) {
      const float *pos = a.surfpt + ((size_t) smp * a.L + lux) * 3;
      const float px = pos[0], py = pos[1], pz = pos[2];
      const int leaf = point_in_leaf(a.nodes, px, py, pz);
      const int cluster = a.leaf_cluster[leaf];
      if (cluster < 0) continue;  // PvsForOrigin fails in solid, qrad3.c:171-172
      const uint32_t *row = a.pvs + (size_t) cluster * a.pvs_words;
      for (int wi = 0; wi < a.pvs_words; wi++) {
        uint32_t bits = __ldg(row + wi);
        while (bits) {
          const int cl = wi * 32 + (__ffs(bits) - 1);
          bits &= bits - 1;
          if (cl >= a.numclusters) break;
          const int l0 = a.light_first[cl], l1 = a.light_first[cl + 1];
          for (int l = l0; l < l1; l++) {
            const float4 lo = __ldg(a.l_origin_int + l);
            float dx = lo.x - px, dy = lo.y - py, dz = lo.z - pz;
            const float dist = vnormalize(dx, dy, dz);
            const float dot = dx * nx + dy * ny + dz * nz;
            if ((double) dot <= 0.001) continue;  // behind sample surface
            const int type = a.l_type[l];
            float scale;
            if (type == 1) {
              scale = (lo.w - dist) * dot;
            } else {
              const float4 ln = __ldg(a.l_normal + l);
              const float dot2 = -(dx * ln.x + dy * ln.y + dz * ln.z);
              if (type == 0) {
                if ((double) dot2 <= 0.001) continue;  // behind light surface
                scale = (lo.w / (dist * dist)) * dot * dot2;
              } else {
                if (dot2 <= __ldg(a.l_color + l).w) continue;  // outside light cone
                scale = (lo.w - dist) * dot;
              }
            }
            rays++;  // the reference traces before it looks at scale (lightmap.c:898-902)
            if (scale <= 0) continue;
            unsigned v = 0;
            const int blocked = test_line<STACK, COUNT>(a.nodes, px, py, pz, lo.x, lo.y, lo.z, v);
            if (COUNT) visits_total += v;
            if (blocked) continue;
            const float4 col = __ldg(a.l_color + l);
            const double sc = (double) (scale * lightscale);  // VectorMA(dest, scale * lightscale, color, dest)
            const int slot = a.l_styleslot[l];
            if (slot == 0) {
              acc0 = (float) ((double) acc0 + sc * (double) col.x);
              acc1 = (float) ((double) acc1 + sc * (double) col.y);
              acc2 = (float) ((double) acc2 + sc * (double) col.z);
              touched0 = true;
            } else {
              float *dst = a.samples + ((size_t) slot * a.L + lux) * 3;
              dst[0] = (float) ((double) dst[0] + sc * (double) col.x);
              dst[1] = (float) ((double) dst[1] + sc * (double) col.y);
              dst[2] = (float) ((double) dst[2] + sc * (double) col.z);
              a.touched[(size_t) f * a.numstyles + slot] = 1u;
            }
          }
        }
      }
    }
    float *dst = a.samples + (size_t) lux * 3;
    dst[0] = acc0;
    dst[1] = acc1;
    dst[2] = acc2;
    (void) touched0;
  }
  for (int o = 16; o; o >>= 1) {
    rays += __shfl_xor_sync(0xffffffffu, rays, o);
    if (COUNT) visits_total += __shfl_xor_sync(0xffffffffu, visits_total, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (rays) atomicAdd(a.counters + 0, rays);
    if (COUNT && visits_total) atomicAdd(a.counters + 1, visits_total);
  }
}

// Warp-cooperative GatherSampleLight: one warp per luxel.  The lights a sample sees are, in the reference's order
// (clusters ascending, each cluster's list in order), exactly the flattened light array restricted to the clusters
// of the sample's PVS; the warp compacts those light indices into a small queue and evaluates 32 of them at a time
// (same sample point: the shadow rays of a batch share their origin and walk the top of the tree together).  The
// contributions are then folded into the style tables in queue order by all lanes redundantly, which keeps the
// serial chain of rounded double multiply-adds of the reference.
template <int STACK, bool COUNT>
__global__ void __launch_bounds__(128) k_gather_warp(const FaceArgs a) {
  __shared__ int32_t s_q[4][64];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1;
  const int lux = a.lux_lo + blockIdx.x * 4 + wid;
  unsigned long long rays = 0, visits_total = 0;
  if (lux < a.lux_lo + a.lux_n) {
    const int f = a.luxel_face[lux];
    const float nx = a.f_normal[f * 3], ny = a.f_normal[f * 3 + 1], nz = a.f_normal[f * 3 + 2];
    const float lightscale = a.numsamples == 5 ? (float) (1.0 / 5) : 1.0f;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;  // style slot 0 (style 0), identical in every lane
    for (int smp = 0; smp < a.numsamples; smp++) {
      const float *pos = a.surfpt + ((size_t) smp * a.L + lux) * 3;
      const float px = pos[0], py = pos[1], pz = pos[2];
      const int leaf = point_in_leaf(a.nodes, px, py, pz);
      const int cluster = a.leaf_cluster[leaf];
      if (cluster < 0) continue;  // PvsForOrigin fails in solid, qrad3.c:171-172
      const uint32_t *row = a.pvs + (size_t) cluster * a.pvs_words;
      // evaluate the first `cnt` lights of the queue, then fold their contributions in order
      auto batch = [&](const int cnt) {
        bool contrib = false;
        float scale = 0.f;
        int l = 0;
        if (lane < cnt) {
          l = s_q[wid][lane];
          const float4 lo = __ldg(a.l_origin_int + l);
          float dx = lo.x - px, dy = lo.y - py, dz = lo.z - pz;
          const float dist = vnormalize(dx, dy, dz);
          const float dot = dx * nx + dy * ny + dz * nz;
          bool pass = !((double) dot <= 0.001);  // behind sample surface
          if (pass) {
            const int type = a.l_type[l];
            if (type == 1) {
              scale = (lo.w - dist) * dot;
            } else {
              const float4 ln = __ldg(a.l_normal + l);
              const float dot2 = -(dx * ln.x + dy * ln.y + dz * ln.z);
              if (type == 0) {
                if ((double) dot2 <= 0.001) pass = false;  // behind light surface
                else scale = (lo.w / (dist * dist)) * dot * dot2;
              } else {
                if (dot2 <= __ldg(a.l_color + l).w) pass = false;  // outside light cone
                else scale = (lo.w - dist) * dot;
              }
            }
          }
          if (pass) {
            rays++;  // the reference traces before it looks at scale (lightmap.c:898-902)
            if (scale > 0) {
              unsigned v = 0;
              const int blocked = test_line<STACK, COUNT>(a.nodes, px, py, pz, lo.x, lo.y, lo.z, v);
              if (COUNT) visits_total += v;
              contrib = !blocked;
            }
          }
        }
        unsigned m = __ballot_sync(0xffffffffu, contrib);
        const float scaled = scale * lightscale;  // VectorMA(dest, scale * lightscale, color, dest)
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const double sc = (double) __shfl_sync(0xffffffffu, scaled, src);
          const int ll = __shfl_sync(0xffffffffu, l, src);
          const float4 col = __ldg(a.l_color + ll);
          const int slot = __ldg(a.l_styleslot + ll);
          if (slot == 0) {
            acc0 = (float) ((double) acc0 + sc * (double) col.x);
            acc1 = (float) ((double) acc1 + sc * (double) col.y);
            acc2 = (float) ((double) acc2 + sc * (double) col.z);
          } else if (lane == 0) {
            float *dst = a.samples + ((size_t) slot * a.L + lux) * 3;
            dst[0] = (float) ((double) dst[0] + sc * (double) col.x);
            dst[1] = (float) ((double) dst[1] + sc * (double) col.y);
            dst[2] = (float) ((double) dst[2] + sc * (double) col.z);
            a.touched[(size_t) f * a.numstyles + slot] = 1u;
          }
        }
        __syncwarp();
      };
      int qn = 0;
      for (int lb = 0; lb < a.numlights; lb += 32) {
        const int l = lb + lane;
        bool vis = false;
        if (l < a.numlights) {
          const int cl = __ldg(a.l_cluster + l);
          vis = (__ldg(row + (cl >> 5)) >> (cl & 31)) & 1u;
        }
        const unsigned m = __ballot_sync(0xffffffffu, vis);
        if (vis) s_q[wid][qn + __popc(m & lt)] = l;
        qn += __popc(m);
        __syncwarp();
        if (qn >= 32) {
          batch(32);
          int t = 0;
          if (lane < qn - 32) t = s_q[wid][32 + lane];
          __syncwarp();
          if (lane < qn - 32) s_q[wid][lane] = t;
          qn -= 32;
          __syncwarp();
        }
      }
      if (qn > 0) batch(qn);
    }
    if (lane == 0) {
      float *dst = a.samples + (size_t) lux * 3;
      dst[0] = acc0;
      dst[1] = acc1;
      dst[2] = acc2;
    }
  }
  for (int o = 16; o; o >>= 1) {
    rays += __shfl_xor_sync(0xffffffffu, rays, o);
    if (COUNT) visits_total += __shfl_xor_sync(0xffffffffu, visits_total, o);
  }
  if (lane == 0) {
    if (rays) atomicAdd(a.counters + 0, rays);
    if (COUNT && visits_total) atomicAdd(a.counters + 1, visits_total);
  }
}

// AddSampleToPatch + the averaging loop, lightmap.c:932-958, 1028-1034: one thread per patch walks
// the luxels of its face in order, so samplelight is the same serial float sum.
__global__ void k_add_samples(int N, const int32_t *__restrict__ p_face, const float *__restrict__ p_bounds,
                              const uint8_t *__restrict__ f_lit, const int32_t *__restrict__ luxel_first,
                              const float *__restrict__ surfpt0, const float *__restrict__ samples0,
                              float *__restrict__ samplelight, int32_t *__restrict__ nsamples, const int face_lo,
                              const int face_hi) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  const int f = p_face[p];
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  int n = 0;
  if (f_lit[f] && f >= face_lo && f < face_hi) {   // several GPUs: the other faces' patches stay 0 here and are summed in
    const float *mn = p_bounds + (size_t) p * 6, *mx = mn + 3;
    const float mn0 = mn[0], mn1 = mn[1], mn2 = mn[2], mx0 = mx[0], mx1 = mx[1], mx2 = mx[2];
    for (int l = luxel_first[f]; l < luxel_first[f + 1]; l++) {
      const float c0 = samples0[l * 3], c1 = samples0[l * 3 + 1], c2 = samples0[l * 3 + 2];
      if (c0 + c1 + c2 < 3) continue;
      const float x = surfpt0[l * 3], y = surfpt0[l * 3 + 1], z = surfpt0[l * 3 + 2];
      if (mn0 > x + 16 || mx0 < x - 16) continue;
      if (mn1 > y + 16 || mx1 < y - 16) continue;
      if (mn2 > z + 16 || mx2 < z - 16) continue;
      n++;
      s0 = s0 + c0;
      s1 = s1 + c1;
      s2 = s2 + c2;
    }
  }
  if (n) {
    const double inv = 1.0 / n;  // VectorScale(samplelight, 1.0 / samples, samplelight): double scale
    s0 = (float) (inv * (double) s0);
    s1 = (float) (inv * (double) s1);
    s2 = (float) (inv * (double) s2);
  }
  samplelight[p * 3] = s0;
  samplelight[p * 3 + 1] = s1;
  samplelight[p * 3 + 2] = s2;
  nsamples[p] = n;
}

// emissive faces stay full bright: lightmap.c:1046-1055
__global__ void k_add_baselight(int L, const int32_t *__restrict__ luxel_face, const int32_t *__restrict__ face_head,
                                const float *__restrict__ p_baselight, float *__restrict__ samples0) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  const int head = face_head[luxel_face[l]];
  if (head < 0) return;
  const float b0 = p_baselight[head * 3], b1 = p_baselight[head * 3 + 1], b2 = p_baselight[head * 3 + 2];
  if (b0 >= 3 || b1 >= 3 || b2 >= 3) {
    samples0[l * 3] = samples0[l * 3] + b0;
    samples0[l * 3 + 1] = samples0[l * 3 + 1] + b1;
    samples0[l * 3 + 2] = samples0[l * 3 + 2] + b2;
  }
}

}  // namespace

// Several GPUs: faces are lit (BuildFacelights) and finished (FinalLightFace) by contiguous runs of faces with (nearly)
// equal luxel counts; one GPU owns them all.
void face_range(qrad_cuda_ctx *c) {
  const int F = c->F;
  c->face_lo = 0;
  c->face_hi = F;
  if (c->world <= 1) return;
  const int64_t L = c->h_luxel_first[F];
  auto bound = [&](int r) {   // first face whose luxels start at or after r / world of all luxels
    if (r <= 0) return 0;
    if (r >= c->world) return F;
    const int64_t want = L * r / c->world;
    return (int) (std::lower_bound(c->h_luxel_first.begin(), c->h_luxel_first.begin() + F, (int32_t) want) - c->h_luxel_first.begin());
  };
  c->face_lo = bound(c->rank);
  c->face_hi = bound(c->rank + 1);
}

void build_facelights_impl(qrad_cuda_ctx *c, int extrasamples, int numbounce) {
  if (!c->numnodes || !c->N || !c->F) dfail("build_facelights: BSP, patches, lights and faces must be uploaded first");
  Timer t;
  t.start(c->stream);
  cudaStream_t st = c->stream;
  const int L = c->num_luxels;
  const int ns = extrasamples ? 5 : 1;
  c->fl_numsamples = ns;
  c->fl_surfpt.alloc((size_t) ns * L * 3 + 1);
  c->fl_samples.alloc((size_t) c->numstyles_map * L * 3 + 1);
  c->fl_samples.zero(st);
  c->fl_touched.alloc((size_t) c->F * c->numstyles_map + 1);
  c->fl_touched.zero(st);
  DBuf<unsigned long long> counters;
  counters.alloc(2);
  counters.zero(st);
  FaceArgs a{};
  a.nodes = c->tree();
  a.leaf_contents = c->leaf_contents.p;
  a.leaf_cluster = c->leaf_cluster.p;
  a.pvs = c->pvs.p;
  a.pvs_words = c->pvs_words;
  a.numclusters = c->numclusters;
  a.luxel_face = c->luxel_face.p;
  a.luxel_first = c->f_luxel_first.p;
  a.f_normal = c->f_normal.p;
  a.f_texorg = c->f_texorg.p;
  a.f_textoworld = c->f_textoworld.p;
  a.f_exactmins = c->f_exactmins.p;
  a.f_exactmaxs = c->f_exactmaxs.p;
  a.f_texmins = c->f_texmins.p;
  a.f_texsize = c->f_texsize.p;
  a.L = L;
  a.numsamples = ns;
  face_range(c);
  const int lux_lo = c->h_luxel_first[c->face_lo], lux_n = c->h_luxel_first[c->face_hi] - lux_lo;
  a.lux_lo = lux_lo;
  a.lux_n = lux_n;
  a.surfpt = c->fl_surfpt.p;
  a.light_first = c->light_first.p;
  a.l_type = c->l_type.p;
  a.l_styleslot = c->l_styleslot.p;
  a.l_cluster = c->l_cluster.p;
  a.numlights = c->numlights;
  a.l_origin_int = c->l_origin_int.p;
  a.l_color = c->l_color.p;
  a.l_normal = c->l_normal.p;
  a.numstyles = c->numstyles_map;
  a.samples = c->fl_samples.p;
  a.touched = c->fl_touched.p;
  a.counters = counters.p;
  const bool big = c->tree_depth > kTraceStack;
  if (lux_n > 0) {
    const int64_t npts = (int64_t) lux_n * ns;
    const int pb = (int) ((npts + 127) / 128), gb = (lux_n + 127) / 128;
#define QRAD_LAUNCH2(K, G)                                                                  \
  do {                                                                                      \
    if (c->count_nodes) {                                                                   \
      if (big) K<kTraceStackBig, true><<<G, 128, 0, st>>>(a);                               \
      else K<kTraceStack, true><<<G, 128, 0, st>>>(a);                                      \
    } else {                                                                                \
      if (big) K<kTraceStackBig, false><<<G, 128, 0, st>>>(a);                              \
      else K<kTraceStack, false><<<G, 128, 0, st>>>(a);                                     \
    }                                                                                       \
  } while (0)
    QRAD_LAUNCH2(k_calc_points, pb);
    c->launch_check("k_calc_points");
    // an un-vised map has dvis->numclusters == 0: GatherSampleLight's loop never runs (appendix C)
    if (c->numclusters > 0 && c->numlights > 0) {
      // many lights per luxel: one warp per luxel (coherent shadow rays); few: one thread per luxel
      const bool warp_mode = getenv("QRAD_GATHER_THREAD") == nullptr && c->numlights >= 64;
      if (warp_mode) {
        QRAD_LAUNCH2(k_gather_warp, (lux_n + 3) / 4);
        c->launch_check("k_gather_warp");
      } else {
        QRAD_LAUNCH2(k_gather, gb);
        c->launch_check("k_gather");
      }
    }
#undef QRAD_LAUNCH2
  }
  if (numbounce > 0) {
    // AddSampleToPatch sees the luxel colours BEFORE the emissive baselight is added (lightmap.c:1024 vs 1046-1055).
    // Several GPUs: patches of faces lit elsewhere get 0 here and the sum over ranks below fills them in.
    k_add_samples<<<(c->N + 127) / 128, 128, 0, st>>>(c->N, c->p_face.p, c->p_bounds.p, c->f_lit.p, c->f_luxel_first.p,
                                                      c->fl_surfpt.p, c->fl_samples.p, c->p_samplelight.p, c->p_samples.p,
                                                      c->face_lo, c->face_hi);
    c->launch_check("k_add_samples");
    if (c->world > 1) {
      const NcclApi &n = nccl();
      ncclResult_t r = n.AllReduce(c->p_samplelight.p, c->p_samplelight.p, (size_t) c->N * 3, ncclFloat, ncclSum, c->comm, st);
      if (r == ncclSuccess) r = n.AllReduce(c->p_samples.p, c->p_samples.p, (size_t) c->N, ncclInt32, ncclSum, c->comm, st);
      if (r != ncclSuccess) dfail("ncclAllReduce(samplelight) failed: %s", n.GetErrorString(r));
    }
  }
  if (L > 0) {
    k_add_baselight<<<(L + 255) / 256, 256, 0, st>>>(L, c->luxel_face.p, c->face_head.p, c->p_baselight.p, c->fl_samples.p);
    c->launch_check("k_add_baselight");
  }
  if (c->world > 1) {   // which style tables exist per face (any rank may have lit the face: exactly one did)
    const NcclApi &n = nccl();
    ncclResult_t r = n.AllReduce(c->fl_touched.p, c->fl_touched.p, (size_t) c->F * c->numstyles_map, ncclUint32, ncclMax, c->comm, st);
    if (r != ncclSuccess) dfail("ncclAllReduce(styles) failed: %s", n.GetErrorString(r));
  }
  // style bookkeeping of facelight_t (lightmap.c:1036-1044): style 0 always, others when touched; ascending ids
  std::vector<uint32_t> touched((size_t) c->F * c->numstyles_map + 1);
  c->fl_touched.download(touched.data(), touched.size(), st);
  c->h_fl_numstyles.assign(c->F, 0);
  c->h_fl_stylenums.assign((size_t) c->F * 32, 0);
  for (int f = 0; f < c->F; f++) {
    if (!c->h_lit[f]) continue;
    int n = 0;
    for (int s = 0; s < c->numstyles_map && n < 32; s++)
      if (s == 0 || touched[(size_t) f * c->numstyles_map + s]) c->h_fl_stylenums[(size_t) f * 32 + n++] = s;
    c->h_fl_numstyles[f] = n;
  }
  unsigned long long hc[2];
  counters.download(hc, 2, st);
  c->stats.rays_direct = (int64_t) hc[0];
  c->stats.node_visits += (int64_t) hc[1];
  c->facelights_done = true;
  c->stats.ms_facelights = t.stop();
}

}  // namespace qrad
