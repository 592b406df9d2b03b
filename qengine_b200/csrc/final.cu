// FinalLightFace (lightmap.c:1066-1201) with its patch-point triangulation (lightmap.c:128-439)
// on the GPU.  One warp owns one face: lane 0 keeps the (inherently serial) edge / triangle
// bookkeeping of TriEdge_r in the reference's depth-first order, all 32 lanes share the
// "point with the best angle" scans, then every lane shades luxels.  Compiled with -fmad=false.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>

#include "comm.hpp"
#include "qrad_device.cuh"

namespace qrad {
namespace {

constexpr int kMaxTriPoints = 1024;            // MAX_TRI_POINTS, lightmap.c:107
constexpr int kMaxTriEdges = kMaxTriPoints * 6;  // MAX_TRI_EDGES
constexpr int kMaxTriTris = kMaxTriPoints * 2;   // MAX_TRI_TRIS
constexpr int kStackCap = 2 * kMaxTriTris + 8;
constexpr int kWarps = 4;                      // warps (faces in flight) per block

enum { kErrNone = 0, kErrEdges = 1, kErrTris = 2, kErrStack = 3 };

// ---- gather the triangulation points of each face (lightmap.c:1119-1135) ---------------------
struct PointArgs {
  int F;
  int face_lo, face_hi;           // faces this rank finishes (the others report 0 points)
  const uint8_t *f_lit;
  const int32_t *group_of_face;   // [F] plane group or -1
  const int32_t *group_first;     // [G+1] into group_patch
  const int32_t *group_patch;     // patches of the group's faces in chain order
  const float4 *p_origin_area;
  const float *f_bounds;          // [F][6]
  float subdiv2;                  // subdiv * 2
  int32_t *tri_count;             // [F] points of the face (may exceed kMaxTriPoints: the host reports the reference's error)
  int32_t *tri_points;            // [F][kMaxTriPoints] patch indices in chain order
  int32_t *pmax;                  // largest count
};

// One block per face of this rank.  The chain of a large plane (the floor of a whole map) holds hundreds of thousands of
// patches and the points must come out in chain order: the block takes 256 candidates at a time and compacts them with
// a ballot per warp and a prefix over the warps.  (One warp per face, and a counting pass before the filling pass, took
// 14 ms on C5 however many GPUs shared the faces: the time of the longest chain.)
__global__ void __launch_bounds__(256) k_tri_points(const PointArgs a) {
  __shared__ int s_cnt[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int f = a.face_lo + blockIdx.x;
  if (f >= a.face_hi) return;
  const int g = a.f_lit[f] ? a.group_of_face[f] : -1;
  int total = 0;
  if (g >= 0) {
    const float *mn = a.f_bounds + (size_t) f * 6, *mx = mn + 3;
    const float mn0 = mn[0], mn1 = mn[1], mn2 = mn[2], mx0 = mx[0], mx1 = mx[1], mx2 = mx[2];
    const int k0 = a.group_first[g], k1 = a.group_first[g + 1];
    int32_t *out = a.tri_points + (size_t) f * kMaxTriPoints;
    for (int k = k0; k < k1; k += 256) {
      bool keep = false;
      int p = -1;
      if (k + (int) threadIdx.x < k1) {
        p = a.group_patch[k + threadIdx.x];
        const float4 o = a.p_origin_area[p];
        keep = !(mn0 - o.x > a.subdiv2 || o.x - mx0 > a.subdiv2 || mn1 - o.y > a.subdiv2 || o.y - mx1 > a.subdiv2 ||
                 mn2 - o.z > a.subdiv2 || o.z - mx2 > a.subdiv2);
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) s_cnt[wid] = __popc(m);
      __syncthreads();
      int before = 0, all = 0;
#pragma unroll
      for (int w = 0; w < 8; w++) {
        const int cw = s_cnt[w];
        before += w < wid ? cw : 0;
        all += cw;
      }
      const int at = total + before + __popc(m & ((1u << lane) - 1));
      if (keep && at < kMaxTriPoints) out[at] = p;
      total += all;
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    a.tri_count[f] = total;
    atomicMax(a.pmax, total);
  }
}

// ---- per-warp scratch --------------------------------------------------------------------------
struct Scratch {
  uint16_t *matrix;   // [pmax][pmax] edge index + 1, 0 = none   (edgematrix)
  float4 *pts;        // [pmax] origins
  float4 *e_plane;    // [kMaxTriEdges + 2] normal.xyz, dist
  uint16_t *e_p0, *e_p1;
  int16_t *e_tri;     // -1 = exterior
  uint16_t *t_edges;  // [kMaxTriTris][3]
  uint16_t *stack;    // [kStackCap]
};

struct FinalArgs {
  int F, pmax;
  int face_hi;                   // faces are drawn from a counter that starts at this rank's first face
  const uint8_t *f_lit;
  const int32_t *luxel_first;
  const int32_t *tri_count;
  const int32_t *tri_points;      // [F][kMaxTriPoints]
  const float4 *p_origin_area;
  const float *p_totallight;
  const float *f_plane_normal;
  const float *origins;          // facelight origins = surfpt of sample 0
  const float *samples;          // [numstyles_map][L][3]
  int L;
  const int32_t *fl_numstyles;   // [F]
  const int32_t *fl_styleslot;   // [F][32]
  const int32_t *lightofs;       // [F]
  const float *f_minlight;
  float ambient, maxlight, lightscale;
  int numbounce;
  uint8_t *out;
  float *lb_out;                 // [4][L][3] pre-scaling colours for faces with _minlight (host finishes), may be null
  int *error;
  // scratch pools, one slice per warp in the grid
  uint16_t *matrix_pool;
  float4 *pts_pool, *eplane_pool;
  uint16_t *ep0_pool, *ep1_pool, *tedges_pool, *stack_pool;
  int16_t *etri_pool;
  int *face_counter;
};

__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
  return ax * bx + ay * by + az * bz;
}

// FindEdge, lightmap.c:146-186.  Uniform across the warp; lane 0 writes.
__device__ int find_edge(const Scratch &s, int pmax, int p0, int p1, int &numedges, const float pnx, const float pny,
                         const float pnz, int lane, int *error) {
  const int have = s.matrix[(size_t) p0 * pmax + p1];
  if (have) return have - 1;
  if (numedges > kMaxTriEdges - 2) {
    if (lane == 0) atomicMax(error, kErrEdges);
    return -1;
  }
  const float4 a = s.pts[p0], b = s.pts[p1];
  float vx = b.x - a.x, vy = b.y - a.y, vz = b.z - a.z;
  vnormalize(vx, vy, vz);
  const float nx = vy * pnz - vz * pny;
  const float ny = vz * pnx - vx * pnz;
  const float nz = vx * pny - vy * pnx;
  const float dist = dot3(a.x, a.y, a.z, nx, ny, nz);
  const int e = numedges;
  if (lane == 0) {
    s.e_p0[e] = (uint16_t) p0;
    s.e_p1[e] = (uint16_t) p1;
    s.e_tri[e] = -1;
    s.e_plane[e] = make_float4(nx, ny, nz, dist);
    s.matrix[(size_t) p0 * pmax + p1] = (uint16_t) (e + 1);
    s.e_p0[e + 1] = (uint16_t) p1;
    s.e_p1[e + 1] = (uint16_t) p0;
    s.e_tri[e + 1] = -1;
    s.e_plane[e + 1] = make_float4(0 - nx, 0 - ny, 0 - nz, -dist);
    s.matrix[(size_t) p1 * pmax + p0] = (uint16_t) (e + 2);
  }
  numedges += 2;
  __syncwarp();
  return e;
}

__global__ void __launch_bounds__(kWarps * 32) k_final(const FinalArgs a) {
  const int lane = threadIdx.x & 31;
  const int wslot = blockIdx.x * kWarps + (threadIdx.x >> 5);
  Scratch s;
  s.matrix = a.matrix_pool + (size_t) wslot * a.pmax * a.pmax;
  s.pts = a.pts_pool + (size_t) wslot * a.pmax;
  s.e_plane = a.eplane_pool + (size_t) wslot * (kMaxTriEdges + 2);
  s.e_p0 = a.ep0_pool + (size_t) wslot * (kMaxTriEdges + 2);
  s.e_p1 = a.ep1_pool + (size_t) wslot * (kMaxTriEdges + 2);
  s.e_tri = a.etri_pool + (size_t) wslot * (kMaxTriEdges + 2);
  s.t_edges = a.tedges_pool + (size_t) wslot * kMaxTriTris * 3;
  s.stack = a.stack_pool + (size_t) wslot * kStackCap;

  for (;;) {
    int f = 0;
    if (lane == 0) f = atomicAdd(a.face_counter, 1);
    f = __shfl_sync(0xffffffffu, f, 0);
    if (f >= a.face_hi) break;   // the counter starts at this rank's first face
    if (!a.f_lit[f]) continue;
    const int P = a.numbounce > 0 ? a.tri_count[f] : 0;
    const int32_t *pidx = a.tri_points + (size_t) f * kMaxTriPoints;
    int numedges = 0, numtris = 0;
    const float pnx = a.f_plane_normal[f * 3], pny = a.f_plane_normal[f * 3 + 1], pnz = a.f_plane_normal[f * 3 + 2];

    if (P >= 2) {
      // ---- TriangulatePoints, lightmap.c:262-293 ----
      for (int i = lane; i < P; i += 32) s.pts[i] = a.p_origin_area[pidx[i]];
      for (int i = 0; i < P; i++)
        for (int j = lane; j < P; j += 32) s.matrix[(size_t) i * a.pmax + j] = 0;
      __syncwarp();
      // the two closest points: first minimum in (i, j) lexicographic order
      float bestd = 9999;
      int bi = -1, bj = -1;
      for (int i = 0; i < P; i++) {
        const float4 p1 = s.pts[i];
        for (int j = i + 1 + lane; j < P; j += 32) {
          const float4 p2 = s.pts[j];
          const float vx = p2.x - p1.x, vy = p2.y - p1.y, vz = p2.z - p1.z;
          // VectorLength: float squares accumulated in double (mathlib.c:30-41), then rounded to vec_t
          const float d = (float) sqrt(((double) (vx * vx) + (double) (vy * vy)) + (double) (vz * vz));
          if (d < bestd) {
            bestd = d;
            bi = i;
            bj = j;
          }
        }
      }
      for (int o = 16; o; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, bestd, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
        const bool better = oi >= 0 && (bi < 0 || od < bestd || (od == bestd && (oi < bi || (oi == bi && oj < bj))));
        if (better) {
          bestd = od;
          bi = oi;
          bj = oj;
        }
      }
      if (bi >= 0) {
        const int e = find_edge(s, a.pmax, bi, bj, numedges, pnx, pny, pnz, lane, a.error);
        const int e2 = find_edge(s, a.pmax, bj, bi, numedges, pnx, pny, pnz, lane, a.error);
        int sp = 0;
        if (lane == 0) {
          s.stack[0] = (uint16_t) e2;
          s.stack[1] = (uint16_t) e;
        }
        sp = 2;
        __syncwarp();
        // ---- TriEdge_r, lightmap.c:211-260, depth first with an explicit stack ----
        bool fail = false;
        while (sp > 0 && !fail) {
          const int e = s.stack[--sp];
          if (s.e_tri[e] >= 0) continue;  // already connected by someone
          const int ep0 = s.e_p0[e], ep1 = s.e_p1[e];
          const float4 p0 = s.pts[ep0], p1 = s.pts[ep1];
          const float4 en = s.e_plane[e];
          float best = 1.1f;
          int bestp = -1;
          for (int i = lane; i < P; i += 32) {
            const float4 p = s.pts[i];
            if (dot3(p.x, p.y, p.z, en.x, en.y, en.z) - en.w < 0) continue;  // behind edge
            float v1x = p0.x - p.x, v1y = p0.y - p.y, v1z = p0.z - p.z;
            float v2x = p1.x - p.x, v2y = p1.y - p.y, v2z = p1.z - p.z;
            if (!vnormalize(v1x, v1y, v1z)) continue;
            if (!vnormalize(v2x, v2y, v2z)) continue;
            const float ang = dot3(v1x, v1y, v1z, v2x, v2y, v2z);
            if (ang < best) {
              best = ang;
              bestp = i;
            }
          }
          for (int o = 16; o; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int op = __shfl_xor_sync(0xffffffffu, bestp, o);
            if (op >= 0 && (bestp < 0 || ob < best || (ob == best && op < bestp))) {
              best = ob;
              bestp = op;
            }
          }
          if (best >= 1 || bestp < 0) continue;  // edge doesn't match anything
          if (numtris >= kMaxTriTris) {
            if (lane == 0) atomicMax(a.error, kErrTris);
            fail = true;
            break;
          }
          const int nt = numtris++;
          const int e1 = find_edge(s, a.pmax, ep1, bestp, numedges, pnx, pny, pnz, lane, a.error);
          const int e2b = e1 < 0 ? -1 : find_edge(s, a.pmax, bestp, ep0, numedges, pnx, pny, pnz, lane, a.error);
          if (e1 < 0 || e2b < 0) {
            fail = true;
            break;
          }
          const int ea = find_edge(s, a.pmax, bestp, ep1, numedges, pnx, pny, pnz, lane, a.error);
          const int eb = find_edge(s, a.pmax, ep0, bestp, numedges, pnx, pny, pnz, lane, a.error);
          if (sp + 2 > kStackCap || ea < 0 || eb < 0) {
            if (lane == 0) atomicMax(a.error, kErrStack);
            fail = true;
            break;
          }
          if (lane == 0) {
            s.t_edges[nt * 3] = (uint16_t) e;
            s.t_edges[nt * 3 + 1] = (uint16_t) e1;
            s.t_edges[nt * 3 + 2] = (uint16_t) e2b;
            s.e_tri[e] = s.e_tri[e1] = s.e_tri[e2b] = (int16_t) nt;
            s.stack[sp] = (uint16_t) eb;       // visited second
            s.stack[sp + 1] = (uint16_t) ea;   // visited first
          }
          sp += 2;
          __syncwarp();
        }
      }
    } else if (P == 1) {
      if (lane == 0) s.pts[0] = a.p_origin_area[pidx[0]];
      __syncwarp();
    }

    // ---- sample the triangulation and pack bytes, lightmap.c:1152-1196 ----
    const int l0 = a.luxel_first[f], nlux = a.luxel_first[f + 1] - l0;
    int nst = a.fl_numstyles[f];
    if (nst > 4) nst = 4;  // MAXLIGHTMAPS
    const float minlight = a.f_minlight[f];
    for (int j = lane; j < nlux; j += 32) {
      const int lux = l0 + j;
      float add0 = 0.f, add1 = 0.f, add2 = 0.f;
      if (a.numbounce > 0) {
        const float px = a.origins[lux * 3], py = a.origins[lux * 3 + 1], pz = a.origins[lux * 3 + 2];
        if (P == 1) {
          const float *tl = a.p_totallight + (size_t) pidx[0] * 3;
          add0 = tl[0];
          add1 = tl[1];
          add2 = tl[2];
        } else if (P >= 2) {
          bool found = false;
          // SampleTriangulation, lightmap.c:369-439
          for (int t = 0; t < numtris && !found; t++) {
            const int e0 = s.t_edges[t * 3], e1 = s.t_edges[t * 3 + 1], e2 = s.t_edges[t * 3 + 2];
            const float4 n0 = s.e_plane[e0];
            if (dot3(n0.x, n0.y, n0.z, px, py, pz) - n0.w < 0) continue;
            const float4 n1 = s.e_plane[e1];
            if (dot3(n1.x, n1.y, n1.z, px, py, pz) - n1.w < 0) continue;
            const float4 n2 = s.e_plane[e2];
            if (dot3(n2.x, n2.y, n2.z, px, py, pz) - n2.w < 0) continue;
            found = true;
            // LerpTriangle, lightmap.c:315-349
            const int q1 = s.e_p0[e0], q2 = s.e_p0[e1], q3 = s.e_p0[e2];
            const float *b = a.p_totallight + (size_t) pidx[q1] * 3;
            const float *t2 = a.p_totallight + (size_t) pidx[q2] * 3;
            const float *t3 = a.p_totallight + (size_t) pidx[q3] * 3;
            const float d1x = t2[0] - b[0], d1y = t2[1] - b[1], d1z = t2[2] - b[2];
            const float d2x = t3[0] - b[0], d2y = t3[1] - b[1], d2z = t3[2] - b[2];
            const float x = dot3(px, py, pz, n0.x, n0.y, n0.z) - n0.w;
            const float y = dot3(px, py, pz, n2.x, n2.y, n2.z) - n2.w;
            const float4 o2 = s.pts[q2], o3 = s.pts[q3];
            const float y1 = dot3(o2.x, o2.y, o2.z, n2.x, n2.y, n2.z) - n2.w;
            const float x2 = dot3(o3.x, o3.y, o3.z, n0.x, n0.y, n0.z) - n0.w;
            if (fabs((double) y1) < 0.1 || fabs((double) x2) < 0.1) {
              add0 = b[0];
              add1 = b[1];
              add2 = b[2];
            } else {
              const double sx = (double) (x / x2), sy = (double) (y / y1);
              float c0 = (float) ((double) b[0] + sx * (double) d2x);
              float c1 = (float) ((double) b[1] + sx * (double) d2y);
              float c2 = (float) ((double) b[2] + sx * (double) d2z);
              add0 = (float) ((double) c0 + sy * (double) d1x);
              add1 = (float) ((double) c1 + sy * (double) d1y);
              add2 = (float) ((double) c2 + sy * (double) d1z);
            }
          }
          for (int e = 0; e < numedges && !found; e++) {  // search for exterior edge
            if (s.e_tri[e] >= 0) continue;
            const float4 en = s.e_plane[e];
            float d = dot3(px, py, pz, en.x, en.y, en.z) - en.w;
            if (d < 0) continue;
            const int q0 = s.e_p0[e], q1 = s.e_p1[e];
            const float4 o0 = s.pts[q0], o1 = s.pts[q1];
            float v1x = o1.x - o0.x, v1y = o1.y - o0.y, v1z = o1.z - o0.z;
            vnormalize(v1x, v1y, v1z);
            d = dot3(px - o0.x, py - o0.y, pz - o0.z, v1x, v1y, v1z);
            if (d < 0) continue;
            if (d > 1) continue;
            const float *t0 = a.p_totallight + (size_t) pidx[q0] * 3;
            const float *t1 = a.p_totallight + (size_t) pidx[q1] * 3;
            add0 = t0[0] + d * (t1[0] - t0[0]);
            add1 = t0[1] + d * (t1[1] - t0[1]);
            add2 = t0[2] + d * (t1[2] - t0[2]);
            found = true;
          }
          if (!found) {  // search for nearest point
            float best = 99999;
            int bp = -1;
            for (int q = 0; q < P; q++) {
              const float4 o = s.pts[q];
              const float vx = px - o.x, vy = py - o.y, vz = pz - o.z;
              const float d = (float) sqrt(((double) (vx * vx) + (double) (vy * vy)) + (double) (vz * vz));
              if (d < best) {
                best = d;
                bp = q;
              }
            }
            if (bp >= 0) {
              const float *t0 = a.p_totallight + (size_t) pidx[bp] * 3;
              add0 = t0[0];
              add1 = t0[1];
              add2 = t0[2];
            }
          }
        }
      }
      for (int st = 0; st < nst; st++) {
        const int slot = a.fl_styleslot[(size_t) f * 32 + st];
        const float *smp = a.samples + ((size_t) slot * a.L + lux) * 3;
        float lb0 = smp[0], lb1 = smp[1], lb2 = smp[2];
        if (a.numbounce > 0 && st == 0) {
          lb0 = lb0 + add0;
          lb1 = lb1 + add1;
          lb2 = lb2 + add2;
        }
        lb0 += a.ambient;
        lb1 += a.ambient;
        lb2 += a.ambient;
        lb0 = a.lightscale * lb0;
        lb1 = a.lightscale * lb1;
        lb2 = a.lightscale * lb2;
        if (lb0 < 1) lb0 = 1;
        if (lb1 < 1) lb1 = 1;
        if (lb2 < 1) lb2 = 1;
        if (a.lb_out && minlight > 0) {  // host applies the rand()-mottled _minlight (lightmap.c:1187-1189)
          float *o = a.lb_out + ((size_t) st * a.L + lux) * 3;
          o[0] = lb0;
          o[1] = lb1;
          o[2] = lb2;
        }
        float mx = lb0;
        if (lb1 > mx) mx = lb1;
        if (lb2 > mx) mx = lb2;
        float newmax = mx;
        if (newmax < 0) newmax = 0;
        if (newmax > a.maxlight) newmax = a.maxlight;
        uint8_t *dst = a.out + a.lightofs[f] + ((size_t) st * nlux + j) * 3;
        dst[0] = (uint8_t) (int) (lb0 * newmax / mx);
        dst[1] = (uint8_t) (int) (lb1 * newmax / mx);
        dst[2] = (uint8_t) (int) (lb2 * newmax / mx);
      }
    }
    __syncwarp();
  }
}

}  // namespace

// plane groups: the chain FinalLightFace walks from planelinks[side][planenum] (lightmap.c:1123); it stops at face index 0,
// so face 0 never contributes (appendix C).  Pure host work on what was uploaded.
std::shared_ptr<qrad_cuda_ctx::PlaneGroups> compute_plane_groups(const qrad_cuda_ctx *c) {
  auto pg = std::make_shared<qrad_cuda_ctx::PlaneGroups>();
  const int F = c->F;
  pg->group_of_face.assign((size_t) F, -1);
  pg->group_first.assign(1, 0);
  std::map<int32_t, int32_t> group_of_head;
  for (int f = 0; f < F; f++) {
    if (!c->h_lit[f]) continue;
    const int head = c->h_planelink_head[f];
    auto it = group_of_head.find(head);
    if (it == group_of_head.end()) {
      const int g = (int) pg->group_first.size() - 1;
      int guard = 0;
      for (int pf = head; pf; pf = c->h_facelinks[pf]) {
        if (pf < 0 || pf >= F || ++guard > F) {
          pg->error = "corrupt plane face links";
          return pg;
        }
        for (int pt = c->h_face_head[pf]; pt >= 0; pt = c->h_next[pt]) pg->group_patch.push_back(pt);
      }
      pg->group_first.push_back((int32_t) pg->group_patch.size());
      it = group_of_head.emplace(head, g).first;
    }
    pg->group_of_face[f] = it->second;
  }
  if (pg->group_patch.empty()) pg->group_patch.push_back(0);
  return pg;
}

void final_light_impl(qrad_cuda_ctx *c, const qrad_final_params *p, uint8_t *out, int32_t cap, int32_t *size,
                      int32_t *lightofs_out, uint8_t *styles_out) {
  if (!c->facelights_done) dfail("final_light: build_facelights has not run");
  face_range(c);
  Timer t;
  t.start(c->stream);
  cudaStream_t st = c->stream;
  DebugLaps laps("final_light", st);
  const int F = c->F, L = c->num_luxels;
  // lightofs in face order with the UNCLAMPED style count (lightmap.c:1086-1099), styles[] (1101-1102, 1158)
  std::vector<int32_t> lightofs(F, 0);
  int64_t total = 0;
  bool any_minlight = false;
  for (int f = 0; f < F; f++) {
    if (!c->h_lit[f]) {
      lightofs_out[f] = 0;
      memset(styles_out + f * 4, 0, 4);
      continue;
    }
    lightofs[f] = (int32_t) total;
    const int ns = c->h_luxel_first[f + 1] - c->h_luxel_first[f];
    total += (int64_t) c->h_fl_numstyles[f] * ns * 3;
    if (total > 0x200000 || total > cap) dfail("MAX_MAP_LIGHTING");
    uint8_t *sty = styles_out + f * 4;
    sty[0] = 0;
    sty[1] = sty[2] = sty[3] = 0xff;
    const int nst = std::min(c->h_fl_numstyles[f], 4);
    if (c->h_fl_numstyles[f] > 4) printf("face with too many lightstyles: face %d\n", f);
    for (int s = 0; s < nst; s++) sty[s] = (uint8_t) c->h_style_ids[c->h_fl_stylenums[(size_t) f * 32 + s]];
    lightofs_out[f] = lightofs[f];
    if (c->h_minlight[f] > 0) any_minlight = true;
  }
  *size = (int32_t) total;
  laps.lap("lightofs / styles");

  // plane groups: the chain FinalLightFace walks from planelinks[side][planenum] (lightmap.c:1123); it stops
  // at face index 0, so face 0 never contributes (appendix C)
  // device buffers live in the context so that repeated calls reuse the allocations
  DBuf<int32_t> &d_group_of_face = c->fin_i32[0], &d_group_first = c->fin_i32[1], &d_group_patch = c->fin_i32[2],
                &tri_count = c->fin_i32[3], &tri_points = c->fin_i32[4];
  DBuf<int32_t> &d_pmax = c->fin_i32[10];
  int pmax = 1;
  tri_count.alloc((size_t) F);
  tri_count.zero(st);
  if (p->numbounce > 0) {
    std::shared_ptr<qrad_cuda_ctx::PlaneGroups> pg;
    if (c->groups_future.valid()) pg = c->groups_future.get();   // started by upload_faces / build_facelights
    else pg = compute_plane_groups(c);
    c->groups_future = {};
    if (!pg->error.empty()) dfail("%s", pg->error.c_str());
    laps.lap("plane groups (wait)");
    d_group_of_face.upload(pg->group_of_face, st);
    d_group_first.upload(pg->group_first, st);
    d_group_patch.upload(pg->group_patch, st);
    tri_points.alloc((size_t) F * kMaxTriPoints + 1);
    d_pmax.alloc(1);
    d_pmax.zero(st);
    PointArgs pa{};
    pa.F = F;
    pa.face_lo = c->face_lo;
    pa.face_hi = c->face_hi;
    pa.f_lit = c->f_lit.p;
    pa.group_of_face = d_group_of_face.p;
    pa.group_first = d_group_first.p;
    pa.group_patch = d_group_patch.p;
    pa.p_origin_area = c->p_origin_area.p;
    pa.f_bounds = c->f_bounds.p;
    pa.subdiv2 = p->subdiv * 2;
    pa.tri_count = tri_count.p;
    pa.tri_points = tri_points.p;
    pa.pmax = d_pmax.p;
    if (c->face_hi > c->face_lo) {
      k_tri_points<<<c->face_hi - c->face_lo, 256, 0, st>>>(pa);
      c->launch_check("k_tri_points");
    }
    d_pmax.download(&pmax, 1, st);
    if (pmax > kMaxTriPoints) dfail("trian->numpoints == MAX_TRI_POINTS");
    pmax = std::max(pmax, 1);
    laps.lap("tri points");
  } else {
    tri_points.alloc(1);
  }

  // face style slots
  DBuf<int32_t> &d_numstyles = c->fin_i32[5], &d_styleslot = c->fin_i32[6], &d_lightofs = c->fin_i32[7];
  d_numstyles.upload(c->h_fl_numstyles, st);
  d_styleslot.upload(c->h_fl_stylenums, st);
  d_lightofs.upload(lightofs, st);
  DBuf<uint8_t> &d_out = c->fin_out;
  d_out.alloc((size_t) total + 4);
  d_out.zero(st);
  DBuf<float> &lb_out = c->fin_lb;
  if (any_minlight) {
    lb_out.alloc((size_t) 4 * L * 3 + 1);
    if (c->world > 1) lb_out.zero(st);
  }
  DBuf<int32_t> &d_err = c->fin_i32[8], &d_counter = c->fin_i32[9];
  d_err.alloc(1);
  d_err.zero(st);
  d_counter.upload(&c->face_lo, 1, st);   // k_final draws faces from here

  // the triangulation is a serial walk per face (lane 0 keeps the edge / triangle books): latency, not throughput, so as
  // many faces in flight as the scratch pools allow
  int per_sm = 6;
  if (const char *e = getenv("QRAD_FINAL_BLOCKS")) per_sm = std::max(1, atoi(e));
  while (per_sm > 2 && (size_t) c->sm_count * per_sm * kWarps * pmax * pmax * 2 > ((size_t) 2 << 30)) per_sm--;   // edge matrices <= 2 GB
  const int blocks = std::max(1, std::min((c->face_hi - c->face_lo + kWarps - 1) / kWarps, c->sm_count * per_sm));
  const size_t nw = (size_t) blocks * kWarps;
  // scratch pools live in the context so that repeated calls reuse the allocations
  DBuf<uint16_t> &matrix_pool = c->fin_matrix, &ep0 = c->fin_ep0, &ep1 = c->fin_ep1, &tedges = c->fin_tedges, &stack = c->fin_stack;
  DBuf<int16_t> &etri = c->fin_etri;
  DBuf<float4> &pts_pool = c->fin_pts, &eplane = c->fin_eplane;
  matrix_pool.alloc(nw * pmax * pmax);
  pts_pool.alloc(nw * pmax);
  eplane.alloc(nw * (kMaxTriEdges + 2));
  ep0.alloc(nw * (kMaxTriEdges + 2));
  ep1.alloc(nw * (kMaxTriEdges + 2));
  etri.alloc(nw * (kMaxTriEdges + 2));
  tedges.alloc(nw * kMaxTriTris * 3);
  stack.alloc(nw * kStackCap);
  laps.lap("uploads, pools");

  FinalArgs fa{};
  fa.F = F;
  fa.face_hi = c->face_hi;
  fa.pmax = pmax;
  fa.f_lit = c->f_lit.p;
  fa.luxel_first = c->f_luxel_first.p;
  fa.tri_count = tri_count.p;
  fa.tri_points = tri_points.p;
  fa.p_origin_area = c->p_origin_area.p;
  fa.p_totallight = c->p_totallight.p;
  fa.f_plane_normal = c->f_plane_normal.p;
  fa.origins = c->fl_surfpt.p;
  fa.samples = c->fl_samples.p;
  fa.L = L;
  fa.fl_numstyles = d_numstyles.p;
  fa.fl_styleslot = d_styleslot.p;
  fa.lightofs = d_lightofs.p;
  fa.f_minlight = c->f_minlight.p;
  fa.ambient = p->ambient;
  fa.maxlight = p->maxlight;
  fa.lightscale = p->lightscale;
  fa.numbounce = p->numbounce;
  fa.out = d_out.p;
  fa.lb_out = any_minlight ? lb_out.p : nullptr;
  fa.error = d_err.p;
  fa.matrix_pool = matrix_pool.p;
  fa.pts_pool = pts_pool.p;
  fa.eplane_pool = eplane.p;
  fa.ep0_pool = ep0.p;
  fa.ep1_pool = ep1.p;
  fa.tedges_pool = tedges.p;
  fa.stack_pool = stack.p;
  fa.etri_pool = etri.p;
  fa.face_counter = d_counter.p;
  k_final<<<blocks, kWarps * 32, 0, st>>>(fa);
  c->launch_check("k_final");
  laps.lap("k_final");
  if (c->world > 1) {   // all ranks must agree on failure before anybody enters the sums below
    ncclResult_t r = nccl().AllReduce(d_err.p, d_err.p, 1, ncclInt32, ncclMax, c->comm, st);
    if (r != ncclSuccess) dfail("ncclAllReduce(error flag) failed: %s", nccl().GetErrorString(r));
  }
  int err = 0;
  d_err.download(&err, 1, st);
  if (err == kErrEdges) dfail("trian->numedges > MAX_TRI_EDGES-2");
  if (err == kErrTris) dfail("trian->numtris >= MAX_TRI_TRIS");
  if (err) dfail("triangulation stack overflow");
  if (c->world > 1 && total) {   // every face was finished by one rank; the others hold zeros there
    const NcclApi &n = nccl();
    ncclResult_t r = n.AllReduce(d_out.p, d_out.p, (size_t) total, ncclUint8, ncclSum, c->comm, st);
    if (r == ncclSuccess && any_minlight)
      r = n.AllReduce(lb_out.p, lb_out.p, (size_t) 4 * L * 3, ncclFloat, ncclSum, c->comm, st);
    if (r != ncclSuccess) dfail("ncclAllReduce(lightmap) failed: %s", n.GetErrorString(r));
  }
  if (total) d_out.download(out, (size_t) total, st);
  laps.lap("all-reduce, read-back");

  if (any_minlight) {
    // _minlight faces: the mottling uses the C library rand() in face / style / luxel order on one thread
    // (lightmap.c:1187-1189), so those faces are finished here with the same call sequence.
    std::vector<float> lb((size_t) 4 * L * 3 + 1);
    lb_out.download(lb.data(), lb.size(), st);
    // the reference process never seeds: rand() is glibc's TYPE_3 additive generator from seed 1.  A private
    // random_r state with the same parameters reproduces it without touching the host program's rand() state.
    char rstate[128];
    struct random_data rdata;
    memset(&rdata, 0, sizeof rdata);
    memset(rstate, 0, sizeof rstate);
    initstate_r(1, rstate, sizeof rstate, &rdata);
    for (int f = 0; f < F; f++) {
      if (!c->h_lit[f] || !(c->h_minlight[f] > 0)) continue;
      const float minlight = c->h_minlight[f];
      const int l0 = c->h_luxel_first[f], nlux = c->h_luxel_first[f + 1] - l0;
      const int nst = std::min(c->h_fl_numstyles[f], 4);
      for (int s = 0; s < nst; s++)
        for (int j = 0; j < nlux; j++) {
          const float *v = &lb[((size_t) s * L + l0 + j) * 3];
          float mx = v[0];
          if (v[1] > mx) mx = v[1];
          if (v[2] > mx) mx = v[2];
          float newmax = mx;
          if (newmax < 0) newmax = 0;
          if (newmax < minlight) {
            int32_t r = 0;
            random_r(&rdata, &r);
            newmax = minlight + (r % 48);
          }
          if (newmax > p->maxlight) newmax = p->maxlight;
          uint8_t *dst = out + lightofs[f] + ((size_t) s * nlux + j) * 3;
          for (int k = 0; k < 3; k++) dst[k] = (uint8_t) (int) (v[k] * newmax / mx);
        }
    }
  }
  c->stats.ms_final = t.stop();
}

}  // namespace qrad
