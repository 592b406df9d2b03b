// The BSP walkers and the exact shortcuts around them: TestLine_r, PointInLeafnum, the common start node of a bundle of
// rays, and the cell proofs.  Device code of libqrad_cuda (included by qrad_device.cuh) that ALSO compiles for the host:
// host_walkers.cpp builds the very same functions with g++ so that the CPU tests can hold them against the reference's
// golden rays and against each other (tests/test_walkers_cpu.py) -- the proofs must never call a clear ray blocked.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>

#if defined(__CUDACC__)
#define QRAD_HD __host__ __device__ __forceinline__
#define QRAD_HD_NOINLINE __host__ __device__ __noinline__
#else
#define QRAD_HD inline
#define QRAD_HD_NOINLINE
#endif
#ifdef __CUDA_ARCH__
#define QRAD_LDG(p) __ldg(p)
#define QRAD_TRAP() __trap()
#else
#define QRAD_LDG(p) (*(p))
#define QRAD_TRAP() abort()
#endif

namespace qrad {

QRAD_HD float qrad_int_as_float(const int v) {
#ifdef __CUDA_ARCH__
  return __int_as_float(v);
#else
  float f;
  memcpy(&f, &v, 4);
  return f;
#endif
}
QRAD_HD int qrad_float_as_int(const float f) {
#ifdef __CUDA_ARCH__
  return __float_as_int(f);
#else
  int v;
  memcpy(&v, &f, 4);
  return v;
#endif
}

// ---- trace tree ------------------------------------------------------------------
// One 16-byte record per BSP node {dist, plane type, child0, child1} so that a node visit is ONE 128-bit
// load; the plane normal (needed only by non-axial planes in TestLine_r and by PointInLeaf) sits in a
// separate float4 array.  Nodes are numbered breadth-first: the two children of a node are adjacent
// (same 32-byte sector) and the top levels, which every ray visits, share a few cache lines (the whole tree is
// <= 2 MB and lives in L1/L2; staging its top in shared memory was measured and rejected).  Layout is
// free, only results matter (MakeTnode, trace.c:47-71, uses preorder).
// child >= 0: node index.  child < 0: leaf, encoded kLeafFlag | solid << 30 | leaf number.
struct Tree {
  const int4 *nodes;      // x = float bits of dist, y = plane type, z = child 0, w = child 1
  const float4 *normals;  // xyz = plane normal, w = dist
};
constexpr int kLeafFlag = int(0x80000000u);
constexpr int kSolidBit = 1 << 30;
constexpr int kLeafMask = 0x3fffffff;
constexpr int kTraceStack = 64;    // pending far halves; deeper trees use the big-stack instantiation
constexpr int kTraceStackBig = 256;

// Distances of both segment ends to a NON-axial plane (trace.c:119-121).  Kept out of line on purpose: when it is
// inlined ptxas if-converts it and every node visit issues the (predicated-off) dot products and the normal load
// even on maps whose planes are all axial -- 10 % of k_visibility's instructions in profiles/r01f_full_c5chop16.txt.
static QRAD_HD_NOINLINE float2 nonaxial_dists(const float4 *__restrict__ normals, const int node, const float dist,
                                              const float sx, const float sy, const float sz, const float ex,
                                              const float ey, const float ez) {
  const float4 pl = QRAD_LDG(normals + node);
  return make_float2((sx * pl.x + sy * pl.y + sz * pl.z) - dist, (ex * pl.x + ey * pl.y + ez * pl.z) - dist);
}

// Coordinate `type` (0, 1, 2) of a point, as two selects.  Written in PTX on purpose: from the C ternaries ptxas builds
// a three-way branch, and since neighbouring lanes walk different paths every node visit then diverges by plane axis
// (profiles/r01g_full_c5chop16.txt: 25.4 of 32 lanes active, the ALU pipe the busiest at 58 %).
QRAD_HD float axis_pick(const int type, const float x, const float y, const float z) {
#ifdef __CUDA_ARCH__
  float r;
  asm("{\n\t.reg .pred p0, p1;\n\t.reg .f32 t;\n\t"
      "setp.eq.s32 p0, %1, 0;\n\tsetp.eq.s32 p1, %1, 1;\n\t"
      "selp.f32 t, %3, %4, p1;\n\tselp.f32 %0, %2, t, p0;\n\t}"
      : "=f"(r) : "r"(type), "f"(x), "f"(y), "f"(z));
  return r;
#else
  return type == 0 ? x : (type == 1 ? y : z);
#endif
}

// TestLine_r(0, start, stop), trace.c:92-142, as an explicit-stack walk.  The near half is
// followed first and the far half (node, stop) is parked; on a pop the new start is the
// current stop (= the split point), exactly the segments the recursion passes down.
// Compares: `front >= -ON_EPSILON` with ON_EPSILON the double 0.1 (trace.c:27) is, for a float
// front, the same predicate as `front > -0.1f`, and `front < ON_EPSILON` the same as
// `front < 0.1f` (-0.1f < -0.1 and 0.1f > 0.1; checked exhaustively around the constants).
// Returns 0 (clear) or 1 (blocked); *blocker, when given, receives the solid leaf that ended the walk.
// `node0`: where to start.  Any node that the walk from the root provably reaches with the UNCLIPPED segment will do
// (common_start_node() below): the nodes above it only pass the segment on.
template <int STACK, bool COUNT>
QRAD_HD int test_line(const Tree tree, float sx, float sy, float sz, float ex, float ey, float ez,
                                         unsigned &visits, int *blocker = nullptr, const int node0 = 0) {
  float4 stack[STACK];
  int sp = 0;
  int node = node0;
  for (;;) {
    while (node >= 0) {
      const int4 nd = QRAD_LDG(tree.nodes + node);
      const float dist = qrad_int_as_float(nd.x);
      if (COUNT) visits++;
      float front, back;
      if (nd.y < 3) {
        front = axis_pick(nd.y, sx, sy, sz) - dist;
        back = axis_pick(nd.y, ex, ey, ez) - dist;
      } else {
        const float2 fb = nonaxial_dists(tree.normals, node, dist, sx, sy, sz, ex, ey, ez);
        front = fb.x;
        back = fb.y;
      }
      if (front > -0.1f && back > -0.1f) {
        node = nd.z;
        continue;
      }
      if (front < 0.1f && back < 0.1f) {
        node = nd.w;
        continue;
      }
      const bool side = front < 0;
      const float frac = front / (front - back);
      const float mx = sx + (ex - sx) * frac;
      const float my = sy + (ey - sy) * frac;
      const float mz = sz + (ez - sz) * frac;
      if (sp < STACK) stack[sp] = make_float4(qrad_int_as_float(side ? nd.z : nd.w), ex, ey, ez);
      sp++;
      ex = mx;
      ey = my;
      ez = mz;
      node = side ? nd.w : nd.z;
    }
    if (node & kSolidBit) {
      if (blocker) *blocker = node & kLeafMask;
      return 1;
    }
    if (sp == 0) return 0;
    sp--;
    if (sp >= STACK) QRAD_TRAP();   // deeper than the uploaded tree allows (upload_bsp checks the depth): never a wrong answer
    const float4 t = stack[sp];
    sx = ex;
    sy = ey;
    sz = ez;
    node = qrad_float_as_int(t.x);
    ex = t.y;
    ey = t.z;
    ez = t.w;
  }
}

// The deepest node that TestLine_r(0, s, e) reaches with the unclipped segment for EVERY s in box A and e in box B:
// while both boxes lie on one side of the node's (axial) plane far enough that every segment passes the same one-sided
// test of trace.c:124-128, the walk just descends.  `coordinate - dist` is one rounded subtraction, monotone in the
// coordinate, so the box ends bound front / back of every segment exactly, with the very predicates test_line uses:
//   first test  (front > -0.1f && back > -0.1f) for all  <=>  both box minima pass it            -> child 0
//   second test (front <  0.1f && back <  0.1f) for all, and the first test fails for all: guaranteed when additionally
//   one whole box lies at or below -0.1 (then front or back of every segment fails the first test)  -> child 1
// Starting the walk there instead of at the root changes no result.
QRAD_HD int common_start_node(const Tree tree, const float4 amn, const float4 amx, const float4 bmn,
                                                 const float4 bmx) {
  int node = 0;
  for (;;) {
    const int4 nd = QRAD_LDG(tree.nodes + node);
    if (nd.y >= 3) return node;
    const float dist = qrad_int_as_float(nd.x);
    const float a0 = axis_pick(nd.y, amn.x, amn.y, amn.z) - dist, a1 = axis_pick(nd.y, amx.x, amx.y, amx.z) - dist;
    const float b0 = axis_pick(nd.y, bmn.x, bmn.y, bmn.z) - dist, b1 = axis_pick(nd.y, bmx.x, bmx.y, bmx.z) - dist;
    int next;
    if (a0 > -0.1f && b0 > -0.1f) next = nd.z;
    else if (a1 < 0.1f && b1 < 0.1f && (!(a1 > -0.1f) || !(b1 > -0.1f))) next = nd.w;
    else return node;
    if (next < 0) return node;   // a leaf: let the walk itself read it
    node = next;
  }
}

// Shaft proof: does EVERY segment from a point of box A = [amn, amx] to a point of box B = [bmn, bmx] make TestLine_r
// return 1 because of solid leaf X?  X must be a leaf whose ancestors are all axial planes, so that its cell is the box
// [clo, chi] (clo.w = 1 marks such a solid leaf).  Sufficient condition used here: the segment contains a point q that
// lies inside the cell by more than `margin` > ON_EPSILON on every side.  Then TestLine_r(0, s, e) reaches X:
//  * at an ancestor where both ends pass the first test (front, back >= -0.1) the plane distance of q, a convex
//    combination of front and back, is >= -0.1 too, so X cannot lie behind that plane by more than 0.1: the walk turns
//    towards X; the same for the second test (< 0.1);
//  * at a split, mid lies on the plane (up to float rounding) and q, more than `margin` away from it, stays on the half
//    that is handed to the child on X's side; the recursion visits both halves unless the first already returned 1.
// By induction the sub-segment handed down always contains q (up to the accumulated rounding of the clip points,
// ~1e-3 per split at |x| <= 8192; `margin` = 0.1 + slack covers hundreds of splits) until leaf X itself is reached.
// The box version: A and B lie on opposite sides of the (shrunken) cell along some axis k; every segment then crosses
// the plane x_k = c through the middle of the cell at parameter t in [tmin, tmax] (t is monotone in a_k and b_k), and the
// other two coordinates of the crossing point, convex combinations of a_j and b_j, are bounded by their values at the
// interval ends; if that rectangle lies inside the shrunken cell, every segment passes through it.  No tree walk.
QRAD_HD bool shaft_blocked(const float4 clo, const float4 chi, const float margin, const float4 amn,
                                              const float4 amx, const float4 bmn, const float4 bmx) {
  if (clo.w == 0.f) return false;
  const float lo[3] = {clo.x + margin, clo.y + margin, clo.z + margin};
  const float hi[3] = {chi.x - margin, chi.y - margin, chi.z - margin};
  const float an[3] = {amn.x, amn.y, amn.z}, ax[3] = {amx.x, amx.y, amx.z};
  const float bn[3] = {bmn.x, bmn.y, bmn.z}, bx[3] = {bmx.x, bmx.y, bmx.z};
  bool ok = false;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const bool below = ax[k] < lo[k] && bn[k] > hi[k];   // A below the cell, B above it
    const bool above = bx[k] < lo[k] && an[k] > hi[k];
    if (!(below || above) || !(hi[k] > lo[k])) continue;
    const float cc = 0.5f * (lo[k] + hi[k]);
    // t = (c - a_k) / (b_k - a_k): smallest / largest over the boxes
    float tmin, tmax;
    if (below) {
      tmin = (cc - ax[k]) / (bx[k] - ax[k]);
      tmax = (cc - an[k]) / (bn[k] - an[k]);
    } else {
      tmin = (cc - an[k]) / (bn[k] - an[k]);
      tmax = (cc - ax[k]) / (bx[k] - ax[k]);
    }
    tmin = fmaxf(tmin - 1e-5f, 0.f);
    tmax = fminf(tmax + 1e-5f, 1.f);
    bool inside = true;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      if (j == k) continue;
      const float q0 = fminf(an[j] + (bn[j] - an[j]) * tmin, an[j] + (bn[j] - an[j]) * tmax) - 0.01f;
      const float q1 = fmaxf(ax[j] + (bx[j] - ax[j]) * tmin, ax[j] + (bx[j] - ax[j]) * tmax) + 0.01f;
      inside = inside && q0 >= lo[j] && q1 <= hi[j];
    }
    ok = ok || inside;
  }
  return ok;
}

// The same proof for ONE segment s -> e, complete for a box cell: the segment is clipped against the shrunken cell (slab
// method) and the middle of the clipped piece is taken as the witness q; q is then CHECKED to lie inside the shrunken
// cell, so inaccuracies of the clipping can only lose a proof, never invent one (q, evaluated in float, is within ~1e-3 of
// the exact segment; `margin` has 0.15 of slack).
QRAD_HD bool segment_through_cell(const float4 clo, const float4 chi, const float margin, const float4 s,
                                                     const float4 e) {
  if (clo.w == 0.f) return false;
  const float lo[3] = {clo.x + margin, clo.y + margin, clo.z + margin};
  const float hi[3] = {chi.x - margin, chi.y - margin, chi.z - margin};
  const float p[3] = {s.x, s.y, s.z}, d[3] = {e.x - s.x, e.y - s.y, e.z - s.z};
  float t0 = 0.f, t1 = 1.f;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    if (d[k] != 0.f) {
      const float inv = 1.0f / d[k];
      const float ta = (lo[k] - p[k]) * inv, tb = (hi[k] - p[k]) * inv;
      t0 = fmaxf(t0, fminf(ta, tb));
      t1 = fminf(t1, fmaxf(ta, tb));
    }
  }
  if (!(t0 <= t1)) return false;
  const float t = 0.5f * (t0 + t1);
  bool inside = true;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float q = p[k] + d[k] * t;
    inside = inside && q >= lo[k] && q <= hi[k];
  }
  return inside;
}

// PointInLeafnum, qrad3.c:131-150: full dot product on every node, `dist > 0` picks child 0.
QRAD_HD int point_in_leaf(const Tree tree, float x, float y, float z) {
  int node = 0;
  while (node >= 0) {
    const float4 pl = QRAD_LDG(tree.normals + node);
    const int4 nd = QRAD_LDG(tree.nodes + node);
    const float d = (x * pl.x + y * pl.y + z * pl.z) - pl.w;
    node = d > 0 ? nd.z : nd.w;
  }
  return node & kLeafMask;
}

}  // namespace qrad
