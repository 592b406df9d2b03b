/* TEST INFRASTRUCTURE ONLY (oracle/): plain-C restatement of the second stage of the reference's qvis3 -- PortalFlow,
 * src/tools/qvis3/flow.c:112-523 -- for the next stage outward of this project.  Nothing in the
 * product links or calls it.  Pinned by tests/test_qvis_oracle_cpu.py against per-portal dumps of the reference itself
 * (oracle/qvis_harness.c).
 *
 * One call computes the portalvis vector of ONE portal k.  The reference finishes the portals one after the other in
 * the order of SortPortals, and a portal that is crossed contributes its final vector when it is finished and its flood
 * vector otherwise (flow.c:391-396); here "finished" is DEFINED as rank[j] < rank[k], so the call depends on nothing but
 * its arguments -- which is the schedule a parallel version needs, and what the tests check against the reference.
 *
 * Arithmetic: vec_t is float; products and sums are separately rounded floats in the reference's order (compile with
 * -ffp-contract=off, baseline x86-64); comparisons against ON_EPSILON are against the DOUBLE 0.1 (vis.h:31).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ON_EPS 0.1
#define FIXED_POINTS 12 /* MAX_POINTS_ON_FIXED_WINDING, vis.h:40 */
#define MAX_POINTS 64   /* MAX_POINTS_ON_WINDING */

typedef struct { float n[3], d; } plane_t;
typedef struct { int np; float (*p)[3]; } wind_t; /* a view: the points of an original winding, or a frame's slot */

typedef struct {
  int P, nc, pb;
  const float *planes;          /* [P][4] */
  const int32_t *leaf;          /* [P] cluster behind the portal */
  const wind_t *wind;           /* [P] */
  const float *origin, *radius; /* bounding spheres of the windings */
  const int32_t *leaf_first, *leaf_portals;
  const uint8_t *flood, *done_vis;
  const int32_t *rank;
  int base;
  uint8_t *vis; /* the base portal's vector, growing */
  long chains;
} ctx_t;

typedef struct frame_s {
  uint8_t *might;
  const wind_t *source, *pass;
  wind_t slot[3];
  float store[3][FIXED_POINTS][3];
  int free_slot[3];
  plane_t portalplane;
} frame_t;

static float dot3(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

static wind_t *take_slot(frame_t *f) {
  for (int i = 0; i < 3; i++)
    if (f->free_slot[i]) {
      f->free_slot[i] = 0;
      return &f->slot[i];
    }
  return 0; /* the reference stops with an error here; it never happens with three slots */
}
static void give_slot(frame_t *f, const wind_t *w) {
  for (int i = 0; i < 3; i++)
    if (w == &f->slot[i]) f->free_slot[i] = 1;
}

/* ChopWindingVis, flow.c:112-206: the part of `in` in front of `split`; `in` itself when nothing is behind (or when the
 * result would not fit a fixed winding), NULL when nothing is in front */
static const wind_t *chop(const wind_t *in, frame_t *f, const plane_t *split) {
  float dists[MAX_POINTS * 2]; /* (128 in the reference) */
  int sides[MAX_POINTS * 2], counts[3] = {0, 0, 0};
  int i;
  for (i = 0; i < in->np; i++) {
    float d = dot3(in->p[i], split->n);
    d -= split->d;
    dists[i] = d;
    sides[i] = (double) d > ON_EPS ? 0 : ((double) d < -ON_EPS ? 1 : 2);
    counts[sides[i]]++;
  }
  if (!counts[1]) return in;
  if (!counts[0]) {
    give_slot(f, in);
    return 0;
  }
  sides[i] = sides[0];
  dists[i] = dists[0];
  wind_t *out = take_slot(f);
  out->np = 0;
  for (i = 0; i < in->np; i++) {
    const float *p1 = in->p[i];
    if (out->np == FIXED_POINTS) {
      give_slot(f, out);
      return in;
    }
    if (sides[i] == 2) {
      memcpy(out->p[out->np++], p1, 12);
      continue;
    }
    if (sides[i] == 0) memcpy(out->p[out->np++], p1, 12);
    if (sides[i + 1] == 2 || sides[i + 1] == sides[i]) continue;
    if (out->np == FIXED_POINTS) {
      give_slot(f, out);
      return in;
    }
    const float *p2 = in->p[(i + 1) % in->np];
    const float t = dists[i] / (dists[i] - dists[i + 1]);
    float mid[3];
    for (int j = 0; j < 3; j++) {
      if (split->n[j] == 1) mid[j] = split->d;
      else if (split->n[j] == -1) mid[j] = -split->d;
      else mid[j] = p1[j] + t * (p2[j] - p1[j]);
    }
    memcpy(out->p[out->np++], mid, 12);
  }
  give_slot(f, in);
  return out;
}

/* ClipToSeperators, flow.c:224-348 */
static const wind_t *clip_to_separators(const wind_t *source, const wind_t *pass, const wind_t *target, int flipclip, frame_t *f) {
  for (int i = 0; i < source->np; i++) {
    const int l = (i + 1) % source->np;
    float v1[3], v2[3];
    for (int c = 0; c < 3; c++) v1[c] = source->p[l][c] - source->p[i][c];
    for (int j = 0; j < pass->np; j++) {
      for (int c = 0; c < 3; c++) v2[c] = pass->p[j][c] - source->p[i][c];
      plane_t pl;
      pl.n[0] = v1[1] * v2[2] - v1[2] * v2[1];
      pl.n[1] = v1[2] * v2[0] - v1[0] * v2[2];
      pl.n[2] = v1[0] * v2[1] - v1[1] * v2[0];
      float length = pl.n[0] * pl.n[0] + pl.n[1] * pl.n[1] + pl.n[2] * pl.n[2];
      if ((double) length < ON_EPS) continue;
      length = (float) (1 / sqrt((double) length));
      pl.n[0] *= length;
      pl.n[1] *= length;
      pl.n[2] *= length;
      pl.d = dot3(pass->p[j], pl.n);
      /* which side of the candidate holds the source portal */
      int fliptest = 0, k;
      for (k = 0; k < source->np; k++) {
        if (k == i || k == l) continue;
        const float d = dot3(source->p[k], pl.n) - pl.d;
        if ((double) d < -ON_EPS) {
          fliptest = 0;
          break;
        } else if ((double) d > ON_EPS) {
          fliptest = 1;
          break;
        }
      }
      if (k == source->np) continue; /* planar with the source portal */
      if (fliptest) {
        for (int c = 0; c < 3; c++) pl.n[c] = 0.f - pl.n[c];
        pl.d = -pl.d;
      }
      /* a separating plane has every other point of pass on its positive side */
      int front = 0;
      for (k = 0; k < pass->np; k++) {
        if (k == j) continue;
        const float d = dot3(pass->p[k], pl.n) - pl.d;
        if ((double) d < -ON_EPS) break;
        else if ((double) d > ON_EPS) front++;
      }
      if (k != pass->np) continue;
      if (!front) continue;
      if (flipclip) {
        for (int c = 0; c < 3; c++) pl.n[c] = 0.f - pl.n[c];
        pl.d = -pl.d;
      }
      target = chop(target, f, &pl);
      if (!target) return 0;
    }
  }
  return target;
}

static int has_bit(const uint8_t *v, int i) { return (v[i >> 3] >> (i & 7)) & 1; }

/* RecursiveLeafFlow, flow.c:358-484 */
static void leaf_flow(ctx_t *c, int leafnum, const frame_t *head, const frame_t *prev) {
  frame_t st;
  uint8_t might[c->pb];
  st.might = might;
  for (int k = 0; k < 3; k++) {
    st.slot[k].np = 0;
    st.slot[k].p = st.store[k];
  }
  c->chains++;
  for (int e = c->leaf_first[leafnum]; e < c->leaf_first[leafnum + 1]; e++) {
    const int pnum = c->leaf_portals[e];
    if (!has_bit(prev->might, pnum)) continue;
    /* a finished portal contributes what it really sees, an unfinished one what it might see */
    const uint8_t *test = (c->rank[pnum] < c->rank[c->base] ? c->done_vis : c->flood) + (size_t) pnum * c->pb;
    int more = 0;
    for (int j = 0; j < c->pb; j++) {
      might[j] = prev->might[j] & test[j];
      more |= might[j] & ~c->vis[j];
    }
    if (!more && has_bit(c->vis, pnum)) continue; /* nothing new behind it */
    const float *pp = c->planes + 4 * pnum;
    memcpy(st.portalplane.n, pp, 12);
    st.portalplane.d = pp[3];
    plane_t back;
    for (int k = 0; k < 3; k++) back.n[k] = 0.f - pp[k];
    back.d = -pp[3];
    st.free_slot[0] = st.free_slot[1] = st.free_slot[2] = 1;
    float d = dot3(c->origin + 3 * pnum, head->portalplane.n);
    d -= head->portalplane.d;
    if (d < -c->radius[pnum]) continue;
    else if (d > c->radius[pnum]) st.pass = &c->wind[pnum];
    else {
      st.pass = chop(&c->wind[pnum], &st, &head->portalplane);
      if (!st.pass) continue;
    }
    d = dot3(c->origin + 3 * c->base, pp);
    d -= pp[3];
    if (d > c->radius[pnum]) continue;
    else if (d < -c->radius[pnum]) st.source = prev->source;
    else {
      st.source = chop(prev->source, &st, &back);
      if (!st.source) continue;
    }
    if (!prev->pass) { /* the second leaf can only be blocked if coplanar */
      c->vis[pnum >> 3] |= (uint8_t) (1 << (pnum & 7));
      leaf_flow(c, c->leaf[pnum], head, &st);
      continue;
    }
    st.pass = clip_to_separators(st.source, prev->pass, st.pass, 0, &st);
    if (!st.pass) continue;
    st.pass = clip_to_separators(prev->pass, st.source, st.pass, 1, &st);
    if (!st.pass) continue;
    c->vis[pnum >> 3] |= (uint8_t) (1 << (pnum & 7));
    leaf_flow(c, c->leaf[pnum], head, &st);
  }
}

/* SetPortalSphere, qvis3.c:272-298 */
static void sphere(const wind_t *w, float *origin, float *radius) {
  float total[3] = {0, 0, 0};
  for (int i = 0; i < w->np; i++)
    for (int c = 0; c < 3; c++) total[c] = total[c] + w->p[i][c];
  for (int c = 0; c < 3; c++) total[c] /= w->np;
  float best = 0;
  for (int i = 0; i < w->np; i++) {
    float dv[3];
    for (int c = 0; c < 3; c++) dv[c] = w->p[i][c] - total[c];
    double len = 0;
    for (int c = 0; c < 3; c++) len += dv[c] * dv[c];
    const float r = (float) sqrt(len);
    if (r > best) best = r;
  }
  memcpy(origin, total, 12);
  *radius = best;
}

/* PortalFlow for portal `base`, flow.c:497-523.  Windings: points[wfirst[k] .. wfirst[k + 1]) of portal k.  vis_out:
 * portalbytes, cleared here.  Returns the number of chains (RecursiveLeafFlow calls), -1 when out of memory. */
long qvo_portal_flow(int P, int nc, int pb, const float *planes, const int32_t *leaf, const int32_t *wfirst, float *points,
                     const int32_t *leaf_first, const int32_t *leaf_portals, const uint8_t *flood, const uint8_t *done_vis,
                     const int32_t *rank, int base, uint8_t *vis_out) {
  wind_t *wind = malloc((size_t) P * sizeof(wind_t));
  float *origin = malloc((size_t) P * 12), *radius = malloc((size_t) P * 4);
  if (!wind || !origin || !radius) return -1;
  for (int k = 0; k < P; k++) {
    wind[k].np = wfirst[k + 1] - wfirst[k];
    wind[k].p = (float(*)[3]) (points + 3 * (size_t) wfirst[k]);
    sphere(&wind[k], origin + 3 * k, radius + k);
  }
  ctx_t c = {P, nc, pb, planes, leaf, wind, origin, radius, leaf_first, leaf_portals, flood, done_vis, rank, base, vis_out, 0};
  memset(vis_out, 0, (size_t) pb);
  frame_t head;
  memset(&head, 0, sizeof head);
  uint8_t might[pb];
  memcpy(might, flood + (size_t) base * pb, (size_t) pb);
  head.might = might;
  head.source = &wind[base];
  head.pass = 0;
  memcpy(head.portalplane.n, planes + 4 * base, 12);
  head.portalplane.d = planes[4 * base + 3];
  leaf_flow(&c, leaf[base], &head, &head);
  free(wind);
  free(origin);
  free(radius);
  return c.chains;
}
