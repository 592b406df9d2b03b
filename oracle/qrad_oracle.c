/*
 * qrad_oracle.c -- TEST INFRASTRUCTURE ONLY.  Nothing in the product (qengine_b200/, include/) links,
 * imports or executes this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may.
 *
 * A plain-C, single-threaded CPU restatement of the qrad3 radiosity path of klaussilveira/qengine, written
 * from the algorithm (not copied): every function cites the reference file:line it follows
 * (paths relative to /root/reference/src/tools).  vec_t is float, FLT_EVAL_METHOD 0, no FMA
 * (build: gcc -O2 -ffp-contract=off, baseline x86-64), see SURVEY.md appendix A.
 *
 * Parity status: PINNED.  tests/test_oracle_cpu.py checks this restatement against the golden vectors in
 * tests/golden/ (outputs of the reference itself, compiled and run in the build container by
 * tests/golden/make_golden.py through oracle/_ref/ref_harness) -- TestLine results, per-patch numtransfers,
 * every transfer record, totallight, `added`, and the lighting lump byte for byte -- and, when oracle/_ref
 * is present, against fresh reference runs.
 *
 * Interface: a process-global scene (like the reference), filled through oq_* calls from oracle/oracle.py,
 * which does the file/entity parsing.  All arrays are copied in.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qrad_oracle.h"

typedef float vec3[3];

/* ---- on-disk records, common/qfiles.h:283-447 ---- */
typedef struct { float normal[3]; float dist; int32_t type; } dplane_o;
typedef struct { int32_t planenum; int32_t children[2]; int16_t mins[3], maxs[3]; uint16_t firstface, numfaces; } dnode_o;
typedef struct { int32_t contents; int16_t cluster, area; int16_t mins[3], maxs[3]; uint16_t a, b, c, d; } dleaf_o;
typedef struct { uint16_t planenum; int16_t side; int32_t firstedge; int16_t numedges; int16_t texinfo; uint8_t styles[4]; int32_t lightofs; } dface_o;
typedef struct { float vecs[2][4]; int32_t flags; int32_t value; char texture[32]; int32_t nexttexinfo; } texinfo_o;
typedef struct { float mins[3], maxs[3], origin[3]; int32_t headnode, firstface, numfaces; } dmodel_o;
typedef struct { uint16_t v[2]; } dedge_o;

#define MAXPTS 64
typedef struct { int n; vec3 p[MAXPTS + 8]; } wind_t;

typedef struct {
  wind_t *w;
  int next, face, cluster, sky, numtransfers, samples;
  vec3 origin, normal, totallight, reflectivity, baselight, samplelight;
  float dist, area;
  uint32_t *tpatch;
  uint16_t *ttrans;
} opatch;

typedef struct { int cluster, type, style; float intensity, stopdot; vec3 origin, color, normal; int nextl; } olight;

typedef struct { int numsamples, numstyles; int stylenums[32]; float *origins; float *samples[32]; } ofacelight;

static struct {
  dplane_o *planes; int numplanes;
  dnode_o *nodes; int numnodes;
  dleaf_o *leafs; int numleafs;
  uint8_t *vis; int vissize, numclusters;
  dface_o *faces; int numfaces;
  texinfo_o *texinfo; int numtexinfo;
  float *verts; dedge_o *edges; int32_t *surfedges; dmodel_o *models; int nummodels;
  /* params */
  int numbounce, extrasamples, nopvs;
  float subdiv, ambient, maxlight, lightscale, direct_scale, entity_scale;
  float *refl;                 /* [numtexinfo][3] */
  float *model_origin;         /* [nummodels][3] */
  float *model_minlight;       /* [nummodels] */
  opatch *patches; int num_patches, cap_patches;
  int *face_head; float *face_offset; int *face_model;
  olight *lights; int numlights, cap_lights; int *light_head;   /* per cluster list heads (LIFO) */
  ofacelight *fl;
  float *radiosity, *illumination;
  uint8_t *lightdata; int lightdatasize;
  long long rays_direct, rays_transfers, total_transfer;
} S;

/* ------------------------------------------------------------------ math, common/mathlib.c */
static float dot3(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static float vnorm(const float *in, float *out) /* mathlib.c:106-122 */
{
  float len = (float) sqrt((double) (in[0] * in[0] + in[1] * in[1] + in[2] * in[2]));
  float il;
  if (len == 0) { out[0] = out[1] = out[2] = 0; return 0; }
  il = (float) (1.0 / (double) len);
  out[0] = in[0] * il; out[1] = in[1] * il; out[2] = in[2] * il;
  return len;
}
static double vlen(const float *v) /* mathlib.c:30-41 */
{
  double l = 0; int i;
  for (i = 0; i < 3; i++) l += (double) (v[i] * v[i]);
  return sqrt(l);
}
static void vma(const float *a, double s, const float *b, float *c) /* mathlib.c:59-64 */
{
  c[0] = (float) ((double) a[0] + s * (double) b[0]);
  c[1] = (float) ((double) a[1] + s * (double) b[1]);
  c[2] = (float) ((double) a[2] + s * (double) b[2]);
}
static float cnorm(const float *in, float *out) /* ColorNormalize mathlib.c:124-141 */
{
  float m = in[0], sc;
  if (in[1] > m) m = in[1];
  if (in[2] > m) m = in[2];
  if (m == 0) return 0;
  sc = (float) (1.0 / (double) m);
  out[0] = sc * in[0]; out[1] = sc * in[1]; out[2] = sc * in[2];
  return m;
}
static void cross3(const float *a, const float *b, float *c)
{
  c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}

/* ------------------------------------------------------------------ BSP queries */
/* PointInLeafnum qrad3.c:131-150 */
static int point_in_leaf(const float *p)
{
  int n = 0;
  while (n >= 0) {
    dnode_o *nd = &S.nodes[n];
    dplane_o *pl = &S.planes[nd->planenum];
    float d = dot3(p, pl->normal) - pl->dist;
    n = d > 0 ? nd->children[0] : nd->children[1];
  }
  return -n - 1;
}

/* TestLine_r trace.c:92-142, recursion on the disk nodes (MakeTnode trace.c:47-71 only re-packs them:
   a leaf child is `contents & CONTENTS_SOLID`) */
static int test_line_r(int node, const float *start, const float *stop)
{
  dnode_o *nd; dplane_o *pl; float front, back, frac; vec3 mid; int side, r, i;
  if (node < 0) return S.leafs[-node - 1].contents & 1;
  nd = &S.nodes[node];
  pl = &S.planes[nd->planenum];
  switch (pl->type) {
    case 0: front = start[0] - pl->dist; back = stop[0] - pl->dist; break;
    case 1: front = start[1] - pl->dist; back = stop[1] - pl->dist; break;
    case 2: front = start[2] - pl->dist; back = stop[2] - pl->dist; break;
    default:
      front = (start[0] * pl->normal[0] + start[1] * pl->normal[1] + start[2] * pl->normal[2]) - pl->dist;
      back = (stop[0] * pl->normal[0] + stop[1] * pl->normal[1] + stop[2] * pl->normal[2]) - pl->dist;
  }
  if (front >= -0.1 && back >= -0.1) return test_line_r(nd->children[0], start, stop);
  if (front < 0.1 && back < 0.1) return test_line_r(nd->children[1], start, stop);
  side = front < 0;
  frac = front / (front - back);
  for (i = 0; i < 3; i++) mid[i] = start[i] + (stop[i] - start[i]) * frac;
  r = test_line_r(nd->children[side], start, mid);
  if (r) return r;
  return test_line_r(nd->children[!side], mid, stop);
}
int oq_test_line(const float *start, const float *stop) { return test_line_r(0, start, stop); }
int oq_point_in_leaf(const float *p) { return point_in_leaf(p); }

/* DecompressVis bspfile.c:129-154 + PvsForOrigin qrad3.c:160-175 */
static int pvs_for_origin(const float *org, uint8_t *pvs)
{
  int row = (S.numclusters + 7) >> 3, o = 0, cl;
  const uint8_t *in;
  if (!S.vissize) { memset(pvs, 255, (S.numleafs + 7) / 8); return 1; }
  cl = S.leafs[point_in_leaf(org)].cluster;
  if (cl == -1) return 0;
  in = S.vis + *(int32_t *) (S.vis + 4 + cl * 8);
  do {
    if (*in) { pvs[o++] = *in++; continue; }
    { int c = in[1]; in += 2; while (c) { pvs[o++] = 0; c--; } }
  } while (o < row);
  return 1;
}

/* ------------------------------------------------------------------ scene set-up */
static void *dup_mem(const void *p, size_t n) { void *q = malloc(n ? n : 1); if (n) memcpy(q, p, n); return q; }

void oq_reset(void)
{
  int i;
  for (i = 0; i < S.num_patches; i++) { free(S.patches[i].w); free(S.patches[i].tpatch); free(S.patches[i].ttrans); }
  if (S.fl) for (i = 0; i < S.numfaces; i++) { int s; free(S.fl[i].origins); for (s = 0; s < 32; s++) free(S.fl[i].samples[s]); }
  free(S.planes); free(S.nodes); free(S.leafs); free(S.vis); free(S.faces); free(S.texinfo); free(S.verts); free(S.edges);
  free(S.surfedges); free(S.models); free(S.refl); free(S.model_origin); free(S.model_minlight); free(S.patches);
  free(S.face_head); free(S.face_offset); free(S.face_model); free(S.lights); free(S.light_head); free(S.fl);
  free(S.radiosity); free(S.illumination); free(S.lightdata);
  memset(&S, 0, sizeof S);
}

void oq_set_bsp(const void *planes, int numplanes, const void *nodes, int numnodes, const void *leafs, int numleafs,
                const void *vis, int vissize, const void *faces, int numfaces, const void *texinfo, int numtexinfo,
                const void *verts, int numverts, const void *edges, int numedges, const void *surfedges, int numsurfedges,
                const void *models, int nummodels)
{
  oq_reset();
  S.planes = dup_mem(planes, (size_t) numplanes * 20); S.numplanes = numplanes;
  S.nodes = dup_mem(nodes, (size_t) numnodes * 28); S.numnodes = numnodes;
  S.leafs = dup_mem(leafs, (size_t) numleafs * 28); S.numleafs = numleafs;
  S.vis = dup_mem(vis, (size_t) vissize); S.vissize = vissize;
  S.numclusters = vissize >= 4 ? *(int32_t *) S.vis : 0;
  S.faces = dup_mem(faces, (size_t) numfaces * 20); S.numfaces = numfaces;
  S.texinfo = dup_mem(texinfo, (size_t) numtexinfo * 76); S.numtexinfo = numtexinfo;
  S.verts = dup_mem(verts, (size_t) numverts * 12);
  S.edges = dup_mem(edges, (size_t) numedges * 4);
  S.surfedges = dup_mem(surfedges, (size_t) numsurfedges * 4);
  S.models = dup_mem(models, (size_t) nummodels * 48); S.nummodels = nummodels;
  S.refl = calloc((size_t) (numtexinfo ? numtexinfo : 1) * 3, 4);
  S.model_origin = calloc((size_t) nummodels * 3 + 3, 4);
  S.model_minlight = calloc((size_t) nummodels + 1, 4);
  /* defaults, qrad3.c:49-74 */
  S.numbounce = 8; S.subdiv = 64; S.maxlight = 196; S.lightscale = 1.0f; S.direct_scale = 0.4f; S.entity_scale = 1.0f;
}

void oq_set_params(int numbounce, int extrasamples, float subdiv, float ambient, float maxlight, float lightscale,
                   float direct_scale, float entity_scale, int nopvs)
{
  S.numbounce = numbounce; S.extrasamples = extrasamples; S.subdiv = subdiv; S.ambient = ambient; S.maxlight = maxlight;
  S.lightscale = lightscale; S.direct_scale = direct_scale; S.entity_scale = entity_scale; S.nopvs = nopvs;
  if (!S.vissize) { S.numbounce = 0; S.ambient = 0.1f; } /* qrad3.c:627-631 */
}

/* CalcTextureReflectivity patches.c:40-112 for one texinfo: texels == NULL means the .wal is missing */
void oq_texture_reflectivity(int ti, const uint8_t *palette, const uint8_t *texels, int ntexels)
{
  float *r = S.refl + ti * 3, scale; int color[3] = {0, 0, 0}, j, k;
  if (!texels) { r[0] = r[1] = r[2] = 0.5f; return; }
  for (j = 0; j < ntexels; j++) for (k = 0; k < 3; k++) color[k] += palette[texels[j] * 3 + k];
  for (k = 0; k < 3; k++) r[k] = (float) ((double) (color[k] / ntexels) / 255.0);
  scale = cnorm(r, r);
  if (scale < 0.5) { scale *= 2; for (k = 0; k < 3; k++) r[k] = scale * r[k]; }
}
void oq_copy_reflectivity(int dst, int src) { memcpy(S.refl + dst * 3, S.refl + src * 3, 12); }
void oq_get_reflectivity(float *out) { memcpy(out, S.refl, (size_t) S.numtexinfo * 12); }

/* EntityForModel + GetVectorForKey(ent, "origin") patches.c:253-293, and _minlight lightmap.c:1150 */
void oq_set_model_entity(int model, const char *origin_str, const char *minlight_str)
{
  double a = 0, b = 0, c = 0;
  sscanf(origin_str, "%lf %lf %lf", &a, &b, &c);
  S.model_origin[model * 3] = (float) a; S.model_origin[model * 3 + 1] = (float) b; S.model_origin[model * 3 + 2] = (float) c;
  S.model_minlight[model] = (float) atof(minlight_str) * 128;
}

/* ------------------------------------------------------------------ windings, common/polylib.c */
static float wind_area(const wind_t *w) /* :137-151 */
{
  float total = 0; int i, k; vec3 d1, d2, c;
  for (i = 2; i < w->n; i++) {
    for (k = 0; k < 3; k++) { d1[k] = w->p[i - 1][k] - w->p[0][k]; d2[k] = w->p[i][k] - w->p[0][k]; }
    cross3(d1, d2, c);
    total = (float) ((double) total + 0.5 * vlen(c));
  }
  return total;
}
static void wind_bounds(const wind_t *w, float *mins, float *maxs) /* :153-170 */
{
  int i, j;
  mins[0] = mins[1] = mins[2] = 99999; maxs[0] = maxs[1] = maxs[2] = -99999;
  for (i = 0; i < w->n; i++) for (j = 0; j < 3; j++) {
    float v = w->p[i][j];
    if (v < mins[j]) mins[j] = v;
    if (v > maxs[j]) maxs[j] = v;
  }
}
static void wind_center(const wind_t *w, float *c) /* :177-188 */
{
  int i, j; float sc;
  c[0] = c[1] = c[2] = 0;
  for (i = 0; i < w->n; i++) for (j = 0; j < 3; j++) c[j] = w->p[i][j] + c[j];
  sc = (float) (1.0 / w->n);
  for (j = 0; j < 3; j++) c[j] = sc * c[j];
}
static void wind_remove_colinear(wind_t *w) /* :86-117 */
{
  vec3 keep[MAXPTS + 8]; int n = 0, i, a;
  for (i = 0; i < w->n; i++) {
    int j = (i + 1) % w->n, k = (i + w->n - 1) % w->n; vec3 v1, v2;
    for (a = 0; a < 3; a++) { v1[a] = w->p[j][a] - w->p[i][a]; v2[a] = w->p[i][a] - w->p[k][a]; }
    vnorm(v1, v1); vnorm(v2, v2);
    if (dot3(v1, v2) < 0.999) { memcpy(keep[n], w->p[i], 12); n++; }
  }
  if (n == w->n) return;
  w->n = n; memcpy(w->p, keep, (size_t) n * 12);
}
/* ClipWindingEpsilon :297-392 */
static void wind_clip(const wind_t *in, const float *normal, float dist, float eps, wind_t *f, wind_t *b)
{
  float dists[MAXPTS + 8]; int sides[MAXPTS + 8], counts[3] = {0, 0, 0}, i, j;
  for (i = 0; i < in->n; i++) {
    float d = dot3(in->p[i], normal); d -= dist; dists[i] = d;
    sides[i] = d > eps ? 0 : (d < -eps ? 1 : 2);
    counts[sides[i]]++;
  }
  sides[i] = sides[0]; dists[i] = dists[0];
  f->n = b->n = 0;
  if (!counts[0]) { *b = *in; return; }
  if (!counts[1]) { *f = *in; return; }
  for (i = 0; i < in->n; i++) {
    const float *p1 = in->p[i], *p2; float d; vec3 mid;
    if (sides[i] == 2) { memcpy(f->p[f->n++], p1, 12); memcpy(b->p[b->n++], p1, 12); continue; }
    if (sides[i] == 0) memcpy(f->p[f->n++], p1, 12);
    if (sides[i] == 1) memcpy(b->p[b->n++], p1, 12);
    if (sides[i + 1] == 2 || sides[i + 1] == sides[i]) continue;
    p2 = in->p[(i + 1) % in->n];
    d = dists[i] / (dists[i] - dists[i + 1]);
    for (j = 0; j < 3; j++) {
      if (normal[j] == 1) mid[j] = dist;
      else if (normal[j] == -1) mid[j] = -dist;
      else mid[j] = p1[j] + d * (p2[j] - p1[j]);
    }
    memcpy(f->p[f->n++], mid, 12); memcpy(b->p[b->n++], mid, 12);
  }
}

/* ------------------------------------------------------------------ patches, qrad3/patches.c */
static int new_patch(void)
{
  if (S.num_patches == S.cap_patches) {
    S.cap_patches = S.cap_patches ? S.cap_patches * 2 : 4096;
    S.patches = realloc(S.patches, (size_t) S.cap_patches * sizeof(opatch));
  }
  memset(&S.patches[S.num_patches], 0, sizeof(opatch));
  S.patches[S.num_patches].next = -1;
  return S.num_patches++;
}
static void place_patch(opatch *p) /* patches.c:222-230, 337-352 */
{
  vec3 c; int k;
  wind_center(p->w, c);
  for (k = 0; k < 3; k++) p->origin[k] = c[k] + p->normal[k];
  p->cluster = S.leafs[point_in_leaf(p->origin)].cluster;
}

/* MakePatches + MakePatchForFace patches.c:187-309; backplanes qrad3.c:88-96 */
void oq_make_patches(void)
{
  int m, j, e, k;
  S.face_head = malloc((size_t) S.numfaces * 4);
  S.face_offset = calloc((size_t) S.numfaces * 3, 4);
  S.face_model = calloc((size_t) S.numfaces, 4);
  for (j = 0; j < S.numfaces; j++) S.face_head[j] = -1;
  for (m = 0; m < S.nummodels; m++) {
    dmodel_o *mod = &S.models[m];
    const float *origin = S.model_origin + m * 3;
    for (j = 0; j < mod->numfaces; j++) {
      int fn = mod->firstface + j, pi; dface_o *f = &S.faces[fn]; opatch *p; dplane_o *pl; texinfo_o *tx; float area;
      wind_t *w = calloc(1, sizeof(wind_t));
      S.face_model[fn] = m;
      memcpy(S.face_offset + fn * 3, origin, 12);
      w->n = f->numedges;   /* WindingFromFace patches.c:122-145 */
      for (e = 0; e < f->numedges; e++) {
        int se = S.surfedges[f->firstedge + e];
        int v = se < 0 ? S.edges[-se].v[1] : S.edges[se].v[0];
        memcpy(w->p[e], S.verts + v * 3, 12);
      }
      wind_remove_colinear(w);
      for (e = 0; e < w->n; e++) for (k = 0; k < 3; k++) w->p[e][k] = w->p[e][k] + origin[k];
      area = wind_area(w);
      pi = new_patch(); p = &S.patches[pi];
      p->next = S.face_head[fn]; S.face_head[fn] = pi; p->face = fn; p->w = w;
      pl = &S.planes[f->planenum];
      if (f->side) { for (k = 0; k < 3; k++) p->normal[k] = 0 - pl->normal[k]; p->dist = -pl->dist; }
      else { memcpy(p->normal, pl->normal, 12); p->dist = pl->dist; }
      if (origin[0] || origin[1] || origin[2]) p->dist += dot3(origin, p->normal);
      place_patch(p);
      p->area = area; if (p->area <= 1) p->area = 1;
      tx = &S.texinfo[f->texinfo];
      p->sky = (tx->flags & 4) ? 1 : 0;
      memcpy(p->reflectivity, S.refl + f->texinfo * 3, 12);
      if (fn < S.models[0].numfaces) {
        vec3 color = {0, 0, 0};
        if (!(tx->flags & 1) || tx->value == 0) p->baselight[0] = p->baselight[1] = p->baselight[2] = 0;
        else for (k = 0; k < 3; k++) p->baselight[k] = (float) tx->value * S.refl[f->texinfo * 3 + k];
        cnorm(p->reflectivity, color);
        for (k = 0; k < 3; k++) p->baselight[k] *= color[k];
        memcpy(p->totallight, p->baselight, 12);
      }
    }
  }
}

/* DicePatch + FinishSplit patches.c:311-470 */
static void dice(int pi)
{
  vec3 mins, maxs, split = {0, 0, 0}; int i, ni; float dist; wind_t *o1, *o2; opatch *p, *np;
  wind_bounds(S.patches[pi].w, mins, maxs);
  for (i = 0; i < 3; i++)
    if (floor((mins[i] + 1) / S.subdiv) < floor((maxs[i] - 1) / S.subdiv)) break;
  if (i == 3) return;
  split[i] = 1;
  dist = S.subdiv * (1 + floor((mins[i] + 1) / S.subdiv));
  o1 = calloc(1, sizeof(wind_t)); o2 = calloc(1, sizeof(wind_t));
  wind_clip(S.patches[pi].w, split, dist, 0.1f, o1, o2);
  ni = new_patch();
  p = &S.patches[pi]; np = &S.patches[ni];
  np->next = p->next; p->next = ni;
  p->w = o1; np->w = o2;   /* (the old winding is leaked, as in the reference) */
  memcpy(np->baselight, p->baselight, 12); memcpy(np->totallight, p->totallight, 12);
  memcpy(np->reflectivity, p->reflectivity, 12); memcpy(np->normal, p->normal, 12);
  np->dist = p->dist; np->sky = p->sky; np->face = p->face;
  p->area = wind_area(p->w); np->area = wind_area(np->w);
  if (p->area <= 1) p->area = 1;
  if (np->area <= 1) np->area = 1;
  place_patch(p); place_patch(np);
  dice(pi); dice(ni);
}
void oq_subdivide(void) /* SubdividePatches patches.c:477-492 */
{
  int i, n = S.num_patches;
  if (S.subdiv < 1) return;
  for (i = 0; i < n; i++) dice(i);
}
int oq_num_patches(void) { return S.num_patches; }
void oq_get_patches(float *origin, float *normal, float *area, int *cluster, float *totallight, float *samplelight,
                    int *samples, int *numtransfers, int *face, int *next)
{
  int i;
  for (i = 0; i < S.num_patches; i++) {
    opatch *p = &S.patches[i];
    if (origin) memcpy(origin + i * 3, p->origin, 12);
    if (normal) memcpy(normal + i * 3, p->normal, 12);
    if (area) area[i] = p->area;
    if (cluster) cluster[i] = p->cluster;
    if (totallight) memcpy(totallight + i * 3, p->totallight, 12);
    if (samplelight) memcpy(samplelight + i * 3, p->samplelight, 12);
    if (samples) samples[i] = p->samples;
    if (numtransfers) numtransfers[i] = p->numtransfers;
    if (face) face[i] = p->face;
    if (next) next[i] = p->next;
  }
}

/* ------------------------------------------------------------------ direct lights, lightmap.c:721-838 */
static olight *push_light(int cluster)
{
  olight *l;
  if (S.numlights == S.cap_lights) {
    S.cap_lights = S.cap_lights ? S.cap_lights * 2 : 256;
    S.lights = realloc(S.lights, (size_t) S.cap_lights * sizeof(olight));
  }
  l = &S.lights[S.numlights];
  memset(l, 0, sizeof *l);
  l->cluster = cluster;
  if (cluster >= 0 && cluster < S.numclusters) { l->nextl = S.light_head[cluster]; S.light_head[cluster] = S.numlights; }
  else l->nextl = -2; /* the reference would index directlights[-1]; such a light is never gathered */
  S.numlights++;
  return l;
}
void oq_lights_from_patches(void) /* the "surfaces" loop, :741-760 */
{
  int i;
  S.light_head = malloc((size_t) (S.numclusters + 1) * 4);
  for (i = 0; i <= S.numclusters; i++) S.light_head[i] = -1;
  for (i = 0; i < S.num_patches; i++) {
    opatch *p = &S.patches[i]; olight *l;
    if (p->totallight[0] < 3 && p->totallight[1] < 3 && p->totallight[2] < 3) continue;
    l = push_light(S.leafs[point_in_leaf(p->origin)].cluster);
    memcpy(l->origin, p->origin, 12);
    l->type = 0;
    memcpy(l->normal, p->normal, 12);
    l->intensity = cnorm(p->totallight, l->color);
    l->intensity *= p->area * S.direct_scale;
    p->totallight[0] = p->totallight[1] = p->totallight[2] = 0;
  }
}
/* one "light*" entity, :765-836; strings are the raw key values ("" when absent); target_origin NULL = no target found */
void oq_add_light_entity(const char *classname, const char *origin, const char *light, const char *_light,
                         const char *style, const char *_style, const char *color, const char *target,
                         const char *target_origin, const char *cone, const char *angle_s)
{
  olight *l; double a = 0, b = 0, c = 0; vec3 org; float intensity;
  sscanf(origin, "%lf %lf %lf", &a, &b, &c);
  org[0] = (float) a; org[1] = (float) b; org[2] = (float) c;
  l = push_light(S.leafs[point_in_leaf(org)].cluster);
  memcpy(l->origin, org, 12);
  l->style = (int) (float) atof(_style);
  if (!l->style) l->style = (int) (float) atof(style);
  if (l->style < 0 || l->style >= 256) l->style = 0;
  intensity = (float) atof(light);
  if (!intensity) intensity = (float) atof(_light);
  if (!intensity) intensity = 300;
  if (color[0] && color[1]) { sscanf(color, "%f %f %f", &l->color[0], &l->color[1], &l->color[2]); cnorm(l->color, l->color); }
  else l->color[0] = l->color[1] = l->color[2] = 1.0f;
  l->intensity = intensity * S.entity_scale;
  l->type = 1;
  if (!strcmp(classname, "light_spot") || target[0]) {
    l->type = 2;
    l->stopdot = (float) atof(cone);
    if (!l->stopdot) l->stopdot = 10;
    l->stopdot = (float) cos(l->stopdot / 180 * 3.14159);
    if (target[0]) {
      if (target_origin) {
        double x = 0, y = 0, z = 0; vec3 dest; int k;
        sscanf(target_origin, "%lf %lf %lf", &x, &y, &z);
        dest[0] = (float) x; dest[1] = (float) y; dest[2] = (float) z;
        for (k = 0; k < 3; k++) l->normal[k] = dest[k] - l->origin[k];
        vnorm(l->normal, l->normal);
      }
    } else {
      float angle = (float) atof(angle_s);
      if (angle == -1) { l->normal[0] = l->normal[1] = 0; l->normal[2] = 1; }
      else if (angle == -2) { l->normal[0] = l->normal[1] = 0; l->normal[2] = -1; }
      else { l->normal[2] = 0; l->normal[0] = (float) cos(angle / 180 * 3.14159); l->normal[1] = (float) sin(angle / 180 * 3.14159); }
    }
  }
}
int oq_num_lights(void) { return S.numlights; }

/* ------------------------------------------------------------------ facelights, lightmap.c:480-1056 */
typedef struct {
  float facedist; vec3 facenormal; int numsurfpt; float *surfpt; vec3 modelorg, texorg; vec3 worldtotex[2], textoworld[2];
  float exactmins[2], exactmaxs[2]; int texmins[2], texsize[2]; dface_o *face;
} linfo;

static void calc_face_extents(linfo *l) /* :480-528 */
{
  dface_o *s = l->face; float mins[2] = {999999, 999999}, maxs[2] = {-99999, -99999}; texinfo_o *tex = &S.texinfo[s->texinfo];
  int i, j;
  for (i = 0; i < s->numedges; i++) {
    int e = S.surfedges[s->firstedge + i];
    const float *vt = S.verts + 3 * (e >= 0 ? S.edges[e].v[0] : S.edges[-e].v[1]);
    for (j = 0; j < 2; j++) {
      float val = dot3(vt, tex->vecs[j]) + tex->vecs[j][3];
      if (val < mins[j]) mins[j] = val;
      if (val > maxs[j]) maxs[j] = val;
    }
  }
  for (i = 0; i < 2; i++) {
    l->exactmins[i] = mins[i]; l->exactmaxs[i] = maxs[i];
    mins[i] = floor(mins[i] / 16); maxs[i] = ceil(maxs[i] / 16);
    l->texmins[i] = mins[i]; l->texsize[i] = maxs[i] - mins[i];
  }
}
static void calc_face_vectors(linfo *l) /* :536-603 */
{
  texinfo_o *tex = &S.texinfo[l->face->texinfo]; vec3 texnormal; float distscale, dist, len; int i, j;
  for (i = 0; i < 2; i++) for (j = 0; j < 3; j++) l->worldtotex[i][j] = tex->vecs[i][j];
  texnormal[0] = tex->vecs[1][1] * tex->vecs[0][2] - tex->vecs[1][2] * tex->vecs[0][1];
  texnormal[1] = tex->vecs[1][2] * tex->vecs[0][0] - tex->vecs[1][0] * tex->vecs[0][2];
  texnormal[2] = tex->vecs[1][0] * tex->vecs[0][1] - tex->vecs[1][1] * tex->vecs[0][0];
  vnorm(texnormal, texnormal);
  distscale = dot3(texnormal, l->facenormal);
  if (!distscale) distscale = 1;
  if (distscale < 0) { distscale = -distscale; for (j = 0; j < 3; j++) texnormal[j] = 0 - texnormal[j]; }
  distscale = 1 / distscale;
  for (i = 0; i < 2; i++) {
    float sc;
    len = vlen(l->worldtotex[i]);
    dist = dot3(l->worldtotex[i], l->facenormal);
    dist *= distscale;
    vma(l->worldtotex[i], -dist, texnormal, l->textoworld[i]);
    sc = (1 / len) * (1 / len);
    for (j = 0; j < 3; j++) l->textoworld[i][j] = sc * l->textoworld[i][j];
  }
  for (i = 0; i < 3; i++) l->texorg[i] = -tex->vecs[0][3] * l->textoworld[0][i] - tex->vecs[1][3] * l->textoworld[1][i];
  dist = dot3(l->texorg, l->facenormal) - l->facedist - 1;
  dist *= distscale;
  vma(l->texorg, -dist, texnormal, l->texorg);
  for (i = 0; i < 3; i++) l->texorg[i] = l->texorg[i] + l->modelorg[i];
}
static void calc_points(linfo *l, float sofs, float tofs) /* :608-677 */
{
  int s, t, j, i, w, h, step = 16; float starts, startt, us, ut, mids, midt; vec3 facemid; float *surf;
  mids = (l->exactmaxs[0] + l->exactmins[0]) / 2;
  midt = (l->exactmaxs[1] + l->exactmins[1]) / 2;
  for (j = 0; j < 3; j++) facemid[j] = l->texorg[j] + l->textoworld[0][j] * mids + l->textoworld[1][j] * midt;
  h = l->texsize[1] + 1; w = l->texsize[0] + 1;
  l->numsurfpt = w * h;
  l->surfpt = malloc((size_t) w * h * 12 + 12);
  surf = l->surfpt;
  starts = l->texmins[0] * 16; startt = l->texmins[1] * 16;
  for (t = 0; t < h; t++) for (s = 0; s < w; s++, surf += 3) {
    us = starts + (s + sofs) * step;
    ut = startt + (t + tofs) * step;
    for (i = 0; i < 6; i++) {
      for (j = 0; j < 3; j++) surf[j] = l->texorg[j] + l->textoworld[0][j] * us + l->textoworld[1][j] * ut;
      if (S.leafs[point_in_leaf(surf)].contents != 1) {
        S.rays_direct++;
        if (!test_line_r(0, facemid, surf)) break;
      }
      if (i & 1) {
        if (us > mids) { us -= 8; if (us < mids) us = mids; } else { us += 8; if (us > mids) us = mids; }
      } else {
        if (ut > midt) { ut -= 8; if (ut < midt) ut = midt; } else { ut += 8; if (ut > midt) ut = midt; }
      }
    }
  }
}
/* GatherSampleLight :847-917 */
static void gather(const float *pos, const float *normal, float **styletable, int offset, int mapsize, float lightscale)
{
  static uint8_t pvs[(65536 + 7) / 8]; int i, li;
  if (!pvs_for_origin(pos, pvs)) return;
  for (i = 0; i < S.numclusters; i++) {
    if (!(pvs[i >> 3] & (1 << (i & 7)))) continue;
    for (li = S.light_head[i]; li >= 0; li = S.lights[li].nextl) {
      olight *l = &S.lights[li]; vec3 delta; float dist, dot, dot2, scale, *dest; int k;
      for (k = 0; k < 3; k++) delta[k] = l->origin[k] - pos[k];
      dist = vnorm(delta, delta);
      dot = dot3(delta, normal);
      if (dot <= 0.001) continue;
      if (l->type == 1) scale = (l->intensity - dist) * dot;
      else if (l->type == 0) {
        dot2 = -dot3(delta, l->normal);
        if (dot2 <= 0.001) continue;
        scale = (l->intensity / (dist * dist)) * dot * dot2;
      } else {
        dot2 = -dot3(delta, l->normal);
        if (dot2 <= l->stopdot) continue;
        scale = (l->intensity - dist) * dot;
      }
      S.rays_direct++;
      if (test_line_r(0, pos, l->origin)) continue;
      if (scale <= 0) continue;
      if (!styletable[l->style]) styletable[l->style] = calloc(1, (size_t) mapsize);
      dest = styletable[l->style] + offset;
      vma(dest, scale * lightscale, l->color, dest);
    }
  }
}
/* BuildFacelights :967-1056 (with AddSampleToPatch :932-958) */
void oq_build_facelights(void)
{
  static const float sampleofs[5][2] = {{0, 0}, {-0.25f, -0.25f}, {0.25f, -0.25f}, {0.25f, 0.25f}, {-0.25f, 0.25f}};
  int fn;
  S.fl = calloc((size_t) S.numfaces, sizeof(ofacelight));
  for (fn = 0; fn < S.numfaces; fn++) {
    dface_o *f = &S.faces[fn]; linfo l[5]; float *styletable[256]; ofacelight *fl = &S.fl[fn];
    int numsamples = S.extrasamples ? 5 : 1, i, j, k, tablesize, pi, head;
    if (S.texinfo[f->texinfo].flags & (8 | 4)) continue;
    memset(styletable, 0, sizeof styletable);
    for (i = 0; i < numsamples; i++) {
      memset(&l[i], 0, sizeof l[i]);
      l[i].face = f;
      memcpy(l[i].facenormal, S.planes[f->planenum].normal, 12);
      l[i].facedist = S.planes[f->planenum].dist;
      if (f->side) { for (k = 0; k < 3; k++) l[i].facenormal[k] = 0 - l[i].facenormal[k]; l[i].facedist = -l[i].facedist; }
      memcpy(l[i].modelorg, S.face_offset + fn * 3, 12);
      calc_face_vectors(&l[i]);
      calc_face_extents(&l[i]);
      calc_points(&l[i], sampleofs[i][0], sampleofs[i][1]);
    }
    tablesize = l[0].numsurfpt * 12;
    styletable[0] = calloc(1, (size_t) tablesize + 12);
    fl->numsamples = l[0].numsurfpt;
    fl->origins = malloc((size_t) tablesize + 12);
    memcpy(fl->origins, l[0].surfpt, (size_t) tablesize);
    for (i = 0; i < l[0].numsurfpt; i++) {
      const float *pos = l[0].surfpt + i * 3, *color;
      for (j = 0; j < numsamples; j++) gather(l[j].surfpt + i * 3, l[0].facenormal, styletable, i * 3, tablesize, 1.0 / numsamples);
      color = styletable[0] + i * 3;
      if (S.numbounce == 0) continue;
      if (color[0] + color[1] + color[2] < 3) continue;
      for (pi = S.face_head[fn]; pi >= 0; pi = S.patches[pi].next) {
        opatch *p = &S.patches[pi]; vec3 mins, maxs;
        wind_bounds(p->w, mins, maxs);
        for (k = 0; k < 3; k++) { if (mins[k] > pos[k] + 16) break; if (maxs[k] < pos[k] - 16) break; }
        if (k != 3) continue;
        p->samples++;
        for (k = 0; k < 3; k++) p->samplelight[k] = p->samplelight[k] + color[k];
      }
    }
    for (pi = S.face_head[fn]; pi >= 0; pi = S.patches[pi].next) {
      opatch *p = &S.patches[pi];
      if (p->samples) for (k = 0; k < 3; k++) p->samplelight[k] = (1.0 / p->samples) * p->samplelight[k];
    }
    for (i = 0; i < 256; i++) {
      if (!styletable[i]) continue;
      if (fl->numstyles == 32) break;
      fl->samples[fl->numstyles] = styletable[i];
      fl->stylenums[fl->numstyles] = i;
      fl->numstyles++;
    }
    head = S.face_head[fn];
    if (S.patches[head].baselight[0] >= 3 || S.patches[head].baselight[1] >= 3 || S.patches[head].baselight[2] >= 3) {
      float *spot = fl->samples[0];
      for (i = 0; i < l[0].numsurfpt; i++, spot += 3) for (k = 0; k < 3; k++) spot[k] = spot[k] + S.patches[head].baselight[k];
    }
    for (i = 0; i < numsamples; i++) free(l[i].surfpt);
  }
}
int oq_facelight(int face, int *numsamples, int *stylenums, float **origins, float **samples0)
{
  ofacelight *fl = &S.fl[face]; int s;
  *numsamples = fl->numsamples;
  for (s = 0; s < fl->numstyles; s++) stylenums[s] = fl->stylenums[s];
  *origins = fl->origins; *samples0 = fl->samples[0];
  return fl->numstyles;
}

/* ------------------------------------------------------------------ MakeTransfers qrad3.c:185-293 */
void oq_make_transfers(void)
{
  int i, j, N = S.num_patches;
  float *transfers = malloc((size_t) N * 4 + 4);
  static uint8_t pvs[(65536 + 7) / 8];
  S.total_transfer = 0;
  for (i = 0; i < N; i++) {
    opatch *patch = &S.patches[i]; float total = 0;
    free(patch->tpatch); free(patch->ttrans); patch->tpatch = NULL; patch->ttrans = NULL; patch->numtransfers = 0;
    if (!pvs_for_origin(patch->origin, pvs)) continue;
    for (j = 0; j < N; j++) {
      opatch *p2 = &S.patches[j]; vec3 delta; float dist, scale, trans; int k;
      transfers[j] = 0;
      if (j == i) continue;
      if (!S.nopvs) {
        int cl = p2->cluster;
        if (cl == -1) continue;
        if (!(pvs[cl >> 3] & (1 << (cl & 7)))) continue;
      }
      for (k = 0; k < 3; k++) delta[k] = p2->origin[k] - patch->origin[k];
      dist = vnorm(delta, delta);
      if (!dist) continue;
      scale = dot3(delta, patch->normal);
      scale *= -dot3(delta, p2->normal);
      if (scale <= 0) continue;
      S.rays_transfers++;
      if (test_line_r(0, patch->origin, p2->origin)) continue;
      trans = scale * p2->area / (dist * dist);
      if (trans < 0) trans = 0;
      transfers[j] = trans;
      if (trans > 0) { total += trans; patch->numtransfers++; }
    }
    if (patch->numtransfers) {
      int n = 0;
      patch->tpatch = malloc((size_t) patch->numtransfers * 4);
      patch->ttrans = malloc((size_t) patch->numtransfers * 2);
      for (j = 0; j < N; j++) {
        int itrans;
        if (transfers[j] <= 0) continue;
        itrans = transfers[j] * 0x10000 / total;
        patch->ttrans[n] = (uint16_t) itrans;
        patch->tpatch[n] = (uint32_t) j;
        n++;
      }
    }
    S.total_transfer += patch->numtransfers;
  }
  free(transfers);
}
long long oq_total_transfer(void) { return S.total_transfer; }
void oq_get_transfers(uint32_t *patch, uint16_t *transfer)
{
  int i; size_t o = 0;
  for (i = 0; i < S.num_patches; i++) {
    opatch *p = &S.patches[i];
    if (!p->numtransfers) continue;
    memcpy(patch + o, p->tpatch, (size_t) p->numtransfers * 4);
    memcpy(transfer + o, p->ttrans, (size_t) p->numtransfers * 2);
    o += p->numtransfers;
  }
}
void oq_ray_counts(long long *direct, long long *transfers) { *direct = S.rays_direct; *transfers = S.rays_transfers; }

/* ------------------------------------------------------------------ BounceLight qrad3.c:383-473 */
void oq_bounce(float *added)
{
  int i, j, b, k, N = S.num_patches;
  free(S.radiosity); free(S.illumination);
  S.radiosity = calloc((size_t) N * 3 + 3, 4); S.illumination = calloc((size_t) N * 3 + 3, 4);
  for (i = 0; i < N; i++) for (j = 0; j < 3; j++)
    S.radiosity[i * 3 + j] = S.patches[i].samplelight[j] * S.patches[i].reflectivity[j] * S.patches[i].area;
  for (b = 0; b < S.numbounce; b++) {
    float total = 0;
    for (i = 0; i < N; i++) { /* ShootLight :419-441 */
      opatch *p = &S.patches[i]; vec3 send;
      for (k = 0; k < 3; k++) send[k] = S.radiosity[i * 3 + k] / 0x10000;
      for (k = 0; k < p->numtransfers; k++) for (j = 0; j < 3; j++)
        S.illumination[p->tpatch[k] * 3 + j] += send[j] * p->ttrans[k];
    }
    for (i = 0; i < N; i++) { /* CollectLight :383-409 */
      opatch *p = &S.patches[i];
      if (p->sky) { for (j = 0; j < 3; j++) S.radiosity[i * 3 + j] = S.illumination[i * 3 + j] = 0; continue; }
      for (j = 0; j < 3; j++) {
        p->totallight[j] += S.illumination[i * 3 + j] / p->area;
        S.radiosity[i * 3 + j] = S.illumination[i * 3 + j] * p->reflectivity[j];
      }
      total += S.radiosity[i * 3] + S.radiosity[i * 3 + 1] + S.radiosity[i * 3 + 2];
      for (j = 0; j < 3; j++) S.illumination[i * 3 + j] = 0;
    }
    if (added) added[b] = total;
  }
}

/* ------------------------------------------------------------------ FinalLightFace lightmap.c:1066-1201 */
#define MAX_TRI_POINTS 1024
#define MAX_TRI_EDGES (MAX_TRI_POINTS * 6)
#define MAX_TRI_TRIS (MAX_TRI_POINTS * 2)
typedef struct { int p0, p1, tri; vec3 normal; float dist; } tedge;
typedef struct { int e[3]; } ttri;
typedef struct {
  int numpoints, numedges, numtris; const float *plane_normal;
  int points[MAX_TRI_POINTS]; tedge edges[MAX_TRI_EDGES]; ttri tris[MAX_TRI_TRIS];
  int *matrix; /* [numpoints][numpoints] edge index + 1 */
} trian_t;
static int oq_error;

static int find_edge(trian_t *t, int p0, int p1) /* :146-186 */
{
  vec3 v1, normal; float dist; tedge *e, *be; int k, n = t->numpoints;
  if (t->matrix[p0 * n + p1]) return t->matrix[p0 * n + p1] - 1;
  if (t->numedges > MAX_TRI_EDGES - 2) { oq_error = 1; return 0; }
  for (k = 0; k < 3; k++) v1[k] = S.patches[t->points[p1]].origin[k] - S.patches[t->points[p0]].origin[k];
  vnorm(v1, v1);
  cross3(v1, t->plane_normal, normal);
  dist = dot3(S.patches[t->points[p0]].origin, normal);
  e = &t->edges[t->numedges]; e->p0 = p0; e->p1 = p1; e->tri = -1; memcpy(e->normal, normal, 12); e->dist = dist;
  t->matrix[p0 * n + p1] = ++t->numedges;
  be = &t->edges[t->numedges]; be->p0 = p1; be->p1 = p0; be->tri = -1;
  for (k = 0; k < 3; k++) be->normal[k] = 0 - normal[k];
  be->dist = -dist;
  t->matrix[p1 * n + p0] = ++t->numedges;
  return (int) (e - t->edges);
}
static void tri_edge_r(trian_t *t, int ei) /* :211-260 */
{
  tedge *e = &t->edges[ei]; int i, bestp = -1, nt, e1, e2, k; float best, ang; const float *p0, *p1;
  if (e->tri >= 0 || oq_error) return;
  p0 = S.patches[t->points[e->p0]].origin; p1 = S.patches[t->points[e->p1]].origin;
  best = 1.1;
  for (i = 0; i < t->numpoints; i++) {
    const float *p = S.patches[t->points[i]].origin; vec3 v1, v2;
    if (dot3(p, e->normal) - e->dist < 0) continue;
    for (k = 0; k < 3; k++) { v1[k] = p0[k] - p[k]; v2[k] = p1[k] - p[k]; }
    if (!vnorm(v1, v1)) continue;
    if (!vnorm(v2, v2)) continue;
    ang = dot3(v1, v2);
    if (ang < best) { best = ang; bestp = i; }
  }
  if (best >= 1) return;
  if (t->numtris >= MAX_TRI_TRIS) { oq_error = 2; return; }
  nt = t->numtris++;
  e1 = find_edge(t, e->p1, bestp);
  e2 = find_edge(t, bestp, e->p0);
  t->tris[nt].e[0] = ei; t->tris[nt].e[1] = e1; t->tris[nt].e[2] = e2;
  t->edges[ei].tri = t->edges[e1].tri = t->edges[e2].tri = nt;
  tri_edge_r(t, find_edge(t, bestp, t->edges[ei].p1));
  tri_edge_r(t, find_edge(t, t->edges[ei].p0, bestp));
}
static void triangulate(trian_t *t) /* :262-293 */
{
  float d, bestd = 9999; int i, j, k, bp1 = 0, bp2 = 0, e, e2;
  if (t->numpoints < 2) return;
  for (i = 0; i < t->numpoints; i++) for (j = i + 1; j < t->numpoints; j++) {
    vec3 v1;
    for (k = 0; k < 3; k++) v1[k] = S.patches[t->points[j]].origin[k] - S.patches[t->points[i]].origin[k];
    d = vlen(v1);
    if (d < bestd) { bestd = d; bp1 = i; bp2 = j; }
  }
  e = find_edge(t, bp1, bp2);
  e2 = find_edge(t, bp2, bp1);
  tri_edge_r(t, e);
  tri_edge_r(t, e2);
}
static void sample_trian(const float *point, trian_t *t, float *color) /* :315-439 */
{
  int j, i, k; float d, best; int bp;
  if (t->numpoints == 0) { color[0] = color[1] = color[2] = 0; return; }
  if (t->numpoints == 1) { memcpy(color, S.patches[t->points[0]].totallight, 12); return; }
  for (j = 0; j < t->numtris; j++) {
    ttri *tr = &t->tris[j]; tedge *e0, *e2; opatch *p1, *p2, *p3; vec3 base, d1, d2; float x, y, y1, x2;
    for (i = 0; i < 3; i++) { tedge *e = &t->edges[tr->e[i]]; if (dot3(e->normal, point) - e->dist < 0) break; }
    if (i != 3) continue;
    e0 = &t->edges[tr->e[0]]; e2 = &t->edges[tr->e[2]];
    p1 = &S.patches[t->points[t->edges[tr->e[0]].p0]];
    p2 = &S.patches[t->points[t->edges[tr->e[1]].p0]];
    p3 = &S.patches[t->points[t->edges[tr->e[2]].p0]];
    memcpy(base, p1->totallight, 12);
    for (k = 0; k < 3; k++) { d1[k] = p2->totallight[k] - base[k]; d2[k] = p3->totallight[k] - base[k]; }
    x = dot3(point, e0->normal) - e0->dist;
    y = dot3(point, e2->normal) - e2->dist;
    y1 = dot3(p2->origin, e2->normal) - e2->dist;
    x2 = dot3(p3->origin, e0->normal) - e0->dist;
    if (fabs(y1) < 0.1 || fabs(x2) < 0.1) { memcpy(color, base, 12); return; }
    vma(base, x / x2, d2, color);
    vma(color, y / y1, d1, color);
    return;
  }
  for (j = 0; j < t->numedges; j++) {
    tedge *e = &t->edges[j]; opatch *p0, *p1; vec3 v1, v2;
    if (e->tri >= 0) continue;
    d = dot3(point, e->normal) - e->dist;
    if (d < 0) continue;
    p0 = &S.patches[t->points[e->p0]]; p1 = &S.patches[t->points[e->p1]];
    for (k = 0; k < 3; k++) v1[k] = p1->origin[k] - p0->origin[k];
    vnorm(v1, v1);
    for (k = 0; k < 3; k++) v2[k] = point[k] - p0->origin[k];
    d = dot3(v2, v1);
    if (d < 0) continue;
    if (d > 1) continue;
    for (i = 0; i < 3; i++) color[i] = p0->totallight[i] + d * (p1->totallight[i] - p0->totallight[i]);
    return;
  }
  best = 99999; bp = -1;
  for (j = 0; j < t->numpoints; j++) {
    vec3 v1;
    for (k = 0; k < 3; k++) v1[k] = point[k] - S.patches[t->points[j]].origin[k];
    d = vlen(v1);
    if (d < best) { best = d; bp = j; }
  }
  if (bp >= 0) memcpy(color, S.patches[t->points[bp]].totallight, 12);
}

/* returns lightdatasize; fills lightofs[numfaces], styles[numfaces][4] (only for lit faces); <0 on Error() */
int oq_final_light(uint8_t *out, int cap, int *lightofs, uint8_t *styles)
{
  int fn, i, j, k, st;
  int *facelinks = calloc((size_t) S.numfaces + 1, 4);
  int *planelinks = calloc((size_t) S.numplanes * 2 + 2, 4);
  trian_t *trian = malloc(sizeof(trian_t));
  int lightdatasize = 0;
  oq_error = 0;
  srand(1);
  for (i = 0; i < S.numfaces; i++) { /* LinkPlaneFaces :42-52 */
    dface_o *f = &S.faces[i]; int idx = (f->side ? 1 : 0) * S.numplanes + f->planenum;
    facelinks[i] = planelinks[idx]; planelinks[idx] = i;
  }
  for (fn = 0; fn < S.numfaces; fn++) {
    dface_o *f = &S.faces[fn]; ofacelight *fl = &S.fl[fn]; float minlight; uint8_t *dest; int ofs, pf, pi, nst;
    vec3 facemins, facemaxs;
    if (S.texinfo[f->texinfo].flags & (8 | 4)) continue;
    ofs = lightdatasize;
    lightdatasize += fl->numstyles * (fl->numsamples * 3);
    if (lightdatasize > 0x200000 || lightdatasize > cap) { lightdatasize = -1; break; }
    lightofs[fn] = ofs;
    styles[fn * 4] = 0; styles[fn * 4 + 1] = styles[fn * 4 + 2] = styles[fn * 4 + 3] = 0xff;
    trian->numpoints = trian->numedges = trian->numtris = 0; trian->matrix = NULL;
    if (S.numbounce > 0) {
      facemins[0] = facemins[1] = facemins[2] = 99999; facemaxs[0] = facemaxs[1] = facemaxs[2] = -99999;
      for (i = 0; i < f->numedges; i++) {
        int e = S.surfedges[f->firstedge + i];
        const float *v = S.verts + 3 * (e >= 0 ? S.edges[e].v[0] : S.edges[-e].v[1]);
        for (k = 0; k < 3; k++) { if (v[k] < facemins[k]) facemins[k] = v[k]; if (v[k] > facemaxs[k]) facemaxs[k] = v[k]; }
      }
      trian->plane_normal = S.planes[f->planenum].normal;
      for (pf = planelinks[(f->side ? 1 : 0) * S.numplanes + f->planenum]; pf; pf = facelinks[pf])
        for (pi = S.face_head[pf]; pi >= 0; pi = S.patches[pi].next) {
          opatch *p = &S.patches[pi];
          for (i = 0; i < 3; i++) {
            if (facemins[i] - p->origin[i] > S.subdiv * 2) break;
            if (p->origin[i] - facemaxs[i] > S.subdiv * 2) break;
          }
          if (i != 3) continue;
          if (trian->numpoints == MAX_TRI_POINTS) { oq_error = 3; break; }
          trian->points[trian->numpoints++] = pi;
        }
      trian->matrix = calloc((size_t) trian->numpoints * trian->numpoints + 1, 4);
      triangulate(trian);
    }
    if (oq_error) { free(trian->matrix); lightdatasize = -1; break; }
    minlight = S.model_minlight[S.face_model[fn]];
    dest = out + ofs;
    nst = fl->numstyles > 4 ? 4 : fl->numstyles;
    for (st = 0; st < nst; st++) {
      styles[fn * 4 + st] = (uint8_t) fl->stylenums[st];
      for (j = 0; j < fl->numsamples; j++) {
        vec3 lb; float max, newmax;
        memcpy(lb, fl->samples[st] + j * 3, 12);
        if (S.numbounce > 0 && st == 0) {
          vec3 add;
          sample_trian(fl->origins + j * 3, trian, add);
          for (k = 0; k < 3; k++) lb[k] = lb[k] + add[k];
        }
        for (k = 0; k < 3; k++) lb[k] += S.ambient;
        for (k = 0; k < 3; k++) lb[k] = S.lightscale * lb[k];
        for (k = 0; k < 3; k++) if (lb[k] < 1) lb[k] = 1;
        max = lb[0];
        if (lb[1] > max) max = lb[1];
        if (lb[2] > max) max = lb[2];
        newmax = max;
        if (newmax < 0) newmax = 0;
        if (newmax < minlight) newmax = minlight + (rand() % 48);
        if (newmax > S.maxlight) newmax = S.maxlight;
        for (k = 0; k < 3; k++) *dest++ = lb[k] * newmax / max;
      }
    }
    free(trian->matrix);
  }
  free(trian); free(facelinks); free(planelinks);
  return lightdatasize;
}
