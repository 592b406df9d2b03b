// Host-side set-up of the qrad3 drop-in: texture reflectivity, patches and their
// dicing, the direct-light lists and the per-face lightmap frames.  These run once
// per map in milliseconds; they stay on the host but must be result-identical to
// the reference because patch origins / areas / clusters feed every kernel
// (SURVEY.md section 8a, row a12).  The float/double mix of each expression follows
// the reference's C promotion rules (vec_t = float, FLT_EVAL_METHOD 0, no FMA):
// see the comments citing common/mathlib.c and common/polylib.c below.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "host_scene.hpp"

namespace qrad {
namespace {

// ---- math with the reference's rounding points --------------------------------
inline float dot3(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// VectorNormalize, mathlib.c:106-122: sqrt and reciprocal are evaluated in double and rounded to float
inline float normalize(const float *in, float *out) {
  float length = (float) sqrt((double) (in[0] * in[0] + in[1] * in[1] + in[2] * in[2]));
  if (length == 0) {
    out[0] = out[1] = out[2] = 0;
    return 0;
  }
  float ilength = (float) (1.0 / (double) length);
  out[0] = in[0] * ilength;
  out[1] = in[1] * ilength;
  out[2] = in[2] * ilength;
  return length;
}
// VectorLength, mathlib.c:30-41: float products accumulated in double
inline double length_d(const float *v) {
  double l = 0;
  for (int i = 0; i < 3; i++) l += (double) (v[i] * v[i]);
  return sqrt(l);
}
// VectorMA, mathlib.c:59-64: double multiply-add, rounded to float on store
inline void vector_ma(const float *va, double scale, const float *vb, float *vc) {
  vc[0] = (float) ((double) va[0] + scale * (double) vb[0]);
  vc[1] = (float) ((double) va[1] + scale * (double) vb[1]);
  vc[2] = (float) ((double) va[2] + scale * (double) vb[2]);
}
// ColorNormalize, mathlib.c:124-141; returns max, leaves out untouched when max == 0
inline float color_normalize(const float *in, float *out) {
  float max = in[0];
  if (in[1] > max) max = in[1];
  if (in[2] > max) max = in[2];
  if (max == 0) return 0;
  float scale = (float) (1.0 / (double) max);
  out[0] = scale * in[0];
  out[1] = scale * in[1];
  out[2] = scale * in[2];
  return max;
}
inline void cross(const float *a, const float *b, float *c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// ---- windings (common/polylib.c) --------------------------------------------
float winding_area(const Winding &w) {  // :137-151
  float total = 0;
  for (size_t i = 2; i < w.p.size(); i++) {
    float d1[3], d2[3], c[3];
    for (int k = 0; k < 3; k++) {
      d1[k] = w.p[i - 1][k] - w.p[0][k];
      d2[k] = w.p[i][k] - w.p[0][k];
    }
    cross(d1, d2, c);
    total = (float) ((double) total + 0.5 * length_d(c));
  }
  return total;
}
void winding_bounds(const Winding &w, float *mins, float *maxs) {  // :153-170
  mins[0] = mins[1] = mins[2] = 99999;
  maxs[0] = maxs[1] = maxs[2] = -99999;
  for (auto &p : w.p)
    for (int j = 0; j < 3; j++) {
      if (p[j] < mins[j]) mins[j] = p[j];
      if (p[j] > maxs[j]) maxs[j] = p[j];
    }
}
void winding_center(const Winding &w, float *c) {  // :177-188
  c[0] = c[1] = c[2] = 0;
  for (auto &p : w.p)
    for (int j = 0; j < 3; j++) c[j] = p[j] + c[j];
  float scale = (float) (1.0 / (double) (int) w.p.size());
  for (int j = 0; j < 3; j++) c[j] = scale * c[j];
}
void remove_colinear(Winding &w) {  // :86-117
  int n = (int) w.p.size();
  std::vector<Vec3> keep;
  for (int i = 0; i < n; i++) {
    int j = (i + 1) % n, k = (i + n - 1) % n;
    float v1[3], v2[3];
    for (int a = 0; a < 3; a++) {
      v1[a] = w.p[j][a] - w.p[i][a];
      v2[a] = w.p[i][a] - w.p[k][a];
    }
    normalize(v1, v1);
    normalize(v2, v2);
    if ((double) dot3(v1, v2) < 0.999) keep.push_back(w.p[i]);
  }
  if ((int) keep.size() != n) w.p.swap(keep);
}
// ClipWindingEpsilon, :297-392, for an axial split normal; epsilon is a float parameter (0.1f)
void clip_winding(const Winding &in, const float *normal, float dist, float epsilon, Winding &front, Winding &back,
                  bool &has_front, bool &has_back) {
  int n = (int) in.p.size();
  std::vector<float> dists(n + 1);
  std::vector<int> sides(n + 1);
  int counts[3] = {0, 0, 0};
  for (int i = 0; i < n; i++) {
    float d = dot3(in.p[i].v, normal);
    d -= dist;
    dists[i] = d;
    sides[i] = d > epsilon ? 0 : (d < -epsilon ? 1 : 2);
    counts[sides[i]]++;
  }
  sides[n] = sides[0];
  dists[n] = dists[0];
  front.p.clear();
  back.p.clear();
  has_front = has_back = false;
  if (!counts[0]) {
    back = in;
    has_back = true;
    return;
  }
  if (!counts[1]) {
    front = in;
    has_front = true;
    return;
  }
  has_front = has_back = true;
  for (int i = 0; i < n; i++) {
    const Vec3 &p1 = in.p[i];
    if (sides[i] == 2) {
      front.p.push_back(p1);
      back.p.push_back(p1);
      continue;
    }
    if (sides[i] == 0) front.p.push_back(p1);
    if (sides[i] == 1) back.p.push_back(p1);
    if (sides[i + 1] == 2 || sides[i + 1] == sides[i]) continue;
    const Vec3 &p2 = in.p[(i + 1) % n];
    float d = dists[i] / (dists[i] - dists[i + 1]);
    Vec3 mid;
    for (int j = 0; j < 3; j++) {
      if (normal[j] == 1)
        mid[j] = dist;
      else if (normal[j] == -1)
        mid[j] = -dist;
      else
        mid[j] = p1[j] + d * (p2[j] - p1[j]);
    }
    front.p.push_back(mid);
    back.p.push_back(mid);
  }
  if (front.p.size() > 64 || back.p.size() > 64) fail("ClipWinding: MAX_POINTS_ON_WINDING");
}

std::vector<uint8_t> read_file(const std::string &path, bool required) {
  std::vector<uint8_t> out;
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) {
    if (required) fail("Error opening %s: %s", path.c_str(), strerror(errno));
    return out;
  }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  out.resize((size_t) n + 1);  // + NUL like LoadFile (cmdlib.c:560-575); size()-1 is the length
  if (n && fread(out.data(), 1, (size_t) n, f) != (size_t) n) {
    fclose(f);
    fail("File read failure");
  }
  fclose(f);
  out[(size_t) n] = 0;
  return out;
}

}  // namespace

// ---------------------------------------------------------------------------
// CalcTextureReflectivity, patches.c:40-112
void calc_texture_reflectivity(Scene &s) {
  const Bsp &b = s.bsp;
  std::string pal_path = b.gamedir + "../pics/colormap.pcx";
  std::vector<uint8_t> pcx = read_file(pal_path, true);
  size_t len = pcx.size() - 1;
  // LoadPCX header checks, lbmlib.c:373-376
  if (len < 128 + 768 || pcx[0] != 0x0a || pcx[1] != 5 || pcx[2] != 1 || pcx[3] != 8 ||
      (pcx[8] | (pcx[9] << 8)) >= 640 || (pcx[10] | (pcx[11] << 8)) >= 480)
    fail("Bad pcx file %s", pal_path.c_str());
  const uint8_t *palette = pcx.data() + len - 768;

  int nt = b.numtexinfo();
  const qrad_texinfo_t *ti = b.as<qrad_texinfo_t>(kLumpTexinfo);
  s.texture_reflectivity.assign(nt > 0 ? nt : 1, Vec3{{0, 0, 0}});
  s.texture_reflectivity[0] = Vec3{{0.5f, 0.5f, 0.5f}};
  for (int i = 0; i < nt; i++) {
    int j;
    for (j = 0; j < i; j++)
      if (!strncmp(ti[i].texture, ti[j].texture, 32)) {
        s.texture_reflectivity[i] = s.texture_reflectivity[j];
        break;
      }
    if (j != i) continue;
    char name[33];
    memcpy(name, ti[i].texture, 32);
    name[32] = 0;
    std::string path = b.gamedir + "../textures/" + name + ".wal";
    std::vector<uint8_t> wal = read_file(path, false);
    if (wal.empty()) {
      printf("Couldn't load %s\n", path.c_str());
      s.texture_reflectivity[i] = Vec3{{0.5f, 0.5f, 0.5f}};
      continue;
    }
    size_t wlen = wal.size() - 1;
    if (wlen < 100) fail("%s: truncated miptex", path.c_str());
    uint32_t width, height, ofs0;
    memcpy(&width, wal.data() + 32, 4);
    memcpy(&height, wal.data() + 36, 4);
    memcpy(&ofs0, wal.data() + 40, 4);
    int texels = (int) (width * height);
    if (texels <= 0 || (size_t) ofs0 + (size_t) texels > wlen) fail("%s: bad miptex", path.c_str());
    int color[3] = {0, 0, 0};
    for (int t = 0; t < texels; t++) {
      int texel = wal[ofs0 + t];
      for (int k = 0; k < 3; k++) color[k] += palette[texel * 3 + k];
    }
    float *r = s.texture_reflectivity[i].v;
    for (int k = 0; k < 3; k++) r[k] = (float) ((double) (color[k] / texels) / 255.0);
    float scale = color_normalize(r, r);
    if ((double) scale < 0.5) {
      scale *= 2;
      for (int k = 0; k < 3; k++) r[k] = scale * r[k];
    }
  }
}

// ---------------------------------------------------------------------------
namespace {

int entity_for_model(const Scene &s, int modnum) {  // patches.c:253-268
  char name[16];
  snprintf(name, sizeof name, "*%i", modnum);
  for (size_t i = 0; i < s.bsp.entities.size(); i++)
    if (!strcmp(s.bsp.entities[i].value("model"), name)) return (int) i;
  return 0;
}

// the tail shared by MakePatchForFace (:222-236) and FinishSplit (:326-352)
void place_patch(Scene &s, int p) {
  float c[3];
  winding_center(s.winding[p], c);
  for (int k = 0; k < 3; k++) s.origin[p][k] = c[k] + s.normal[p][k];
  int leaf = s.bsp.point_in_leaf(s.origin[p].v);
  s.cluster[p] = s.bsp.as<qrad_dleaf_t>(kLumpLeafs)[leaf].cluster;
  if (s.cluster[p] == -1 && s.params.verbose) printf("patch->cluster == -1\n");
}

int new_patch(Scene &s) {
  if (s.params.max_patches && s.num_patches() == s.params.max_patches) fail("MAX_PATCHES");
  s.winding.emplace_back();
  s.origin.push_back(Vec3{{0, 0, 0}});
  s.normal.push_back(Vec3{{0, 0, 0}});
  s.reflectivity.push_back(Vec3{{0, 0, 0}});
  s.baselight.push_back(Vec3{{0, 0, 0}});
  s.totallight.push_back(Vec3{{0, 0, 0}});
  s.planedist.push_back(0);
  s.area.push_back(0);
  s.cluster.push_back(-1);
  s.face.push_back(-1);
  s.next.push_back(-1);
  s.sky.push_back(0);
  return s.num_patches() - 1;
}

// ---- DicePatch, patches.c:425-470 -------------------------------------------------------------------------
// The dicing of one face patch touches nothing but that patch and its own descendants, so the faces are diced in
// parallel, each into thread-local arrays, and the results are stitched together afterwards with the indices the
// reference's serial loop would have produced: the original keeps its index i, and the k-th patch created while
// dicing it (creation order = the order of the `num_patches++` in DicePatch: split, recurse into the front half,
// then into the back half) gets  numoriginals + (patches created by the originals before i) + k.
constexpr int kMaxWindingPoints = 64;   // MAX_POINTS_ON_WINDING, polylib.h

struct LocalPatches {          // everything one thread has produced so far
  std::vector<Vec3> pts;       // winding points, pooled
  std::vector<int64_t> wfirst;
  std::vector<int32_t> wcount, next, cluster;   // next: local index within the same original, -1 = the original's next
  std::vector<Vec3> origin;
  std::vector<float> area;
  int add() {
    wfirst.push_back(0); wcount.push_back(0); next.push_back(-1); cluster.push_back(-1);
    origin.push_back(Vec3{{0, 0, 0}}); area.push_back(0);
    return (int) wcount.size() - 1;
  }
};

struct Dicer {
  const Scene &s;
  LocalPatches &L;
  const float *normal;        // of the face being diced
  int base = 0;               // index in L of the original being diced
  const qrad_dleaf_t *leafs;

  void set_winding(int p, const Vec3 *w, int n) {
    L.wfirst[p] = (int64_t) L.pts.size();
    L.wcount[p] = n;
    L.pts.insert(L.pts.end(), w, w + n);
  }
  // the tail shared by MakePatchForFace (:222-236) and FinishSplit (:326-352)
  void finish(int p) {
    const Vec3 *w = &L.pts[(size_t) L.wfirst[p]];
    const int n = L.wcount[p];
    float total = 0;   // WindingArea, polylib.c:137-151
    for (int i = 2; i < n; i++) {
      float d1[3], d2[3], c[3];
      for (int k = 0; k < 3; k++) {
        d1[k] = w[i - 1][k] - w[0][k];
        d2[k] = w[i][k] - w[0][k];
      }
      cross(d1, d2, c);
      total = (float) ((double) total + 0.5 * length_d(c));
    }
    L.area[p] = total <= 1 ? 1 : total;
    float c[3] = {0, 0, 0};   // WindingCenter, polylib.c:177-188
    for (int i = 0; i < n; i++)
      for (int j = 0; j < 3; j++) c[j] = w[i][j] + c[j];
    const float scale = (float) (1.0 / (double) n);
    for (int j = 0; j < 3; j++) L.origin[p][j] = scale * c[j] + normal[j];
    L.cluster[p] = leafs[s.bsp.point_in_leaf(L.origin[p].v)].cluster;
  }
  // `dirty`: the winding of p has changed since its area / origin / cluster were computed.  The reference recomputes
  // them after every split (FinishSplit); only the values of the final patches are ever read, so they are computed once,
  // when a patch turns out not to be split any further.
  void dice(int p, bool dirty) {
    const Vec3 *w = &L.pts[(size_t) L.wfirst[p]];
    const int n = L.wcount[p];
    float mins[3] = {99999, 99999, 99999}, maxs[3] = {-99999, -99999, -99999};
    for (int i = 0; i < n; i++)
      for (int j = 0; j < 3; j++) {
        if (w[i][j] < mins[j]) mins[j] = w[i][j];
        if (w[i][j] > maxs[j]) maxs[j] = w[i][j];
      }
    const float subdiv = s.params.subdiv;
    int ax;
    for (ax = 0; ax < 3; ax++)
      if (floor((double) ((mins[ax] + 1) / subdiv)) < floor((double) ((maxs[ax] - 1) / subdiv))) break;
    if (ax == 3) {
      if (dirty) finish(p);
      return;
    }
    const float dist = (float) ((double) subdiv * (1 + floor((double) ((mins[ax] + 1) / subdiv))));
    // ClipWindingEpsilon (polylib.c:297-392) against the axial plane x[ax] = dist, epsilon 0.1f
    Vec3 in[kMaxWindingPoints + 4], front[kMaxWindingPoints + 4], back[kMaxWindingPoints + 4];
    float dists[kMaxWindingPoints + 5];
    int sides[kMaxWindingPoints + 5], counts[3] = {0, 0, 0};
    if (n > kMaxWindingPoints) fail("ClipWinding: MAX_POINTS_ON_WINDING");
    for (int i = 0; i < n; i++) {
      in[i] = w[i];
      float d = in[i][ax];   // DotProduct with the axial normal: x*0 + y*0 + z*1 adds exact zeros
      d -= dist;
      dists[i] = d;
      sides[i] = d > 0.1f ? 0 : (d < -0.1f ? 1 : 2);
      counts[sides[i]]++;
    }
    sides[n] = sides[0];
    dists[n] = dists[0];
    if (!counts[0] || !counts[1]) fail("DicePatch: split produced an empty side");
    int nf = 0, nb = 0;
    for (int i = 0; i < n; i++) {
      const Vec3 &p1 = in[i];
      if (sides[i] == 2) {
        front[nf++] = p1;
        back[nb++] = p1;
        continue;
      }
      if (sides[i] == 0) front[nf++] = p1;
      if (sides[i] == 1) back[nb++] = p1;
      if (sides[i + 1] == 2 || sides[i + 1] == sides[i]) continue;
      const Vec3 &p2 = in[(i + 1) % n];
      const float d = dists[i] / (dists[i] - dists[i + 1]);
      Vec3 mid;
      for (int j = 0; j < 3; j++) mid[j] = j == ax ? dist : p1[j] + d * (p2[j] - p1[j]);
      front[nf++] = mid;
      back[nb++] = mid;
      if (nf > kMaxWindingPoints || nb > kMaxWindingPoints) fail("ClipWinding: MAX_POINTS_ON_WINDING");
    }
    const int np = L.add();
    L.next[np] = L.next[p];
    L.next[p] = np - base;
    set_winding(p, front, nf);
    set_winding(np, back, nb);
    dice(p, true);
    dice(np, true);
  }
};

int host_threads(const Scene &s, int work_items) {
  int n = s.params.threads > 0 ? s.params.threads : (int) std::thread::hardware_concurrency();
  if (n < 1) n = 1;
  if (n > 64) n = 64;
  if (n > work_items) n = work_items > 0 ? work_items : 1;
  return n;
}

template <class F> void parallel_for(int nthreads, F &&f) {   // f(thread index); exceptions are carried over
  std::vector<std::thread> th;
  std::vector<std::string> err((size_t) nthreads);
  for (int t = 0; t < nthreads; t++)
    th.emplace_back([&, t] {
      try {
        f(t);
      } catch (const std::exception &e) {
        err[t] = e.what()[0] ? e.what() : "error";
      }
    });
  for (auto &t : th) t.join();
  for (auto &e : err)
    if (!e.empty()) fail("%s", e.c_str());
}

}  // namespace

// MakePatches + MakePatchForFace, patches.c:187-309
void make_patches(Scene &s) {
  const Bsp &b = s.bsp;
  const qrad_dface_t *faces = b.as<qrad_dface_t>(kLumpFaces);
  const qrad_dplane_t *planes = b.as<qrad_dplane_t>(kLumpPlanes);
  const qrad_texinfo_t *ti = b.as<qrad_texinfo_t>(kLumpTexinfo);
  const DModel *models = b.as<DModel>(kLumpModels);
  const int32_t *surfedges = b.as<int32_t>(kLumpSurfEdges);
  const DEdge *edges = b.as<DEdge>(kLumpEdges);
  const DVertex *verts = b.as<DVertex>(kLumpVertexes);
  int nf = b.numfaces();
  s.face_head.assign(nf, -1);
  s.face_offset.assign(nf, Vec3{{0, 0, 0}});
  s.face_entity.assign(nf, 0);
  s.totalarea = 0;
  float totalarea = 0;
  if (s.params.verbose) printf("%i faces\n", nf);
  for (int m = 0; m < b.nummodels(); m++) {
    const DModel &mod = models[m];
    int ent = entity_for_model(s, m);
    float origin[3];
    s.bsp.entities.empty() ? (void) (origin[0] = origin[1] = origin[2] = 0) : s.bsp.entities[ent].vector("origin", origin);
    for (int j = 0; j < mod.numfaces; j++) {
      int fn = mod.firstface + j;
      if (fn < 0 || fn >= nf) fail("model %d references face %d out of range", m, fn);
      s.face_entity[fn] = ent;
      for (int k = 0; k < 3; k++) s.face_offset[fn][k] = origin[k];
      const qrad_dface_t &f = faces[fn];
      // WindingFromFace, :122-145
      Winding w;
      w.p.resize(f.numedges);
      for (int e = 0; e < f.numedges; e++) {
        int se = surfedges[f.firstedge + e];
        int v = se < 0 ? edges[-se].v[1] : edges[se].v[0];
        for (int k = 0; k < 3; k++) w.p[e][k] = verts[v].point[k];
      }
      remove_colinear(w);
      for (auto &p : w.p)
        for (int k = 0; k < 3; k++) p[k] = p[k] + origin[k];
      // MakePatchForFace
      float area = winding_area(w);
      totalarea += area;
      int p = new_patch(s);
      s.next[p] = s.face_head[fn];
      s.face_head[fn] = p;
      s.face[p] = fn;
      s.winding[p] = std::move(w);
      const qrad_dplane_t &pl = planes[f.planenum];
      if (f.side) {  // backplanes[], qrad3.c:88-96
        for (int k = 0; k < 3; k++) s.normal[p][k] = 0 - pl.normal[k];
        s.planedist[p] = -pl.dist;
      } else {
        for (int k = 0; k < 3; k++) s.normal[p][k] = pl.normal[k];
        s.planedist[p] = pl.dist;
      }
      if (origin[0] || origin[1] || origin[2])  // origin-offset bmodel faces get a shifted plane (:214-223)
        s.planedist[p] += dot3(origin, s.normal[p].v);
      place_patch(s, p);
      s.area[p] = area;
      if (s.area[p] <= 1) s.area[p] = 1;
      s.sky[p] = (ti[f.texinfo].flags & kSurfSky) ? 1 : 0;
      s.reflectivity[p] = s.texture_reflectivity[f.texinfo];
      if (fn < models[0].numfaces) {  // non-bmodel patches can emit light
        const qrad_texinfo_t &tx = ti[f.texinfo];
        float *bl = s.baselight[p].v;
        if (!(tx.flags & kSurfLight) || tx.value == 0) {
          bl[0] = bl[1] = bl[2] = 0;
        } else {
          float v = (float) tx.value;
          for (int k = 0; k < 3; k++) bl[k] = v * s.texture_reflectivity[f.texinfo][k];
        }
        float color[3] = {0, 0, 0};  // (the reference leaves this uninitialised for an all-black texture)
        color_normalize(s.reflectivity[p].v, color);
        for (int k = 0; k < 3; k++) bl[k] *= color[k];
        s.totallight[p] = s.baselight[p];
      }
    }
  }
  s.totalarea = totalarea;
  if (s.params.verbose) printf("%i sqaure feet\n", (int) (totalarea / 64));
}

namespace {
struct Where { int32_t pool, first, count; };   // original i: its patches are `count` consecutive entries of pools[pool]
struct PoolView {              // what stitch_patches reads of a pool: a thread's LocalPatches, or another process's blob
  const Vec3 *pts, *origin;
  const int64_t *wfirst;
  const int32_t *wcount, *next, *cluster;
  const float *area;
};
PoolView view_of(const LocalPatches &L) {
  return PoolView{L.pts.data(), L.origin.data(), L.wfirst.data(), L.wcount.data(), L.next.data(), L.cluster.data(), L.area.data()};
}

// DicePatch for the originals [lo, hi) on `nth` threads, each into a pool of its own
void dice_range(Scene &s, int lo, int hi, int nth, std::vector<LocalPatches> &locals, std::vector<Where> &where) {
  const qrad_dleaf_t *leafs = s.bsp.as<qrad_dleaf_t>(kLumpLeafs);
  const bool dice = !(s.params.subdiv < 1);
  locals.assign((size_t) nth, LocalPatches());
  std::atomic<int> cursor{lo};
  parallel_for(nth, [&](int t) {
    LocalPatches &L = locals[t];
    for (;;) {
      const int i0 = cursor.fetch_add(16);
      if (i0 >= hi) break;
      for (int i = i0; i < std::min(hi, i0 + 16); i++) {
        Dicer d{s, L, s.normal[i].v, 0, leafs};
        d.base = L.add();
        d.set_winding(d.base, s.winding[i].p.data(), (int) s.winding[i].p.size());
        L.origin[d.base] = s.origin[i];
        L.area[d.base] = s.area[i];
        L.cluster[d.base] = s.cluster[i];
        if (dice) d.dice(d.base, false);
        where[i] = Where{t, d.base, (int32_t) L.wcount.size() - d.base};
      }
    }
  });
}

// The results in the reference's order: original i stays i, its created patches follow the originals in creation order
void stitch_patches(Scene &s, const std::vector<PoolView> &locals, const std::vector<Where> &where, int nth) {
  const int num = s.num_patches();
  std::vector<int64_t> created_before((size_t) num + 1, 0), pts_before((size_t) num + 1, 0);
  for (int i = 0; i < num; i++) {
    const Where &w = where[i];
    created_before[i + 1] = created_before[i] + (w.count - 1);
    int64_t np = 0;
    const PoolView &L = locals[w.pool];
    for (int k = 0; k < w.count; k++) np += L.wcount[w.first + k];
    pts_before[i + 1] = pts_before[i] + np;
  }
  const int64_t total = num + created_before[num];
  if (s.params.max_patches && total > s.params.max_patches) fail("MAX_PATCHES");
  if (total > 0x7fffffff) fail("MAX_PATCHES");
  const size_t N = (size_t) total;
  s.origin.resize(N); s.normal.resize(N); s.reflectivity.resize(N); s.baselight.resize(N); s.totallight.resize(N);
  s.planedist.resize(N); s.area.resize(N); s.cluster.resize(N); s.face.resize(N); s.next.resize(N); s.sky.resize(N);
  s.wfirst.resize(N); s.wcount.resize(N); s.bounds.resize(N * 6);
  s.wpts.resize((size_t) pts_before[num]);
  std::atomic<int> cursor2{0};
  parallel_for(nth, [&](int) {
    for (;;) {
      const int i0 = cursor2.fetch_add(64);
      if (i0 >= num) break;
      for (int i = i0; i < std::min(num, i0 + 64); i++) {
        const Where &w = where[i];
        const PoolView &L = locals[w.pool];
        const int64_t gbase = num + created_before[i];
        auto gidx = [&](int k) -> int64_t { return k == 0 ? i : gbase + k - 1; };
        const int32_t orig_next = s.next[i];
        int64_t po = pts_before[i];
        for (int k = 0; k < w.count; k++) {
          const int l = w.first + k;
          const size_t g = (size_t) gidx(k);
          s.origin[g] = L.origin[l];
          s.area[g] = L.area[l];
          s.cluster[g] = L.cluster[l];
          s.next[g] = L.next[l] < 0 ? orig_next : (int32_t) gidx(L.next[l]);
          if (k) {   // FinishSplit copies the parent's face-constant fields, patches.c:326-336
            s.normal[g] = s.normal[i]; s.reflectivity[g] = s.reflectivity[i]; s.baselight[g] = s.baselight[i];
            s.totallight[g] = s.totallight[i]; s.planedist[g] = s.planedist[i]; s.sky[g] = s.sky[i]; s.face[g] = s.face[i];
          }
          s.wfirst[g] = po;
          s.wcount[g] = L.wcount[l];
          const Vec3 *src = &L.pts[(size_t) L.wfirst[l]];
          float *mn = &s.bounds[g * 6], *mx = mn + 3;   // WindingBounds, polylib.c:153-170
          mn[0] = mn[1] = mn[2] = 99999;
          mx[0] = mx[1] = mx[2] = -99999;
          for (int q = 0; q < L.wcount[l]; q++) {
            s.wpts[(size_t) po + q] = src[q];
            for (int j = 0; j < 3; j++) {
              if (src[q][j] < mn[j]) mn[j] = src[q][j];
              if (src[q][j] > mx[j]) mx[j] = src[q][j];
            }
          }
          po += L.wcount[l];
        }
      }
    }
  });
  if (s.params.verbose) {
    const bool dice = !(s.params.subdiv < 1);
    for (size_t g = 0; g < N; g++)
      if (s.cluster[g] == -1 && (int) g >= num) printf("patch->cluster == -1\n");
    if (dice) printf("%i patches after subdivision\n", s.num_patches());
  }
  s.winding.clear();
  s.winding.shrink_to_fit();
}

// Several processes that light the same map (one per GPU) share the dicing: part k of n dices a contiguous range of the
// originals -- ranges of equal winding area, the same on every process -- and hands its results on as one flat blob;
// every process then stitches all the blobs together exactly as it stitches its own threads' pools.
void part_range(const Scene &s, int part, int parts, int &lo, int &hi) {
  const int num = s.num_patches();
  std::vector<double> cum((size_t) num + 1, 0.0);
  for (int i = 0; i < num; i++) cum[i + 1] = cum[i] + (double) std::max(s.area[i], 1.0f);
  auto cut = [&](int k) -> int {
    if (k <= 0) return 0;
    if (k >= parts) return num;
    const double target = cum[num] * k / parts;
    return (int) (std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
  };
  lo = std::min(cut(part), num);
  hi = std::max(lo, std::min(cut(part + 1), num));
}

struct BlobHeader {
  uint32_t magic;
  int32_t lo, hi, num;       // originals [lo, hi) of num
  int64_t patches, points;   // entries that follow
};
constexpr uint32_t kBlobMagic = 0x51504431;   // "QPD1"
}  // namespace

void subdivide_patches(Scene &s) {  // patches.c:477-492
  const int num = s.num_patches();   // the originals: one per face
  const int nth = host_threads(s, num / 16);
  const bool timing = getenv("QRAD_DEBUG_TIMING") != nullptr;
  auto tp = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!timing) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "  subdivide: %-18s %7.1f ms (%d threads)\n", what, std::chrono::duration<double, std::milli>(now - tp).count(), nth);
    tp = now;
  };
  std::vector<LocalPatches> locals;
  std::vector<Where> where((size_t) num);
  dice_range(s, 0, num, nth, locals, where);
  lap("dice");
  std::vector<PoolView> views;
  for (const LocalPatches &L : locals) views.push_back(view_of(L));
  stitch_patches(s, views, where, nth);
  lap("merge");
}

// part `part` of `parts`: dices its range of the originals; `blob` = {header, patches per original, then per patch origin,
// area, cluster, next, winding point count, and all winding points} in the order of the originals
void subdivide_part(Scene &s, int part, int parts, std::vector<uint8_t> &blob) {
  if (parts < 1 || part < 0 || part >= parts) fail("prepare_part: bad part");
  const int num = s.num_patches();
  int lo, hi;
  part_range(s, part, parts, lo, hi);
  const int nth = host_threads(s, (hi - lo) / 16);
  std::vector<LocalPatches> locals;
  std::vector<Where> where((size_t) num);
  dice_range(s, lo, hi, nth, locals, where);
  BlobHeader h{kBlobMagic, lo, hi, num, 0, 0};
  for (int i = lo; i < hi; i++) {
    h.patches += where[i].count;
    const LocalPatches &L = locals[where[i].pool];
    for (int k = 0; k < where[i].count; k++) h.points += L.wcount[where[i].first + k];
  }
  const size_t n = (size_t) h.patches;
  blob.resize(sizeof h + (size_t) (hi - lo) * 4 + n * (12 + 4 + 4 + 4 + 4) + (size_t) h.points * 12);
  uint8_t *o = blob.data();
  memcpy(o, &h, sizeof h);
  int32_t *counts = reinterpret_cast<int32_t *>(o + sizeof h);
  Vec3 *origin = reinterpret_cast<Vec3 *>(counts + (hi - lo));
  float *area = reinterpret_cast<float *>(origin + n);
  int32_t *cluster = reinterpret_cast<int32_t *>(area + n), *next = cluster + n, *wcount = next + n;
  Vec3 *pts = reinterpret_cast<Vec3 *>(wcount + n);
  std::vector<int64_t> at_of((size_t) (hi - lo) + 1, 0), pat_of((size_t) (hi - lo) + 1, 0);
  for (int i = lo; i < hi; i++) {
    const Where &w = where[i];
    const LocalPatches &L = locals[w.pool];
    int64_t np = 0;
    for (int k = 0; k < w.count; k++) np += L.wcount[w.first + k];
    at_of[i - lo + 1] = at_of[i - lo] + w.count;
    pat_of[i - lo + 1] = pat_of[i - lo] + np;
  }
  std::atomic<int> cursor{lo};
  parallel_for(nth, [&](int) {
    for (;;) {
      const int i0 = cursor.fetch_add(64);
      if (i0 >= hi) break;
      for (int i = i0; i < std::min(hi, i0 + 64); i++) {
        const Where &w = where[i];
        const LocalPatches &L = locals[w.pool];
        counts[i - lo] = w.count;
        size_t at = (size_t) at_of[i - lo], pat = (size_t) pat_of[i - lo];
        for (int k = 0; k < w.count; k++, at++) {
          const int l = w.first + k;
          origin[at] = L.origin[l];
          area[at] = L.area[l];
          cluster[at] = L.cluster[l];
          next[at] = L.next[l];
          wcount[at] = L.wcount[l];
          memcpy(pts + pat, &L.pts[(size_t) L.wfirst[l]], (size_t) L.wcount[l] * 12);
          pat += (size_t) L.wcount[l];
        }
      }
    }
  });
}

// all parts' blobs (own one included) -> the patches, as subdivide_patches leaves them
void subdivide_merge(Scene &s, int parts, const uint8_t *const *blobs, const int64_t *bytes) {
  const int num = s.num_patches();
  std::vector<PoolView> views((size_t) parts);        // straight into the blobs (4-byte aligned: the header is 32 bytes)
  std::vector<std::vector<int64_t>> wfirsts((size_t) parts);
  std::vector<Where> where((size_t) num, Where{-1, 0, 0});
  for (int q = 0; q < parts; q++) {
    BlobHeader h;
    if (bytes[q] < (int64_t) sizeof h) fail("prepare_merge: blob too small");
    if ((uintptr_t) blobs[q] & 3) fail("prepare_merge: blobs must be 4-byte aligned");
    memcpy(&h, blobs[q], sizeof h);
    if (h.magic != kBlobMagic || h.num != num || h.lo < 0 || h.hi < h.lo || h.hi > num || h.patches < h.hi - h.lo || h.points < 0)
      fail("prepare_merge: blob of part %d does not belong to this map", q);
    const size_t n = (size_t) h.patches;
    if ((uint64_t) bytes[q] != sizeof h + (size_t) (h.hi - h.lo) * 4 + n * 28 + (size_t) h.points * 12) fail("prepare_merge: blob size");
    const uint8_t *o = blobs[q] + sizeof h;
    const int32_t *counts = reinterpret_cast<const int32_t *>(o);
    const Vec3 *origin = reinterpret_cast<const Vec3 *>(counts + (h.hi - h.lo));
    const float *area = reinterpret_cast<const float *>(origin + n);
    const int32_t *cluster = reinterpret_cast<const int32_t *>(area + n), *next = cluster + n, *wcount = next + n;
    const Vec3 *pts = reinterpret_cast<const Vec3 *>(wcount + n);
    std::vector<int64_t> &wf = wfirsts[q];
    wf.resize(n);
    int64_t pat = 0;
    for (size_t k = 0; k < n; k++) {
      if (wcount[k] < 0 || wcount[k] > kMaxWindingPoints + 4) fail("prepare_merge: bad winding");
      wf[k] = pat;
      pat += wcount[k];
    }
    if (pat != h.points) fail("prepare_merge: point count");
    views[q] = PoolView{pts, origin, wf.data(), wcount, next, cluster, area};
    int64_t at = 0;
    for (int i = h.lo; i < h.hi; i++) {
      const int cnt = counts[i - h.lo];
      if (cnt < 1 || at + cnt > h.patches || where[i].pool >= 0) fail("prepare_merge: parts overlap or are malformed");
      where[i] = Where{q, (int32_t) at, cnt};
      at += cnt;
    }
    if (at != h.patches) fail("prepare_merge: patch count");
  }
  for (int i = 0; i < num; i++)
    if (where[i].pool < 0) fail("prepare_merge: original %d is in no part", i);
  stitch_patches(s, views, where, host_threads(s, num / 16));
}

// ---------------------------------------------------------------------------
// CreateDirectLights, lightmap.c:721-838.  Lists are LIFO per cluster.
void create_direct_lights(Scene &s) {
  const Bsp &b = s.bsp;
  const qrad_dleaf_t *leafs = b.as<qrad_dleaf_t>(kLumpLeafs);
  int nc = b.numclusters();
  struct L {
    int cluster, type, style;
    float intensity, stopdot;
    Vec3 origin, color, normal;
  };
  std::vector<L> made;  // in creation order
  s.numdlights = 0;
  for (int i = 0; i < s.num_patches(); i++) {
    float *tl = s.totallight[i].v;
    if (tl[0] < 3 && tl[1] < 3 && tl[2] < 3) continue;
    s.numdlights++;
    L l{};
    l.origin = s.origin[i];
    l.cluster = leafs[b.point_in_leaf(l.origin.v)].cluster;
    l.type = 0;
    l.normal = s.normal[i];
    l.intensity = color_normalize(tl, l.color.v);
    l.intensity *= s.area[i] * s.params.direct_scale;
    tl[0] = tl[1] = tl[2] = 0;
    made.push_back(l);
  }
  for (size_t i = 0; i < b.entities.size(); i++) {
    const Entity &e = b.entities[i];
    const char *name = e.value("classname");
    if (strncmp(name, "light", 5)) continue;
    s.numdlights++;
    L l{};
    e.vector("origin", l.origin.v);
    l.style = (int) e.number("_style");
    if (!l.style) l.style = (int) e.number("style");
    if (l.style < 0 || l.style >= 256) l.style = 0;
    l.cluster = leafs[b.point_in_leaf(l.origin.v)].cluster;
    float intensity = e.number("light");
    if (!intensity) intensity = e.number("_light");
    if (!intensity) intensity = 300;
    const char *col = e.value("_color");
    if (col[0] && col[1]) {  // finding F4: the reference reads col[1] of "" (UB); this is the intended test
      sscanf(col, "%f %f %f", &l.color[0], &l.color[1], &l.color[2]);
      color_normalize(l.color.v, l.color.v);
    } else
      l.color[0] = l.color[1] = l.color[2] = 1.0f;
    l.intensity = intensity * s.params.entity_scale;
    l.type = 1;
    const char *target = e.value("target");
    if (!strcmp(name, "light_spot") || target[0]) {
      l.type = 2;
      l.stopdot = e.number("_cone");
      if (!l.stopdot) l.stopdot = 10;
      l.stopdot = (float) cos((double) (l.stopdot / 180) * 3.14159);
      if (target[0]) {
        const Entity *e2 = nullptr;
        for (auto &c : b.entities)
          if (!strcmp(c.value("targetname"), target)) {
            e2 = &c;
            break;
          }
        if (!e2)
          printf("WARNING: light at (%i %i %i) has missing target\n", (int) l.origin[0], (int) l.origin[1], (int) l.origin[2]);
        else {
          float dest[3];
          e2->vector("origin", dest);
          for (int k = 0; k < 3; k++) l.normal[k] = dest[k] - l.origin[k];
          normalize(l.normal.v, l.normal.v);
        }
      } else {
        float angle = e.number("angle");
        if (angle == -1) {
          l.normal[0] = l.normal[1] = 0;
          l.normal[2] = 1;
        } else if (angle == -2) {
          l.normal[0] = l.normal[1] = 0;
          l.normal[2] = -1;
        } else {
          l.normal[2] = 0;
          l.normal[0] = (float) cos((double) (angle / 180) * 3.14159);
          l.normal[1] = (float) sin((double) (angle / 180) * 3.14159);
        }
      }
    }
    made.push_back(l);
  }
  if (s.params.verbose) printf("%i direct lights\n", s.numdlights);

  // flatten: cluster ascending, within a cluster most recently created first
  s.light_first.assign(nc + 1, 0);
  int dropped = 0;
  for (auto &l : made) {
    if (l.cluster < 0 || l.cluster >= nc) {
      dropped++;  // the reference indexes directlights[-1] here (out of bounds); such a light can never be gathered
      continue;
    }
    s.light_first[l.cluster + 1]++;
  }
  if (dropped && nc == 0)   // no vis lump: leaves carry no cluster, and the reference's GatherSampleLight finds no light either
    printf("WARNING: the map has no vis information, its %d direct lights cannot be gathered\n", dropped);
  else if (dropped)
    printf("WARNING: %d direct lights sit in solid (cluster -1) and are ignored\n", dropped);
  for (int c = 0; c < nc; c++) s.light_first[c + 1] += s.light_first[c];
  int total = s.light_first[nc];
  s.light_type.assign(total, 0);
  s.light_style.assign(total, 0);
  s.light_intensity.assign(total, 0);
  s.light_stopdot.assign(total, 0);
  s.light_origin.assign(total, Vec3{{0, 0, 0}});
  s.light_color.assign(total, Vec3{{0, 0, 0}});
  s.light_normal.assign(total, Vec3{{0, 0, 0}});
  std::vector<int> fill(s.light_first.begin(), s.light_first.end() - 1);
  for (auto it = made.rbegin(); it != made.rend(); ++it) {
    if (it->cluster < 0 || it->cluster >= nc) continue;
    int k = fill[it->cluster]++;
    s.light_type[k] = it->type;
    s.light_style[k] = it->style;
    s.light_intensity[k] = it->intensity;
    s.light_stopdot[k] = it->stopdot;
    s.light_origin[k] = it->origin;
    s.light_color[k] = it->color;
    s.light_normal[k] = it->normal;
  }
}

// ---------------------------------------------------------------------------
// Per-face frames: CalcFaceVectors :536-603, CalcFaceExtents :480-528 (as called from
// BuildFacelights :993-1008), LinkPlaneFaces :42-52, FinalLightFace's vertex bounds :1109-1117.
void build_face_frames(Scene &s) {
  const Bsp &b = s.bsp;
  int nf = b.numfaces();
  const qrad_dface_t *faces = b.as<qrad_dface_t>(kLumpFaces);
  const qrad_dplane_t *planes = b.as<qrad_dplane_t>(kLumpPlanes);
  const qrad_texinfo_t *ti = b.as<qrad_texinfo_t>(kLumpTexinfo);
  const int32_t *surfedges = b.as<int32_t>(kLumpSurfEdges);
  const DEdge *edges = b.as<DEdge>(kLumpEdges);
  const DVertex *verts = b.as<DVertex>(kLumpVertexes);
  s.face_lit.assign(nf, 0);
  s.facenormal.assign(nf, Vec3{{0, 0, 0}});
  s.texorg.assign(nf, Vec3{{0, 0, 0}});
  s.plane_normal.assign(nf, Vec3{{0, 0, 0}});
  s.textoworld.assign((size_t) nf * 6, 0);
  s.exactmins.assign((size_t) nf * 2, 0);
  s.exactmaxs.assign((size_t) nf * 2, 0);
  s.face_bounds.assign((size_t) nf * 6, 0);
  s.minlight.assign(nf, 0);
  s.texmins.assign((size_t) nf * 2, 0);
  s.texsize.assign((size_t) nf * 2, 0);
  s.planelink_head.assign(nf, 0);
  s.facelinks.assign(nf, 0);

  std::vector<int32_t> planelinks[2];
  planelinks[0].assign(b.numplanes(), 0);
  planelinks[1].assign(b.numplanes(), 0);
  for (int i = 0; i < nf; i++) {
    int side = faces[i].side ? 1 : 0;
    s.facelinks[i] = planelinks[side][faces[i].planenum];
    planelinks[side][faces[i].planenum] = i;
  }

  for (int fn = 0; fn < nf; fn++) {
    const qrad_dface_t &f = faces[fn];
    const qrad_texinfo_t &tex = ti[f.texinfo];
    s.planelink_head[fn] = planelinks[f.side ? 1 : 0][f.planenum];
    s.plane_normal[fn] = Vec3{{planes[f.planenum].normal[0], planes[f.planenum].normal[1], planes[f.planenum].normal[2]}};
    s.minlight[fn] = s.bsp.entities.empty() ? 0.0f : s.bsp.entities[s.face_entity[fn]].number("_minlight") * 128;
    {
      float *mn = &s.face_bounds[fn * 6], *mx = mn + 3;
      mn[0] = mn[1] = mn[2] = 99999;
      mx[0] = mx[1] = mx[2] = -99999;
      for (int e = 0; e < f.numedges; e++) {
        int se = surfedges[f.firstedge + e];
        const float *v = verts[se >= 0 ? edges[se].v[0] : edges[-se].v[1]].point;
        for (int k = 0; k < 3; k++) {
          if (v[k] < mn[k]) mn[k] = v[k];
          if (v[k] > mx[k]) mx[k] = v[k];
        }
      }
    }
    if (tex.flags & (kSurfWarp | kSurfSky)) continue;
    s.face_lit[fn] = 1;

    float facenormal[3], facedist;
    for (int k = 0; k < 3; k++) facenormal[k] = planes[f.planenum].normal[k];
    facedist = planes[f.planenum].dist;
    if (f.side) {
      for (int k = 0; k < 3; k++) facenormal[k] = 0 - facenormal[k];
      facedist = -facedist;
    }
    for (int k = 0; k < 3; k++) s.facenormal[fn][k] = facenormal[k];
    const float *modelorg = s.face_offset[fn].v;

    // CalcFaceVectors
    float worldtotex[2][3], textoworld[2][3], texnormal[3], texorg[3];
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 3; j++) worldtotex[i][j] = tex.vecs[i][j];
    texnormal[0] = tex.vecs[1][1] * tex.vecs[0][2] - tex.vecs[1][2] * tex.vecs[0][1];
    texnormal[1] = tex.vecs[1][2] * tex.vecs[0][0] - tex.vecs[1][0] * tex.vecs[0][2];
    texnormal[2] = tex.vecs[1][0] * tex.vecs[0][1] - tex.vecs[1][1] * tex.vecs[0][0];
    normalize(texnormal, texnormal);
    float distscale = dot3(texnormal, facenormal);
    if (!distscale) {
      if (s.params.verbose) printf("WARNING: Texture axis perpendicular to face\n");
      distscale = 1;
    }
    if (distscale < 0) {
      distscale = -distscale;
      for (int k = 0; k < 3; k++) texnormal[k] = 0 - texnormal[k];
    }
    distscale = 1 / distscale;
    for (int i = 0; i < 2; i++) {
      float len = (float) length_d(worldtotex[i]);
      float dist = dot3(worldtotex[i], facenormal);
      dist *= distscale;
      vector_ma(worldtotex[i], (double) -dist, texnormal, textoworld[i]);
      float sc = (1 / len) * (1 / len);
      for (int k = 0; k < 3; k++) textoworld[i][k] = sc * textoworld[i][k];
    }
    for (int i = 0; i < 3; i++) texorg[i] = -tex.vecs[0][3] * textoworld[0][i] - tex.vecs[1][3] * textoworld[1][i];
    float dist = dot3(texorg, facenormal) - facedist - 1;
    dist *= distscale;
    vector_ma(texorg, (double) -dist, texnormal, texorg);
    for (int k = 0; k < 3; k++) texorg[k] = texorg[k] + modelorg[k];
    for (int k = 0; k < 3; k++) s.texorg[fn][k] = texorg[k];
    for (int i = 0; i < 2; i++)
      for (int k = 0; k < 3; k++) s.textoworld[fn * 6 + i * 3 + k] = textoworld[i][k];

    // CalcFaceExtents
    float mins[2] = {999999, 999999}, maxs[2] = {-99999, -99999};
    for (int e = 0; e < f.numedges; e++) {
      int se = surfedges[f.firstedge + e];
      const float *vt = verts[se >= 0 ? edges[se].v[0] : edges[-se].v[1]].point;
      for (int j = 0; j < 2; j++) {
        float val = dot3(vt, tex.vecs[j]) + tex.vecs[j][3];
        if (val < mins[j]) mins[j] = val;
        if (val > maxs[j]) maxs[j] = val;
      }
    }
    int texsize[2] = {0, 0};
    for (int i = 0; i < 2; i++) {
      s.exactmins[fn * 2 + i] = mins[i];
      s.exactmaxs[fn * 2 + i] = maxs[i];
      float mn = (float) floor((double) (mins[i] / 16));
      float mx = (float) ceil((double) (maxs[i] / 16));
      s.texmins[fn * 2 + i] = (int) mn;
      texsize[i] = (int) (mx - mn);
      s.texsize[fn * 2 + i] = texsize[i];
      if (texsize[0] * texsize[1] > (64 * 64 * 4) / 4) fail("Surface to large to map");
    }
  }
}

}  // namespace qrad
