// Host half of the qvis3 first stage (include/qvis_cuda.h): LoadPortals and PlaneFromWinding.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/qvis_cuda.h"

namespace {
thread_local std::string g_qvis_error;
constexpr int kMaxPointsOnWinding = 64;   // MAX_POINTS_ON_WINDING, vis.h:39
constexpr int kMaxPortalsOnLeaf = 128;    // MAX_PORTALS_ON_LEAF, vis.h:91
constexpr int kMaxPortals = 32768;        // MAX_PORTALS, vis.h:27 (memory portals)
}  // namespace

void qvis_set_error(const std::string &m) { g_qvis_error = m; }

struct qvis_host_portals {
  int numportals = 0, numclusters = 0;
  std::vector<float> planes, points;
  std::vector<int32_t> leaf, winding_first, leaf_first, leaf_portals;
};

extern "C" {

const char *qvis_last_error(void) { return g_qvis_error.c_str(); }

int qvis_portal_bytes(int32_t numportals) { return ((numportals + 63) & ~63) >> 3; }

// LoadPortals, qvis3.c:304-412.  The file: "PRT1", clusters, portals, then per portal
// "numpoints cluster0 cluster1 (x y z ) (x y z ) ...".
int qvis_host_portals_load(const char *path, qvis_host_portals **out) {
  *out = nullptr;
  FILE *f = fopen(path, "r");
  if (!f) {
    g_qvis_error = std::string("LoadPortals: couldn't read ") + path;
    return 1;
  }
  auto bad = [&](const std::string &m) {
    g_qvis_error = m;
    fclose(f);
    return 1;
  };
  char magic[80];
  int nc = 0, nfile = 0;
  if (fscanf(f, "%79s\n%i\n%i\n", magic, &nc, &nfile) != 3) return bad("LoadPortals: failed to read header");
  if (strcmp(magic, "PRT1")) return bad("LoadPortals: not a portal file");
  if (nc < 0 || nfile < 0 || 2 * (int64_t) nfile > kMaxPortals) return bad("LoadPortals: too many portals");
  auto *p = new qvis_host_portals;
  p->numclusters = nc;
  p->numportals = 2 * nfile;
  p->planes.resize((size_t) 8 * nfile);
  p->leaf.resize((size_t) 2 * nfile);
  p->winding_first.assign(1, 0);
  std::vector<std::vector<int32_t>> per_leaf((size_t) nc);
  for (int i = 0; i < nfile; i++) {
    int np, l0, l1;
    if (fscanf(f, "%i %i %i ", &np, &l0, &l1) != 3) { delete p; return bad("LoadPortals: reading portal " + std::to_string(i)); }
    if (np > kMaxPointsOnWinding || np < 3) { delete p; return bad("LoadPortals: portal " + std::to_string(i) + " has too many points"); }
    // (the reference's bound check lets a cluster number EQUAL to portalclusters pass and then writes past its leaf array)
    if ((unsigned) l0 >= (unsigned) nc || (unsigned) l1 >= (unsigned) nc) { delete p; return bad("LoadPortals: reading portal " + std::to_string(i)); }
    float w[kMaxPointsOnWinding][3];
    for (int j = 0; j < np; j++) {
      double v[3];   // scanned into doubles, then assigned to the float vec_t
      if (fscanf(f, "(%lf %lf %lf ) ", &v[0], &v[1], &v[2]) != 3) { delete p; return bad("LoadPortals: reading portal " + std::to_string(i)); }
      for (int k = 0; k < 3; k++) w[j][k] = (float) v[k];
    }
    if (fscanf(f, "\n") != 0) {}
    // PlaneFromWinding, qvis3.c:58-69
    float v1[3], v2[3], n[3];
    for (int k = 0; k < 3; k++) {
      v1[k] = w[2][k] - w[1][k];
      v2[k] = w[0][k] - w[1][k];
    }
    n[0] = v2[1] * v1[2] - v2[2] * v1[1];
    n[1] = v2[2] * v1[0] - v2[0] * v1[2];
    n[2] = v2[0] * v1[1] - v2[1] * v1[0];
    const float length = (float) sqrt((double) (n[0] * n[0] + n[1] * n[1] + n[2] * n[2]));   // VectorNormalize, mathlib.c:106-122
    if (length == 0) n[0] = n[1] = n[2] = 0;
    else {
      const float il = (float) (1.0 / (double) length);
      for (int k = 0; k < 3; k++) n[k] = n[k] * il;
    }
    const float dist = w[0][0] * n[0] + w[0][1] * n[1] + w[0][2] * n[2];
    if ((int) per_leaf[l0].size() == kMaxPortalsOnLeaf || (int) per_leaf[l1].size() == kMaxPortalsOnLeaf) {
      delete p;
      return bad("Leaf with too many portals");
    }
    // forward portal: flipped plane, leads into cluster 1, leaves cluster 0
    float *pl = &p->planes[(size_t) 8 * i];
    for (int k = 0; k < 3; k++) pl[k] = 0.f - n[k];
    pl[3] = -dist;
    p->leaf[2 * i] = l1;
    per_leaf[l0].push_back(2 * i);
    for (int j = 0; j < np; j++) p->points.insert(p->points.end(), w[j], w[j] + 3);
    p->winding_first.push_back((int32_t) (p->points.size() / 3));
    // backward portal: the plane itself, the winding reversed
    for (int k = 0; k < 3; k++) pl[4 + k] = n[k];
    pl[7] = dist;
    p->leaf[2 * i + 1] = l0;
    per_leaf[l1].push_back(2 * i + 1);
    for (int j = 0; j < np; j++) p->points.insert(p->points.end(), w[np - 1 - j], w[np - 1 - j] + 3);
    p->winding_first.push_back((int32_t) (p->points.size() / 3));
  }
  fclose(f);
  p->leaf_first.assign(1, 0);
  for (auto &l : per_leaf) {
    p->leaf_portals.insert(p->leaf_portals.end(), l.begin(), l.end());
    p->leaf_first.push_back((int32_t) p->leaf_portals.size());
  }
  *out = p;
  return 0;
}

void qvis_host_portals_free(qvis_host_portals *p) { delete p; }

int qvis_host_portals_view(const qvis_host_portals *p, qvis_portals_view *o) {
  o->numportals = p->numportals;
  o->numclusters = p->numclusters;
  o->planes = p->planes.data();
  o->leaf = p->leaf.data();
  o->winding_first = p->winding_first.data();
  o->points = p->points.data();
  o->leaf_first = p->leaf_first.data();
  o->leaf_portals = p->leaf_portals.data();
  return 0;
}

}  // extern "C"
