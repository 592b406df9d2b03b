// Host half of the qvis3 first stage (include/qvis_cuda.h): LoadPortals and PlaneFromWinding.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/qvis_cuda.h"
#include "host_scene.hpp"

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


// ClusterMerge for every cluster + CalcPHS, qvis3.c:168-224, 427-486: the visibility lump {numclusters, bitofs[][2], the
// compressed PVS rows in cluster order, then the PHS rows} from the portals' final vectors.  Host work: a few OR-ed bit
// rows per cluster, serial in the lump offsets.  Returns 0, or 1 with the reference's error text.
int qvis_host_merge_lump(const qvis_host_portals *p, const uint8_t *portalvis, uint8_t *lump, int32_t cap, int32_t *size,
                         int32_t *avg_visible, int32_t *avg_hearable, int32_t *saw_into_leaf) {
  const int P = p->numportals, nc = p->numclusters;
  const int pb = qvis_portal_bytes(P), leafbytes = ((nc + 63) & ~63) >> 3, visrow = (nc + 7) >> 3;
  if (nc <= 0) {
    g_qvis_error = "ClusterMerge: no clusters";
    return 1;
  }
  int64_t at = 4 + (int64_t) 8 * nc;   // vismap_p starts behind bitofs[portalclusters], qvis3.c:352
  if (at > cap) {
    g_qvis_error = "Vismap expansion overflow";
    return 1;
  }
  memset(lump, 0, (size_t) at);
  int32_t n32 = nc;
  memcpy(lump, &n32, 4);
  std::vector<uint8_t> uncompressedvis((size_t) nc * leafbytes, 0), portalvector((size_t) pb), row((size_t) leafbytes), packed((size_t) 2 * visrow + 4);
  auto compress = [&](const uint8_t *vis) -> int {   // CompressVis, bspfile.c:95-124
    int o = 0;
    for (int j = 0; j < visrow; j++) {
      packed[o++] = vis[j];
      if (vis[j]) continue;
      int rep = 1;
      for (j++; j < visrow; j++)
        if (vis[j] || rep == 255) break;
        else rep++;
      packed[o++] = (uint8_t) rep;
      j--;
    }
    return o;
  };
  auto append = [&](int cluster, int which, int n) -> bool {
    if (at + n > cap) return false;
    const int32_t ofs = (int32_t) at;
    memcpy(lump + 4 + (size_t) 8 * cluster + 4 * which, &ofs, 4);
    memcpy(lump + at, packed.data(), (size_t) n);
    at += n;
    return true;
  };
  int64_t totalvis = 0;
  int warned = 0;
  for (int leafnum = 0; leafnum < nc; leafnum++) {
    std::fill(portalvector.begin(), portalvector.end(), 0);
    for (int e = p->leaf_first[leafnum]; e < p->leaf_first[leafnum + 1]; e++) {
      const int pnum = p->leaf_portals[e];
      const uint8_t *v = portalvis + (size_t) pnum * pb;
      for (int j = 0; j < pb; j++) portalvector[j] |= v[j];
      portalvector[pnum >> 3] |= (uint8_t) (1 << (pnum & 7));
    }
    // LeafVectorFromPortalVector, qvis3.c:138-156
    std::fill(row.begin(), row.end(), 0);
    for (int i = 0; i < P; i++)
      if (portalvector[i >> 3] & (1 << (i & 7))) row[p->leaf[i] >> 3] |= (uint8_t) (1 << (p->leaf[i] & 7));
    int numvis = 0;
    for (int i = 0; i < nc; i++) numvis += (row[i >> 3] >> (i & 7)) & 1;
    if (row[leafnum >> 3] & (1 << (leafnum & 7))) {
      printf("WARNING: Leaf portals saw into leaf\n");
      warned++;
    }
    row[leafnum >> 3] |= (uint8_t) (1 << (leafnum & 7));
    numvis++;   // (the reference counts the leaf itself again even when its portals saw into it)
    memcpy(&uncompressedvis[(size_t) leafnum * leafbytes], row.data(), (size_t) leafbytes);
    totalvis += numvis;
    if (!append(leafnum, 0, compress(row.data()))) {
      g_qvis_error = "Vismap expansion overflow";
      return 1;
    }
  }
  // CalcPHS: OR the PVS rows of everything a cluster sees
  int64_t count = 0;
  for (int i = 0; i < nc; i++) {
    const uint8_t *scan = &uncompressedvis[(size_t) i * leafbytes];
    memcpy(row.data(), scan, (size_t) leafbytes);
    for (int j = 0; j < leafbytes; j++) {
      const int bitbyte = scan[j];
      if (!bitbyte) continue;
      for (int k = 0; k < 8; k++) {
        if (!(bitbyte & (1 << k))) continue;
        const int index = (j << 3) + k;
        if (index >= nc) {
          g_qvis_error = "Bad bit in PVS";
          return 1;
        }
        const uint8_t *src = &uncompressedvis[(size_t) index * leafbytes];
        for (int l = 0; l < leafbytes; l++) row[l] |= src[l];
      }
    }
    for (int j = 0; j < nc; j++) count += (row[j >> 3] >> (j & 7)) & 1;
    if (!append(i, 1, compress(row.data()))) {
      g_qvis_error = "Vismap expansion overflow";
      return 1;
    }
  }
  *size = (int32_t) at;
  if (avg_visible) *avg_visible = (int32_t) (totalvis / nc);
  if (avg_hearable) *avg_hearable = (int32_t) (count / nc);
  if (saw_into_leaf) *saw_into_leaf = warned;
  return 0;
}

// qvis3 -fast (qvis3.c:493-560 with fastvis: CalcPortalVis takes portalflood for portalvis, :235-241): reads map.bsp and
// map.prt, BasePortalVis on the GPU, ClusterMerge + CalcPHS on the host, writes the BSP with the new visibility lump.
int qvis_host_fast_vis(const char *bsp_in, const char *prt_path, const char *bsp_out, int device) {
  try {
    qrad::Bsp bsp;
    qrad::load_bsp(bsp_in, bsp);
    if (bsp.numnodes() == 0 || bsp.numfaces() == 0) throw std::runtime_error("Empty map");
    qvis_host_portals *p = nullptr;
    printf("reading %s\n", prt_path);
    if (qvis_host_portals_load(prt_path, &p)) return 1;
    struct Guard {
      qvis_host_portals *p;
      ~Guard() { qvis_host_portals_free(p); }
    } guard{p};
    printf("%4i portalclusters\n", p->numclusters);
    printf("%4i numportals\n", p->numportals / 2);
    qvis_portals_view v;
    qvis_host_portals_view(p, &v);
    const int pb = qvis_portal_bytes(p->numportals);
    std::vector<uint8_t> front((size_t) p->numportals * pb), flood((size_t) p->numportals * pb);
    std::vector<int32_t> might((size_t) p->numportals);
    if (qvis_cuda_base_portal_vis(device, &v, front.data(), flood.data(), might.data(), nullptr, nullptr)) return 1;
    constexpr int kMaxMapVisibility = 0x100000;   // MAX_MAP_VISIBILITY, qfiles.h
    std::vector<uint8_t> lump((size_t) kMaxMapVisibility);
    int32_t size = 0, vis = 0, hear = 0, warned = 0;
    if (qvis_host_merge_lump(p, flood.data(), lump.data(), kMaxMapVisibility, &size, &vis, &hear, &warned)) return 1;
    printf("Average clusters visible: %i\n", vis);
    printf("Building PHS...\n");
    printf("Average clusters hearable: %i\n", hear);
    const int leafbytes = ((p->numclusters + 63) & ~63) >> 3;
    printf("visdatasize:%i  compressed from %i\n", size, p->numclusters * leafbytes * 2);
    bsp.lump[qrad::kLumpVisibility].assign(lump.begin(), lump.begin() + size);
    printf("writing %s\n", bsp_out);
    qrad::write_bsp(bsp, bsp_out);
    return 0;
  } catch (const std::exception &e) {
    g_qvis_error = e.what();
    return 1;
  }
}


// main() of the reference's qvis3 (qvis3.c:493-566) for what is built: the same flags are accepted, the same path rules
// apply (the map argument expanded against the working directory, extension replaced by .bsp / .prt, -tmpin / -tmpout
// prefix "/tmp"), the same chatter and error banner are printed.  Without -fast the tool stops: PortalFlow is not built.
int qvis3_main(int argc, char **argv) {
  printf("---- vis ----\n");
  bool fast = false, tmpin = false, tmpout = false;
  int device = 0, i;
  auto die = [](const std::string &msg) {   // Error(), cmdlib.c:142-155
    printf("\n************ ERROR ************\n%s\n", msg.c_str());
    return 1;
  };
  for (i = 1; i < argc; i++) {
    const char *a = argv[i];
    if (!strcmp(a, "-threads")) i++;   // (the reference's Linux build runs one thread whatever is asked, threads.c:379-420)
    else if (!strcmp(a, "-fast")) {
      printf("fastvis = true\n");
      fast = true;
    } else if (!strcmp(a, "-level")) {
      printf("testlevel = %i\n", i + 1 < argc ? atoi(argv[i + 1]) : 0);
      i++;
    } else if (!strcmp(a, "-v")) printf("verbose = true\n");
    else if (!strcmp(a, "-nosort")) printf("nosort = true\n");
    else if (!strcmp(a, "-tmpin")) tmpin = true;
    else if (!strcmp(a, "-tmpout")) tmpout = true;
    else if (!strcmp(a, "-gpu")) device = i + 1 < argc ? atoi(argv[++i]) : 0;   // extension: CUDA device ordinal
    else if (a[0] == '-') return die(std::string("Unknown option \"") + a + "\"");
    else break;
  }
  if (i != argc - 1) return die("usage: vis [-threads #] [-level 0-4] [-fast] [-v] bspfile");
  if (!fast)
    return die("PortalFlow is not built in this library (DESIGN.md section 9): run with -fast, or use the reference's qvis3");
  std::string source = argv[i];
  if (source[0] != '/') {   // ExpandArg, cmdlib.c:219-229
    char cwd[1024];
    if (!getcwd(cwd, sizeof cwd - 1)) return die("Couldn't determine current directory");
    source = std::string(cwd) + "/" + source;
  }
  const size_t slash = source.find_last_of('/'), dot = source.find_last_of('.');
  if (dot != std::string::npos && (slash == std::string::npos || dot > slash)) source.resize(dot);
  const std::string inname = (tmpin ? "/tmp" : "") + source + ".bsp", prt = (tmpin ? "/tmp" : "") + source + ".prt",
                    outname = (tmpout ? "/tmp" : "") + source + ".bsp";
  printf("reading %s\n", inname.c_str());
  if (qvis_host_fast_vis(inname.c_str(), prt.c_str(), outname.c_str(), device)) return die(g_qvis_error);
  return 0;
}

}  // extern "C"
