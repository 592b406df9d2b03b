// Host side of the trace tree: flatten_bsp() (used by qrad_cuda_upload_bsp), and the walkers of walkers.cuh compiled
// for the HOST behind C entry points (include/qrad_cuda.h, "host builds of the walkers") so that CPU tests can hold the
// product's own TestLine / PointInLeaf / cell proofs against the reference's golden rays and against each other.
#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <string>

#include "bsp_flat.hpp"
#include "walkers.cuh"

namespace qrad {

namespace {
[[noreturn]] void bfail(const std::string &m) { throw std::runtime_error(m); }
}  // namespace

void flatten_bsp(const qrad_bsp_view *b, FlatBsp &o) {
  if (b->numnodes <= 0 || b->numleafs <= 0) bfail("Empty map");
  if (b->numnodes > 65536 || b->numleafs > 65536) bfail("BSP exceeds MAX_MAP_NODES / MAX_MAP_LEAFS");
  o.nodes.assign((size_t) b->numnodes, make_int4(0, 0, 0, 0));
  o.normals.assign((size_t) b->numnodes, make_float4(0, 0, 0, 0));
  o.leaf_contents.resize(b->numleafs);
  o.leaf_cluster.resize(b->numleafs);
  for (int i = 0; i < b->numleafs; i++) {
    o.leaf_contents[i] = b->leafs[i].contents;
    o.leaf_cluster[i] = b->leafs[i].cluster;
  }
  // breadth-first numbering: queue of (disk node, depth); children get consecutive new indices
  std::vector<int> order, depth_of, newidx((size_t) b->numnodes, -1);
  order.push_back(0);
  depth_of.push_back(1);
  newidx[0] = 0;
  int maxdepth = 0;
  for (size_t q = 0; q < order.size(); q++) {
    const int dn_i = order[q];
    if (depth_of[q] > maxdepth) maxdepth = depth_of[q];
    const qrad_dnode_t &dn = b->nodes[dn_i];
    for (int k = 0; k < 2; k++) {
      const int ch = dn.children[k];
      if (ch < 0) continue;
      if (ch >= b->numnodes) bfail("node index " + std::to_string(ch) + " out of range");
      if (newidx[ch] >= 0) bfail("BSP node " + std::to_string(ch) + " reached twice (not a tree)");
      newidx[ch] = (int) order.size();
      order.push_back(ch);
      depth_of.push_back(depth_of[q] + 1);
    }
  }
  const int next_new = (int) order.size();
  o.nodes.resize((size_t) next_new);
  o.normals.resize((size_t) next_new);
  for (int me = 0; me < next_new; me++) {
    const qrad_dnode_t &dn = b->nodes[order[me]];
    if (dn.planenum < 0 || dn.planenum >= b->numplanes) bfail("plane index " + std::to_string(dn.planenum) + " out of range");
    const qrad_dplane_t &pl = b->planes[dn.planenum];
    int child[2];
    for (int k = 0; k < 2; k++) {
      const int ch = dn.children[k];
      if (ch < 0) {
        const int leaf = -ch - 1;
        if (leaf >= b->numleafs) bfail("leaf index " + std::to_string(leaf) + " out of range");
        child[k] = kLeafFlag | ((o.leaf_contents[leaf] & 1) ? kSolidBit : 0) | leaf;
      } else
        child[k] = newidx[ch];
    }
    int distbits;
    memcpy(&distbits, &pl.dist, 4);
    o.nodes[me] = make_int4(distbits, pl.type, child[0], child[1]);
    o.normals[me] = make_float4(pl.normal[0], pl.normal[1], pl.normal[2], pl.dist);
  }
  o.depth = maxdepth;
  // parent links in the new numbering, for the cells of the leaves
  std::vector<int> parent((size_t) next_new, -1), pside((size_t) next_new, 0);
  std::vector<int> lparent((size_t) b->numleafs, -1), lside((size_t) b->numleafs, 0);
  for (int me = 0; me < next_new; me++) {
    const int ch[2] = {o.nodes[me].z, o.nodes[me].w};
    for (int k = 0; k < 2; k++) {
      if (ch[k] >= 0) {
        parent[ch[k]] = me;
        pside[ch[k]] = k;
      } else {
        lparent[ch[k] & kLeafMask] = me;
        lside[ch[k] & kLeafMask] = k;
      }
    }
  }
  // cells of the leaves: a leaf whose ancestors are all axial planes is the box between the tightest of them
  // (shaft_blocked() in walkers.cuh); lo.w = 1 marks a SOLID leaf with such a cell
  o.cell_lo.resize((size_t) b->numleafs);
  o.cell_hi.resize((size_t) b->numleafs);
  float extent = 0;
  for (int l = 0; l < b->numleafs; l++) {
    float lo[3] = {-1e30f, -1e30f, -1e30f}, hi[3] = {1e30f, 1e30f, 1e30f};
    bool axial = true;
    int node = lparent[l], sd = lside[l];
    while (node >= 0) {
      const int type = o.nodes[node].y;
      float dist;
      memcpy(&dist, &o.nodes[node].x, 4);
      if (type < 3) {
        if (sd == 0) lo[type] = std::max(lo[type], dist);   // child 0 = in front of the plane
        else hi[type] = std::min(hi[type], dist);
        extent = std::max(extent, fabsf(dist));
      } else
        axial = false;
      sd = pside[node];
      node = parent[node];
    }
    const bool solid = (o.leaf_contents[l] & 1) != 0;
    const bool in_tree = lparent[l] >= 0;   // leaf 0 of an IBSP file is a dummy no node points to: it has no cell
    o.cell_lo[l] = make_float4(lo[0], lo[1], lo[2], (axial && solid && in_tree) ? 1.f : 0.f);
    o.cell_hi[l] = make_float4(hi[0], hi[1], hi[2], 0.f);
  }
  o.extent = extent;
}

}  // namespace qrad

// ---------------------------------------------------------------------------------------------------------------------
using namespace qrad;

struct qrad_host_tree {
  FlatBsp f;
  Tree tree() const { return Tree{f.nodes.data(), f.normals.data()}; }
};

static thread_local std::string g_walk_error;

extern "C" {

const char *qrad_host_tree_last_error(void) { return g_walk_error.c_str(); }

int qrad_host_tree_create(const qrad_bsp_view *bsp, qrad_host_tree **out) {
  *out = nullptr;
  try {
    qrad_host_tree *t = new qrad_host_tree;
    try {
      flatten_bsp(bsp, t->f);
    } catch (...) {
      delete t;
      throw;
    }
    *out = t;
    return 0;
  } catch (const std::exception &e) {
    g_walk_error = e.what();
    return 1;
  }
}

void qrad_host_tree_destroy(qrad_host_tree *t) { delete t; }

// out[0] nodes, [1] leafs, [2] depth, [3] solid leaves with a box cell
int qrad_host_tree_info(const qrad_host_tree *t, int32_t *out4, float *extent) {
  out4[0] = (int32_t) t->f.nodes.size();
  out4[1] = (int32_t) t->f.leaf_contents.size();
  out4[2] = t->f.depth;
  int n = 0;
  for (auto &c : t->f.cell_lo) n += c.w != 0.f;
  out4[3] = n;
  *extent = t->f.extent;
  return 0;
}

// the product's TestLine (walkers.cuh: test_line) on the host: results[i] = 1 blocked / 0 clear, blockers[i] (may be NULL)
// = the solid leaf that ended the walk or -1; start_nodes (may be NULL) = node to start each walk at
int qrad_host_test_lines(const qrad_host_tree *t, int64_t n, const float *starts, const float *stops, const int32_t *start_nodes,
                         uint8_t *results, int32_t *blockers) {
  const Tree tr = t->tree();
  for (int64_t i = 0; i < n; i++) {
    unsigned visits = 0;
    int leaf = -1;
    const float *s = starts + 3 * i, *e = stops + 3 * i;
    results[i] = (uint8_t) test_line<kTraceStackBig, false>(tr, s[0], s[1], s[2], e[0], e[1], e[2], visits, &leaf,
                                                           start_nodes ? start_nodes[i] : 0);
    if (blockers) blockers[i] = leaf;
  }
  return 0;
}

int qrad_host_point_in_leaf(const qrad_host_tree *t, int64_t n, const float *pts, int32_t *leafs) {
  const Tree tr = t->tree();
  for (int64_t i = 0; i < n; i++) leafs[i] = point_in_leaf(tr, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
  return 0;
}

// the node every segment from box A = [amn, amx] to box B = [bmn, bmx] reaches unclipped (walkers.cuh: common_start_node)
int qrad_host_common_start_node(const qrad_host_tree *t, const float *amn, const float *amx, const float *bmn, const float *bmx) {
  return common_start_node(t->tree(), make_float4(amn[0], amn[1], amn[2], 0), make_float4(amx[0], amx[1], amx[2], 0),
                           make_float4(bmn[0], bmn[1], bmn[2], 0), make_float4(bmx[0], bmx[1], bmx[2], 0));
}

// Cell proofs (walkers.cuh).  For every segment i: the first leaf in [leaf_lo, leaf_hi) whose cell proves it blocked
// (segment_through_cell), or -1.  margin <= 0 uses the library's own margin for this tree.
int qrad_host_prove_segments(const qrad_host_tree *t, float margin, int64_t n, const float *starts, const float *stops,
                             int32_t leaf_lo, int32_t leaf_hi, int32_t *proving_leaf) {
  if (!(margin > 0)) margin = proof_margin(t->f.extent);
  const int nl = (int) t->f.cell_lo.size();
  leaf_lo = std::max(0, leaf_lo);
  leaf_hi = std::min(nl, leaf_hi);
  for (int64_t i = 0; i < n; i++) {
    const float4 s = make_float4(starts[3 * i], starts[3 * i + 1], starts[3 * i + 2], 0);
    const float4 e = make_float4(stops[3 * i], stops[3 * i + 1], stops[3 * i + 2], 0);
    proving_leaf[i] = -1;
    for (int l = leaf_lo; l < leaf_hi; l++)
      if (segment_through_cell(t->f.cell_lo[l], t->f.cell_hi[l], margin, s, e)) {
        proving_leaf[i] = l;
        break;
      }
  }
  return 0;
}

// box A x box B against every leaf cell: the first leaf that proves ALL segments blocked (shaft_blocked), or -1
int qrad_host_prove_boxes(const qrad_host_tree *t, float margin, const float *amn, const float *amx, const float *bmn,
                          const float *bmx) {
  if (!(margin > 0)) margin = proof_margin(t->f.extent);
  const float4 a0 = make_float4(amn[0], amn[1], amn[2], 0), a1 = make_float4(amx[0], amx[1], amx[2], 0);
  const float4 b0 = make_float4(bmn[0], bmn[1], bmn[2], 0), b1 = make_float4(bmx[0], bmx[1], bmx[2], 0);
  for (size_t l = 0; l < t->f.cell_lo.size(); l++)
    if (shaft_blocked(t->f.cell_lo[l], t->f.cell_hi[l], margin, a0, a1, b0, b1)) return (int) l;
  return -1;
}

}  // extern "C"
