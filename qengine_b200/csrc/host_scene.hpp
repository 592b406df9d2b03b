// Host-side scene model of the qrad3 drop-in (internal to libqrad_cuda.so).
// Everything the reference keeps in process globals (common/bspfile.h:25-80,
// qrad3/qrad.h:93-139) lives in one Scene object here, as flat vectors ready to
// be handed to the device through the views of include/qrad_cuda.h.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/qrad_host.h"

namespace qrad {

struct HostError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
[[noreturn]] void fail(const char *fmt, ...);

// ---- IBSP v38 -------------------------------------------------------------
constexpr int kLumpEntities = 0, kLumpPlanes = 1, kLumpVertexes = 2, kLumpVisibility = 3, kLumpNodes = 4,
              kLumpTexinfo = 5, kLumpFaces = 6, kLumpLighting = 7, kLumpLeafs = 8, kLumpLeafFaces = 9,
              kLumpLeafBrushes = 10, kLumpEdges = 11, kLumpSurfEdges = 12, kLumpModels = 13, kLumpBrushes = 14,
              kLumpBrushSides = 15, kLumpPop = 16, kLumpAreas = 17, kLumpAreaPortals = 18, kNumLumps = 19;
constexpr int kMaxMapLighting = 0x200000;  // MAX_MAP_LIGHTING, qfiles.h:242
constexpr int kSurfLight = 0x1, kSurfSky = 0x4, kSurfWarp = 0x8;

struct DModel { float mins[3], maxs[3], origin[3]; int32_t headnode, firstface, numfaces; };  // 48 B
struct DVertex { float point[3]; };
struct DEdge { uint16_t v[2]; };

struct Vec3 {
  float v[3];
  float &operator[](int i) { return v[i]; }
  const float &operator[](int i) const { return v[i]; }
};

struct Entity {
  std::vector<std::pair<std::string, std::string>> epairs;  // in text order
  const char *value(const char *key) const;                   // ValueForKey: last occurrence wins, "" if absent
  float number(const char *key) const;                        // FloatForKey
  void vector(const char *key, float out[3]) const;           // GetVectorForKey
};

struct Bsp {
  std::vector<uint8_t> lump[kNumLumps];  // raw bytes as on disk
  std::vector<uint8_t> lighting;         // dlightdata[] with the reference's static-buffer semantics
  int32_t lightdatasize = 0;
  std::vector<Entity> entities;
  std::string gamedir;                   // ".../assets/<dir>/"

  template <class T> const T *as(int l) const { return reinterpret_cast<const T *>(lump[l].data()); }
  template <class T> T *as_mut(int l) { return reinterpret_cast<T *>(lump[l].data()); }
  template <class T> int count(int l) const { return int(lump[l].size() / sizeof(T)); }
  int numfaces() const { return count<qrad_dface_t>(kLumpFaces); }
  int numnodes() const { return count<qrad_dnode_t>(kLumpNodes); }
  int numleafs() const { return count<qrad_dleaf_t>(kLumpLeafs); }
  int numplanes() const { return count<qrad_dplane_t>(kLumpPlanes); }
  int numtexinfo() const { return count<qrad_texinfo_t>(kLumpTexinfo); }
  int nummodels() const { return count<DModel>(kLumpModels); }
  int visdatasize() const { return int(lump[kLumpVisibility].size()); }
  int numclusters() const {
    return visdatasize() >= 4 ? *reinterpret_cast<const int32_t *>(lump[kLumpVisibility].data()) : 0;
  }
  int point_in_leaf(const float p[3]) const;  // PointInLeafnum, qrad3.c:131-150
};

void load_bsp(const std::string &path, Bsp &bsp);        // LoadBSPFile + ParseEntities
void write_bsp(const Bsp &bsp, const std::string &path);  // WriteBSPFile
void set_gamedir_from_path(const std::string &path, Bsp &bsp);

// std::vector whose resize() leaves new elements uninitialised: the per-patch arrays of a 1M-patch map are hundreds of
// MB that are completely overwritten (in parallel) right after they are sized.
template <class T> struct DefaultInitAlloc : std::allocator<T> {
  template <class U> struct rebind { using other = DefaultInitAlloc<U>; };
  using std::allocator<T>::allocator;
  template <class U> void construct(U *p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new (static_cast<void *>(p)) U; }
  template <class U, class... A> void construct(U *p, A &&...a) { ::new (static_cast<void *>(p)) U(std::forward<A>(a)...); }
};
template <class T> using RawVec = std::vector<T, DefaultInitAlloc<T>>;

struct Winding {
  std::vector<Vec3> p;
};

struct Scene {
  Bsp bsp;
  qrad_host_params params{};
  bool prepared = false;

  std::vector<Vec3> texture_reflectivity;

  // patches (SoA)
  std::vector<Winding> winding;         // the face patches while MakePatches runs; emptied by subdivide_patches
  RawVec<Vec3> wpts;                    // winding points of all patches, pooled ...
  RawVec<int64_t> wfirst;               // ... first point and
  RawVec<int32_t> wcount;               // ... number of points of patch i
  RawVec<Vec3> origin, normal, reflectivity, baselight, totallight;
  RawVec<float> planedist, area, bounds;  // bounds: 6 per patch
  RawVec<int32_t> cluster, face, next;
  RawVec<uint8_t> sky;
  std::vector<int32_t> face_head;
  std::vector<Vec3> face_offset;
  std::vector<int32_t> face_entity;
  double totalarea = 0;

  // direct lights, flattened per cluster in list order
  std::vector<int32_t> light_first, light_type, light_style;
  std::vector<float> light_intensity, light_stopdot;
  std::vector<Vec3> light_origin, light_color, light_normal;
  int numdlights = 0;

  // per-face lightmap frames
  std::vector<uint8_t> face_lit;
  std::vector<Vec3> facenormal, texorg, plane_normal;
  std::vector<float> textoworld, exactmins, exactmaxs, face_bounds, minlight;
  std::vector<int32_t> texmins, texsize, planelink_head, facelinks;

  int num_patches() const { return int(origin.size()); }
};

void calc_texture_reflectivity(Scene &s);  // patches.c:40-112
void make_patches(Scene &s);               // patches.c:275-309 (+ MakeBackplanes, qrad3.c:88-96)
void subdivide_patches(Scene &s);          // patches.c:425-492
void subdivide_part(Scene &s, int part, int parts, std::vector<uint8_t> &blob);   // the same, shared by several processes
void subdivide_merge(Scene &s, int parts, const uint8_t *const *blobs, const int64_t *bytes);
void create_direct_lights(Scene &s);       // lightmap.c:721-838
void build_face_frames(Scene &s);          // lightmap.c:480-603, 42-52, 1109-1117

}  // namespace qrad

struct qrad_host_scene {
  qrad::Scene scene;
  std::vector<uint8_t> part_blob;   // qrad_host_prepare_part: this process's share of the dicing
};

namespace qrad {
// RadWorld on one context (host_api.cpp); the scene is only read unless `store`
void rad_world_impl(qrad_host_scene *hs, qrad_cuda_ctx *ctx, const qrad_host_params *p, float *added_out, bool store);
void set_host_error(const std::string &m);   // what qrad_host_last_error() returns on this thread
}
