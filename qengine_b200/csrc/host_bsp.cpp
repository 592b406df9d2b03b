// IBSP v38 reader / writer and entity-string parser for the qrad3 drop-in.
// Behaviour follows common/bspfile.c (LoadBSPFile :354, WriteBSPFile :466, ParseEntities :625,
// ValueForKey :707, FloatForKey :717, GetVectorForKey :725), common/scriplib.c (GetToken :157)
// and common/cmdlib.c (SetQdirFromPath :187).  New code; nothing here is shared with oracle/.
#include <strings.h>
#include <unistd.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "host_scene.hpp"

namespace qrad {

void fail(const char *fmt, ...) {
  char buf[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw HostError(buf);
}

// ---------------------------------------------------------------------------
const char *Entity::value(const char *key) const {
  for (auto it = epairs.rbegin(); it != epairs.rend(); ++it)
    if (it->first == key) return it->second.c_str();
  return "";
}
float Entity::number(const char *key) const { return (float) atof(value(key)); }
void Entity::vector(const char *key, float out[3]) const {
  double a = 0, b = 0, c = 0;
  sscanf(value(key), "%lf %lf %lf", &a, &b, &c);
  out[0] = (float) a;
  out[1] = (float) b;
  out[2] = (float) c;
}

namespace {

struct Header {
  int32_t ident, version;
  struct { int32_t fileofs, filelen; } lumps[kNumLumps];
};
constexpr int32_t kIdent = ('P' << 24) + ('S' << 16) + ('B' << 8) + 'I';
constexpr int32_t kVersion = 38;

// element size per lump for the "odd lump size" check of CopyLump (bspfile.c:340-350)
int lump_elem_size(int l) {
  switch (l) {
    case kLumpPlanes: return 20;
    case kLumpVertexes: return 12;
    case kLumpNodes: return 28;
    case kLumpTexinfo: return 76;
    case kLumpFaces: return 20;
    case kLumpLeafs: return 28;
    case kLumpLeafFaces: return 2;
    case kLumpLeafBrushes: return 2;
    case kLumpEdges: return 4;
    case kLumpSurfEdges: return 4;
    case kLumpModels: return 48;
    case kLumpBrushes: return 12;
    case kLumpBrushSides: return 4;
    case kLumpAreas: return 8;
    case kLumpAreaPortals: return 8;
    default: return 1;
  }
}

// Tokenizer with the reference's rules: whitespace <= 32, ';' '#' '//' line comments, quoted tokens.
struct Tokens {
  const char *p, *end;
  std::string tok;
  bool next() {
    for (;;) {
      while (p < end && (unsigned char) *p <= 32) p++;
      if (p >= end) return false;
      if (*p == ';' || *p == '#' || (p + 1 < end && p[0] == '/' && p[1] == '/')) {
        while (p < end && *p != '\n') p++;
        continue;
      }
      break;
    }
    tok.clear();
    if (*p == '"') {
      p++;
      while (p < end && *p != '"') tok.push_back(*p++);
      p++;
    } else {
      while (p < end && (unsigned char) *p > 32 && *p != ';') tok.push_back(*p++);
    }
    // the reference stores tokens in a NUL-terminated buffer: anything after an embedded NUL is lost
    size_t z = tok.find('\0');
    if (z != std::string::npos) tok.resize(z);
    return true;
  }
};

void strip_trailing(std::string &s) {
  while (!s.empty() && (signed char) s.back() <= 32) s.pop_back();
}

void parse_entities(Bsp &bsp) {
  bsp.entities.clear();
  const auto &ed = bsp.lump[kLumpEntities];
  Tokens t{reinterpret_cast<const char *>(ed.data()), reinterpret_cast<const char *>(ed.data()) + ed.size(), {}};
  while (t.next()) {
    if (t.tok.empty()) break;  // trailing NUL of the lump
    if (t.tok != "{") fail("ParseEntity: { not found");
    if (bsp.entities.size() == 2048) fail("num_entities == MAX_MAP_ENTITIES");
    Entity e;
    for (;;) {
      if (!t.next()) fail("ParseEntity: EOF without closing brace");
      if (t.tok == "}") break;
      std::string key = t.tok;
      if (key.size() >= 31) fail("ParseEpar: token too long");
      if (!t.next()) fail("ParseEntity: EOF without closing brace");
      std::string val = t.tok;
      if (val.size() >= 1023) fail("ParseEpar: token too long");
      strip_trailing(key);
      strip_trailing(val);
      e.epairs.emplace_back(std::move(key), std::move(val));
    }
    bsp.entities.push_back(std::move(e));
  }
}

}  // namespace

// ---------------------------------------------------------------------------
void set_gamedir_from_path(const std::string &path_in, Bsp &bsp) {
  std::string path = path_in;
  if (!(path.size() && (path[0] == '/' || path[0] == '\\' || (path.size() > 1 && path[1] == ':')))) {
    char cwd[1024];
    if (!getcwd(cwd, sizeof cwd - 1)) fail("getcwd failed");
    std::string c = cwd;
    if (c.empty() || c.back() != '/') c += '/';
    path = c + path;
  }
  const char *base = path.c_str();
  for (const char *c = base + path.size() - 1; c != base; c--) {
    if (!strncasecmp(c, "assets", 6)) {
      c += 7;
      if (c > base + path.size()) break;
      while (*c) {
        if (*c == '/' || *c == '\\') {
          bsp.gamedir.assign(base, c + 1 - base);
          return;
        }
        c++;
      }
      fail("No gamedir in %s", base);
    }
  }
  fail("SetQdirFromPath: no 'assets' in %s", base);
}

void load_bsp(const std::string &path, Bsp &bsp) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) fail("Error opening %s: %s", path.c_str(), strerror(errno));
  fseek(f, 0, SEEK_END);
  long len = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> file((size_t) len);
  if (len && fread(file.data(), 1, (size_t) len, f) != (size_t) len) {
    fclose(f);
    fail("File read failure");
  }
  fclose(f);
  if ((size_t) len < sizeof(Header)) fail("%s is not a IBSP file", path.c_str());
  Header h;
  memcpy(&h, file.data(), sizeof h);
  if (h.ident != kIdent) fail("%s is not a IBSP file", path.c_str());
  if (h.version != kVersion) fail("%s is version %i, not %i", path.c_str(), h.version, kVersion);
  for (int l = 0; l < kNumLumps; l++) {
    int32_t ofs = h.lumps[l].fileofs, n = h.lumps[l].filelen;
    if (n < 0 || ofs < 0 || (long) ofs + n > len) fail("LoadBSPFile: lump %d out of file bounds", l);
    if (n % lump_elem_size(l)) fail("LoadBSPFile: odd lump size");
    bsp.lump[l].assign(file.begin() + ofs, file.begin() + ofs + n);
  }
  // dlightdata is a static MAX_MAP_LIGHTING buffer in the reference: bytes past the new
  // lightdatasize keep whatever the input file held there (only visible as lump padding).
  bsp.lighting.assign(kMaxMapLighting + 4, 0);
  bsp.lightdatasize = (int32_t) bsp.lump[kLumpLighting].size();
  if (bsp.lightdatasize > kMaxMapLighting) fail("MAX_MAP_LIGHTING");
  memcpy(bsp.lighting.data(), bsp.lump[kLumpLighting].data(), bsp.lump[kLumpLighting].size());
  parse_entities(bsp);
}

void write_bsp(const Bsp &bsp, const std::string &path) {
  Header h;
  memset(&h, 0, sizeof h);
  h.ident = kIdent;
  h.version = kVersion;
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) fail("Error opening %s: %s", path.c_str(), strerror(errno));
  auto put = [&](const void *p, size_t n) {
    if (n && fwrite(p, 1, n, f) != n) {
      fclose(f);
      fail("File write failure");
    }
  };
  put(&h, sizeof h);
  auto add = [&](int l, const uint8_t *data, size_t n) {
    h.lumps[l].fileofs = (int32_t) ftell(f);
    h.lumps[l].filelen = (int32_t) n;
    put(data, n);
    static const uint8_t zeros[4] = {0, 0, 0, 0};
    put(zeros, ((n + 3) & ~size_t(3)) - n);
  };
  // fixed order of WriteBSPFile, bspfile.c:479-498
  static const int order[] = {kLumpPlanes, kLumpLeafs, kLumpVertexes, kLumpNodes, kLumpTexinfo, kLumpFaces,
                              kLumpBrushes, kLumpBrushSides, kLumpLeafFaces, kLumpLeafBrushes, kLumpSurfEdges,
                              kLumpEdges, kLumpModels, kLumpAreas, kLumpAreaPortals};
  for (int l : order) add(l, bsp.lump[l].data(), bsp.lump[l].size());
  {
    // lighting: padding bytes come from the static buffer, like SafeWrite(data, (len + 3) & ~3)
    size_t n = (size_t) bsp.lightdatasize;
    h.lumps[kLumpLighting].fileofs = (int32_t) ftell(f);
    h.lumps[kLumpLighting].filelen = (int32_t) n;
    put(bsp.lighting.data(), (n + 3) & ~size_t(3));
  }
  add(kLumpVisibility, bsp.lump[kLumpVisibility].data(), bsp.lump[kLumpVisibility].size());
  add(kLumpEntities, bsp.lump[kLumpEntities].data(), bsp.lump[kLumpEntities].size());
  {
    uint8_t pop[256];
    memset(pop, 0, sizeof pop);
    memcpy(pop, bsp.lump[kLumpPop].data(), bsp.lump[kLumpPop].size() < 256 ? bsp.lump[kLumpPop].size() : 256);
    add(kLumpPop, pop, sizeof pop);
  }
  fseek(f, 0, SEEK_SET);
  put(&h, sizeof h);
  fclose(f);
}

int Bsp::point_in_leaf(const float p[3]) const {
  const qrad_dnode_t *nodes = as<qrad_dnode_t>(kLumpNodes);
  const qrad_dplane_t *planes = as<qrad_dplane_t>(kLumpPlanes);
  int n = 0;
  while (n >= 0) {
    const qrad_dnode_t &nd = nodes[n];
    const qrad_dplane_t &pl = planes[nd.planenum];
    float d = (p[0] * pl.normal[0] + p[1] * pl.normal[1] + p[2] * pl.normal[2]) - pl.dist;
    n = d > 0 ? nd.children[0] : nd.children[1];
  }
  return -n - 1;
}

}  // namespace qrad
