// extern "C" surface of the host half (include/qrad_host.h): scene life-cycle, the
// views handed to libqrad_cuda's upload calls, RadWorld and the qrad3 command line.
#include <chrono>
#include <memory>
#include <cstdio>
#include <cstdlib>
#include <unistd.h>

#include "host_scene.hpp"

using namespace qrad;

static thread_local std::string g_host_error;
namespace qrad {
void set_host_error(const std::string &m) { g_host_error = m; }
}

template <class F> static int guarded(F &&f) {
  try {
    f();
    return 0;
  } catch (const std::exception &e) {
    g_host_error = e.what();
    return 1;
  }
}

extern "C" {

const char *qrad_host_last_error(void) { return g_host_error.c_str(); }

void qrad_host_default_params(qrad_host_params *p) {  // qrad3.c:49-74
  memset(p, 0, sizeof *p);
  p->numbounce = 8;
  p->subdiv = 64;
  p->ambient = 0;
  p->maxlight = 196;
  p->lightscale = 1.0f;
  p->direct_scale = 0.4f;
  p->entity_scale = 1.0f;
}

// file_path is read; qdir_path (the command-line argument in the reference) decides where pics/ and textures/ are
// looked up (SetQdirFromPath, cmdlib.c:187-217) -- the two differ only under -tmpin.
int qrad_host_load_as(const char *file_path, const char *qdir_path, qrad_host_scene **out) {
  *out = nullptr;
  return guarded([&] {
    std::unique_ptr<qrad_host_scene> s(new qrad_host_scene);
    set_gamedir_from_path(qdir_path, s->scene.bsp);
    load_bsp(file_path, s->scene.bsp);
    calc_texture_reflectivity(s->scene);
    *out = s.release();
  });
}

int qrad_host_load(const char *bsp_path, qrad_host_scene **out) { return qrad_host_load_as(bsp_path, bsp_path, out); }

void qrad_host_free(qrad_host_scene *s) { delete s; }

// what every flavour of prepare starts with: qrad3.c:627-631, and a clean slate
static void prepare_begin(Scene &s, qrad_host_params *params) {
  if (!s.bsp.visdatasize()) {  // qrad3.c:627-631
    printf("No vis information, direct lighting only.\n");
    params->numbounce = 0;
    params->ambient = 0.1f;
  }
  s.params = *params;
  if (s.bsp.numnodes() == 0 || s.bsp.numfaces() == 0) fail("Empty map");
  s.winding.clear(); s.origin.clear(); s.normal.clear(); s.reflectivity.clear(); s.baselight.clear();
  s.totallight.clear(); s.planedist.clear(); s.area.clear(); s.cluster.clear(); s.face.clear();
  s.next.clear(); s.sky.clear(); s.wpts.clear(); s.wfirst.clear(); s.wcount.clear(); s.bounds.clear();
  s.prepared = false;
}

int qrad_host_prepare(qrad_host_scene *hs, qrad_host_params *params) {
  return guarded([&] {
    Scene &s = hs->scene;
    prepare_begin(s, params);
    const bool timing = getenv("QRAD_DEBUG_TIMING") != nullptr;
    auto tp = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
      if (!timing) return;
      auto now = std::chrono::steady_clock::now();
      fprintf(stderr, "prepare: %-22s %7.1f ms\n", what, std::chrono::duration<double, std::milli>(now - tp).count());
      tp = now;
    };
    make_patches(s);
    lap("MakePatches");
    subdivide_patches(s);
    lap("SubdividePatches");
    create_direct_lights(s);
    lap("CreateDirectLights");
    build_face_frames(s);
    lap("face frames");
    s.prepared = true;
  });
}

int qrad_host_prepare_part(qrad_host_scene *hs, qrad_host_params *params, int part, int parts, const uint8_t **blob,
                           int64_t *bytes) {
  *blob = nullptr;
  *bytes = 0;
  return guarded([&] {
    Scene &s = hs->scene;
    prepare_begin(s, params);
    make_patches(s);
    subdivide_part(s, part, parts, hs->part_blob);
    *blob = hs->part_blob.data();
    *bytes = (int64_t) hs->part_blob.size();
  });
}

int qrad_host_prepare_merge(qrad_host_scene *hs, int parts, const uint8_t *const *blobs, const int64_t *bytes) {
  return guarded([&] {
    Scene &s = hs->scene;
    if (s.prepared || s.winding.empty()) fail("prepare_merge: prepare_part has not run on this scene");
    if (parts < 1) fail("prepare_merge: no parts");
    subdivide_merge(s, parts, blobs, bytes);
    hs->part_blob.clear();
    hs->part_blob.shrink_to_fit();
    create_direct_lights(s);
    build_face_frames(s);
    s.prepared = true;
  });
}

int qrad_host_bsp_view(qrad_host_scene *hs, qrad_bsp_view *o) {
  const Bsp &b = hs->scene.bsp;
  o->planes = b.as<qrad_dplane_t>(kLumpPlanes);
  o->numplanes = b.numplanes();
  o->nodes = b.as<qrad_dnode_t>(kLumpNodes);
  o->numnodes = b.numnodes();
  o->leafs = b.as<qrad_dleaf_t>(kLumpLeafs);
  o->numleafs = b.numleafs();
  o->visdata = b.lump[kLumpVisibility].data();
  o->visdatasize = b.visdatasize();
  return 0;
}

int qrad_host_patches_view(qrad_host_scene *hs, qrad_patches_view *o) {
  Scene &s = hs->scene;
  if (!s.prepared) { g_host_error = "scene not prepared"; return 1; }
  o->num_patches = s.num_patches();
  o->origin = s.origin.data()->v;
  o->normal = s.normal.data()->v;
  o->area = s.area.data();
  o->cluster = s.cluster.data();
  o->sky = s.sky.data();
  o->reflectivity = s.reflectivity.data()->v;
  o->baselight = s.baselight.data()->v;
  o->totallight = s.totallight.data()->v;
  o->bounds = s.bounds.data();
  o->face = s.face.data();
  o->next = s.next.data();
  o->numfaces = s.bsp.numfaces();
  o->face_head = s.face_head.data();
  return 0;
}

int qrad_host_lights_view(qrad_host_scene *hs, qrad_lights_view *o) {
  Scene &s = hs->scene;
  if (!s.prepared) { g_host_error = "scene not prepared"; return 1; }
  o->numclusters = s.bsp.numclusters();
  o->light_first = s.light_first.data();
  o->numlights = (int32_t) s.light_type.size();
  o->type = s.light_type.data();
  o->intensity = s.light_intensity.data();
  o->style = s.light_style.data();
  o->origin = s.light_origin.data()->v;
  o->color = s.light_color.data()->v;
  o->normal = s.light_normal.data()->v;
  o->stopdot = s.light_stopdot.data();
  return 0;
}

int qrad_host_faces_view(qrad_host_scene *hs, qrad_faces_view *o) {
  Scene &s = hs->scene;
  if (!s.prepared) { g_host_error = "scene not prepared"; return 1; }
  o->numfaces = s.bsp.numfaces();
  o->lit = s.face_lit.data();
  o->facenormal = s.facenormal.data()->v;
  o->texorg = s.texorg.data()->v;
  o->textoworld = s.textoworld.data();
  o->exactmins = s.exactmins.data();
  o->exactmaxs = s.exactmaxs.data();
  o->texmins = s.texmins.data();
  o->texsize = s.texsize.data();
  o->face_bounds = s.face_bounds.data();
  o->planelink_head = s.planelink_head.data();
  o->facelinks = s.facelinks.data();
  o->minlight = s.minlight.data();
  o->plane_normal = s.plane_normal.data()->v;
  return 0;
}

int qrad_host_counts(qrad_host_scene *hs, int32_t *o) {
  Scene &s = hs->scene;
  o[0] = s.bsp.numfaces();
  o[1] = s.num_patches();
  o[2] = s.numdlights;
  o[3] = s.bsp.numclusters();
  o[4] = s.bsp.numnodes();
  o[5] = s.bsp.numleafs();
  o[6] = s.bsp.numtexinfo();
  o[7] = (int32_t) s.bsp.entities.size();
  return 0;
}

int qrad_host_reflectivity(qrad_host_scene *hs, float *out) {
  Scene &s = hs->scene;
  for (int i = 0; i < s.bsp.numtexinfo(); i++)
    for (int k = 0; k < 3; k++) out[i * 3 + k] = s.texture_reflectivity[i][k];
  return 0;
}

int qrad_host_windings(qrad_host_scene *hs, int32_t *numpoints, float *pts) {
  Scene &s = hs->scene;
  size_t o = 0;
  for (int i = 0; i < s.num_patches(); i++) {
    numpoints[i] = s.wcount[i];
    if (pts)
      for (int k = 0; k < s.wcount[i]; k++) {
        const Vec3 &p = s.wpts[(size_t) s.wfirst[i] + k];
        pts[o++] = p[0];
        pts[o++] = p[1];
        pts[o++] = p[2];
      }
  }
  return 0;
}

int qrad_host_store_lighting(qrad_host_scene *hs, const uint8_t *data, int32_t size, const int32_t *lightofs,
                             const uint8_t *styles) {
  return guarded([&] {
    Scene &s = hs->scene;
    if (size < 0 || size > kMaxMapLighting) fail("MAX_MAP_LIGHTING");
    memcpy(s.bsp.lighting.data(), data, (size_t) size);
    s.bsp.lightdatasize = size;
    qrad_dface_t *faces = s.bsp.as_mut<qrad_dface_t>(kLumpFaces);
    for (int f = 0; f < s.bsp.numfaces(); f++) {
      if (!s.face_lit[f]) continue;  // skipped faces keep their input lightofs/styles (lightmap.c:1083)
      faces[f].lightofs = lightofs[f];
      memcpy(faces[f].styles, styles + f * 4, 4);
    }
  });
}

int qrad_host_lighting(qrad_host_scene *hs, const uint8_t **data, int32_t *size) {
  *data = hs->scene.bsp.lighting.data();
  *size = hs->scene.bsp.lightdatasize;
  return 0;
}

int qrad_host_write(qrad_host_scene *hs, const char *path) {
  return guarded([&] { write_bsp(hs->scene.bsp, path); });
}

// RadWorld, qrad3.c:494-536, with the five dispatch sites replaced by libqrad_cuda calls.  `store`: keep the result in
// the scene (on several GPUs every rank ends with the same lump; one of them stores it).
int qrad_rad_world(qrad_host_scene *hs, qrad_cuda_ctx *ctx, const qrad_host_params *p, float *added_out) {
  return guarded([&] { rad_world_impl(hs, ctx, p, added_out, true); });
}

}  // extern "C"

namespace qrad {
void rad_world_impl(qrad_host_scene *hs, qrad_cuda_ctx *ctx, const qrad_host_params *p, float *added_out, bool store) {
  Scene &s = hs->scene;
  auto ck = [&](int rc) {
    if (rc) fail("%s", qrad_cuda_last_error());
  };
  qrad_bsp_view bv;
  qrad_patches_view pv;
  qrad_lights_view lv;
  qrad_faces_view fv;
  qrad_host_bsp_view(hs, &bv);
  if (qrad_host_patches_view(hs, &pv) || qrad_host_lights_view(hs, &lv) || qrad_host_faces_view(hs, &fv))
    fail("%s", qrad_host_last_error());
  ck(qrad_cuda_upload_bsp(ctx, &bv));
  ck(qrad_cuda_upload_patches(ctx, &pv));
  ck(qrad_cuda_upload_lights(ctx, &lv));
  ck(qrad_cuda_upload_faces(ctx, &fv));
  ck(qrad_cuda_build_facelights(ctx, p->extrasamples, p->numbounce));
  if (p->numbounce > 0) {
    int64_t total = 0;
    ck(qrad_cuda_make_transfers(ctx, p->nopvs, &total));
    if (p->verbose && store) printf("transfer lists: %5.1f megs\n", (float) total * 4 / (1024 * 1024));
    std::vector<float> added(p->numbounce);
    ck(qrad_cuda_bounce(ctx, p->numbounce, (p->verbose || added_out) ? added.data() : nullptr));
    for (int i = 0; i < p->numbounce; i++) {
      if (p->verbose && store) printf("bounce:%i added:%f\n", i, added[i]);
      if (added_out) added_out[i] = added[i];
    }
  }
  qrad_final_params fp;
  fp.ambient = p->ambient;
  fp.maxlight = p->maxlight;
  fp.lightscale = p->lightscale;
  fp.subdiv = p->subdiv;
  fp.numbounce = p->numbounce;
  std::vector<uint8_t> light(kMaxMapLighting);
  std::vector<int32_t> lightofs(s.bsp.numfaces());
  std::vector<uint8_t> styles((size_t) s.bsp.numfaces() * 4);
  int32_t size = 0;
  ck(qrad_cuda_final_light(ctx, &fp, light.data(), kMaxMapLighting, &size, lightofs.data(), styles.data()));
  if (store && qrad_host_store_lighting(hs, light.data(), size, lightofs.data(), styles.data()))
    fail("%s", qrad_host_last_error());
}
}  // namespace qrad

extern "C" {

// The flag loop of main(), qrad3.c:555-601: same flags, same side effects on the tunables; silent (qrad3_main prints
// the reference's chatter).  Parsing stops at the first argument that is not a known flag (cli->next_arg), like the reference's `else break`.
int qrad_host_parse_flags(int argc, char **argv, qrad_host_params *p, qrad_host_cli *cli) {
  memset(cli, 0, sizeof *cli);
  cli->ngpus = 1;
  int i;
  for (i = 1; i < argc; i++) {
    const char *a = argv[i];
    auto arg = [&]() -> const char * { return (i + 1 < argc) ? argv[++i] : "0"; };
    if (!strcmp(a, "-dump"))
      cli->dump = 1;
    else if (!strcmp(a, "-bounce"))
      p->numbounce = atoi(arg());
    else if (!strcmp(a, "-v"))
      p->verbose = 1;
    else if (!strcmp(a, "-extra"))
      p->extrasamples = 1;
    else if (!strcmp(a, "-threads"))
      cli->threads = p->threads = atoi(arg());  // the reference's Linux build then forces 1 (threads.c:379-420); here: host set-up threads
    else if (!strcmp(a, "-chop"))
      p->subdiv = (float) atoi(arg());
    else if (!strcmp(a, "-scale"))
      p->lightscale = (float) atof(arg());
    else if (!strcmp(a, "-direct")) {
      p->direct_scale *= atof(arg());
      cli->direct_given = 1;
    } else if (!strcmp(a, "-entity")) {
      p->entity_scale *= atof(arg());
      cli->entity_given = 1;
    } else if (!strcmp(a, "-glview"))
      cli->glview = 1;
    else if (!strcmp(a, "-nopvs"))
      p->nopvs = 1;
    else if (!strcmp(a, "-ambient"))
      p->ambient = (float) (atof(arg()) * 128);
    else if (!strcmp(a, "-maxlight"))
      p->maxlight = (float) (atof(arg()) * 128);
    else if (!strcmp(a, "-tmpin"))
      cli->tmpin = 1;
    else if (!strcmp(a, "-tmpout"))
      cli->tmpout = 1;
    else if (!strcmp(a, "-gpu"))   // extension: CUDA device ordinal of the (first) GPU
      cli->device = atoi(arg());
    else if (!strcmp(a, "-gpus"))  // extension: number of GPUs of this box to spread the map over
      cli->ngpus = atoi(arg());
    else
      break;
  }
  if (p->maxlight > 255) p->maxlight = 255;
  cli->next_arg = i;
  return 0;
}

// main(), qrad3.c:545-643.  Same flags, same messages where they are part of the contract.
int qrad3_main(int argc, char **argv) {
  printf("----- Radiosity ----\n");
  qrad_host_params p;
  qrad_host_default_params(&p);
  qrad_host_cli cli;
  qrad_host_parse_flags(argc, argv, &p, &cli);
  if (p.extrasamples) printf("extrasamples = true\n");
  if (cli.direct_given) printf("direct light scaling at %f\n", p.direct_scale);
  if (cli.entity_given) printf("entity light scaling at %f\n", p.entity_scale);
  if (cli.glview) printf("glview = true\n");
  if (p.nopvs) printf("nopvs = true\n");
  const int i = cli.next_arg;
  auto die = [](const std::string &msg) {  // Error(), cmdlib.c:142-155
    printf("\n************ ERROR ************\n%s\n", msg.c_str());
    return 1;
  };
  if (i != argc - 1)
    return die("usage: qrad [-v] [-chop num] [-scale num] [-ambient num] [-maxlight num] [-threads num] bspfile");
  auto t0 = std::chrono::steady_clock::now();
  // SetQdirFromPath(argv[i]) works on the argument as given; source = ExpandArg + StripExtension + DefaultExtension
  std::string source = argv[i];
  if (source[0] != '/') {   // ExpandArg, cmdlib.c:219-229
    char cwd[1024];
    if (!getcwd(cwd, sizeof cwd - 1)) return die("Couldn't determine current directory");
    source = std::string(cwd) + "/" + source;
  }
  {
    size_t slash = source.find_last_of('/');
    size_t dot = source.find_last_of('.');
    if (dot != std::string::npos && (slash == std::string::npos || dot > slash)) source.resize(dot);
    source += ".bsp";
  }
  // -tmpin / -tmpout prefix the expanded path with "/tmp" (inbase / outbase, qrad3.c:595-598, 619, 635)
  const std::string inname = (cli.tmpin ? "/tmp" : "") + source, outname = (cli.tmpout ? "/tmp" : "") + source;
  printf("reading %s\n", inname.c_str());
  qrad_host_scene *hs = nullptr;
  if (qrad_host_load_as(inname.c_str(), argv[i], &hs)) return die(g_host_error);
  if (qrad_host_prepare(hs, &p)) {
    qrad_host_free(hs);
    return die(g_host_error);
  }
  if (cli.dump || cli.glview)
    printf("-dump / -glview: the debugging dumps (bounceN.txt, .glr) are not produced by this build\n");
  int rc;
  qrad_cuda_stats st{};
  if (cli.ngpus > 1) {
    rc = qrad_rad_world_multi(hs, cli.device, cli.ngpus, &p, nullptr, &st);
  } else {
    qrad_cuda_ctx *ctx = nullptr;
    if (qrad_cuda_create(cli.device, &ctx)) {
      qrad_host_free(hs);
      return die(qrad_cuda_last_error());
    }
    rc = qrad_rad_world(hs, ctx, &p, nullptr);
    if (rc == 0) qrad_cuda_get_stats(ctx, &st);
    qrad_cuda_destroy(ctx);
  }
  if (rc == 0) {
    printf("writing %s\n", outname.c_str());
    rc = qrad_host_write(hs, outname.c_str());
  }
  if (p.verbose && rc == 0)
    printf("gpu: facelights %.2f ms, transfers %.2f ms, bounce %.2f ms, final %.2f ms; %lld rays\n", st.ms_facelights,
           st.ms_transfers, st.ms_bounce, st.ms_final, (long long) (st.rays_transfers + st.rays_direct));
  qrad_host_free(hs);
  if (rc) return die(g_host_error);
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  printf("%5.0f seconds elapsed\n", secs);
  return 0;
}

}  // extern "C"
