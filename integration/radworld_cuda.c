/*
 * radworld_cuda.c -- what a maintainer of the reference adds to src/tools/qrad3/qrad3.c to light maps on the GPU:
 * RadWorld() (qrad3.c:494-536) with its five dispatch sites replaced by calls into libqrad_cuda (include/qrad_cuda.h).
 * Everything else of the tool stays the reference's own code: LoadBSPFile, ParseEntities, CalcTextureReflectivity,
 * MakePatches, SubdividePatches, CreateDirectLights, WriteBSPFile, the flag parser.
 *
 * integration/Makefile proves the binding by building it: a scratch copy of qrad3.c with the body of RadWorld cut out
 * and this file included instead, lightmap.c with faceview_cuda.c appended (lightinfo_t is private to that file), the
 * rest of the reference compiled in place, linked with -lqrad_cuda.  tests/test_gpu_integration.py runs the result.
 */
#include "qrad_cuda.h"

void FillFaceView(qrad_faces_view *v); /* faceview_cuda.c, inside lightmap.c */

static void CudaCheck(int rc)
{
  if (rc)
    Error("%s", qrad_cuda_last_error()); /* cmdlib.c:142 */
}

/* patch_t[] (AoS with pointers, qrad.h:65-91) -> the SoA view the library takes */
static void FillPatchView(qrad_patches_view *v)
{
  unsigned i;
  int f;
  float *origin = malloc(num_patches * 12), *normal = malloc(num_patches * 12), *area = malloc(num_patches * 4);
  float *refl = malloc(num_patches * 12), *base = malloc(num_patches * 12), *total = malloc(num_patches * 12);
  float *bounds = malloc(num_patches * 24);
  int *cluster = malloc(num_patches * 4), *face = malloc(num_patches * 4), *next = malloc(num_patches * 4);
  unsigned char *sky = malloc(num_patches);
  int *head = malloc(numfaces * 4);
  for (f = 0; f < numfaces; f++) {
    patch_t *p;
    head[f] = face_patches[f] ? (int) (face_patches[f] - patches) : -1;
    for (p = face_patches[f]; p; p = p->next)
      face[p - patches] = f;
  }
  for (i = 0; i < num_patches; i++) {
    patch_t *p = &patches[i];
    VectorCopy(p->origin, (origin + 3 * i));
    VectorCopy(p->plane->normal, (normal + 3 * i));
    VectorCopy(p->reflectivity, (refl + 3 * i));
    VectorCopy(p->baselight, (base + 3 * i));
    VectorCopy(p->totallight, (total + 3 * i));
    WindingBounds(p->winding, bounds + 6 * i, bounds + 6 * i + 3);
    area[i] = p->area;
    cluster[i] = p->cluster;
    sky[i] = p->sky;
    next[i] = p->next ? (int) (p->next - patches) : -1;
  }
  v->num_patches = num_patches;
  v->origin = origin;
  v->normal = normal;
  v->area = area;
  v->cluster = cluster;
  v->sky = sky;
  v->reflectivity = refl;
  v->baselight = base;
  v->totallight = total;
  v->bounds = bounds;
  v->face = face;
  v->next = next;
  v->numfaces = numfaces;
  v->face_head = head;
}

/* directlights[cluster] (lightmap.c:690, 751, 785) flattened in list order: the order GatherSampleLight walks them */
static void FillLightView(qrad_lights_view *v)
{
  int c, n = 0, k = 0, nc = dvis->numclusters;
  directlight_t *l;
  int *first = malloc((nc + 1) * 4), *type, *style;
  float *intensity, *origin, *color, *normal, *stopdot;
  for (c = 0; c < nc; c++)
    for (l = directlights[c]; l; l = l->next)
      n++;
  type = malloc((n + 1) * 4);
  style = malloc((n + 1) * 4);
  intensity = malloc((n + 1) * 4);
  stopdot = malloc((n + 1) * 4);
  origin = malloc((n + 1) * 12);
  color = malloc((n + 1) * 12);
  normal = malloc((n + 1) * 12);
  for (c = 0; c < nc; c++) {
    first[c] = k;
    for (l = directlights[c]; l; l = l->next, k++) {
      type[k] = l->type;
      style[k] = l->style;
      intensity[k] = l->intensity;
      stopdot[k] = l->stopdot;
      VectorCopy(l->origin, (origin + 3 * k));
      VectorCopy(l->color, (color + 3 * k));
      VectorCopy(l->normal, (normal + 3 * k));
    }
  }
  first[nc] = k;
  v->numclusters = nc;
  v->light_first = first;
  v->numlights = n;
  v->type = type;
  v->intensity = intensity;
  v->style = style;
  v->origin = origin;
  v->color = color;
  v->normal = normal;
  v->stopdot = stopdot;
}

void RadWorld(void)
{
  qrad_cuda_ctx *gpu;
  qrad_bsp_view bsp;
  qrad_patches_view pv;
  qrad_lights_view lv;
  qrad_faces_view fv;
  qrad_final_params fp;
  long long total = 0;
  static int lightofs[MAX_MAP_FACES];
  static byte styles[MAX_MAP_FACES][4];
  int f;

  if (numnodes == 0 || numfaces == 0)
    Error("Empty map");
  MakeBackplanes();
  MakePatches();
  SubdividePatches();
  CreateDirectLights(); /* unchanged host set-up, qrad3.c:498-509 */
  LinkPlaneFaces();     /* qrad3.c:532: the plane chains FinalLightFace walks */

  bsp.planes = (const qrad_dplane_t *) dplanes;
  bsp.numplanes = numplanes;
  bsp.nodes = (const qrad_dnode_t *) dnodes;
  bsp.numnodes = numnodes;
  bsp.leafs = (const qrad_dleaf_t *) dleafs;
  bsp.numleafs = numleafs;
  bsp.visdata = dvisdata;
  bsp.visdatasize = visdatasize;

  CudaCheck(qrad_cuda_create(0, &gpu));
  CudaCheck(qrad_cuda_upload_bsp(gpu, &bsp)); /* was MakeTnodes(&dmodels[0])                          :500 */
  FillPatchView(&pv);
  CudaCheck(qrad_cuda_upload_patches(gpu, &pv));
  FillLightView(&lv);
  CudaCheck(qrad_cuda_upload_lights(gpu, &lv));
  FillFaceView(&fv);
  CudaCheck(qrad_cuda_upload_faces(gpu, &fv));
  CudaCheck(qrad_cuda_build_facelights(gpu, extrasamples, numbounce)); /* was RunThreadsOnIndividual(.., BuildFacelights) :512 */
  if (numbounce > 0) {
    CudaCheck(qrad_cuda_make_transfers(gpu, nopvs, (int64_t *) &total)); /* was RunThreadsOnIndividual(.., MakeTransfers) :516 */
    qprintf("transfer lists: %5.1f megs\n", (float) total * sizeof(transfer_t) / (1024 * 1024));
    CudaCheck(qrad_cuda_bounce(gpu, numbounce, NULL)); /* was BounceLight() + CheckPatches()            :520-524 */
  }
  fp.ambient = ambient;
  fp.maxlight = maxlight;
  fp.lightscale = lightscale;
  fp.subdiv = subdiv;
  fp.numbounce = numbounce;
  CudaCheck(qrad_cuda_final_light(gpu, &fp, dlightdata, MAX_MAP_LIGHTING, &lightdatasize, lightofs, &styles[0][0]));
  for (f = 0; f < numfaces; f++) /* was RunThreadsOnIndividual(.., FinalLightFace)               :535 */
    if (!(texinfo[dfaces[f].texinfo].flags & (SURF_WARP | SURF_SKY))) {
      dfaces[f].lightofs = lightofs[f];
      memcpy(dfaces[f].styles, styles[f], 4);
    }
  qrad_cuda_destroy(gpu);
}
