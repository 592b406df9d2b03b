/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Drives the UNMODIFIED reference qrad3 translation units (compiled in place
 * from /root/reference/src/tools/{common,qrad3}, see oracle/Makefile) phase by
 * phase and dumps the intermediate state the reference keeps in its process
 * globals, so that every seam of the hot path can be
 * compared against libqrad_cuda:
 *
 *   phase order follows RadWorld()            qrad3.c:494-536
 *   flag parsing follows main()               qrad3.c:555-610
 *   the bounce loop follows BounceLight()     qrad3.c:448-473 (re-stated here only
 *   so that `added` and totallight can be captured after each bounce; the work
 *   itself is done by the reference's ShootLight / CollectLight).
 *
 * usage: ref_harness [qrad3 flags] [-dumpdir DIR] [-stopafter PHASE] [sampling mode] bspfile
 * Dump files are flat little-endian binaries documented in tests/refdump.py.
 *
 * Sampling modes (for maps on which the whole O(N^2) MakeTransfers is hours of CPU; they need the
 * limit-lifted build ref_harness_big for more than 65 000 patches, see oracle/Makefile):
 *   -rows F:S -rowsout FILE        MakeTransfers(i) (qrad3.c:185-293) for i = F, F+S, ...; every row is
 *                                  written out with its root TestLine_r count
 *   -cols A:N [-part k:K] -colsout FILE
 *                                  every source that can see one of the receivers A .. A+N-1: its
 *                                  radiosity (BuildFacelights of its face, qrad3.c:455-459) and its
 *                                  weights towards those receivers -- what ShootLight (qrad3.c:419-441)
 *                                  adds to them during the first bounce; -part splits the sources
 *                                  over K processes
 *   -benchsteps K -rows F:S -facesample F2:S2
 *                                  K timed steps for bench.py's reference arm: step s lights the faces
 *                                  f = (F2+s) mod S2 (BuildFacelights) and the rows i = (F+s) mod S
 *                                  (MakeTransfers) and prints one "REFSTEP {json}" line; S = 0 picks
 *                                  S = S2 = ceil(num_patches / 128), i.e. about 128 rows per step
 * Root TestLine_r calls are counted without touching the reference: the link wraps the symbol
 * (-Wl,--wrap=TestLine_r); the recursion inside trace.c binds locally and is not counted.
 * Phase timings (clock_gettime, finding F10: the reference timer has 1 s
 * resolution) are printed as one JSON line prefixed "REFTIMING ".
 */
#include <malloc.h>
#include <time.h>
#include "qrad.h"

/* globals defined in the reference TUs but not declared in qrad.h */
extern int numbounce;
extern qboolean extrasamples, dumppatches, glview, nopvs;
extern float subdiv, ambient, maxlight, lightscale, direct_scale, entity_scale;
extern char source[1024];
extern char inbase[32], outbase[32];
extern vec3_t radiosity[MAX_PATCHES];
extern vec3_t illumination[MAX_PATCHES];
extern int total_transfer;
extern int numdlights;
extern vec3_t texture_reflectivity[MAX_MAP_TEXINFO];

void MakeBackplanes(void);
void MakeParents(int nodenum, int parent);
void MakeTransfers(int i);
void ShootLight(int patchnum);
float CollectLight(void);
void FreeTransfers(void);
void CheckPatches(void);
void WriteGlView(void);
int PointInLeafnum(vec3_t point);

/* layout of lightmap.c's file-local facelight_t (lightmap.c:681-688) */
#define H_MAX_STYLES 32
typedef struct
{
  int numsamples;
  float *origins;
  int numstyles;
  int stylenums[H_MAX_STYLES];
  float *samples[H_MAX_STYLES];
} h_facelight_t;
extern h_facelight_t facelight[MAX_MAP_FACES];

/* root TestLine_r calls (qrad3.c:244, lightmap.c:647,898), counted through ld --wrap */
int __real_TestLine_r(int node, vec3_t start, vec3_t stop);
static long long harness_rays;
int __wrap_TestLine_r(int node, vec3_t start, vec3_t stop)
{
  harness_rays++;
  return __real_TestLine_r(node, start, stop);
}

static char dumpdir[1024];
static double now(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + ts.tv_nsec * 1e-9;
}
static FILE *dopen(const char *name)
{
  char p[2048];
  FILE *f;
  sprintf(p, "%s/%s", dumpdir, name);
  f = fopen(p, "wb");
  if (!f)
    Error("harness: cannot open %s", p);
  return f;
}
static void wi(FILE *f, int v) { fwrite(&v, 4, 1, f); }
static void wf(FILE *f, const float *v, int n) { fwrite(v, 4, n, f); }

static int patch_face(patch_t *p)
{
  /* which face list is this patch on?  resolved through face_patches chains */
  int fn;
  patch_t *q;
  for (fn = 0; fn < numfaces; fn++)
    for (q = face_patches[fn]; q; q = q->next)
      if (q == p)
        return fn;
  return -1;
}

/* patches.bin: i32 N, then N records of 36 x 4 bytes (see tests/refdump.py) */
static void dump_patches(const char *name)
{
  FILE *f = dopen(name);
  unsigned i;
  int *faceof = malloc(sizeof(int) * (num_patches + 1));
  int fn;
  patch_t *q;
  for (i = 0; i < num_patches; i++)
    faceof[i] = -1;
  for (fn = 0; fn < numfaces; fn++)
    for (q = face_patches[fn]; q; q = q->next)
      faceof[q - patches] = fn;
  wi(f, num_patches);
  for (i = 0; i < num_patches; i++) {
    patch_t *p = &patches[i];
    vec3_t mins, maxs;
    WindingBounds(p->winding, mins, maxs);
    wf(f, p->origin, 3);
    wf(f, p->plane->normal, 3);
    wf(f, &p->plane->dist, 1);
    wf(f, &p->area, 1);
    wi(f, p->cluster);
    wi(f, p->sky);
    wf(f, p->reflectivity, 3);
    wf(f, p->baselight, 3);
    wf(f, p->totallight, 3);
    wf(f, p->samplelight, 3);
    wi(f, p->samples);
    wi(f, faceof[i]);
    wi(f, p->next ? (int) (p->next - patches) : -1);
    wf(f, mins, 3);
    wf(f, maxs, 3);
    wi(f, p->numtransfers);
    wi(f, p->winding->numpoints);
  }
  /* face heads */
  wi(f, numfaces);
  for (fn = 0; fn < numfaces; fn++)
    wi(f, face_patches[fn] ? (int) (face_patches[fn] - patches) : -1);
  fclose(f);
  free(faceof);
}

/* windings.bin: per patch i32 n, n*3 f32 */
static void dump_windings(void)
{
  FILE *f = dopen("windings.bin");
  unsigned i;
  wi(f, num_patches);
  for (i = 0; i < num_patches; i++) {
    wi(f, patches[i].winding->numpoints);
    wf(f, patches[i].winding->p[0], 3 * patches[i].winding->numpoints);
  }
  fclose(f);
}

/* lights.bin: i32 count, then per light (in per-cluster list order, clusters ascending):
   i32 cluster, i32 type, f32 intensity, i32 style, origin[3], color[3], normal[3], f32 stopdot */
static void dump_lights(void)
{
  FILE *f = dopen("lights.bin");
  int c, n = 0;
  directlight_t *l;
  for (c = 0; c < dvis->numclusters; c++)
    for (l = directlights[c]; l; l = l->next)
      n++;
  wi(f, n);
  for (c = 0; c < dvis->numclusters; c++)
    for (l = directlights[c]; l; l = l->next) {
      wi(f, c);
      wi(f, l->type);
      wf(f, &l->intensity, 1);
      wi(f, l->style);
      wf(f, l->origin, 3);
      wf(f, l->color, 3);
      wf(f, l->normal, 3);
      wf(f, &l->stopdot, 1);
    }
  fclose(f);
}

/* facelight.bin: i32 numfaces; per face i32 numsamples, i32 numstyles, stylenums[numstyles],
   origins[numsamples*3], then numstyles * samples[numsamples*3] */
static void dump_facelight(void)
{
  FILE *f = dopen("facelight.bin");
  int fn, s;
  wi(f, numfaces);
  for (fn = 0; fn < numfaces; fn++) {
    h_facelight_t *fl = &facelight[fn];
    wi(f, fl->numsamples);
    wi(f, fl->numstyles);
    for (s = 0; s < fl->numstyles; s++)
      wi(f, fl->stylenums[s]);
    if (fl->numsamples)
      wf(f, fl->origins, fl->numsamples * 3);
    for (s = 0; s < fl->numstyles; s++)
      wf(f, fl->samples[s], fl->numsamples * 3);
  }
  fclose(f);
}

/* transfers.bin: i32 N, i32 numtransfers[N], then all records {u16 patch, u16 transfer} row-major */
static void dump_transfers(void)
{
  FILE *f = dopen("transfers.bin");
  unsigned i;
  wi(f, num_patches);
  for (i = 0; i < num_patches; i++)
    wi(f, patches[i].numtransfers);
  for (i = 0; i < num_patches; i++)
    if (patches[i].numtransfers)
      fwrite(patches[i].transfers, sizeof(transfer_t), patches[i].numtransfers, f);
  fclose(f);
}

static void dump_vec3_array(const char *name, int which)
{
  FILE *f = dopen(name);
  unsigned i;
  wi(f, num_patches);
  for (i = 0; i < num_patches; i++)
    wf(f, which == 0 ? patches[i].totallight : radiosity[i], 3);
  fclose(f);
}

/* rays.bin: sample of TestLine_r results: i32 n, then n * {start[3], stop[3], i32 result} */
static void dump_rays(int want)
{
  FILE *f = dopen("rays.bin");
  FILE *fl = dopen("rayleafs.bin"); /* i32 per ray of rays.bin */
  unsigned a, b, n = 0, step;
  unsigned long long pairs = (unsigned long long) num_patches * num_patches;
  step = (unsigned) (pairs / (want ? want : 1));
  if (step < 1)
    step = 1;
  step |= 1;
  wi(f, 0);
  {
    unsigned long long k;
    for (k = 0; k < pairs; k += step) {
      int r;
      a = (unsigned) (k / num_patches);
      b = (unsigned) (k % num_patches);
      r = TestLine_r(0, patches[a].origin, patches[b].origin);
      wf(f, patches[a].origin, 3);
      wf(f, patches[b].origin, 3);
      wi(f, r);
      /* PointInLeafnum (qrad3.c:131-150) of a point next to the ray's end: pushed onto the 16-unit grid in x, where
         many node planes of the maps lie, so that the `dist > 0` tie is exercised */
      {
        vec3_t q;
        VectorCopy(patches[b].origin, q);
        if (n & 1)
          q[0] = (float) (16 * (int) (q[0] / 16));
        wi(fl, PointInLeafnum(q));
      }
      n++;
    }
  }
  fseek(f, 0, SEEK_SET);
  wi(f, n);
  fclose(f);
}


/* ---- sampling modes -------------------------------------------------------------------------- */
static int *face_of_patch(void)
{
  int *faceof = malloc(sizeof(int) * (num_patches + 1));
  unsigned i;
  int fn;
  patch_t *q;
  for (i = 0; i < num_patches; i++)
    faceof[i] = -1;
  for (fn = 0; fn < numfaces; fn++)
    for (q = face_patches[fn]; q; q = q->next)
      faceof[q - patches] = fn;
  return faceof;
}

/* rows file: i32 count; per row: i32 patch, i32 numtransfers, i64 rays, u32 patch[numtransfers], u16 transfer[numtransfers] */
static void run_rows(int first, int stride, const char *out)
{
  FILE *f = fopen(out, "wb");
  unsigned i;
  int n = 0, k;
  double t0 = now();
  long long rays_total = 0;
  if (!f)
    Error("harness: cannot open %s", out);
  wi(f, 0);
  for (i = first; i < num_patches; i += stride) {
    patch_t *p = &patches[i];
    long long r0 = harness_rays, r;
    MakeTransfers(i);
    r = harness_rays - r0;
    rays_total += r;
    wi(f, (int) i);
    wi(f, p->numtransfers);
    fwrite(&r, 8, 1, f);
    for (k = 0; k < p->numtransfers; k++) {
      unsigned pj = p->transfers[k].patch;
      fwrite(&pj, 4, 1, f);
    }
    for (k = 0; k < p->numtransfers; k++) {
      unsigned short w = p->transfers[k].transfer;
      fwrite(&w, 2, 1, f);
    }
    if (p->numtransfers)
      free(p->transfers);
    p->transfers = NULL;
    p->numtransfers = 0;
    n++;
  }
  fseek(f, 0, SEEK_SET);
  wi(f, n);
  fclose(f);
  printf("REFTIMING {\"rows\": %d, \"rows_s\": %.6f, \"rays_transfers\": %lld, \"num_patches\": %u}\n", n, now() - t0,
         rays_total, num_patches);
}

/* cols file: i32 A, i32 N; N x {f32 area, i32 sky, f32 totallight[3], f32 reflectivity[3]};
   then until EOF: i32 source, f32 radiosity[3], i32 n, n x {i32 receiver, i32 transfer} */
static void run_cols(int a, int n, int part, int parts, const char *out)
{
  FILE *f = fopen(out, "wb");
  byte pvs[(MAX_MAP_LEAFS + 7) / 8];
  byte *facedone = calloc(numfaces + 1, 1);
  int *faceof = face_of_patch();
  unsigned i;
  int j, k, seen = 0, mine = 0;
  double t0 = now();
  if (!f)
    Error("harness: cannot open %s", out);
  if (a < 0 || n < 1 || a + n > (int) num_patches)
    Error("harness: bad -cols window");
  wi(f, a);
  wi(f, n);
  for (j = a; j < a + n; j++) {
    wf(f, &patches[j].area, 1);
    wi(f, patches[j].sky);
    wf(f, patches[j].totallight, 3);
    wf(f, patches[j].reflectivity, 3);
  }
  for (i = 0; i < num_patches; i++) {
    patch_t *p = &patches[i];
    int want = 0, cnt = 0;
    vec3_t rad;
    if (!PvsForOrigin(p->origin, pvs))
      continue; /* MakeTransfers returns at once (qrad3.c:208) */
    for (j = a; j < a + n && !want; j++) {
      int c = patches[j].cluster;
      if (nopvs || (c != -1 && (pvs[c >> 3] & (1 << (c & 7)))))
        want = 1;
    }
    if (!want)
      continue;
    if (seen++ % parts != part)
      continue;
    mine++;
    if (faceof[i] >= 0 && !facedone[faceof[i]]) {
      facedone[faceof[i]] = 1;
      BuildFacelights(faceof[i]);
    }
    for (k = 0; k < 3; k++)
      rad[k] = p->samplelight[k] * p->reflectivity[k] * p->area; /* qrad3.c:455-459 */
    MakeTransfers(i);
    for (k = 0; k < p->numtransfers; k++)
      if ((int) p->transfers[k].patch >= a && (int) p->transfers[k].patch < a + n)
        cnt++;
    wi(f, (int) i);
    wf(f, rad, 3);
    wi(f, cnt);
    for (k = 0; k < p->numtransfers; k++)
      if ((int) p->transfers[k].patch >= a && (int) p->transfers[k].patch < a + n) {
        wi(f, (int) p->transfers[k].patch);
        wi(f, (int) p->transfers[k].transfer);
      }
    if (p->numtransfers)
      free(p->transfers);
    p->transfers = NULL;
    p->numtransfers = 0;
  }
  fclose(f);
  printf("REFTIMING {\"sources\": %d, \"sources_this_part\": %d, \"cols_s\": %.6f, \"num_patches\": %u}\n", seen, mine,
         now() - t0, num_patches);
}

static void run_benchsteps(int steps, int rfirst, int rstride, int ffirst, int fstride, double setup_s)
{
  int s;
  if (rstride <= 0) /* auto: about 128 rows per step, the same fraction of the faces */
    rstride = fstride = (num_patches + 127) / 128;
  for (s = 0; s < steps; s++) {
    double t0 = now(), t1, t2;
    long long r0 = harness_rays, r1, r2;
    int fn, nf = 0, nr = 0;
    unsigned i;
    long long nnz = 0;
    for (fn = (ffirst + s) % fstride; fn < numfaces; fn += fstride, nf++)
      BuildFacelights(fn);
    t1 = now();
    r1 = harness_rays;
    for (i = (rfirst + s) % rstride; i < num_patches; i += rstride, nr++) {
      patch_t *p = &patches[i];
      MakeTransfers(i);
      nnz += p->numtransfers;
      if (p->numtransfers)
        free(p->transfers);
      p->transfers = NULL;
      p->numtransfers = 0;
    }
    t2 = now();
    r2 = harness_rays;
    printf("REFSTEP {\"step\": %d, \"faces\": %d, \"numfaces\": %d, \"rows\": %d, \"num_patches\": %u, \"facelights_s\": %.6f, "
           "\"transfers_s\": %.6f, \"rays_direct\": %lld, \"rays_transfers\": %lld, \"transfers\": %lld, \"setup_s\": %.6f, "
           "\"stride\": %d}\n",
           s, nf, numfaces, nr, num_patches, t1 - t0, t2 - t1, r1 - r0, r2 - r1, nnz, setup_s, rstride);
  }
}

int main(int argc, char **argv)
{
  int i;
  char name[1024];
  double t0, t[8];
  int dump = 0;
  int stopafter = 99;
  float added[64];
  int nadded = 0;
  int rows_first = -1, rows_stride = 1, cols_a = -1, cols_n = 0, part = 0, parts = 1, benchsteps = 0;
  int fs_first = 0, fs_stride = 1;
  const char *rowsout = NULL, *colsout = NULL;
  long long rays_direct = 0, rays_transfers = 0;

  /* finding F1: keep every allocation on the brk heap (< 2 GB with -no-pie) */
  mallopt(M_MMAP_MAX, 0);
  setvbuf(stdout, NULL, _IOLBF, 0);

  verbose = false;
  for (i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "-dump"))
      dumppatches = true;
    else if (!strcmp(argv[i], "-bounce"))
      numbounce = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-v"))
      verbose = true;
    else if (!strcmp(argv[i], "-extra"))
      extrasamples = true;
    else if (!strcmp(argv[i], "-threads"))
      numthreads = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-chop"))
      subdiv = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-scale"))
      lightscale = atof(argv[++i]);
    else if (!strcmp(argv[i], "-direct"))
      direct_scale *= atof(argv[++i]);
    else if (!strcmp(argv[i], "-entity"))
      entity_scale *= atof(argv[++i]);
    else if (!strcmp(argv[i], "-glview"))
      glview = true;
    else if (!strcmp(argv[i], "-nopvs"))
      nopvs = true;
    else if (!strcmp(argv[i], "-ambient"))
      ambient = atof(argv[++i]) * 128;
    else if (!strcmp(argv[i], "-maxlight"))
      maxlight = atof(argv[++i]) * 128;
    else if (!strcmp(argv[i], "-dumpdir")) {
      strcpy(dumpdir, argv[++i]);
      dump = 1;
    } else if (!strcmp(argv[i], "-stopafter"))
      stopafter = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-rows"))
      sscanf(argv[++i], "%d:%d", &rows_first, &rows_stride);
    else if (!strcmp(argv[i], "-rowsout"))
      rowsout = argv[++i];
    else if (!strcmp(argv[i], "-cols"))
      sscanf(argv[++i], "%d:%d", &cols_a, &cols_n);
    else if (!strcmp(argv[i], "-part"))
      sscanf(argv[++i], "%d:%d", &part, &parts);
    else if (!strcmp(argv[i], "-colsout"))
      colsout = argv[++i];
    else if (!strcmp(argv[i], "-benchsteps"))
      benchsteps = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-facesample"))
      sscanf(argv[++i], "%d:%d", &fs_first, &fs_stride);
    else
      break;
  }
  ThreadSetDefault();
  if (maxlight > 255)
    maxlight = 255;
  if (i != argc - 1)
    Error("usage: ref_harness [qrad3 flags] [-dumpdir DIR] [-stopafter N] bspfile");

  SetQdirFromPath(argv[i]);
  strcpy(source, ExpandArg(argv[i]));
  StripExtension(source);
  DefaultExtension(source, ".bsp");
  sprintf(name, "%s%s", inbase, source);
  LoadBSPFile(name);
  ParseEntities();
  CalcTextureReflectivity();
  if (!visdatasize) {
    numbounce = 0;
    ambient = 0.1;
  }
  if (dump) {
    FILE *f = dopen("reflectivity.bin");
    wi(f, numtexinfo);
    wf(f, texture_reflectivity[0], numtexinfo * 3);
    fclose(f);
  }

  /* ---- RadWorld, phase by phase (qrad3.c:494-536) ---- */
  if (numnodes == 0 || numfaces == 0)
    Error("Empty map");
  t0 = now();
  MakeBackplanes();
  MakeParents(0, -1);
  MakeTnodes(&dmodels[0]);
  MakePatches();
  SubdividePatches();
  if (dump) {
    dump_patches("patches_pre.bin"); /* before CreateDirectLights clears emissive totallight */
    dump_windings();
  }
  CreateDirectLights();
  t[0] = now() - t0;
  if (dump) {
    dump_lights();
    dump_rays(20000);
  }
  if (stopafter < 1)
    goto done;
  if ((rows_stride < 1 && !benchsteps) || fs_stride < 1 || parts < 1)
    Error("harness: bad sampling stride");
  if (benchsteps > 0) {
    run_benchsteps(benchsteps, rows_first < 0 ? 0 : rows_first, rows_stride, fs_first, fs_stride, t[0]);
    return 0;
  }
  if (colsout) {
    run_cols(cols_a, cols_n, part, parts, colsout);
    return 0;
  }
  if (rowsout) {
    run_rows(rows_first < 0 ? 0 : rows_first, rows_stride, rowsout);
    return 0;
  }

  t0 = now();
  harness_rays = 0;
  RunThreadsOnIndividual(numfaces, false, BuildFacelights);
  rays_direct = harness_rays;
  t[1] = now() - t0;
  if (dump) {
    dump_facelight();
    dump_patches("patches.bin");
  }
  if (stopafter < 2)
    goto done;

  t[2] = t[3] = 0;
  if (numbounce > 0) {
    unsigned k;
    int j, b;
    t0 = now();
    harness_rays = 0;
    RunThreadsOnIndividual(num_patches, false, MakeTransfers);
    rays_transfers = harness_rays;
    t[2] = now() - t0;
    if (dump)
      dump_transfers();
    if (stopafter < 3)
      goto done;
    t0 = now();
    for (k = 0; k < num_patches; k++)
      for (j = 0; j < 3; j++)
        radiosity[k][j] = patches[k].samplelight[j] * patches[k].reflectivity[j] * patches[k].area;
    if (dump)
      dump_vec3_array("radiosity0.bin", 1);
    for (b = 0; b < numbounce; b++) {
      RunThreadsOnIndividual(num_patches, false, ShootLight);
      added[nadded < 64 ? nadded++ : 63] = CollectLight();
    }
    t[3] = now() - t0;
    if (dump)
      dump_vec3_array("totallight.bin", 0);
    FreeTransfers();
    CheckPatches();
  }
  if (stopafter < 4)
    goto done;

  t0 = now();
  PairEdges();
  LinkPlaneFaces();
  lightdatasize = 0;
  RunThreadsOnIndividual(numfaces, false, FinalLightFace);
  t[4] = now() - t0;

  sprintf(name, "%s%s", outbase, source);
  WriteBSPFile(name);

  printf("REFTIMING {\"setup_s\": %.6f, \"facelights_s\": %.6f, \"transfers_s\": %.6f, \"bounce_s\": %.6f, "
         "\"final_s\": %.6f, \"num_patches\": %u, \"total_transfer\": %d, \"numdlights\": %d, \"lightdatasize\": %d, "
         "\"rays_direct\": %lld, \"rays_transfers\": %lld, \"added\": [",
         t[0], t[1], t[2], t[3], t[4], num_patches, total_transfer, numdlights, lightdatasize, rays_direct, rays_transfers);
  for (i = 0; i < nadded; i++)
    printf("%s%.9g", i ? ", " : "", added[i]);
  printf("]}\n");
  return 0;
done:
  printf("REFTIMING {\"setup_s\": %.6f, \"num_patches\": %u, \"total_transfer\": %d, \"stopped\": %d}\n", t[0], num_patches,
         total_transfer, stopafter);
  return 0;
}
