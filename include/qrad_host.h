/*
 * qrad_host.h -- C ABI of the host-side half of the qrad3 drop-in (C++ inside
 * libqrad_cuda.so).  It mirrors what the reference does on the CPU *around* the
 * five GPU seams, with results identical to the reference's x86 arithmetic:
 *
 *   qrad_host_load       LoadBSPFile + ParseEntities + CalcTextureReflectivity
 *                        (common/bspfile.c:354,625; qrad3/patches.c:40; path rules of
 *                        SetQdirFromPath, common/cmdlib.c:187-217)
 *   qrad_host_prepare    MakeBackplanes, MakePatches, SubdividePatches, CreateDirectLights,
 *                        per-face CalcFaceVectors/CalcFaceExtents, LinkPlaneFaces
 *                        (qrad3.c:88-96; patches.c:187-492; lightmap.c:42-52,480-603,721-838)
 *   qrad_host_*_view     the arrays libqrad_cuda's upload calls take (include/qrad_cuda.h)
 *   qrad_host_store_lighting / qrad_host_write    dlightdata, dface_t.lightofs/.styles and
 *                        WriteBSPFile's fixed lump order (common/bspfile.c:466-503)
 *   qrad3_main           the whole tool: same argv as the reference main() (qrad3.c:545-643)
 *
 * Errors: non-zero return + qrad_host_last_error(); qrad3_main prints them the way
 * Error() does (common/cmdlib.c:142-155) and returns 1.
 */
#ifndef QRAD_HOST_H
#define QRAD_HOST_H

#include "qrad_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qrad_host_scene qrad_host_scene;

/* the reference's tunables (globals of qrad3.c:49-74), defaults as there */
typedef struct {
  int32_t numbounce;      /* -bounce, default 8   */
  int32_t extrasamples;   /* -extra               */
  float   subdiv;         /* -chop, default 64    */
  float   ambient;        /* -ambient * 128       */
  float   maxlight;       /* -maxlight * 128, default 196, capped 255 */
  float   lightscale;     /* -scale, default 1    */
  float   direct_scale;   /* 0.4 * -direct        */
  float   entity_scale;   /* 1.0 * -entity        */
  int32_t nopvs;          /* -nopvs               */
  int32_t verbose;        /* -v                   */
  int32_t max_patches;    /* 0 = unlimited (the reference stops at 65000, qrad.h:63) */
  int32_t threads;        /* -threads n: host threads of the set-up (patch dicing); 0 = all cores.  The reference parses
                             the flag and then runs single-threaded on Linux (threads.c:379-420); results do not depend on it */
} qrad_host_params;

/* what the flag loop of main() (qrad3.c:555-601) collects besides the tunables */
typedef struct {
  int32_t next_arg;       /* index of the first argument that is not a flag (the reference `break`s there) */
  int32_t threads;        /* -threads n: parsed, unused (the reference's Linux build forces 1, threads.c:379-420) */
  int32_t tmpin, tmpout;  /* -tmpin / -tmpout: "/tmp" prefix of the input / output file name (qrad3.c:595-598) */
  int32_t dump, glview;   /* -dump / -glview: accepted; the debugging dumps are not produced */
  int32_t device;         /* -gpu n   (extension) CUDA ordinal of the first GPU */
  int32_t ngpus;          /* -gpus n  (extension) GPUs of this box to spread the map over, default 1 */
  int32_t direct_given, entity_given;   /* -direct / -entity were seen (the reference echoes the resulting scale) */
} qrad_host_cli;

void qrad_host_default_params(qrad_host_params *p);
/* The reference's flag parser applied to *params (which must hold the defaults); never fails. */
int  qrad_host_parse_flags(int argc, char **argv, qrad_host_params *params, qrad_host_cli *cli);
const char *qrad_host_last_error(void);

int  qrad_host_load(const char *bsp_path, qrad_host_scene **out);
/* the same with the file to read and the path that locates pics/ and textures/ given separately (-tmpin) */
int  qrad_host_load_as(const char *file_path, const char *qdir_path, qrad_host_scene **out);
void qrad_host_free(qrad_host_scene *s);
/* After load: numbounce/ambient are forced like qrad3.c:627-631 when the map has no vis data. */
int  qrad_host_prepare(qrad_host_scene *s, qrad_host_params *params);
/* The same, shared by `parts` processes that light the same map (one per GPU): each loads the map, calls prepare_part
   with its own index -- MakePatches, then SubdividePatches for a contiguous range of the faces (ranges of equal face
   area, the same in every process) -- and gets a blob (owned by the scene, valid until the next call on it).  The host
   program gathers all blobs in every process by whatever means it has (MPI_Allgatherv, torch.distributed, a file) and
   calls prepare_merge with all of them in part order: the patches come out exactly as qrad_host_prepare leaves them,
   then CreateDirectLights and the face frames follow.  Dicing is 90 % of prepare and linear in the threads. */
int  qrad_host_prepare_part(qrad_host_scene *s, qrad_host_params *params, int part, int parts, const uint8_t **blob,
                            int64_t *bytes);
int  qrad_host_prepare_merge(qrad_host_scene *s, int parts, const uint8_t *const *blobs, const int64_t *bytes);

int  qrad_host_bsp_view(qrad_host_scene *s, qrad_bsp_view *out);
int  qrad_host_patches_view(qrad_host_scene *s, qrad_patches_view *out);
int  qrad_host_lights_view(qrad_host_scene *s, qrad_lights_view *out);
int  qrad_host_faces_view(qrad_host_scene *s, qrad_faces_view *out);

/* counts: [0] numfaces [1] num_patches [2] numdlights [3] numclusters [4] numnodes [5] numleafs [6] numtexinfo [7] num_entities */
int  qrad_host_counts(qrad_host_scene *s, int32_t *out8);
int  qrad_host_reflectivity(qrad_host_scene *s, float *out /*[numtexinfo][3]*/);
/* windings of all patches: first call with pts==NULL to get sizes */
int  qrad_host_windings(qrad_host_scene *s, int32_t *numpoints /*[N]*/, float *pts /*[sum][3]*/);

int  qrad_host_store_lighting(qrad_host_scene *s, const uint8_t *dlightdata, int32_t lightdatasize,
                              const int32_t *lightofs /*[F]*/, const uint8_t *styles /*[F][4]*/);
int  qrad_host_lighting(qrad_host_scene *s, const uint8_t **data, int32_t *size);
int  qrad_host_write(qrad_host_scene *s, const char *bsp_path);

/* RadWorld (qrad3.c:494-536) over an already loaded+prepared scene on one GPU context. */
int  qrad_rad_world(qrad_host_scene *s, qrad_cuda_ctx *ctx, const qrad_host_params *params, float *added_out);

/* RadWorld spread over `ngpus` GPUs of this box (devices first_device .. first_device+ngpus-1): one host thread and one
   context per GPU, exchanges over NCCL inside the library.  stats_out (may be NULL) receives
   the statistics of GPU 0 with the ray counts summed over the GPUs. */
int  qrad_rad_world_multi(qrad_host_scene *s, int first_device, int ngpus, const qrad_host_params *params,
                          float *added_out, qrad_cuda_stats *stats_out);

/* Drop-in for the reference executable. */
int  qrad3_main(int argc, char **argv);

#ifdef __cplusplus
}
#endif
#endif
