/*
 * qrad_cuda.h -- C ABI of libqrad_cuda, the B200 (sm_100a) implementation of the
 * qrad3 radiosity hot path of klaussilveira/qengine.
 *
 * The reference has no plugin / FFI interface: qrad3 is one executable whose
 * phases talk through process-global arrays.  The entry
 * points below sit exactly at the five seams of RadWorld() and take the same
 * data those globals hold, as plain pointers and sizes:
 *
 *   qrad_cuda_upload_bsp        replaces  MakeTnodes(&dmodels[0])             src/tools/qrad3/qrad3.c:500, trace.c:47-88
 *                               (+ the PVS / leaf data behind PvsForOrigin,   qrad3.c:131-175, common/bspfile.c:129-154)
 *   qrad_cuda_upload_patches    receives  patches[] / face_patches[]          qrad3.c:503-506, qrad.h:65-100
 *   qrad_cuda_upload_lights     receives  directlights[cluster] lists         qrad3.c:509, lightmap.c:721-838
 *   qrad_cuda_build_facelights  replaces  RunThreadsOnIndividual(numfaces, BuildFacelights)      qrad3.c:512, lightmap.c:967-1056
 *   qrad_cuda_make_transfers    replaces  RunThreadsOnIndividual(num_patches, MakeTransfers)     qrad3.c:516, qrad3.c:185-293
 *   qrad_cuda_bounce            replaces  BounceLight() = numbounce x {ShootLight, CollectLight}  qrad3.c:520, 383-473
 *   qrad_cuda_final_light       replaces  RunThreadsOnIndividual(numfaces, FinalLightFace)        qrad3.c:535, lightmap.c:1066-1201
 *
 * Conventions: every function returns 0 on success, non-zero on failure with
 * the message available from qrad_cuda_last_error() so a C host can route it
 * into its Error() (common/cmdlib.c:142).  The library owns all device memory;
 * the caller owns every pointer it passes and it only needs to stay valid for
 * the duration of the call.  One context drives one GPU from one host thread.
 * There is no CPU fallback: without a CUDA device qrad_cuda_create fails.
 *
 * All structs below are plain-old-data with the on-disk IBSP v38 layouts of
 * common/qfiles.h:252-469 where a lump is passed through untouched.
 */
#ifndef QRAD_CUDA_H
#define QRAD_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QRAD_CUDA_ABI_VERSION 2

typedef struct qrad_cuda_ctx qrad_cuda_ctx;

/* ---- IBSP v38 on-disk records (little endian), qfiles.h:283-447 ---- */
typedef struct { float normal[3]; float dist; int32_t type; } qrad_dplane_t;                     /* 20 B */
typedef struct { int32_t planenum; int32_t children[2]; int16_t mins[3], maxs[3];
                 uint16_t firstface, numfaces; } qrad_dnode_t;                                   /* 28 B */
typedef struct { int32_t contents; int16_t cluster, area; int16_t mins[3], maxs[3];
                 uint16_t firstleafface, numleaffaces, firstleafbrush, numleafbrushes; } qrad_dleaf_t; /* 28 B */
typedef struct { uint16_t planenum; int16_t side; int32_t firstedge; int16_t numedges; int16_t texinfo;
                 uint8_t styles[4]; int32_t lightofs; } qrad_dface_t;                            /* 20 B */
typedef struct { float vecs[2][4]; int32_t flags; int32_t value; char texture[32]; int32_t nexttexinfo; } qrad_texinfo_t; /* 76 B */

/* The lumps the device needs for ray tests, point location and PVS. */
typedef struct {
  const qrad_dplane_t *planes; int32_t numplanes;
  const qrad_dnode_t  *nodes;  int32_t numnodes;
  const qrad_dleaf_t  *leafs;  int32_t numleafs;
  const uint8_t       *visdata; int32_t visdatasize;   /* dvis_t header + RLE rows; 0 = un-vised map */
} qrad_bsp_view;

/* patches[] as structure-of-arrays; N = num_patches (qrad.h:65-91). */
typedef struct {
  int32_t      num_patches;
  const float *origin;        /* [N][3] patch->origin (centroid + 1 * normal)            */
  const float *normal;        /* [N][3] patch->plane->normal                              */
  const float *area;          /* [N]                                                      */
  const int32_t *cluster;     /* [N]    -1 = in solid                                     */
  const uint8_t *sky;         /* [N]                                                      */
  const float *reflectivity;  /* [N][3]                                                   */
  const float *baselight;     /* [N][3]                                                   */
  const float *totallight;    /* [N][3] state after CreateDirectLights (emissive cleared) */
  const float *bounds;        /* [N][6] WindingBounds mins, maxs (AddSampleToPatch)       */
  const int32_t *face;        /* [N]    owning face                                       */
  const int32_t *next;        /* [N]    next patch on the face list, -1 = end (patch->next)  */
  int32_t      numfaces;
  const int32_t *face_head;   /* [numfaces] face_patches[f], -1 = none                    */
} qrad_patches_view;

/* directlights[] flattened: lights of cluster c are light_first[c] .. light_first[c+1]-1,
   stored in the reference's list order (LIFO of creation, lightmap.c:751,785). */
typedef struct {
  int32_t      numclusters;
  const int32_t *light_first; /* [numclusters+1] */
  int32_t      numlights;
  const int32_t *type;        /* 0 emit_surface, 1 emit_point, 2 emit_spotlight (qrad.h:34-39) */
  const float *intensity;
  const int32_t *style;
  const float *origin;        /* [L][3] */
  const float *color;         /* [L][3] */
  const float *normal;        /* [L][3] */
  const float *stopdot;
} qrad_lights_view;

/* Per-face lightmap frame = the lightinfo_t fields CalcFaceVectors / CalcFaceExtents
   produce (lightmap.c:480-603); computed by the host in the reference's float/double mix. */
typedef struct {
  int32_t      numfaces;
  const uint8_t *lit;         /* [F] 0 for SURF_WARP|SURF_SKY faces (lightmap.c:981)  */
  const float *facenormal;    /* [F][3] */
  const float *texorg;        /* [F][3] */
  const float *textoworld;    /* [F][2][3] */
  const float *exactmins;     /* [F][2] */
  const float *exactmaxs;     /* [F][2] */
  const int32_t *texmins;     /* [F][2] */
  const int32_t *texsize;     /* [F][2] */
  const float *face_bounds;   /* [F][6] vertex bounds (FinalLightFace facemins/facemaxs, lightmap.c:1109-1117) */
  const int32_t *planelink_head; /* [F] first face on the same (side, plane) list as walked by lightmap.c:1123 (0 = empty) */
  const int32_t *facelinks;   /* [F] facelinks[] (lightmap.c:42-52) */
  const float *minlight;      /* [F] FloatForKey(face_entity, "_minlight") * 128 */
  const float *plane_normal;  /* [F][3] dplanes[f->planenum].normal (triangulation plane, NOT flipped by side) */
} qrad_faces_view;

typedef struct {
  float ambient, maxlight, lightscale, subdiv;
  int32_t numbounce;
} qrad_final_params;

typedef struct {
  int64_t rays_transfers;     /* root TestLine_r calls issued by MakeTransfers (qrad3.c:244)      */
  int64_t rays_direct;        /* root TestLine_r calls of CalcPoints + GatherSampleLight          */
  int64_t node_visits;        /* BSP node visits of all rays (0 unless counting was enabled)      */
  int64_t candidate_pairs;    /* PVS-visible (i, j) pairs examined by MakeTransfers               */
  int64_t total_transfers;    /* nnz of the transfer matrix (total_transfer, qrad3.c:183)         */
  int64_t transfer_bytes;     /* bytes of the gather-form matrix streamed per bounce              */
  int64_t padded_entries;     /* stored entries including slice padding                           */
  int64_t nonzero_words;      /* (source, 32-receiver slice) pairs with at least one transfer     */
  int32_t num_patches, num_luxels, num_clusters, tree_depth;
  float   ms_upload, ms_facelights, ms_transfers, ms_transfers_trace, ms_bounce, ms_bounce_kernel, ms_final;
  int32_t kernel_launches;    /* kernels launched by this context so far                          */
  /* wall-clock breakdown of make_transfers on the host thread (ms): slot order + uploads, ordered lists, pass 2, pass 3 */
  float   ms_tr_setup, ms_tr_lists, ms_tr_rows, ms_tr_columns;
  int64_t full_walks;         /* MakeTransfers rays that needed the full TestLine walk (counted with node counting on) */
  int64_t box_pruned_rays;    /* MakeTransfers rays settled by a whole-word box proof (counted with node counting on)  */
  int64_t h2d_bytes, d2h_bytes;   /* bytes this context's calls copied host -> device / device -> host since its creation */
  /* multi-GPU (a context that joined a communicator): milliseconds spent in the exchanges of the last RadWorld */
  float   ms_x_matrix;        /* all-to-all of the visibility matrix (rows of the owner -> owners of the receiver slices) */
  float   ms_x_small;         /* row totals / counts, samplelight, totallight, lump */
  float   ms_x_radiosity;     /* radiosity vector, all bounces */
  int32_t world, rank;
} qrad_cuda_stats;

int  qrad_cuda_abi_version(void);
const char *qrad_cuda_last_error(void);

int  qrad_cuda_create(int device, qrad_cuda_ctx **out);
void qrad_cuda_destroy(qrad_cuda_ctx *ctx);
/* ---- several GPUs of one box: one context per GPU (one process or one host thread each) ----
   The reference has no multi-process mode; this is the sharding of SURVEY.md section 8e behind the same seams.  Rank 0
   obtains an id, the host hands it to every rank by any means (a file, MPI, torch.distributed, shared memory of its own
   threads), and every rank calls qrad_cuda_comm_init.  From then on the SAME entry points (upload_*, build_facelights,
   make_transfers, bounce, final_light), called on every rank in the same order, each do their share of the faces / source
   rows / receiver slices and exchange inside the library with NCCL over NVLink (loaded at run time: libnccl.so.2): an
   all-to-all of the visibility rows before the transposition, one broadcast round of the radiosity vector per bounce,
   sums of the per-patch light.  Every rank ends with the complete lighting lump.  world = 1 leaves the communicator. */
int  qrad_cuda_comm_unique_id(uint8_t *id128 /*[128]*/);
int  qrad_cuda_comm_init(qrad_cuda_ctx *ctx, int rank, int world, const uint8_t *id128);
int  qrad_cuda_set_count_nodes(qrad_cuda_ctx *ctx, int enable);

int  qrad_cuda_upload_bsp(qrad_cuda_ctx *ctx, const qrad_bsp_view *bsp);
int  qrad_cuda_upload_patches(qrad_cuda_ctx *ctx, const qrad_patches_view *patches);
int  qrad_cuda_upload_lights(qrad_cuda_ctx *ctx, const qrad_lights_view *lights);
int  qrad_cuda_upload_faces(qrad_cuda_ctx *ctx, const qrad_faces_view *faces);

/* Batched TestLine_r(0, start, stop) (trace.c:92-142): results[i] = 1 if blocked. */
int  qrad_cuda_test_lines(qrad_cuda_ctx *ctx, int64_t n, const float *starts, const float *stops, uint8_t *results);
/* PointInLeafnum (qrad3.c:131-150) for a batch of points. */
int  qrad_cuda_point_in_leaf(qrad_cuda_ctx *ctx, int64_t n, const float *points, int32_t *leafnums);

int  qrad_cuda_build_facelights(qrad_cuda_ctx *ctx, int extrasamples, int numbounce);
int  qrad_cuda_make_transfers(qrad_cuda_ctx *ctx, int nopvs, int64_t *total_transfer);
/* numbounce x {ShootLight, CollectLight}; added_out (may be NULL) receives CollectLight()'s return per bounce. */
int  qrad_cuda_bounce(qrad_cuda_ctx *ctx, int numbounce, float *added_out);
/* Writes the lighting lump.  dlightdata_out must hold lightdata_capacity bytes (MAX_MAP_LIGHTING = 0x200000). */
int  qrad_cuda_final_light(qrad_cuda_ctx *ctx, const qrad_final_params *params,
                           uint8_t *dlightdata_out, int32_t lightdata_capacity, int32_t *lightdatasize,
                           int32_t *lightofs_out /*[F]*/, uint8_t *styles_out /*[F][4]*/);

/* ---- planning (host only, no GPU needed): the geometry of the transfer matrix and of its distribution over `world`
   ranks, exactly as make_transfers uses it (qengine_b200/csrc/plan.hpp).  Lets a host size its memory per GPU, and lets
   the CPU tests drive the exchange protocol over any transport with the product's own addressing. ---- */
typedef struct qrad_plan qrad_plan;
const char *qrad_plan_last_error(void);
int  qrad_plan_create(int32_t num_patches, const int32_t *cluster /*[N] patch->cluster*/, const uint8_t *visdata,
                      int32_t visdatasize, int nopvs, int world, int rank, qrad_plan **out);
void qrad_plan_destroy(qrad_plan *plan);
/* out[0] slots, [1] slices (32 receivers each), [2] words of this rank's row buffer, [3] row groups, [4] own row groups,
   [5] pass-1 runs of this rank, [6] clusters, [7] words of the whole matrix */
int  qrad_plan_info(const qrad_plan *plan, int64_t *out8);
int  qrad_plan_slices(const qrad_plan *plan, int32_t *lo /*[world]*/, int32_t *hi /*[world]*/);
int  qrad_plan_slots(const qrad_plan *plan, int32_t *slot_owner /*[slots] or NULL*/, int32_t *patch_slot /*[N] or NULL*/);
/* the contiguous piece of rank q's row buffer that rank r needs for its slices (what q sends r before pass 3) */
int  qrad_plan_piece(const qrad_plan *plan, int q, int r, int64_t *first_word, int64_t *num_words);
/* address of the word (source slot -> slice) in the row buffer of the source's owner; -1 = the pair has no word */
int64_t qrad_plan_word_address(const qrad_plan *plan, int32_t source_slot, int32_t slice);
/* pass 3's walk of a slice: its candidate sources in ascending patch index, who owns each, and where its word lies inside
   the piece the owner sends to the slice's owner */
int  qrad_plan_column_list(const qrad_plan *plan, int32_t slice, int32_t *slots, int32_t *owner, int64_t *addr_in_piece,
                           int64_t capacity, int64_t *count);

/* ---- host builds of the walkers (no GPU needed): the device functions of qengine_b200/csrc/walkers.cuh -- TestLine,
   PointInLeaf, the common start node of a bundle of rays, the cell proofs that settle rays without a walk -- compiled for
   the host from the same source, over the same flattened tree qrad_cuda_upload_bsp builds.  For tests: they can be held
   against the reference's golden rays and against each other without a GPU.  Not a CPU fallback: nothing in the library
   calls them. ---- */
typedef struct qrad_host_tree qrad_host_tree;
const char *qrad_host_tree_last_error(void);
int  qrad_host_tree_create(const qrad_bsp_view *bsp, qrad_host_tree **out);
void qrad_host_tree_destroy(qrad_host_tree *tree);
/* out[0] nodes, [1] leafs, [2] depth, [3] solid leaves that have a box cell; *extent = largest axial plane distance */
int  qrad_host_tree_info(const qrad_host_tree *tree, int32_t *out4, float *extent);
int  qrad_host_test_lines(const qrad_host_tree *tree, int64_t n, const float *starts, const float *stops,
                          const int32_t *start_nodes /*or NULL*/, uint8_t *results, int32_t *blockers /*or NULL*/);
int  qrad_host_point_in_leaf(const qrad_host_tree *tree, int64_t n, const float *points, int32_t *leafs);
int  qrad_host_common_start_node(const qrad_host_tree *tree, const float *amn, const float *amx, const float *bmn, const float *bmx);
/* per segment: the first leaf in [leaf_lo, leaf_hi) whose cell proves it blocked, or -1 (margin <= 0: the library's own) */
int  qrad_host_prove_segments(const qrad_host_tree *tree, float margin, int64_t n, const float *starts, const float *stops,
                              int32_t leaf_lo, int32_t leaf_hi, int32_t *proving_leaf);
/* the first leaf whose cell proves ALL segments from box A to box B blocked, or -1 */
int  qrad_host_prove_boxes(const qrad_host_tree *tree, float margin, const float *amn, const float *amx, const float *bmn,
                           const float *bmx);

/* ---- inspection (parity tests, multi-GPU exchange); not needed by a production host ---- */
int  qrad_cuda_get_stats(qrad_cuda_ctx *ctx, qrad_cuda_stats *out);
/* pass 1 of make_transfers as rank `rank` of `world` would run it, alone on this GPU and without any exchange: time and
   rays of that rank's share (load-balance measurements without the ranks); the context is left without transfers */
int  qrad_cuda_trace_shard(qrad_cuda_ctx *ctx, int nopvs, int rank, int world, float *trace_ms, int64_t *rays);
/* numtransfers per patch (patch->numtransfers). */
int  qrad_cuda_download_numtransfers(qrad_cuda_ctx *ctx, int32_t *out /*[N]*/);
/* All transfer rows in the reference's order (by source patch, receivers ascending): row_ptr[N+1],
   then `patch` / `transfer` arrays of total_transfer entries (transfer_t, qrad.h:57-61; patch widened to u32). */
int  qrad_cuda_download_transfers(qrad_cuda_ctx *ctx, int64_t *row_ptr, uint32_t *patch, uint16_t *transfer);
/* The rows of `nrows` selected source patches only (row-sample checks on maps whose whole matrix is tens of GB):
   row_ptr[nrows+1]; patch / transfer must hold `capacity` >= sum of their numtransfers records; rays (may be NULL)
   receives per row the root TestLine_r calls MakeTransfers issues for it (qrad3.c:244). */
int  qrad_cuda_download_rows(qrad_cuda_ctx *ctx, int32_t nrows, const int32_t *rows /*patch indices*/, int64_t *row_ptr,
                             uint32_t *patch, uint16_t *transfer, int64_t capacity, int64_t *rays);
int  qrad_cuda_download_patch_light(qrad_cuda_ctx *ctx, float *totallight /*[N][3]*/, float *samplelight /*[N][3]*/,
                                    int32_t *samples /*[N]*/, float *radiosity /*[N][3]*/);
int  qrad_cuda_upload_patch_light(qrad_cuda_ctx *ctx, const float *totallight, const float *samplelight, const int32_t *samples);
/* facelight[]: luxel_first[F+1]; origins [num_luxels][3]; numstyles[F]; stylenums[F][32];
   samples for style slot s of face f at samples[(slot_first[f]+s) ...] -- see qengine_b200/qrad.py */
int  qrad_cuda_facelight_layout(qrad_cuda_ctx *ctx, int32_t *luxel_first /*[F+1]*/, int32_t *numstyles /*[F]*/, int32_t *stylenums /*[F][32]*/);
/* distinct light styles of the map, ascending (ids[0] == 0); these index the style tables below */
int  qrad_cuda_style_ids(qrad_cuda_ctx *ctx, int32_t *ids /*[256]*/, int32_t *count);
int  qrad_cuda_download_facelight(qrad_cuda_ctx *ctx, float *origins /*[L][3]*/, int32_t style_slot, float *samples /*[L][3]*/);

#ifdef __cplusplus
}
#endif
#endif /* QRAD_CUDA_H */
