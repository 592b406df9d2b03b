/* C ABI of the next stage outward of the qrad3 path: the reference's qvis3, FIRST STAGE ONLY.
 *
 * qvis3 (src/tools/qvis3) = LoadPortals -> BasePortalVis for every portal (portalfront, SimpleFlood -> portalflood,
 * nummightsee) -> SortPortals -> PortalFlow for every portal in that order -> ClusterMerge, CalcPHS.  What is here:
 *
 *   qvis_host_portals_*          LoadPortals + PlaneFromWinding, qvis3.c:58-69, 304-412 (host, no GPU)
 *   qvis_cuda_base_portal_vis    replaces  RunThreadsOnIndividual(numportals * 2, true, BasePortalVis)  qvis3.c:255
 *                                (BasePortalVis + SimpleFlood, flow.c:576-652) with two CUDA kernels
 *
 * PortalFlow is not built: its result depends on the order the portals are finished in (DESIGN.md section 9 has the
 * schedule that reproduces it and oracle/qvis_oracle.c the pinned restatement to hold kernels against).  There is no CPU
 * fallback: without a CUDA device qvis_cuda_base_portal_vis fails.  Plain pointers and sizes only.
 */
#ifndef QVIS_CUDA_H
#define QVIS_CUDA_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* The memory portals (two per portal of the file: the forward one with the flipped plane, the backward one with the
   reversed winding, qvis3.c:377-407) as flat arrays. */
typedef struct {
  int32_t numportals;            /* memory portals = 2 x the file's count */
  int32_t numclusters;           /* portalclusters */
  const float *planes;           /* [numportals][4] normal (pointing into the neighbour), dist */
  const int32_t *leaf;           /* [numportals] the cluster the portal leads into */
  const int32_t *winding_first;  /* [numportals + 1] into points */
  const float *points;           /* [..][3] */
  const int32_t *leaf_first;     /* [numclusters + 1] into leaf_portals */
  const int32_t *leaf_portals;   /* portals leaving each cluster, in the order LoadPortals adds them */
} qvis_portals_view;

typedef struct qvis_host_portals qvis_host_portals;
const char *qvis_last_error(void);
int  qvis_host_portals_load(const char *prt_path, qvis_host_portals **out);
void qvis_host_portals_free(qvis_host_portals *p);
int  qvis_host_portals_view(const qvis_host_portals *p, qvis_portals_view *out);
/* ((numportals + 63) & ~63) >> 3: bytes of one bit vector over the portals, as the reference sizes them (qvis3.c:336) */
int  qvis_portal_bytes(int32_t numportals);

/* portalfront / portalflood: [numportals][qvis_portal_bytes] bit vectors (bit j of row p: byte j >> 3, bit j & 7);
   nummightsee: [numportals] CountBits(portalflood).  ms_front / ms_flood (may be NULL): CUDA-event times of the kernels. */
int  qvis_cuda_base_portal_vis(int device, const qvis_portals_view *portals, uint8_t *portalfront, uint8_t *portalflood,
                               int32_t *nummightsee, float *ms_front, float *ms_flood);

/* ClusterMerge for every cluster + CalcPHS + CompressVis (qvis3.c:168-224, 427-486, bspfile.c:95): the visibility lump
   from the portals' FINAL vectors [numportals][qvis_portal_bytes] (with -fast: the flood vectors).  Host work.  Errors
   carry the reference's text ("Vismap expansion overflow" when the lump passes `cap`). */
int  qvis_host_merge_lump(const qvis_host_portals *p, const uint8_t *portalvis, uint8_t *lump, int32_t cap, int32_t *size,
                          int32_t *avg_visible, int32_t *avg_hearable, int32_t *saw_into_leaf);
/* `qvis3 -fast map.bsp` (fastvis, qvis3.c:235-241): BasePortalVis on the GPU, the rest on the host; writes bsp_out. */
int  qvis_host_fast_vis(const char *bsp_in, const char *prt_path, const char *bsp_out, int device);
/* main() of the tool for what is built: the reference's flags, path rules, chatter and error banner; stops without -fast. */
int  qvis3_main(int argc, char **argv);

#ifdef __cplusplus
}
#endif
#endif
