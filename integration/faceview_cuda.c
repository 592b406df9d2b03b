/*
 * faceview_cuda.c -- the part of the libqrad_cuda binding that lives in src/tools/qrad3/lightmap.c (appended to it),
 * because lightinfo_t and the plane chains are private to that file: per lit face, the reference's own
 * CalcFaceVectors / CalcFaceExtents (lightmap.c:480-603) are run exactly as BuildFacelights sets them up
 * (lightmap.c:991-1006) and their results are handed to the library as flat arrays (qrad_faces_view).
 */
#include "qrad_cuda.h"

void FillFaceView(qrad_faces_view *v)
{
  static lightinfo_t l;
  int i, j, k;
  int n = numfaces;
  unsigned char *lit = malloc(n);
  float *facenormal = malloc(n * 12), *texorg = malloc(n * 12), *textoworld = malloc(n * 24);
  float *exactmins = malloc(n * 8), *exactmaxs = malloc(n * 8), *bounds = malloc(n * 24);
  float *minlight = malloc(n * 4), *plane_normal = malloc(n * 12);
  int *texmins = malloc(n * 8), *texsize = malloc(n * 8), *head = malloc(n * 4), *links = malloc(n * 4);
  for (i = 0; i < n; i++) {
    dface_t *f = &dfaces[i];
    vec3_t mins, maxs;
    lit[i] = !(texinfo[f->texinfo].flags & (SURF_WARP | SURF_SKY));
    /* BuildFacelights clears the whole lightinfo_t (1.5 MB of sample points); only these fields matter here */
    VectorClear(l.texorg);
    memset(l.textoworld, 0, sizeof(l.textoworld));
    memset(l.worldtotex, 0, sizeof(l.worldtotex));
    l.exactmins[0] = l.exactmins[1] = l.exactmaxs[0] = l.exactmaxs[1] = 0;
    l.texmins[0] = l.texmins[1] = l.texsize[0] = l.texsize[1] = 0;
    l.numsurfpt = 0;
    l.surfnum = i;
    l.face = f;
    VectorCopy(dplanes[f->planenum].normal, l.facenormal);
    l.facedist = dplanes[f->planenum].dist;
    if (f->side) {
      VectorSubtract(vec3_origin, l.facenormal, l.facenormal);
      l.facedist = -l.facedist;
    }
    VectorCopy(face_offset[i], l.modelorg);
    if (lit[i]) {
      CalcFaceVectors(&l);
      CalcFaceExtents(&l);
    }
    VectorCopy(l.facenormal, (facenormal + 3 * i));
    VectorCopy(l.texorg, (texorg + 3 * i));
    for (j = 0; j < 2; j++) {
      for (k = 0; k < 3; k++)
        textoworld[i * 6 + j * 3 + k] = l.textoworld[j][k];
      exactmins[i * 2 + j] = l.exactmins[j];
      exactmaxs[i * 2 + j] = l.exactmaxs[j];
      texmins[i * 2 + j] = l.texmins[j];
      texsize[i * 2 + j] = l.texsize[j];
    }
    /* vertex bounds as FinalLightFace takes them (lightmap.c:1109-1117) */
    ClearBounds(mins, maxs);
    for (j = 0; j < f->numedges; j++) {
      int ednum = dsurfedges[f->firstedge + j];
      if (ednum >= 0)
        AddPointToBounds(dvertexes[dedges[ednum].v[0]].point, mins, maxs);
      else
        AddPointToBounds(dvertexes[dedges[-ednum].v[1]].point, mins, maxs);
    }
    VectorCopy(mins, (bounds + 6 * i));
    VectorCopy(maxs, (bounds + 6 * i + 3));
    head[i] = planelinks[f->side][f->planenum];
    links[i] = facelinks[i];
    minlight[i] = FloatForKey(face_entity[i], "_minlight") * 128;
    VectorCopy(dplanes[f->planenum].normal, (plane_normal + 3 * i));
  }
  v->numfaces = n;
  v->lit = lit;
  v->facenormal = facenormal;
  v->texorg = texorg;
  v->textoworld = textoworld;
  v->exactmins = exactmins;
  v->exactmaxs = exactmaxs;
  v->texmins = texmins;
  v->texsize = texsize;
  v->face_bounds = bounds;
  v->planelink_head = head;
  v->facelinks = links;
  v->minlight = minlight;
  v->plane_normal = plane_normal;
}
