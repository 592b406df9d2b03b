/* TEST INFRASTRUCTURE ONLY (oracle/): drives the REFERENCE qvis3 translation units (compiled in place from
 * /root/reference/src/tools/qvis3, main() renamed) and dumps what the next stage of this project will be held against:
 * per portal the plane, the neighbour cluster, nummightsee, the three bit vectors of vis.h:70-72 (portalfront,
 * portalflood, portalvis), and the order SortPortals (qvis3.c:118-129) puts them in.
 *
 *   qvis_harness [-fast] [-nosort] map.bsp map.prt out.bin     (map.bsp is rewritten with the new visibility lump,
 *                                                               exactly as the reference's main does, qvis3.c:551-560)
 * out.bin: int32 {magic 'QVD1', portalclusters, numportals*2, portalbytes}, then per portal (file order)
 *          float plane[4], int32 leaf, int32 nummightsee, int32 rank in sorted order, then the three vectors. */
#include "vis.h"

extern qboolean fastvis, nosort;
void LoadPortals(char *name);
void CalcVis(void);
void CalcPHS(void);
extern byte *uncompressedvis;

int main(int argc, char **argv)
{
  int i = 1;
  for (; i < argc && argv[i][0] == '-'; i++) {
    if (!strcmp(argv[i], "-fast"))
      fastvis = true;
    else if (!strcmp(argv[i], "-nosort"))
      nosort = true;
    else
      Error("qvis_harness: unknown option %s", argv[i]);
  }
  if (argc - i != 3)
    Error("usage: qvis_harness [-fast] [-nosort] map.bsp map.prt out.bin");
  ThreadSetDefault();
  LoadBSPFile(argv[i]);
  if (numnodes == 0 || numfaces == 0)
    Error("Empty map");
  LoadPortals(argv[i + 1]);
  CalcVis();
  CalcPHS();
  visdatasize = vismap_p - dvisdata;
  WriteBSPFile(argv[i]);

  FILE *f = fopen(argv[i + 2], "wb");
  if (!f)
    Error("qvis_harness: cannot write %s", argv[i + 2]);
  int np = numportals * 2;
  int hdr[4] = {0x31445651, portalclusters, np, portalbytes};
  fwrite(hdr, 4, 4, f);
  int *rank = malloc(np * sizeof(int));
  for (int k = 0; k < np; k++)
    rank[sorted_portals[k] - portals] = k;
  for (int k = 0; k < np; k++) {
    portal_t *p = portals + k;
    float pl[4] = {p->plane.normal[0], p->plane.normal[1], p->plane.normal[2], p->plane.dist};
    int meta[3] = {p->leaf, p->nummightsee, rank[k]};
    fwrite(pl, 4, 4, f);
    fwrite(meta, 4, 3, f);
    fwrite(p->portalfront, 1, portalbytes, f);
    fwrite(p->portalflood, 1, portalbytes, f);
    fwrite(p->portalvis, 1, portalbytes, f);
  }
  fclose(f);
  printf("QVISDUMP portals %d clusters %d visdatasize %d\n", np, portalclusters, visdatasize);
  return 0;
}
