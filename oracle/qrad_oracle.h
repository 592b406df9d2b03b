/* qrad_oracle.h -- TEST INFRASTRUCTURE ONLY: interface of the plain-C restatement (qrad_oracle.c). */
#ifndef QRAD_ORACLE_H
#define QRAD_ORACLE_H
#include <stdint.h>

void oq_reset(void);
void oq_set_bsp(const void *planes, int numplanes, const void *nodes, int numnodes, const void *leafs, int numleafs,
                const void *vis, int vissize, const void *faces, int numfaces, const void *texinfo, int numtexinfo,
                const void *verts, int numverts, const void *edges, int numedges, const void *surfedges, int numsurfedges,
                const void *models, int nummodels);
void oq_set_params(int numbounce, int extrasamples, float subdiv, float ambient, float maxlight, float lightscale,
                   float direct_scale, float entity_scale, int nopvs);
void oq_texture_reflectivity(int ti, const uint8_t *palette, const uint8_t *texels, int ntexels);
void oq_copy_reflectivity(int dst, int src);
void oq_get_reflectivity(float *out);
void oq_set_model_entity(int model, const char *origin_str, const char *minlight_str);
void oq_make_patches(void);
void oq_subdivide(void);
int oq_num_patches(void);
void oq_get_patches(float *origin, float *normal, float *area, int *cluster, float *totallight, float *samplelight,
                    int *samples, int *numtransfers, int *face, int *next);
void oq_lights_from_patches(void);
void oq_add_light_entity(const char *classname, const char *origin, const char *light, const char *_light,
                         const char *style, const char *_style, const char *color, const char *target,
                         const char *target_origin, const char *cone, const char *angle_s);
int oq_num_lights(void);
void oq_build_facelights(void);
int oq_facelight(int face, int *numsamples, int *stylenums, float **origins, float **samples0);
void oq_make_transfers(void);
long long oq_total_transfer(void);
void oq_get_transfers(uint32_t *patch, uint16_t *transfer);
void oq_ray_counts(long long *direct, long long *transfers);
void oq_bounce(float *added);
int oq_final_light(uint8_t *out, int cap, int *lightofs, uint8_t *styles);
int oq_test_line(const float *start, const float *stop);
int oq_point_in_leaf(const float *p);
#endif
