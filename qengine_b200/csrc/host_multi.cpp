// RadWorld over several GPUs of one box (include/qrad_host.h: qrad_rad_world_multi).
#include "host_scene.hpp"

extern "C" int qrad_rad_world_multi(qrad_host_scene *, int, int ngpus, const qrad_host_params *, float *, qrad_cuda_stats *) {
  (void) ngpus;
  return 1;
}
