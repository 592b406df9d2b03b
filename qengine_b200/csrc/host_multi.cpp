// RadWorld over several GPUs of one box from ONE process (include/qrad_host.h: qrad_rad_world_multi): one host thread and
// one libqrad_cuda context per GPU.  The threads share the prepared scene read-only, join one NCCL communicator and then
// make the very same calls a single-GPU host makes (rad_world_impl); every exchange happens inside the CUDA half.
// A multi-PROCESS host (one process per GPU, e.g. under torchrun or MPI) does the same with qrad_cuda_comm_unique_id /
// qrad_cuda_comm_init and qrad_rad_world on every rank: see INTEGRATION.md.
#include <string>
#include <thread>
#include <vector>

#include "host_scene.hpp"

using namespace qrad;

extern "C" int qrad_rad_world_multi(qrad_host_scene *hs, int first_device, int ngpus, const qrad_host_params *p,
                                    float *added_out, qrad_cuda_stats *stats_out) {
  if (ngpus < 1 || ngpus > 16) {
    set_host_error("-gpus must be between 1 and 16");
    return 1;
  }
  uint8_t id[128];
  if (ngpus > 1 && qrad_cuda_comm_unique_id(id)) {
    set_host_error(qrad_cuda_last_error());
    return 1;
  }
  std::vector<std::string> err((size_t) ngpus);
  std::vector<qrad_cuda_stats> stats((size_t) ngpus);
  std::vector<std::thread> th;
  for (int r = 0; r < ngpus; r++)
    th.emplace_back([&, r] {
      qrad_cuda_ctx *ctx = nullptr;
      try {
        if (qrad_cuda_create(first_device + r, &ctx)) fail("%s", qrad_cuda_last_error());
        if (ngpus > 1 && qrad_cuda_comm_init(ctx, r, ngpus, id)) fail("%s", qrad_cuda_last_error());
        rad_world_impl(hs, ctx, p, r == 0 ? added_out : nullptr, r == 0);
        qrad_cuda_get_stats(ctx, &stats[r]);
      } catch (const std::exception &e) {
        err[r] = e.what()[0] ? e.what() : "error";
      }
      if (ctx) qrad_cuda_destroy(ctx);
    });
  for (auto &t : th) t.join();
  for (int r = 0; r < ngpus; r++)
    if (!err[r].empty()) {
      set_host_error("GPU " + std::to_string(first_device + r) + ": " + err[r]);
      return 1;
    }
  if (stats_out) {
    *stats_out = stats[0];
    for (int r = 1; r < ngpus; r++) {
      stats_out->rays_transfers += stats[r].rays_transfers;
      stats_out->rays_direct += stats[r].rays_direct;
      stats_out->node_visits += stats[r].node_visits;
      stats_out->candidate_pairs += stats[r].candidate_pairs;
    }
  }
  return 0;
}
