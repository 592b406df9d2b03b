#include "comm.hpp"

#include <dlfcn.h>

#include <mutex>

#include "qrad_device.cuh"

namespace qrad {

const NcclApi &nccl() {
  static NcclApi api{};
  static std::once_flag once;
  static std::string error;
  std::call_once(once, [] {
    void *h = nullptr;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) {
      error = std::string("multi-GPU needs NCCL and libnccl.so.2 could not be loaded: ") + dlerror();
      return;
    }
    auto need = [&](const char *sym) -> void * {
      void *p = dlsym(h, sym);
      if (!p && error.empty()) error = std::string("libnccl.so.2 lacks ") + sym;
      return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId)) need("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank)) need("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy)) need("ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce)) need("ncclAllReduce");
    api.Broadcast = (decltype(api.Broadcast)) need("ncclBroadcast");
    api.AllGather = (decltype(api.AllGather)) need("ncclAllGather");
    api.Send = (decltype(api.Send)) need("ncclSend");
    api.Recv = (decltype(api.Recv)) need("ncclRecv");
    api.GroupStart = (decltype(api.GroupStart)) need("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd)) need("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString)) need("ncclGetErrorString");
    if (auto gv = (ncclResult_t(*)(int *)) dlsym(h, "ncclGetVersion")) gv(&api.version);
  });
  if (!error.empty()) dfail("%s", error.c_str());
  return api;
}

}  // namespace qrad
