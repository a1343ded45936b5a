// TEST INFRASTRUCTURE: the DeviceAPI half of the mock ICICLE frontend (see icicle_mock.cu).  Separate translation unit
// because the reference's two header families cannot meet in one: include/icicle/device_api.h brings upstream's
// icicle::Device / eIcicleError, include/icicle_types.cuh (behind icicle_backend_api.cuh) declares its own.
// register_deviceAPI and the DeviceAPI vtable order come from bls12-381/include/icicle/device_api.h:54-227.
#include <map>
#include <memory>
#include <string>

#include "icicle/device_api.h"

namespace {
std::map<std::string, std::shared_ptr<icicle::DeviceAPI>>& D() { static std::map<std::string, std::shared_ptr<icicle::DeviceAPI>> d; return d; }
icicle::DeviceAPI* dev() { return D().at("CUDA").get(); }
}  // namespace
namespace icicle {
void register_deviceAPI(const std::string& t, std::shared_ptr<DeviceAPI> api) { D()[t] = std::move(api); }
}  // namespace icicle
#define ERR(e) static_cast<int>(e)

extern "C" {
int mock_device_registered() { return D().count("CUDA") ? 1 : 0; }
int mock_dev_set_device(int id) { icicle::Device d{"CUDA", id}; return ERR(dev()->set_device(d)); }
int mock_dev_count(int* n) { return ERR(dev()->get_device_count(*n)); }
int mock_dev_malloc(void** p, size_t bytes) { return ERR(dev()->allocate_memory(p, bytes)); }
int mock_dev_malloc_async(void** p, size_t bytes, void* st) { return ERR(dev()->allocate_memory_async(p, bytes, st)); }
int mock_dev_free(void* p) { return ERR(dev()->free_memory(p)); }
int mock_dev_free_async(void* p, void* st) { return ERR(dev()->free_memory_async(p, st)); }
int mock_dev_mem(size_t* total, size_t* free_) { return ERR(dev()->get_available_memory(*total, *free_)); }
int mock_dev_memset(void* p, int v, size_t bytes) { return ERR(dev()->memset(p, v, bytes)); }
int mock_dev_memset_async(void* p, int v, size_t bytes, void* st) { return ERR(dev()->memset_async(p, v, bytes, st)); }
int mock_dev_copy(void* dst, const void* src, size_t bytes, int direction) {
  return ERR(dev()->copy(dst, src, bytes, (icicle::eCopyDirection)direction));
}
int mock_dev_copy_async(void* dst, const void* src, size_t bytes, int direction, void* st) {
  return ERR(dev()->copy_async(dst, src, bytes, (icicle::eCopyDirection)direction, st));
}
int mock_dev_synchronize(void* st) { return ERR(dev()->synchronize(st)); }
int mock_dev_create_stream(void** st) { return ERR(dev()->create_stream(st)); }
int mock_dev_destroy_stream(void* st) { return ERR(dev()->destroy_stream(st)); }
int mock_dev_properties(int* using_host_memory, int* num_memory_regions, int* supports_pinned) {
  icicle::DeviceProperties p{};
  int rc = ERR(dev()->get_device_properties(p));
  *using_host_memory = p.using_host_memory;
  *num_memory_regions = p.num_memory_regions;
  *supports_pinned = p.supports_pinned_memory;
  return rc;
}
}
