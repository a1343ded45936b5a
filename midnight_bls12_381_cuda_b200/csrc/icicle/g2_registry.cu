// Backend-local registry for the G2 MSM callbacks (role of bls12-381/src/backend/g2_registry.cu:65-101):
// ICICLE's curve library exports register_g2_msm only when built with G2, so the backend keeps the
// table itself and hands the callbacks out through the getters.
#include <map>
#include <mutex>

#include "icicle_abi.h"

namespace icicle {
namespace {
struct G2Registry {
  std::mutex mu;
  std::map<std::string, MsmG2Impl> msm;
  std::map<std::string, MsmG2PreComputeImpl> precompute;
};
G2Registry& registry() {
  static G2Registry r;   // constructed on first use: safe against static-initialisation order
  return r;
}
}  // namespace

void register_g2_msm(const std::string& deviceType, MsmG2Impl impl) {
  auto& r = registry();
  std::lock_guard<std::mutex> lk(r.mu);
  r.msm[deviceType] = std::move(impl);
}
void register_g2_msm_precompute_bases(const std::string& deviceType, MsmG2PreComputeImpl impl) {
  auto& r = registry();
  std::lock_guard<std::mutex> lk(r.mu);
  r.precompute[deviceType] = std::move(impl);
}
MsmG2Impl get_g2_msm_backend(const std::string& deviceType) {
  auto& r = registry();
  std::lock_guard<std::mutex> lk(r.mu);
  auto it = r.msm.find(deviceType);
  return it == r.msm.end() ? MsmG2Impl() : it->second;
}
MsmG2PreComputeImpl get_g2_msm_precompute_bases_backend(const std::string& deviceType) {
  auto& r = registry();
  std::lock_guard<std::mutex> lk(r.mu);
  auto it = r.precompute.find(deviceType);
  return it == r.precompute.end() ? MsmG2PreComputeImpl() : it->second;
}
}  // namespace icicle
