#include "b200_runtime.h"
#include "lmstep.cuh"

#include <cstdlib>
#include <cstring>

#include "../linalg.cuh"

namespace loam {
namespace b200 {

static int g_device = -1;

int defaultDevice() {
  if (g_device >= 0) return g_device;
  if (const char* e = std::getenv("LOAM_B200_DEVICE")) return std::atoi(e);
  if (const char* e = std::getenv("LOCAL_RANK")) return std::atoi(e);
  return 0;
}
void setDefaultDevice(int device) { g_device = device; }

loam_b200_ctx* Context::get() {
  if (!ctx_) {
    int rc = loam_b200_create(&ctx_, defaultDevice());
    if (rc != LOAM_B200_OK) {
      ctx_ = nullptr;
      throw std::runtime_error(std::string("loam_b200_create: ") + loam_b200_strerror(rc));
    }
    static const bool no_prio = std::getenv("LOAM_B200_NO_PRIORITY") != nullptr;
    if (priority_ != 0 && !no_prio) check(loam_b200_set_priority(ctx_, priority_), "loam_b200_set_priority");
  }
  return ctx_;
}

void Context::check(int status, const char* what) {
  if (status == LOAM_B200_OK) return;
  std::string msg = std::string(what) + ": " + loam_b200_strerror(status);
  if (ctx_) {
    const char* d = loam_b200_last_error(ctx_);
    if (d && *d) msg += std::string(" (") + d + ")";
  }
  throw std::runtime_error(msg);
}

void DualCloud::ensureDevice() {
  if (devValid_) return;
  pack(*host_, buf_);
  ctx_->check(loam_b200_cloud_upload(ctx_->get(), slot_, buf_.data(), (int)host_->points.size()), "loam_b200_cloud_upload");
  devValid_ = true;
  devN_ = (int)host_->points.size();
}

int DualCloud::deviceCount() const {
  if (devN_ < 0) {  // produced asynchronously (deviceWrittenLazy): the context knows the size once the producer is done
    const int n = loam_b200_cloud_size(ctx_->get(), slot_);
    const_cast<DualCloud*>(this)->devN_ = n < 0 ? 0 : n;
  }
  return devN_;
}

void DualCloud::materialise() {
  if (hostValid_) return;
  deviceCount();
  buf_.resize((std::size_t)devN_ * 4 + 4);
  int n = 0;
  ctx_->check(loam_b200_cloud_download(ctx_->get(), slot_, buf_.data(), devN_, &n), "loam_b200_cloud_download");
  unpack(buf_.data(), (std::size_t)n, *host_);
  hostValid_ = true;
}

void DualCloud::dropNonFinite() {
  if (!hostValid_ || host_->is_dense) return;
  auto& pts = host_->points;
  std::size_t j = 0;
  for (std::size_t i = 0; i < pts.size(); i++)
    if (std::isfinite(pts[i].x) && std::isfinite(pts[i].y) && std::isfinite(pts[i].z)) pts[j++] = pts[i];
  if (j != pts.size()) devValid_ = false;
  pts.resize(j);
  host_->width = static_cast<std::uint32_t>(j);
  host_->height = 1;
  host_->is_dense = true;
}

void DualCloud::swap(DualCloud& o) {
  host_.swap(o.host_);
  std::swap(hostValid_, o.hostValid_);
  std::swap(devValid_, o.devValid_);
  std::swap(devN_, o.devN_);
  // the device buffers trade places so each slot keeps meaning "this member"
  if (ctx_ && ctx_ == o.ctx_ && ctx_->created())
    ctx_->check(loam_b200_cloud_swap(ctx_->get(), slot_, o.slot_), "loam_b200_cloud_swap");
}

void GaussNewtonSolver::solve(const loam_b200_normal_eq& ne, bool firstIteration, float eigenThreshold, float x[6]) {
  loamb::GnState g;
  std::memcpy(g.P, P, sizeof P);
  g.degenerate = isDegenerate ? 1 : 0;
  loamb::gn_solve(ne.AtA, ne.AtB, firstIteration, eigenThreshold, g, x);  // csrc/lmstep.cuh: same source as the device loop
  std::memcpy(P, g.P, sizeof P);
  isDegenerate = g.degenerate != 0;
}

bool deviceResidentLoops() {
  static const bool on = [] {
    const char* e = std::getenv("LOAM_B200_DEVICE_LOOP");
    return e && e[0] == '1';
  }();
  return on;
}

void voxelFilter(Context& ctx, const Cloud& in, float leaf, Cloud& out, std::vector<float>& sin, std::vector<float>& sout) {
  out.clear();
  out.is_dense = true;
  if (in.points.empty()) return;
  pack(in, sin);
  sout.resize(sin.size());
  int n_out = 0;
  ctx.check(loam_b200_voxel_grid(ctx.get(), sin.data(), (int)in.points.size(), leaf, sout.data(), (int)in.points.size(),
                                 &n_out),
            "loam_b200_voxel_grid");
  unpack(sout.data(), (std::size_t)n_out, out);
  out.header = in.header;
}

}  // namespace b200
}  // namespace loam
