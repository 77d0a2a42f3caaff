#include "b200_runtime.h"

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

void DualCloud::materialise() {
  if (hostValid_) return;
  buf_.resize((std::size_t)devN_ * 4 + 4);
  int n = 0;
  ctx_->check(loam_b200_cloud_download(ctx_->get(), slot_, buf_.data(), devN_, &n), "loam_b200_cloud_download");
  unpack(buf_.data(), (std::size_t)n, *host_);
  hostValid_ = true;
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
  // column-major copies for the dense kernels
  float A[36], b[6];
  for (int i = 0; i < 6; i++) {
    b[i] = ne.AtB[i];
    for (int j = 0; j < 6; j++) A[i + j * 6] = ne.AtA[i * 6 + j];
  }
  float Aq[36];
  std::memcpy(Aq, A, sizeof A);
  loamb::colpiv_qr_solve<6, 6>(Aq, b, x);

  if (firstIteration) {
    float E[6], V[36], V2[36];
    loamb::sym_eigen<6>(A, E, V);  // ascending eigenvalues, V column-major (column = eigenvector)
    std::memcpy(V2, V, sizeof V);
    isDegenerate = false;
    for (int i = 0; i < 6; i++) {
      if (E[i] < eigenThreshold) {
        for (int j = 0; j < 6; j++) V2[i + j * 6] = 0.f;  // zero ROW i
        isDegenerate = true;
      } else {
        break;
      }
    }
    float Vinv[36];
    loamb::lu_inverse<6>(V, Vinv);
    // P = V^-1 * V2, stored row-major
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++) {
        float acc = 0.f;
        for (int k = 0; k < 6; k++) acc += Vinv[i + k * 6] * V2[k + j * 6];
        P[i * 6 + j] = acc;
      }
  }
  if (isDegenerate) {
    float x2[6];
    for (int i = 0; i < 6; i++) x2[i] = x[i];
    for (int i = 0; i < 6; i++) {
      float acc = 0.f;
      for (int k = 0; k < 6; k++) acc += P[i * 6 + k] * x2[k];
      x[i] = acc;
    }
  }
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
