// pybind entry points for the reference's UNMODIFIED fastba kernels (compiled by path from
// /root/reference/cdvslam/fastba/ba_cuda.cu and block_e.cu).  Replaces the reference's ba.cpp, which cannot be
// built here because solve_system needs Eigen/Sparse (ba.cpp:99-180).  Test infrastructure only.
#include <torch/extension.h>
#include <vector>

std::vector<torch::Tensor> cuda_ba(torch::Tensor poses, torch::Tensor patches, torch::Tensor intrinsics,
                                   torch::Tensor target, torch::Tensor weight, torch::Tensor lmbda,
                                   torch::Tensor ii, torch::Tensor jj, torch::Tensor kk, const int PPF,
                                   int t0, int t1, int iterations, bool eff_impl);
torch::Tensor cuda_reproject(torch::Tensor poses, torch::Tensor patches, torch::Tensor intrinsics,
                             torch::Tensor ii, torch::Tensor jj, torch::Tensor kk);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("forward", &cuda_ba, "reference BA forward");
  m.def("reproject", &cuda_reproject, "reference reproject");
}
