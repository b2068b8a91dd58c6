// pybind entry points for the reference's UNMODIFIED altcorr kernels (compiled by path from
// /root/reference/cdvslam/altcorr/correlation_kernel.cu).  Same table as the reference's correlation.cpp:57-63.
// Test infrastructure only.
#include <torch/extension.h>
#include <vector>

std::vector<torch::Tensor> corr_cuda_forward(torch::Tensor fmap1, torch::Tensor fmap2, torch::Tensor coords,
                                             torch::Tensor ii, torch::Tensor jj, int radius);
std::vector<torch::Tensor> corr_cuda_backward(torch::Tensor fmap1, torch::Tensor fmap2, torch::Tensor coords,
                                              torch::Tensor ii, torch::Tensor jj, torch::Tensor corr_grad, int radius);
std::vector<torch::Tensor> patchify_cuda_forward(torch::Tensor net, torch::Tensor coords, int radius);
std::vector<torch::Tensor> patchify_cuda_backward(torch::Tensor net, torch::Tensor coords, torch::Tensor gradient,
                                                  int radius);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("forward", &corr_cuda_forward, "reference corr forward");
  m.def("backward", &corr_cuda_backward, "reference corr backward");
  m.def("patchify_forward", &patchify_cuda_forward, "reference patchify forward");
  m.def("patchify_backward", &patchify_cuda_backward, "reference patchify backward");
}
