// TEST / MEASUREMENT INFRASTRUCTURE, not product code.
// pybind glue for the REFERENCE's own CUDA RoIAlign v2 kernel
// (/root/reference/mmdet/ops/roi_align/src/cuda/roi_align_kernel_v2.cu, compiled
// unmodified and in place by oracle/build_oracle.py: build_reference_cuda_ext).
// The reference's own binding (roi_align_ext.cpp:126-161) also pulls in the
// legacy v1 kernels, which no longer compile (THC); this file binds only the two
// v2 launchers it declares at roi_align_ext.cpp:27-40.
#include <torch/extension.h>

at::Tensor ROIAlignForwardV2Laucher(const at::Tensor& input, const at::Tensor& rois,
                                    const float spatial_scale, const int pooled_height,
                                    const int pooled_width, const int sampling_ratio, bool aligned);
at::Tensor ROIAlignBackwardV2Laucher(const at::Tensor& grad, const at::Tensor& rois,
                                     const float spatial_scale, const int pooled_height,
                                     const int pooled_width, const int batch_size,
                                     const int channels, const int height, const int width,
                                     const int sampling_ratio, bool aligned);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("forward_v2", &ROIAlignForwardV2Laucher, "reference RoIAlign v2 forward (CUDA)");
  m.def("backward_v2", &ROIAlignBackwardV2Laucher, "reference RoIAlign v2 backward (CUDA)");
}
