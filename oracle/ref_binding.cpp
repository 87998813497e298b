// oracle/ref_binding.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// pybind11 module exposing the REFERENCE's own deformable-convolution entry points under the names its
// `detectron2._C` extension gives them (detectron2/detectron2/layers/csrc/vision.cpp:76-92).  The reference's
// sources are compiled where they lie under /root/reference by oracle/build_ref.py; this file only includes the
// reference's header (deformable/deform_conv.h, found through -I) and binds its five inline dispatchers.  Nothing
// of the reference is copied here.
#include <torch/extension.h>

#include "deformable/deform_conv.h"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("deform_conv_forward", &detectron2::deform_conv_forward, "deform_conv_forward");
  m.def("deform_conv_backward_input", &detectron2::deform_conv_backward_input, "deform_conv_backward_input");
  m.def("deform_conv_backward_filter", &detectron2::deform_conv_backward_filter, "deform_conv_backward_filter");
  m.def("modulated_deform_conv_forward", &detectron2::modulated_deform_conv_forward, "modulated_deform_conv_forward");
  m.def("modulated_deform_conv_backward", &detectron2::modulated_deform_conv_backward, "modulated_deform_conv_backward");
}
