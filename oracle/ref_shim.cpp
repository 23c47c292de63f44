// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// pybind11 shim that exposes the UNMODIFIED reference rasterizer entry points
// (compiled from the sources where they lie under /root/reference by
// oracle/build_ref.py) so that tests/ and bench.py's cpu_baseline /
// `--impl reference` legs can run the real reference next to the C
// restatement in oracle/raster_oracle.c.
//
// Nothing here is copied from the reference: the declarations come from
// including the reference's own dispatch headers
//   pytorch3d/csrc/rasterize_meshes/rasterize_meshes.h   (RasterizeMeshes :513, RasterizeMeshesBackward :211)
//   pytorch3d/csrc/rasterize_points/rasterize_points.h   (RasterizePoints :343, RasterizePointsBackward :281)
//   pytorch3d/csrc/rasterize_coarse/rasterize_coarse.h
// and the registration mirrors pytorch3d/csrc/ext.cpp:53-56,69-73.
#include <torch/extension.h>
#include "rasterize_meshes/rasterize_meshes.h"
#include "rasterize_points/rasterize_points.h"
#include "interp_face_attrs/interp_face_attrs.h"  // CUDA only in the reference (ext.cpp:49-50)
#include "compositing/alpha_composite.h"  // alphaCompositeForward :59, alphaCompositeBackward :84 (ext.cpp:75-76)
#include "compositing/norm_weighted_sum.h"  // weightedSumNormForward / Backward (ext.cpp:77-78)
#include "compositing/weighted_sum.h"  // weightedSumForward / Backward (ext.cpp:79-80)

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("rasterize_meshes", &RasterizeMeshes);
  m.def("rasterize_meshes_backward", &RasterizeMeshesBackward);
  m.def("rasterize_points", &RasterizePoints);
  m.def("rasterize_points_backward", &RasterizePointsBackward);
  m.def("_rasterize_meshes_coarse", &RasterizeMeshesCoarse);
  m.def("_rasterize_meshes_naive", &RasterizeMeshesNaive);
  m.def("_rasterize_meshes_fine", &RasterizeMeshesFine);
  m.def("_rasterize_points_coarse", &RasterizePointsCoarse);
  m.def("_rasterize_points_naive", &RasterizePointsNaive);
  m.def("accum_alphacomposite", &alphaCompositeForward);
  m.def("accum_alphacomposite_backward", &alphaCompositeBackward);
  m.def("accum_weightedsumnorm", &weightedSumNormForward);
  m.def("accum_weightedsumnorm_backward", &weightedSumNormBackward);
  m.def("accum_weightedsum", &weightedSumForward);
  m.def("accum_weightedsum_backward", &weightedSumBackward);
#ifdef WITH_CUDA
  m.def("interp_face_attrs_forward", &InterpFaceAttrsForward);
  m.def("interp_face_attrs_backward", &InterpFaceAttrsBackward);
  m.attr("with_cuda") = true;
#else
  m.attr("with_cuda") = false;
#endif
}
