// The shade kernels are instantiated in separate translation units (shade_tu.cu compiled with -DSG_TU=1..7) so that the
// library builds in parallel; shimmer_gpu.cu (the C ABI, the traversal kernels and the small kernels) reaches them through
// these getters.  Every TU includes the same headers; functions defined there have internal linkage.
#pragma once
#include "sg_wavefront.cuh"

namespace sg {

typedef void (*ShadeKernel)(const DScene, PathState, Queues, RenderConst, int);

ShadeKernel shade_kernel_lean(int kind);        // k_shade<KIND, false, true, false>: untextured scenes lit by triangle emitters          (SG_TU 1)
ShadeKernel shade_kernel_textured_a(int kind);  // k_shade<KIND, true, true, false>: image textures, triangle emitters; first kind group  (SG_TU 6)
ShadeKernel shade_kernel_textured_b(int kind);  //                                                                      second kind group (SG_TU 7)
ShadeKernel shade_kernel_general_a(int kind);   // k_shade<KIND, true, true, true>: + sphere / patch / point / image-infinite lights; KIND in {Diffuse, Conductor, Dielectric, Thin} (SG_TU 2)
ShadeKernel shade_kernel_general_b(int kind);   //                                                          KIND in {CoatedDiffuse, CoatedConductor}                           (SG_TU 3)
ShadeKernel shade_kernel_other_a(int kind);     // k_shade<KIND, true, false> (SimplePath / RandomWalk), first kind group     (SG_TU 4)
ShadeKernel shade_kernel_other_b(int kind);     //                                                         second kind group   (SG_TU 5)
ShadeKernel shade_kernel_force_diffuse_a(int kind);   // k_shade<KIND, true, true, true, FD = true>: Options::force_diffuse (path integrator)  (SG_TU 8)
ShadeKernel shade_kernel_force_diffuse_b(int kind);   //                                                                                  (SG_TU 9)
ShadeKernel resolve_mix_kernel(bool tex);       // k_resolve_mix<TEX>                                                          (SG_TU 2)
ShadeKernel shade_kernel_stage1(int kind);      // k_shade<Diffuse, TEX, .., STAGE 1>: get_bsdf only -> BSDF record; nullptr for other kinds (SG_TU 10)
ShadeKernel shade_kernel_stage2(int kind, bool general_lights);   // k_shade<Diffuse, no TEX, .., STAGE 2>: the rest, from the record     (SG_TU 10)

inline bool shade_kind_in_group_b(int kind) { return kind == SG_MATERIAL_COATED_DIFFUSE || kind == SG_MATERIAL_COATED_CONDUCTOR; }
inline ShadeKernel shade_kernel(int kind, bool textured, bool general_lights, bool path_integrator, bool force_diffuse = false) {
    if (force_diffuse) return shade_kind_in_group_b(kind) ? shade_kernel_force_diffuse_b(kind) : shade_kernel_force_diffuse_a(kind);
    if (!path_integrator) return shade_kind_in_group_b(kind) ? shade_kernel_other_b(kind) : shade_kernel_other_a(kind);
    if (general_lights) return shade_kind_in_group_b(kind) ? shade_kernel_general_b(kind) : shade_kernel_general_a(kind);
    if (textured) return shade_kind_in_group_b(kind) ? shade_kernel_textured_b(kind) : shade_kernel_textured_a(kind);
    return shade_kernel_lean(kind);
}

}  // namespace sg
