// Shade-kernel instantiations, one group per translation unit (see sg_kernels.h): nvcc -DSG_TU=1..10 -c shade_tu.cu
#include "sg_kernels.h"

namespace sg {

#define SG_CASE_A(TEX, PATH, LG) \
    case SG_MATERIAL_DIFFUSE: return k_shade<SG_MATERIAL_DIFFUSE, TEX, PATH, LG>; \
    case SG_MATERIAL_CONDUCTOR: return k_shade<SG_MATERIAL_CONDUCTOR, TEX, PATH, LG>; \
    case SG_MATERIAL_DIELECTRIC: return k_shade<SG_MATERIAL_DIELECTRIC, TEX, PATH, LG>; \
    case SG_MATERIAL_THIN_DIELECTRIC: return k_shade<SG_MATERIAL_THIN_DIELECTRIC, TEX, PATH, LG>;
#define SG_CASE_B(TEX, PATH, LG) \
    case SG_MATERIAL_COATED_DIFFUSE: return k_shade<SG_MATERIAL_COATED_DIFFUSE, TEX, PATH, LG>; \
    case SG_MATERIAL_COATED_CONDUCTOR: return k_shade<SG_MATERIAL_COATED_CONDUCTOR, TEX, PATH, LG>;

#if SG_TU == 1
ShadeKernel shade_kernel_lean(int kind) { switch (kind) { SG_CASE_A(false, true, false) SG_CASE_B(false, true, false) } return nullptr; }
#elif SG_TU == 2
ShadeKernel shade_kernel_general_a(int kind) { switch (kind) { SG_CASE_A(true, true, true) } return nullptr; }
ShadeKernel resolve_mix_kernel(bool tex) { return tex ? k_resolve_mix<true> : k_resolve_mix<false>; }
#elif SG_TU == 3
ShadeKernel shade_kernel_general_b(int kind) { switch (kind) { SG_CASE_B(true, true, true) } return nullptr; }
#elif SG_TU == 4
ShadeKernel shade_kernel_other_a(int kind) { switch (kind) { SG_CASE_A(true, false, true) } return nullptr; }
#elif SG_TU == 5
ShadeKernel shade_kernel_other_b(int kind) { switch (kind) { SG_CASE_B(true, false, true) } return nullptr; }
#elif SG_TU == 6
ShadeKernel shade_kernel_textured_a(int kind) { switch (kind) { SG_CASE_A(true, true, false) } return nullptr; }
#elif SG_TU == 7
ShadeKernel shade_kernel_textured_b(int kind) { switch (kind) { SG_CASE_B(true, true, false) } return nullptr; }
#elif SG_TU == 8
ShadeKernel shade_kernel_force_diffuse_a(int kind) {
    switch (kind) {
    case SG_MATERIAL_DIFFUSE: return k_shade<SG_MATERIAL_DIFFUSE, true, true, true, true>;
    case SG_MATERIAL_CONDUCTOR: return k_shade<SG_MATERIAL_CONDUCTOR, true, true, true, true>;
    case SG_MATERIAL_DIELECTRIC: return k_shade<SG_MATERIAL_DIELECTRIC, true, true, true, true>;
    case SG_MATERIAL_THIN_DIELECTRIC: return k_shade<SG_MATERIAL_THIN_DIELECTRIC, true, true, true, true>;
    }
    return nullptr;
}
#elif SG_TU == 9
ShadeKernel shade_kernel_force_diffuse_b(int kind) {
    switch (kind) {
    case SG_MATERIAL_COATED_DIFFUSE: return k_shade<SG_MATERIAL_COATED_DIFFUSE, true, true, true, true>;
    case SG_MATERIAL_COATED_CONDUCTOR: return k_shade<SG_MATERIAL_COATED_CONDUCTOR, true, true, true, true>;
    }
    return nullptr;
}
#elif SG_TU == 10
// staged shading of textured scenes (k_shade STAGE 1 / 2; Diffuse materials, path integrator): stage 1 carries the texture code
// (LG does not matter there: no light code), stage 2 is compiled without any (TEX = false) for both light-kind variants
ShadeKernel shade_kernel_stage1(int kind) { return kind == SG_MATERIAL_DIFFUSE ? k_shade<SG_MATERIAL_DIFFUSE, true, true, false, false, 1> : nullptr; }
ShadeKernel shade_kernel_stage2(int kind, bool general_lights) {
    if (kind != SG_MATERIAL_DIFFUSE) return nullptr;
    return general_lights ? k_shade<SG_MATERIAL_DIFFUSE, false, true, true, false, 2> : k_shade<SG_MATERIAL_DIFFUSE, false, true, false, false, 2>;
}
#else
#error "compile with -DSG_TU=1..10"
#endif

}  // namespace sg
