// Shade-kernel instantiations, one group per translation unit (see sg_kernels.h): nvcc -DSG_TU=1..9 -c shade_tu.cu
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
#else
#error "compile with -DSG_TU=1..9"
#endif

}  // namespace sg
