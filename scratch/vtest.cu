#include <cstdint>
__global__ void k_vcmpeq4(const uint32_t* a, uint32_t b, uint32_t* o){ o[threadIdx.x] = __vcmpeq4(a[threadIdx.x], b); }
__global__ void k_vseteq4(const uint32_t* a, uint32_t b, uint32_t* o){ o[threadIdx.x] = __vseteq4(a[threadIdx.x], b); }
__global__ void k_vminu4(const uint32_t* a, uint32_t b, uint32_t* o){ o[threadIdx.x] = __vminu4(a[threadIdx.x]^b, 0x01010101u); }
__global__ void k_vcmpne4(const uint32_t* a, uint32_t b, uint32_t* o){ o[threadIdx.x] = __vcmpne4(a[threadIdx.x], b); }
__global__ void k_manual(const uint32_t* a, uint32_t b, uint32_t* o){ uint32_t x = a[threadIdx.x]^b; uint32_t t = (x & 0x7f7f7f7fu) + 0x7f7f7f7fu; o[threadIdx.x] = ~(t|x) & 0x80808080u; }
__global__ void k_vabsdiff(const uint32_t* a, uint32_t b, uint32_t* o){ o[threadIdx.x] = __vabsdiffu4(a[threadIdx.x], b); }
__global__ void k_vsetne2(const uint32_t* a, uint32_t b, uint32_t* o){ o[threadIdx.x] = __vcmpeq2(a[threadIdx.x], b); }
__global__ void k_vimin3(const uint32_t* a, uint32_t b, uint32_t* o){ o[threadIdx.x] = __vimin3_u16x2(a[threadIdx.x], b, o[threadIdx.x]); }
