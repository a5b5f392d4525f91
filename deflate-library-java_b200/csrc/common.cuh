// common.cuh -- shared device helpers for the b2deflate kernels (sm_100a).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b2d {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

constexpr unsigned FULL_MASK = 0xFFFFFFFFu;

__device__ __forceinline__ u32 lane_id() {
	u32 l;
	asm("mov.u32 %0, %%laneid;" : "=r"(l));
	return l;
}
__device__ __forceinline__ u32 lanemask_lt() {
	u32 m;
	asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
	return m;
}
// bit-field extract: (v >> pos) & ((1 << len) - 1), len may be 0
__device__ __forceinline__ u32 bfe(u32 v, u32 pos, u32 len) {
	u32 r;
	asm("bfe.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(v), "r"(pos), "r"(len));
	return r;
}

// DEFLATE symbol tables (RFC 1951 3.2.5; Open.java:843-886 builds the same values)
__device__ __forceinline__ void length_sym_info(int sym, int &base, int &eb) {   // sym 257..285
	if (sym <= 264) { eb = 0; base = sym - 254; }
	else if (sym <= 284) { eb = (sym - 261) >> 2; base = ((((sym - 1) & 3) + 4) << eb) + 3; }
	else { eb = 0; base = 258; }
}
__device__ __forceinline__ void dist_sym_info(int sym, int &base, int &eb) {     // sym 0..29
	if (sym <= 3) { eb = 0; base = sym + 1; }
	else { eb = (sym >> 1) - 1; base = (((sym & 1) + 2) << eb) + 1; }
}

}  // namespace b2d
