// Warp-level bf16 tensor-core helpers (ldmatrix + mma.sync.m16n8k16) for the small per-item products of the
// prologue / gradient finalisation (L x L x d with L <= 128): far too small for a tcgen05 pipeline, far too
// slow on the FMA pipe.
#pragma once
#include <stdint.h>

namespace dvlp {
namespace mm {

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
// D (16x8 fp32) += A (16x16 bf16, row) * B (16x8 bf16, col)
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                          uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// A fragment of the 16x16 block at (row0, k0) of a row-major bf16 matrix in shared memory (ld = row stride, bytes)
__device__ __forceinline__ void load_a(uint32_t base, int ld, int row0, int k0, int lane, uint32_t (&a)[4]) {
  const int r = row0 + (lane & 7) + ((lane >> 3) & 1) * 8, c = k0 + (lane >> 4) * 8;
  ldmatrix_x4(base + r * ld + c * 2, a[0], a[1], a[2], a[3]);
}
// B fragments of two adjacent n8 tiles where B[k][n] = X[n][k] (X row-major, rows n0..n0+15, columns k0..k0+15):
// b[0], b[1] = tile n0..n0+7, b[2], b[3] = tile n0+8..n0+15
__device__ __forceinline__ void load_b_nk(uint32_t base, int ld, int n0, int k0, int lane, uint32_t (&b)[4]) {
  const int r = n0 + (lane & 7) + (lane >> 4) * 8, c = k0 + ((lane >> 3) & 1) * 8;
  ldmatrix_x4(base + r * ld + c * 2, b[0], b[1], b[2], b[3]);
}
// B fragments of two adjacent n8 tiles where B[k][n] = X[k][n] (X row-major, rows k0..k0+15, columns n0..n0+15)
__device__ __forceinline__ void load_b_kn(uint32_t base, int ld, int k0, int n0, int lane, uint32_t (&b)[4]) {
  const int r = k0 + (lane & 7) + ((lane >> 3) & 1) * 8, c = n0 + (lane >> 4) * 8;
  ldmatrix_x4_trans(base + r * ld + c * 2, b[0], b[1], b[2], b[3]);
}

}  // namespace mm
}  // namespace dvlp
