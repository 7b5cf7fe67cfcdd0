// Throughput microbenchmark of the candidate inner-loop instructions on sm_100a (ops per clock per SM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define ITER 4096
template <int OP>
__global__ void k(float* out, uint32_t a0, uint32_t b0, long long* cyc) {
  uint32_t a = a0 + threadIdx.x, b = b0 ^ threadIdx.x;
  float f0 = threadIdx.x, f1 = 1.f, f2 = 2.f, f3 = 3.f, f4 = 4, f5 = 5, f6 = 6, f7 = 7;
  uint32_t u0 = a, u1 = b, u2 = a ^ b, u3 = a + b, u4 = a*3, u5 = b*5, u6=a*7, u7=b*9;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITER; ++i) {
    if (OP == 0) {  // FHFMA
      asm volatile("{ .reg .b16 l, h; mov.b32 {l,h}, %8;\n"
          "fma.rn.f32.f16 %0, l, h, %0; fma.rn.f32.f16 %1, h, l, %1; fma.rn.f32.f16 %2, l, l, %2; fma.rn.f32.f16 %3, h, h, %3;\n"
          "fma.rn.f32.f16 %4, l, h, %4; fma.rn.f32.f16 %5, h, l, %5; fma.rn.f32.f16 %6, l, l, %6; fma.rn.f32.f16 %7, h, h, %7; }"
          : "+f"(f0), "+f"(f1), "+f"(f2), "+f"(f3), "+f"(f4), "+f"(f5), "+f"(f6), "+f"(f7) : "r"(a));
    } else if (OP == 1) {  // FFMA
      asm volatile("fma.rn.f32 %0, %8, %9, %0; fma.rn.f32 %1, %8, %9, %1; fma.rn.f32 %2, %8, %9, %2; fma.rn.f32 %3, %8, %9, %3;\n"
          "fma.rn.f32 %4, %8, %9, %4; fma.rn.f32 %5, %8, %9, %5; fma.rn.f32 %6, %8, %9, %6; fma.rn.f32 %7, %8, %9, %7;"
          : "+f"(f0), "+f"(f1), "+f"(f2), "+f"(f3), "+f"(f4), "+f"(f5), "+f"(f6), "+f"(f7) : "f"(__uint_as_float(a)), "f"(__uint_as_float(b)));
    } else if (OP == 2) {  // HFMA2
      asm volatile("fma.rn.f16x2 %0, %8, %9, %0; fma.rn.f16x2 %1, %8, %9, %1; fma.rn.f16x2 %2, %8, %9, %2; fma.rn.f16x2 %3, %8, %9, %3;\n"
          "fma.rn.f16x2 %4, %8, %9, %4; fma.rn.f16x2 %5, %8, %9, %5; fma.rn.f16x2 %6, %8, %9, %6; fma.rn.f16x2 %7, %8, %9, %7;"
          : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7) : "r"(a), "r"(b));
    } else if (OP == 3) {  // LOP3
      asm volatile("lop3.b32 %0, %0, %8, %9, 0xEA; lop3.b32 %1, %1, %8, %9, 0xEA; lop3.b32 %2, %2, %8, %9, 0xEA; lop3.b32 %3, %3, %8, %9, 0xEA;\n"
          "lop3.b32 %4, %4, %8, %9, 0xEA; lop3.b32 %5, %5, %8, %9, 0xEA; lop3.b32 %6, %6, %8, %9, 0xEA; lop3.b32 %7, %7, %8, %9, 0xEA;"
          : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7) : "r"(a), "r"(b));
    } else if (OP == 4) {  // FFMA2 (f32x2)
      asm volatile("{ .reg .b64 x, y, c0, c1, c2, c3; mov.b64 x, {%8,%9}; mov.b64 y, {%9,%8};\n"
          "mov.b64 c0, {%0,%1}; mov.b64 c1, {%2,%3}; mov.b64 c2, {%4,%5}; mov.b64 c3, {%6,%7};\n"
          "fma.rn.f32x2 c0, x, y, c0; fma.rn.f32x2 c1, x, y, c1; fma.rn.f32x2 c2, x, y, c2; fma.rn.f32x2 c3, x, y, c3;\n"
          "fma.rn.f32x2 c0, x, y, c0; fma.rn.f32x2 c1, x, y, c1; fma.rn.f32x2 c2, x, y, c2; fma.rn.f32x2 c3, x, y, c3;\n"
          "mov.b64 {%0,%1}, c0; mov.b64 {%2,%3}, c1; mov.b64 {%4,%5}, c2; mov.b64 {%6,%7}, c3; }"
          : "+f"(f0), "+f"(f1), "+f"(f2), "+f"(f3), "+f"(f4), "+f"(f5), "+f"(f6), "+f"(f7) : "f"(__uint_as_float(a)), "f"(__uint_as_float(b)));
    } else if (OP == 5) {  // HADD2.F32-style cvt f16->f32 (8 cvts)
      asm volatile("{ .reg .b16 l, h; mov.b32 {l,h}, %8; .reg .f32 t;\n"
          "cvt.f32.f16 t, l; add.f32 %0, %0, t; cvt.f32.f16 t, h; add.f32 %1, %1, t; cvt.f32.f16 t, l; add.f32 %2, %2, t; cvt.f32.f16 t, h; add.f32 %3, %3, t;}"
          : "+f"(f0), "+f"(f1), "+f"(f2), "+f"(f3), "+f"(f4), "+f"(f5), "+f"(f6), "+f"(f7) : "r"(a));
    } else if (OP == 6) {  // PRMT
      asm volatile("prmt.b32 %0, %0, %8, 0x7150; prmt.b32 %1, %1, %8, 0x7150; prmt.b32 %2, %2, %8, 0x7150; prmt.b32 %3, %3, %8, 0x7150;\n"
          "prmt.b32 %4, %4, %8, 0x7352; prmt.b32 %5, %5, %8, 0x7352; prmt.b32 %6, %6, %8, 0x7352; prmt.b32 %7, %7, %8, 0x7352;"
          : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7) : "r"(a));
    } else if (OP == 7) {  // HMMA m16n8k16 f32 acc
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%8,%9,%8,%9}, {%9,%8}, {%0,%1,%2,%3};\n"
                   "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%4,%5,%6,%7}, {%8,%9,%8,%9}, {%9,%8}, {%4,%5,%6,%7};"
          : "+f"(f0), "+f"(f1), "+f"(f2), "+f"(f3), "+f"(f4), "+f"(f5), "+f"(f6), "+f"(f7) : "r"(a), "r"(b));
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7 + __uint_as_float(u0 ^ u1 ^ u2 ^ u3 ^ u4 ^ u5 ^ u6 ^ u7);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP> void run(const char* name, int ops_per_iter, int threads) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  k<OP><<<148, threads>>>(out, 0x3c003c00u, 0x3c003c00u, cyc);
  k<OP><<<148, threads>>>(out, 0x3c003c00u, 0x3c003c00u, cyc);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  double per_clk = (double)ITER * ops_per_iter * threads / (double)c;
  printf("%-28s threads=%4d  %8.1f thread-ops/clk/SM\n", name, threads, per_clk);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int th : {128, 256, 512, 1024}) {
    run<0>("FHFMA (fma.rn.f32.f16)", 8, th);
    run<1>("FFMA", 8, th);
    run<2>("HFMA2", 8, th);
    run<3>("LOP3", 8, th);
    run<4>("FFMA2 (fma.rn.f32x2)", 8, th);
    run<5>("cvt.f32.f16 + FADD (pairs)", 4, th);
    run<6>("PRMT", 8, th);
    run<7>("HMMA m16n8k16 (instr)", 2, th);
  }
  return 0;
}
