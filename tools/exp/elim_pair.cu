// Micro-benchmark: fused right-looking elimination (16x16 SPD block + rhs + 12 sub-diagonal rows), the core of
// k_factor_solve, in two mappings:  A) 1 warp = 1 problem, one row per lane (the shipped mapping);
// B) 1 half-warp = 1 problem, two rows per lane (L row + rhs/E row), 2 problems per warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o elim_pair elim_pair.cu && ./elim_pair
#include <cstdio>
#include <cuda_runtime.h>
constexpr unsigned FULL = 0xffffffffu;
constexpr int NN = 16, NX = 12;
__device__ __forceinline__ double rsqrt_(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double r = fma(-x * y, y, 1.0);
  const double p = fma(0.375, r, 0.5);
  return fma(y * r, p, y);
}
struct alignas(16) V2 { double x, y; };

// ---- A: row per lane
__device__ __forceinline__ void elimA(double (&a)[NN], int lane, double* colbuf) {
  double d = __shfl_sync(FULL, a[0], 0);
  double inv = rsqrt_(d);
#pragma unroll
  for (int k = 0; k < NN; ++k) {
    const double lk = a[k] * inv;
    a[k] = (lane == k) ? inv : lk;
    if (k + 1 < NN) {
      double* col = colbuf + (k & 1) * 32;
      if (lane < NN) col[lane] = lk;
      __syncwarp();
      a[k + 1] -= lk * col[k + 1];
      d = __shfl_sync(FULL, a[k + 1], k + 1);
      inv = rsqrt_(d);
      const int j0 = k + 2, js = (j0 + 1) & ~1;
      if ((j0 & 1) && j0 < NN) a[j0] -= lk * col[j0];
#pragma unroll
      for (int j = 0; j < NN; j += 2)
        if (j >= js && j + 1 < NN) {
          const V2 v = *reinterpret_cast<const V2*>(col + j);
          a[j] -= lk * v.x;
          a[j + 1] -= lk * v.y;
        }
    }
  }
}
__device__ __forceinline__ double gen(int prob, int r, int c) {  // SPD-ish synthetic entry
  return (r == c ? 40.0 : 0.0) + 0.01 * ((prob * 7 + r * 13 + c * 29) % 17);
}
__global__ void __launch_bounds__(128, 6) kA(int nprob, int reps, double* out) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* colbuf = sm + warp * 64;
  double acc = 0;
  for (int p = blockIdx.x * 4 + warp; p < nprob; p += gridDim.x * 4) {
    for (int rep = 0; rep < reps; ++rep) {
      double a[NN];
      const int r = lane < NN ? lane : lane - NN;
#pragma unroll
      for (int c = 0; c < NN; ++c) a[c] = (lane < NN) ? ((c <= lane) ? gen(p + rep, lane, c) : 0.0) : 0.1 * gen(p + rep, r + 3, c + 1);
      elimA(a, lane, colbuf);
#pragma unroll
      for (int c = 0; c < NN; ++c) acc += a[c];
      __syncwarp();
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// ---- B: half-warp per problem, two rows per lane
__device__ __forceinline__ void elimB(double (&a)[NN], double (&b)[NN], int hl, int half, double* colbuf) {
  // a: L row hl (16 lanes), b: rhs (hl == 0) / E row hl-1 (hl = 1..12); colbuf: [2 bufs][2 halves][16]
  double d = __shfl_sync(FULL, a[0], half * 16);
  double inv = rsqrt_(d);
#pragma unroll
  for (int k = 0; k < NN; ++k) {
    const double lk = a[k] * inv, mk = b[k] * inv;
    a[k] = (hl == k) ? inv : lk;
    b[k] = mk;
    if (k + 1 < NN) {
      double* col = colbuf + (k & 1) * 32 + half * 16;
      col[hl] = lk;
      __syncwarp();
      const double c1 = col[k + 1];
      a[k + 1] -= lk * c1;
      b[k + 1] -= mk * c1;
      d = __shfl_sync(FULL, a[k + 1], half * 16 + k + 1);
      inv = rsqrt_(d);
      const int j0 = k + 2, js = (j0 + 1) & ~1;
      if ((j0 & 1) && j0 < NN) { const double c = col[j0]; a[j0] -= lk * c; b[j0] -= mk * c; }
#pragma unroll
      for (int j = 0; j < NN; j += 2)
        if (j >= js && j + 1 < NN) {
          const V2 v = *reinterpret_cast<const V2*>(col + j);
          a[j] -= lk * v.x;
          a[j + 1] -= lk * v.y;
          b[j] -= mk * v.x;
          b[j + 1] -= mk * v.y;
        }
    }
  }
}
template <int MINB>
__global__ void __launch_bounds__(128, MINB) kB(int nprob, int reps, double* out) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
  double* colbuf = sm + warp * 64;
  double acc = 0;
  for (int p0 = (blockIdx.x * 4 + warp) * 2; p0 < nprob; p0 += gridDim.x * 8) {
    const int p = p0 + half;
    for (int rep = 0; rep < reps; ++rep) {
      double a[NN], b[NN];
#pragma unroll
      for (int c = 0; c < NN; ++c) {
        a[c] = (c <= hl) ? gen(p + rep, hl, c) : 0.0;
        b[c] = (hl <= NX) ? 0.1 * gen(p + rep, hl + 2, c + 1) : 0.0;
      }
      elimB(a, b, hl, half, colbuf);
#pragma unroll
      for (int c = 0; c < NN; ++c) acc += a[c] + b[c];
      __syncwarp();
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class K>
float run(K kern, int grid, size_t smem, int nprob, int reps, double* out) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { printf("attr: %s\n", cudaGetErrorString(e)); return -1.f; }
  kern<<<grid, 128, smem>>>(nprob, 1, out);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("warm-up: %s\n", cudaGetErrorString(e)); return -1.f; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kern<<<grid, 128, smem>>>(nprob, reps, out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("   %s\n", cudaGetErrorString(cudaGetLastError()));
  return ms;
}
int main() {
  const int nprob = 65536, reps = 5;  // 5 eliminations per problem = the 5 stages of a T=5 solve
  double* out;
  if (cudaMalloc(&out, sizeof(double) * (size_t)(nprob / 4) * 128) != cudaSuccess) { printf("malloc failed\n"); return 1; }
  const int gridA = nprob / 4, gridB = nprob / 8;
  // shared memory sized like the real kernel: 30.8 KB / CTA (A, 6 CTAs/SM) and 61.6 KB / CTA (B, 3 CTAs/SM)
  float a = run(kA, gridA, 30848, nprob, reps, out);
  printf("A  row/lane, 1 problem/warp, 6 CTAs/SM : %.3f ms for %d x %d eliminations\n", a, nprob, reps);
  float b3 = run(kB<3>, gridB, 61696, nprob, reps, out);
  printf("B  2 rows/lane, 2 problems/warp, 3 CTAs/SM: %.3f ms\n", b3);
  float b4 = run(kB<4>, gridB, 50000, nprob, reps, out);
  printf("B  2 rows/lane, 2 problems/warp, 4 CTAs/SM: %.3f ms\n", b4);
  return 0;
}
