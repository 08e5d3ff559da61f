// Issue-rate microbenchmark: FFMA (3 registers) vs FFMA2 (fma.rn.f32x2) vs HADD2.F32 + FFMA mixes, 4 warps per SM sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu ; prints cycles per warp-instruction per sub-partition.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, int iters, float seed) {
	float a[8], b[8];
	unsigned long long p[8];
	for (int i = 0; i < 8; ++i) a[i] = seed + i + threadIdx.x, b[i] = seed * 0.5f + i;
	for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[i]), "f"(b[i]));
	unsigned long long xm, ym;
	asm("mov.b64 %0, {%1, %2};" : "=l"(xm) : "f"(seed * 1.0001f), "f"(seed * 0.9999f));
	asm("mov.b64 %0, {%1, %2};" : "=l"(ym) : "f"(seed * 0.001f), "f"(seed * 0.002f));
	const float x = seed * 1.0001f, y = seed * 0.001f;
	__syncthreads();
	long long t0 = clock64();
	for (int it = 0; it < iters; ++it) {
		if (MODE == 0) {
#pragma unroll
			for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], x, y);
#pragma unroll
			for (int i = 0; i < 8; ++i) b[i] = fmaf(b[i], x, y);
		} else {
#pragma unroll
			for (int r = 0; r < 2; ++r)
#pragma unroll
				for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(xm), "l"(ym));
		}
	}
	long long t1 = clock64();
	float s = 0;
	for (int i = 0; i < 8; ++i) {
		float u, v;
		asm("mov.b64 {%0, %1}, %2;" : "=f"(u), "=f"(v) : "l"(p[i]));
		s += a[i] + b[i] + u + v;
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
	float* out;
	long long* cyc;
	cudaMalloc(&out, 148 * 512 * 4);
	cudaMallocManaged(&cyc, 148 * 8);
	const int iters = 4096;
	for (int mode = 0; mode < 2; ++mode) {
		for (int rep = 0; rep < 2; ++rep) {
			if (mode == 0) k<0><<<148, 512>>>(out, cyc, iters, 1.0f);
			else k<1><<<148, 512>>>(out, cyc, iters, 1.0f);
			cudaDeviceSynchronize();
		}
		// 16 warp-instructions per iteration per warp, 4 warps per sub-partition
		printf("{\"mode\": \"%s\", \"cycles_per_warp_instr_per_smsp\": %.3f, \"cuda_error\": \"%s\"}\n", mode ? "FFMA2 (2 fp32 FMAs per lane)" : "FFMA (3-register)",
		       (double)cyc[0] / ((double)iters * 16 * 4), cudaGetErrorString(cudaGetLastError()));
	}
	return 0;
}
