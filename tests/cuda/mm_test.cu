// Standalone check of the tcgen05 int8 tile kernel of hu_mm.cuh against a host GEMM (test infrastructure):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I hmmufotu_b200/csrc tests/cuda/mm_test.cu -o mm_test && ./mm_test
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "hu_mm.cuh"

using namespace hu;

struct RawEpi {
	int* out; int mTot;
	static constexpr int SMEM_BYTES = 0;
	__device__ void init(void*, int, int) {}
	__device__ void prefetch(int, int, int) {}
	__device__ void begin(int, int, int) {}
	__device__ void operator()(int mt, int nt, int row, int col0, const uint32_t (&v)[32], int) {
#pragma unroll
		for (int j = 0; j < 32; ++j) out[(size_t) (nt * HU_MM_BN + col0 + j) * mTot + mt * HU_MM_BM + row] = (int) v[j];
	}
};

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

int main(int argc, char** argv) {
	const int numMt = argc > 1 ? atoi(argv[1]) : 3, numNt = argc > 2 ? atoi(argv[2]) : 2, nChunks = argc > 3 ? atoi(argv[3]) : 7;
	const int mTot = numMt * HU_MM_BM, nTot = numNt * HU_MM_BN, kTot = nChunks * HU_MM_BK;
	std::vector<int8_t> a((size_t) mTot * kTot), b((size_t) nTot * kTot);
	srand(1);
	for (auto& x : a) x = (int8_t) (rand() % 256 - 128);
	for (auto& x : b) x = (int8_t) (rand() % 256 - 128);
	std::vector<int8_t> ai((size_t) numMt * nChunks * HU_MM_A_BYTES), bi((size_t) numNt * nChunks * HU_MM_B_BYTES);
	for (int m = 0; m < mTot; ++m)
		for (int k = 0; k < kTot; ++k)
			ai[((size_t) (m / HU_MM_BM) * nChunks + k / HU_MM_BK) * HU_MM_A_BYTES + mmTileOffset(m % HU_MM_BM, k % HU_MM_BK)] = a[(size_t) m * kTot + k];
	for (int n = 0; n < nTot; ++n)
		for (int k = 0; k < kTot; ++k)
			bi[((size_t) (n / HU_MM_BN) * nChunks + k / HU_MM_BK) * HU_MM_B_BYTES + mmTileOffset(n % HU_MM_BN, k % HU_MM_BK)] = b[(size_t) n * kTot + k];
	std::vector<int2> range(numNt);
	for (int t = 0; t < numNt; ++t) range[t] = make_int2(t % 2 ? 0 : (nChunks > 2 ? 1 : 0), t % 2 ? nChunks - 1 : (nChunks > 2 ? nChunks - 2 : nChunks - 1));
	if (numNt > 2) range[2] = make_int2(3, 2); // an empty tile
	int8_t *dA, *dB; int2* dR; int* dO;
	CK(cudaMalloc(&dA, ai.size())); CK(cudaMalloc(&dB, bi.size())); CK(cudaMalloc(&dR, range.size() * sizeof(int2)));
	CK(cudaMalloc(&dO, (size_t) mTot * nTot * sizeof(int)));
	CK(cudaMemcpy(dA, ai.data(), ai.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, bi.data(), bi.size(), cudaMemcpyHostToDevice));
	CK(cudaMemcpy(dR, range.data(), range.size() * sizeof(int2), cudaMemcpyHostToDevice));
	CK(cudaMemset(dO, 0x7f, (size_t) mTot * nTot * sizeof(int)));
	CK(cudaFuncSetAttribute(seedMmKernel<RawEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize, HU_MM_SMEM));
	MmArgs g = {dA, dB, dR, numMt, numNt, nChunks, NULL, -1};
	RawEpi epi = {dO, mTot};
	int sms = 0;
	CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
	const int grid = std::min(sms, numMt * numNt);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	seedMmKernel<RawEpi><<<grid, HU_MM_THREADS, HU_MM_SMEM>>>(g, epi);
	CK(cudaGetLastError());
	CK(cudaDeviceSynchronize());
	cudaEventRecord(e0);
	seedMmKernel<RawEpi><<<grid, HU_MM_THREADS, HU_MM_SMEM>>>(g, epi);
	cudaEventRecord(e1);
	CK(cudaDeviceSynchronize());
	float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
	std::vector<int> out((size_t) mTot * nTot);
	CK(cudaMemcpy(out.data(), dO, out.size() * sizeof(int), cudaMemcpyDeviceToHost));
	long bad = 0, checked = 0;
	const int stepM = mTot > 1024 ? 37 : 1, stepN = nTot > 1024 ? 29 : 1;
	for (int n = 0; n < nTot; n += stepN) {
		const int2 rg = range[n / HU_MM_BN];
		for (int m = 0; m < mTot; m += stepM) {
			const int got = out[(size_t) n * mTot + m];
			if (rg.x > rg.y) { if (got != 0x7f7f7f7f) { if (bad < 10) printf("empty tile written at m=%d n=%d\n", m, n); bad++; } continue; }
			int want = 0;
			for (int k = rg.x * HU_MM_BK; k < (rg.y + 1) * HU_MM_BK; ++k) want += (int) a[(size_t) m * kTot + k] * (int) b[(size_t) n * kTot + k];
			checked++;
			if (got != want) { if (bad < 10) printf("mismatch m=%d n=%d want %d got %d\n", m, n, want, got); bad++; }
		}
	}
	const double ops = 0; (void) ops;
	double macs = 0;
	for (int t = 0; t < numNt; ++t) if (range[t].x <= range[t].y) macs += (double) mTot * HU_MM_BN * (range[t].y - range[t].x + 1) * HU_MM_BK;
	printf("mm_test %dx%dx%d: checked %ld, mismatched %ld, %.3f ms, %.1f TOPS\n", mTot, nTot, kTot, checked, bad, ms, 2 * macs / ms / 1e9);
	return bad ? 1 : 0;
}
