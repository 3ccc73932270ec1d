// sm_100a asynchronous-copy plumbing for the streaming SEP kernels: mbarrier objects in shared memory and
// cp.async.bulk (the 1-D form of the Tensor Memory Accelerator copy: global -> shared, completion counted in
// bytes on an mbarrier).  One producer lane issues whole message windows; consumer warps wait on the barrier
// phase and hand the stage back through a second barrier.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hu {

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbarInit(unsigned long long* bar, int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
// makes freshly initialised barriers visible to the asynchronous proxy (bulk copies complete on them)
__device__ __forceinline__ void mbarInitFence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbarArrive(unsigned long long* bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void mbarArriveExpectTx(unsigned long long* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbarTryWait(unsigned long long* bar, uint32_t parity) {
	uint32_t ok;
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
	             : "=r"(ok) : "r"(smemAddr(bar)), "r"(parity) : "memory");
	return ok != 0;
}
__device__ __forceinline__ void mbarWait(unsigned long long* bar, uint32_t parity) {
	while (!mbarTryWait(bar, parity)) {}
}
// the producer lane's wait: backs off between probes so that its spinning does not take issue slots from the consumers
__device__ __forceinline__ void mbarWaitRelaxed(unsigned long long* bar, uint32_t parity) {
	while (!mbarTryWait(bar, parity)) __nanosleep(400);
}

// global -> shared bulk copy; src, dst 16-byte aligned, bytes a multiple of 16; completes `bytes` on `bar`
__device__ __forceinline__ void bulkLoad(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(smemAddr(dst)), "l"(src), "r"(bytes), "r"(smemAddr(bar)) : "memory");
}

// asks for `bytes` (a multiple of 16) at src to be brought into the L2: no destination, no completion to wait for
__device__ __forceinline__ void bulkPrefetchL2(const void* src, uint32_t bytes) {
	asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// barrier among the first `threads` threads of the CTA (the consumer warps), id 1; the producer warp never joins
__device__ __forceinline__ void consumerSync(int threads) { asm volatile("bar.sync 1, %0;" ::"r"(threads) : "memory"); }

} // namespace hu
