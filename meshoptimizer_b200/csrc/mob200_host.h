// mob200_host.h -- host-side state shared by the C-ABI translation units (mob200_api.cu, mob200_index.cu):
// the per-device context with its staging buffers, and small RAII-free buffer helpers.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <mutex>

#include "../../include/meshopt_b200.h"

#define CUDA_TRY(expr)                                                                                      \
	do                                                                                                      \
	{                                                                                                       \
		cudaError_t err__ = (expr);                                                                         \
		if (err__ != cudaSuccess)                                                                           \
		{                                                                                                   \
			fprintf(stderr, "meshopt_b200: %s failed: %s (%s:%d)\n", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
			return MOB200_ERR_CUDA;                                                                         \
		}                                                                                                   \
	} while (0)

struct DeviceBuffer
{
	void* ptr = nullptr;
	size_t size = 0;

	int reserve(size_t bytes)
	{
		if (bytes <= size)
			return 0;
		if (ptr)
			cudaFree(ptr);
		ptr = nullptr;
		size = 0;
		size_t want = bytes + bytes / 4 + 256;
		CUDA_TRY(cudaMalloc(&ptr, want));
		size = want;
		return 0;
	}
	void release()
	{
		if (ptr)
			cudaFree(ptr);
		ptr = nullptr;
		size = 0;
	}
};

// stream-ordered scratch that is given back on every path out of a function (error returns included)
struct AsyncScratch
{
	void* ptr = nullptr;
	cudaStream_t stream = nullptr;
	~AsyncScratch()
	{
		if (ptr)
			cudaFreeAsync(ptr, stream);
	}
	int alloc(size_t bytes, cudaStream_t st)
	{
		stream = st;
		CUDA_TRY(cudaMallocAsync(&ptr, bytes, st));
		return 0;
	}
};

struct PinnedBuffer
{
	void* ptr = nullptr;
	size_t size = 0;

	int reserve(size_t bytes)
	{
		if (bytes <= size)
			return 0;
		if (ptr)
			cudaFreeHost(ptr);
		ptr = nullptr;
		size = 0;
		size_t want = bytes + bytes / 4 + 256;
		CUDA_TRY(cudaHostAlloc(&ptr, want, cudaHostAllocDefault));
		size = want;
		return 0;
	}
	void release()
	{
		if (ptr)
			cudaFreeHost(ptr);
		ptr = nullptr;
		size = 0;
	}
};

struct mob200_Context
{
	int device = 0;
	int sm_count = 0;
	int decode_ctas_per_sm = 1;
	uint32_t walker_lead = 0;      // see DevTables::walker_lead
	int rounds_mode = 2;           // 0 / 1 force the decoder form, 2 = rounds when most blocks have <= 16-byte vertices (MOB200_ROUNDS)
	int wide_walk_mode = 2;        // 0 / 1 force the walker form, 2 = choose by stream count (MOB200_WIDE_WALK)
	int run_major = 1;             // block mode + rounds: decode order in runs of four consecutive blocks of a stream (MOB200_RUN_MAJOR=0: level-major)
	int team_walk = 1;             // few-stream plans: offsets-only team walk + block-mode decode (two launches) instead of the fused
	                               // one-warp-per-stream walker (MOB200_TEAM_WALK=0 restores that)
	cudaStream_t stream = nullptr; // used by the host-pointer entry points
	std::mutex mu;                 // host-pointer entry points share the staging buffers below
	DeviceBuffer d_in, d_out;
	PinnedBuffer h_in, h_out;
	// mob200_decode_batch_host: chunks in flight
	static const int kHostSlots = 3;
	cudaStream_t slot_stream[kHostSlots] = {};
	cudaEvent_t slot_done[kHostSlots] = {};
	DeviceBuffer slot_arena[kHostSlots];
	PinnedBuffer slot_h_in[kHostSlots], slot_h_out[kHostSlots];
	PinnedBuffer h_status;
};

static inline int set_device(const mob200_Context* ctx)
{
	CUDA_TRY(cudaSetDevice(ctx->device));
	return 0;
}

// Contexts of the drop-in symbols (host pointers, synchronous, callable from many threads at once -- reference
// contract src/meshoptimizer.h, SURVEY.md section 8b "Threading"): a small pool per device, one context per call in
// flight, so that concurrent callers overlap their copies (kernels are chained per device, mob200_api.cu).
// Hidden visibility: shared by the translation units of the library only.
mob200_Context* mob200_pool_acquire();
void mob200_pool_release(mob200_Context* ctx);

struct PoolLease
{
	mob200_Context* ctx;
	PoolLease() : ctx(mob200_pool_acquire()) {}
	~PoolLease()
	{
		if (ctx)
			mob200_pool_release(ctx);
	}
	PoolLease(const PoolLease&) = delete;
	PoolLease& operator=(const PoolLease&) = delete;
};
