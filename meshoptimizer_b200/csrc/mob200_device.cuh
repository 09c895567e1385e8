// mob200_device.cuh -- small device helpers shared by the walker and decoder roles (PTX wrappers, unaligned loads).
#pragma once

#include "mob200_kernels.h"

namespace mob200
{

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------

// cycle counter for the diagnostic counters (mob200_plan_debug_counters); compiled out unless the library is
// built with -DMOB200_DEBUG_COUNTERS (MOB200_DEBUG_COUNTERS=1 python -m meshoptimizer_b200.build)
__device__ __forceinline__ long long dbg_clock()
{
#ifdef MOB200_DEBUG_COUNTERS
	return clock64();
#else
	return 0;
#endif
}

// Event trace of one decode unit (diagnostics; only in a library built with -DMOB200_TRACE, tools/trace_unit.py): the
// producer lane 0 and lane 0 of every decoder warp append (globaltimer ns << 16 | event << 8 | block index & 255).
#ifdef MOB200_TRACE
#ifndef MOB200_TRACE_UNIT
#define MOB200_TRACE_UNIT 300
#endif
constexpr uint32_t kTraceWords = 16384; // u64 entries behind the counters; entry 0 = number of events
__device__ __forceinline__ void trace_event(const DevTables& T, uint32_t unit, uint32_t lane, uint32_t ev, uint32_t i)
{
	if (unit != MOB200_TRACE_UNIT || lane != 0)
		return;
	unsigned long long* tr = reinterpret_cast<unsigned long long*>(T.counters + 64);
	unsigned long long t;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
	const unsigned long long k = atomicAdd(tr, 1ull) + 1ull;
	if (k < kTraceWords)
		tr[k] = (t << 16) | ((unsigned long long)(ev & 255u) << 8) | (i & 255u);
}
#define MOB200_TRACE_EVENT(T, unit, lane, ev, i) trace_event(T, unit, lane, ev, i)
#else
#define MOB200_TRACE_EVENT(T, unit, lane, ev, i) ((void)0)
#endif

__device__ __forceinline__ uint32_t smem_addr(const void* p)
{
	return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
	uint32_t ok;
#ifdef MOB200_TRYWAIT_HINT
	// (variant: an explicit suspend-time hint in nanoseconds)
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
	             : "=r"(ok)
	             : "r"(smem_addr(bar)), "r"(parity), "r"((uint32_t)MOB200_TRYWAIT_HINT)
	             : "memory");
#else
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
	             : "=r"(ok)
	             : "r"(smem_addr(bar)), "r"(parity)
	             : "memory");
#endif
	return ok != 0;
}

// 1-D TMA bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_bulk(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_smem)),
	             "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
	             : "memory");
}

__device__ __forceinline__ void fence_proxy_async()
{
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// barrier among the decoder warps of one unit (named barrier 1 + unit index in the CTA)
__device__ __forceinline__ void decoder_sync(uint32_t bar_id)
{
	asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(kDecodeThreads) : "memory");
}

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p)
{
	unsigned long long v;
	asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v)
{
	asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p)
{
	unsigned long long v;
	asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v)
{
	asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p)
{
	uint32_t v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v)
{
	asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v)
{
	asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// per-block hand-over word of DevTables::block_ready
__device__ __forceinline__ uint32_t ready_word(uint32_t epoch, uint32_t version, bool decodable)
{
	return (epoch << 2) | ((version & 1u) << 1) | (decodable ? 1u : 0u);
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p)
{
	uint32_t v;
	asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

// 32 bits at an arbitrary byte offset of a 4-byte aligned shared-memory array
__device__ __forceinline__ uint32_t lds_u32_at(const uint8_t* base, uint32_t off)
{
	const uint32_t* w = reinterpret_cast<const uint32_t*>(base) + (off >> 2);
	return __funnelshift_r(w[0], w[1], (off & 3u) * 8u);
}

// up to 32 bits at an arbitrary global address, of which the first `needed` bytes matter: two aligned
// loads at most, and never a word that holds none of the needed bytes (memory safety on truncated input)
__device__ __forceinline__ uint32_t ldg_u32_at(const uint8_t* p, uint32_t needed)
{
	uintptr_t a = reinterpret_cast<uintptr_t>(p);
	const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
	uint32_t mis = (uint32_t)(a & 3u);
	uint32_t lo = __ldg(w);
	uint32_t hi = (mis + needed > 4u) ? __ldg(w + 1) : 0u;
	return __funnelshift_r(lo, hi, mis * 8u);
}

// 8 bytes at an arbitrary global address; reads the two aligned 8-byte words around it, i.e. up to 15
// bytes past p -- only for positions with >= 24 readable bytes ahead (the format guarantees that in
// front of every group of a well-formed stream; the walker checks it before taking this path)
__device__ __forceinline__ void ldg_u64_at(const uint8_t* p, uint32_t& w0, uint32_t& w1)
{
	uintptr_t a = reinterpret_cast<uintptr_t>(p);
	const uint2* q = reinterpret_cast<const uint2*>(a & ~uintptr_t(7));
	uint2 lo = __ldg(q);
	uint2 hi = __ldg(q + 1);
	bool up = (a & 4u) != 0;
	uint32_t x0 = up ? lo.y : lo.x;
	uint32_t x1 = up ? hi.x : lo.y;
	uint32_t x2 = up ? hi.y : hi.x;
	uint32_t sh = (uint32_t)(a & 3u) * 8u;
	w0 = __funnelshift_r(x0, x1, sh);
	w1 = __funnelshift_r(x1, x2, sh);
}

// index into the width table {0,1,2,4,8} from version, v1 channel control (0/1) and the 2-bit selector
// (v0 uses {0,2,4,8}; v1 uses a window of the table that starts at ctrl)
__device__ __forceinline__ uint32_t width_index(uint32_t version, uint32_t ctrl, uint32_t sel)
{
	return version ? ctrl + sel : (sel ? sel + 1 : 0);
}

// number of all-ones fields among the 16 fields of a group; idx 1/2/3 = 1/2/4-bit fields
__device__ __forceinline__ uint32_t count_sentinels(uint32_t idx, uint32_t w0, uint32_t w1)
{
	uint32_t a = w0 & (w0 >> 1);
	uint32_t b = w1 & (w1 >> 1);
	uint32_t a4 = a & (a >> 2);
	uint32_t b4 = b & (b >> 2);
	uint32_t m = idx == 1 ? (w0 & 0xffffu) : (idx == 2 ? (a & 0x55555555u) : (a4 & 0x11111111u));
	uint32_t m2 = idx == 3 ? (b4 & 0x11111111u) : 0u;
	return __popc(m) + __popc(m2);
}


} // namespace mob200
