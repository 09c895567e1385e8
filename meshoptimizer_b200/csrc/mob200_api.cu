// mob200_api.cu -- the C ABI (include/meshopt_b200.h): drop-in meshopt_* symbols on host pointers and
// the batched device-pointer variant.  Host C++ only; every byte of decode work happens in
// mob200_kernels.cu.  There is no CPU fallback: without a CUDA device the calls fail.
//
// Reference interfaces mirrored here (argument meaning and return codes):
//   meshopt_decodeVertexBuffer / meshopt_decodeVertexVersion   src/vertexcodec.cpp:1782-1872
//   meshopt_decodeFilterOct/Quat/Exp/Color                     src/vertexfilter.cpp:1211-1274
//   per-bufferView decode loop of a loader                     gltf/parsegltf.cpp:561-627
#include "../../include/meshopt_b200.h"

#include "mob200_host.h"
#include "mob200_kernels.h"

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

using namespace mob200;

struct mob200_Plan
{
	mob200_Context* ctx = nullptr;
	size_t n = 0;
	DevTables T = {};
	int32_t* d_status = nullptr;
	void* arena = nullptr;
	bool owns_arena = true;
	uint32_t grid = 0;
	// block-offset sidecar: position of every caller stream in the sorted device array, and whether the block_offset
	// table holds offsets a block-mode run may start from (imported, or left by a run of the serial walk)
	std::vector<uint32_t> sorted_of_caller;
	std::vector<uint32_t> stream_nblocks, stream_block_base; // per sorted stream
	bool have_offsets = false;
	int wide_walk_choice = 0, rounds_choice = 0;
	bool small_blocks_majority = false;
	const uint2* tickets_level = nullptr; // decode order of the fused walk: level-major inside a wave of streams
	const uint2* tickets_runs = nullptr;  // block mode + rounds: runs of 1 << run_shift consecutive blocks of a stream (small-vertex plans only)
	uint32_t run_shift = 0;
	bool plain_runs = false; // the run-major table is for the plain form (runs of 16), not for rounds
	bool two_phase = false; // the last run was a team walk + block-mode decode (two launches)
	// ring of CUDA-event pairs (before / after the fused walk + decode kernel), one per run, recorded on the
	// launching stream: per-launch durations can be read back after a timed region without any
	// synchronisation inside it
	static const int kRing = 64;
	cudaEvent_t ev[kRing][2] = {};
	float create_ms = 0.f; // host time spent in mob200_plan_create (sort, tables, uploads)
	unsigned long long runs = 0;
};

extern "C" int mob200_context_create(mob200_Context** out, int device)
{
	if (!out)
		return MOB200_ERR_ARGUMENT;
	*out = nullptr;
	int count = 0;
	CUDA_TRY(cudaGetDeviceCount(&count));
	if (count == 0)
		return MOB200_ERR_CUDA;
	if (device < 0)
		CUDA_TRY(cudaGetDevice(&device));
	if (device >= count)
		return MOB200_ERR_ARGUMENT;

	mob200_Context* ctx = new (std::nothrow) mob200_Context();
	if (!ctx)
		return MOB200_ERR_CUDA;
	ctx->device = device;
	CUDA_TRY(cudaSetDevice(device));
	CUDA_TRY(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
	CUDA_TRY(prepare_decode_kernel());
	CUDA_TRY(decode_occupancy(&ctx->decode_ctas_per_sm));
	if (ctx->decode_ctas_per_sm < 1)
		ctx->decode_ctas_per_sm = 1;
	CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
	if (const char* wide = getenv("MOB200_WIDE_WALK"))
		ctx->wide_walk_mode = atoi(wide);
	if (const char* team = getenv("MOB200_TEAM_WALK"))
		ctx->team_walk = atoi(team);
	if (const char* rm = getenv("MOB200_RUN_MAJOR"))
		ctx->run_major = atoi(rm);
	if (const char* rounds = getenv("MOB200_ROUNDS"))
		ctx->rounds_mode = atoi(rounds);
	if (const char* lead = getenv("MOB200_WALKER_LEAD"))
		ctx->walker_lead = (uint32_t)strtoul(lead, nullptr, 10);
	*out = ctx;
	return 0;
}

extern "C" void mob200_context_destroy(mob200_Context* ctx)
{
	if (!ctx)
		return;
	cudaSetDevice(ctx->device);
	ctx->d_in.release();
	ctx->d_out.release();
	ctx->h_in.release();
	ctx->h_out.release();
	ctx->h_status.release();
	for (int k = 0; k < mob200_Context::kHostSlots; ++k)
	{
		ctx->slot_arena[k].release();
		ctx->slot_h_in[k].release();
		ctx->slot_h_out[k].release();
		if (ctx->slot_stream[k])
			cudaStreamDestroy(ctx->slot_stream[k]);
		if (ctx->slot_done[k])
			cudaEventDestroy(ctx->slot_done[k]);
	}
	if (ctx->stream)
		cudaStreamDestroy(ctx->stream);
	delete ctx;
}

extern "C" int mob200_context_sm_count(const mob200_Context* ctx)
{
	return ctx ? ctx->sm_count : 0;
}

extern "C" const char* mob200_version(void)
{
	return "meshopt_b200 0.1 (sm_100a; vertex codec v0/v1 + oct/quat/exp/color filters)";
}

// ------------------------------------------------------------------------------------------------
// plans
// ------------------------------------------------------------------------------------------------

static bool filter_ok(int filter, size_t vs)
{
	switch (filter)
	{
	case MOB200_FILTER_NONE: return true;
	case MOB200_FILTER_OCT:
	case MOB200_FILTER_COLOR: return vs == 4 || vs == 8; // reference asserts, src/vertexfilter.cpp:1215,1261
	case MOB200_FILTER_QUAT: return vs == 8;              // :1234
	case MOB200_FILTER_EXP: return vs % 4 == 0;           // :1248
	default: return false;
	}
}

static size_t align_up(size_t v, size_t a)
{
	return (v + a - 1) / a * a;
}

// ext_arena != NULL: the tables live in a caller-owned, grow-only buffer, their initialisation is only
// ENQUEUED on init_stream (the plan must then run on that stream), and no timing events are created.
static int plan_create_body(mob200_Context* ctx, const mob200_Stream* streams, size_t n, DeviceBuffer* ext_arena, cudaStream_t init_stream, bool timing, mob200_Plan** out,
    const unsigned int* const* sidecars);

static int plan_create_impl(mob200_Context* ctx, const mob200_Stream* streams, size_t n, DeviceBuffer* ext_arena, cudaStream_t init_stream, bool timing, mob200_Plan** out,
    const unsigned int* const* sidecars = nullptr)
{
	const auto t0 = std::chrono::steady_clock::now();
	const int rc = plan_create_body(ctx, streams, n, ext_arena, init_stream, timing, out, sidecars);
	if (rc == 0 && out && *out)
		(*out)->create_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
	return rc;
}

static int plan_create_body(mob200_Context* ctx, const mob200_Stream* streams, size_t n, DeviceBuffer* ext_arena, cudaStream_t init_stream, bool timing, mob200_Plan** out,
    const unsigned int* const* sidecars)
{
	if (!ctx || !out || (n && !streams) || n >= 0xffffffffull)
		return MOB200_ERR_ARGUMENT;
	*out = nullptr;
	if (set_device(ctx))
		return MOB200_ERR_CUDA;

	// validate, and sort by block count (descending; ties by vertex size) so that the 32 lanes of a
	// walker warp advance streams of similar shape and every decode level is a prefix of the array
	std::vector<uint32_t> nblocks(n);
	for (size_t i = 0; i < n; ++i)
	{
		const mob200_Stream& s = streams[i];
		if (s.vertex_size == 0 || s.vertex_size > 256 || s.vertex_size % 4 != 0) // reference asserts, src/vertexcodec.cpp:1803-1804
			return MOB200_ERR_ARGUMENT;
		if (s.vertex_count >= 0xffffffffull || s.src_size >= 0xffffffffull || !filter_ok(s.filter, s.vertex_size))
			return MOB200_ERR_ARGUMENT;
		if (s.vertex_count && !s.dst)
			return MOB200_ERR_ARGUMENT;
		uint32_t bv = block_vertices((uint32_t)s.vertex_size);
		nblocks[i] = (uint32_t)((s.vertex_count + bv - 1) / bv);
	}
	std::vector<uint32_t> order(n);
	for (size_t i = 0; i < n; ++i)
		order[i] = (uint32_t)i;
	std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
		if (nblocks[a] != nblocks[b])
			return nblocks[a] > nblocks[b];
		return streams[a].vertex_size < streams[b].vertex_size;
	});

	std::vector<DevStream> host(n);
	size_t n_with_blocks = 0;
	uint64_t total_blocks = 0, total_chan = 0, small_blocks = 0, tiny_blocks = 0; // small: a block of <= 12-byte vertices (one or two work quanta of the decoders; 16-byte blocks gain nothing from rounds)
	for (size_t i = 0; i < n; ++i)
	{
		const mob200_Stream& s = streams[order[i]];
		DevStream& d = host[i];
		d.src = s.src;
		d.dst = static_cast<uint8_t*>(s.dst);
		d.chan_base = total_chan;
		d.src_size = s.src ? (uint32_t)s.src_size : 0;
		d.vertex_count = (uint32_t)s.vertex_count;
		d.block_base = (uint32_t)total_blocks;
		d.nblocks = nblocks[order[i]];
		d.vertex_size = (uint16_t)s.vertex_size;
		d.filter = (uint8_t)s.filter;
		d.block_groups = (uint8_t)(block_vertices((uint32_t)s.vertex_size) / kGroup);
		d.caller_index = order[i];
		if (d.nblocks)
			n_with_blocks = i + 1;
		total_blocks += d.nblocks;
		small_blocks += s.vertex_size <= 12 ? d.nblocks : 0;
		tiny_blocks += s.vertex_size <= 8 ? d.nblocks : 0; // one work quantum: four of them make a decode round
		total_chan += (uint64_t)d.nblocks * s.vertex_size;
		if (total_blocks >= 0xfffffff0ull)
			return MOB200_ERR_ARGUMENT;
	}

	mob200_Plan* plan = new (std::nothrow) mob200_Plan();
	if (!plan)
		return MOB200_ERR_CUDA;
	plan->ctx = ctx;
	plan->n = n;

	uint64_t resident = (uint64_t)ctx->sm_count * ctx->decode_ctas_per_sm * kUnitsPerCta; // decode units (mob200_kernels.h)
	uint64_t wanted = std::max<uint64_t>(total_blocks, (n + 31) / 32);
	plan->grid = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(wanted, resident));

	// decode order: streams are walked in waves of grid*32 (one lane each); inside a wave the decoders
	// take block b of every stream before block b+1 of any, which is the order the walkers produce them
	std::vector<uint2> ticket_info(total_blocks);
	{
		const size_t wave = (size_t)plan->grid * 32;
		size_t t = 0;
		for (size_t w0 = 0; w0 < n; w0 += wave)
		{
			size_t w1 = std::min(n, w0 + wave);
			uint32_t levels = host[w0].nblocks; // sorted: the first stream of the wave is the longest
			size_t live = w1 - w0;
			for (uint32_t b = 0; b < levels; ++b)
			{
				while (live > 0 && host[w0 + live - 1].nblocks <= b)
					--live;
				for (size_t i = 0; i < live; ++i)
				{
					ticket_info[t] = make_uint2((uint32_t)(w0 + i), b);
					++t;
				}
			}
		}
	}

	// block mode + rounds: the order is free of the walkers (every block is walked on its own), so the four blocks that
	// share a decode round are four CONSECUTIVE blocks of one stream: inside a round the carries chain through the unit's
	// own look-back entries, and the first member's predecessor lies four times as many tickets back
	const bool small_majority = small_blocks * 2 > total_blocks;
	// Plain form (larger vertices), at least as many streams as units: runs of 16 consecutive blocks of a stream per unit.
	// Inside a run the decoder warps hand the running value from block to block in shared memory (BlockParams::chain); only
	// a run's first block looks back, at a block 16 x streams tickets earlier, i.e. finished.  (With fewer streams than
	// units a run would start where another unit's run of the same age ENDS: the units would wait for each other in turn.)
	const bool plain_runs = !small_majority && n_with_blocks >= plan->grid && ctx->run_major;
	const uint32_t run_shift = plain_runs ? kPlainRunShift : (tiny_blocks * 2 > small_blocks ? kRunShiftMax : 1u), kRunBlocks = 1u << run_shift;
	const bool have_runs = small_majority || plain_runs;
	std::vector<uint2> ticket_runs(have_runs ? total_blocks : 0);
	if (have_runs)
	{
		size_t t = 0;
		const uint32_t levels = n ? host[0].nblocks : 0; // sorted: the first stream is the longest
		size_t live = n;
		for (uint32_t b0 = 0; b0 < levels; b0 += kRunBlocks)
		{
			while (live > 0 && host[live - 1].nblocks <= b0)
				--live;
			for (size_t i = 0; i < live; ++i)
				for (uint32_t b = b0; b < std::min<uint32_t>(b0 + kRunBlocks, host[i].nblocks); ++b)
					ticket_runs[t++] = make_uint2((uint32_t)i, b);
		}
	}

	// one arena for every table
	size_t off_streams = 0;
	size_t off_boff = align_up(off_streams + n * sizeof(DevStream), 256);
	size_t off_table = align_up(off_boff + (total_blocks + n) * 4, 256);
	size_t off_ready = align_up(off_table + total_chan * 32, 256);
	size_t off_look = align_up(off_ready + total_blocks * 4, 256);
	size_t off_tinfo = align_up(off_look + (total_chan / 4) * 8, 256);
	size_t off_truns = align_up(off_tinfo + total_blocks * 8, 256);
	size_t off_status = align_up(off_truns + ticket_runs.size() * 8, 256);
	size_t off_counters = align_up(off_status + n * 4, 256);
#ifdef MOB200_TRACE
	size_t arena_bytes = off_counters + 256 + 16384 * 8; // event trace of one unit (diagnostics build)
#else
	size_t arena_bytes = off_counters + 256;
#endif

	if (ext_arena)
	{
		if (ext_arena->reserve(arena_bytes))
		{
			delete plan;
			return MOB200_ERR_CUDA;
		}
		plan->arena = ext_arena->ptr;
		plan->owns_arena = false;
	}
	else if (cudaMalloc(&plan->arena, arena_bytes) != cudaSuccess)
	{
		cudaGetLastError();
		delete plan;
		return MOB200_ERR_CUDA;
	}
	uint8_t* base = static_cast<uint8_t*>(plan->arena);
	plan->T.streams = reinterpret_cast<DevStream*>(base + off_streams);
	plan->T.block_offset = reinterpret_cast<uint32_t*>(base + off_boff);
	plan->T.group_table = reinterpret_cast<uint16_t*>(base + off_table);
	plan->T.block_ready = reinterpret_cast<uint32_t*>(base + off_ready);
	plan->T.lookback = reinterpret_cast<unsigned long long*>(base + off_look);
	plan->T.ticket_info = reinterpret_cast<uint2*>(base + off_tinfo);
	plan->tickets_level = plan->T.ticket_info;
	plan->tickets_runs = have_runs ? reinterpret_cast<uint2*>(base + off_truns) : nullptr;
	plan->plain_runs = plain_runs;
	plan->T.ticket_shift = 0;
	plan->run_shift = run_shift;
	plan->T.status = reinterpret_cast<int32_t*>(base + off_status);
	plan->T.counters = reinterpret_cast<uint32_t*>(base + off_counters);
	plan->T.n_streams = (uint32_t)n;
	plan->T.n_with_blocks = (uint32_t)n_with_blocks;
	plan->T.total_blocks = (uint32_t)total_blocks;
	plan->T.block_mode = 0;
	plan->T.units = plan->grid;
	plan->T.epoch = 0;
	plan->T.walker_lead = ctx->walker_lead;
	// few streams: one walker WARP per stream (6x lower latency per block); many: one lane per stream
	// mostly small-vertex blocks and enough streams for the decoders to be the bottleneck (with fewer the walkers set the
	// pace and a round only waits longer for its members: measured 5-15% slower at 1024 streams, even at 2048, 20% faster
	// at 4096 for 4- and 8-byte vertices): decode in rounds
	plan->T.rounds = ctx->rounds_mode == 2 ? (small_blocks * 2 > total_blocks && n >= (size_t)4 * plan->grid ? 1u : 0u) : (uint32_t)ctx->rounds_mode;
	plan->T.wide_walk = ctx->wide_walk_mode == 2 ? (n < (size_t)2 * resident ? 1u : 0u) : (uint32_t)ctx->wide_walk_mode;
	plan->small_blocks_majority = small_blocks * 2 > total_blocks;
	plan->wide_walk_choice = (int)plan->T.wide_walk;
	plan->rounds_choice = (int)plan->T.rounds;
	plan->sorted_of_caller.assign(n, 0);
	plan->stream_nblocks.resize(n);
	plan->stream_block_base.resize(n);
	for (size_t i = 0; i < n; ++i)
	{
		plan->sorted_of_caller[order[i]] = (uint32_t)i;
		plan->stream_nblocks[i] = host[i].nblocks;
		plan->stream_block_base[i] = host[i].block_base;
	}

	// table initialisation is enqueued on the context's stream and waited for, so that a later
	// mob200_plan_run on any stream sees it (cudaMemset on device memory may return early)
	bool ok = true;
	cudaStream_t st = init_stream;
	ok = ok && cudaMemsetAsync(base + off_ready, 0, total_blocks * 4, st) == cudaSuccess;      // epoch 0 = never published
	ok = ok && cudaMemsetAsync(base + off_look, 0, (total_chan / 4) * 8, st) == cudaSuccess;
	ok = ok && cudaMemsetAsync(base + off_counters, 0, arena_bytes - off_counters, st) == cudaSuccess;
	if (n)
		ok = ok && cudaMemcpyAsync(base + off_streams, host.data(), n * sizeof(DevStream), cudaMemcpyHostToDevice, st) == cudaSuccess;
	if (total_blocks)
	{
		ok = ok && cudaMemcpyAsync(base + off_tinfo, ticket_info.data(), total_blocks * 8, cudaMemcpyHostToDevice, st) == cudaSuccess;
		if (have_runs)
			ok = ok && cudaMemcpyAsync(base + off_truns, ticket_runs.data(), total_blocks * 8, cudaMemcpyHostToDevice, st) == cudaSuccess;
	}
	std::vector<uint32_t> side;
	if (sidecars)
	{
		// block-offset sidecars: stream i (sorted) owns entries [block_base + i, block_base + i + nblocks] of the table;
		// a stream without a sidecar makes the whole plan fall back to the serial walk
		bool all = true;
		side.assign(total_blocks + n, kInvalidOffset);
		for (size_t i = 0; i < n && all; ++i)
		{
			const unsigned int* sc = sidecars[order[i]];
			if (host[i].nblocks == 0)
				continue;
			if (!sc)
			{
				all = false;
				break;
			}
			memcpy(side.data() + host[i].block_base + i, sc, ((size_t)host[i].nblocks + 1) * 4);
		}
		if (all && total_blocks)
		{
			// (pageable source: like the descriptor uploads above, the call returns once the bytes are staged)
			ok = ok && cudaMemcpyAsync(base + off_boff, side.data(), side.size() * 4, cudaMemcpyHostToDevice, st) == cudaSuccess;
			plan->have_offsets = true;
		}
	}
	if (!ext_arena)
		ok = ok && cudaStreamSynchronize(st) == cudaSuccess;
	if (timing)
		for (int r = 0; r < mob200_Plan::kRing; ++r)
			for (int i = 0; i < 2; ++i)
				ok = ok && cudaEventCreate(&plan->ev[r][i]) == cudaSuccess;
	if (!ok)
	{
		mob200_plan_destroy(plan);
		return MOB200_ERR_CUDA;
	}
	*out = plan;
	return 0;
}

extern "C" int mob200_plan_create(mob200_Context* ctx, const mob200_Stream* streams, size_t n, mob200_Plan** out)
{
	if (!ctx)
		return MOB200_ERR_ARGUMENT;
	return plan_create_impl(ctx, streams, n, nullptr, ctx->stream, true, out);
}

extern "C" void mob200_plan_destroy(mob200_Plan* plan)
{
	if (!plan)
		return;
	cudaSetDevice(plan->ctx->device);
	for (int r = 0; r < mob200_Plan::kRing; ++r)
		for (int i = 0; i < 2; ++i)
			if (plan->ev[r][i])
				cudaEventDestroy(plan->ev[r][i]);
	if (plan->arena && plan->owns_arena)
		cudaFree(plan->arena);
	delete plan;
}

extern "C" int mob200_plan_launches(const mob200_Plan* plan)
{
	if (!plan || !plan->n)
		return 0;
	// one fused persistent kernel per run; plans of few long streams run the offsets-only team walk in front of it
	const bool two_phase = plan->runs ? plan->two_phase : (plan->wide_walk_choice == 1 && plan->ctx->team_walk && plan->T.walker_lead == 0);
	return two_phase ? 2 : 1;
}

extern "C" int mob200_plan_create_sidecar(mob200_Context* ctx, const mob200_Stream* streams, size_t n, const unsigned int* const* sidecars, mob200_Plan** out)
{
	if (!ctx)
		return MOB200_ERR_ARGUMENT;
	return plan_create_impl(ctx, streams, n, nullptr, ctx->stream, true, out, sidecars);
}

extern "C" size_t mob200_sidecar_entries(size_t vertex_count, size_t vertex_size)
{
	if (vertex_size == 0 || vertex_size > 256 || vertex_size % 4 != 0)
		return 0;
	const size_t bv = block_vertices((uint32_t)vertex_size);
	const size_t nblocks = (vertex_count + bv - 1) / bv;
	return nblocks ? nblocks + 1 : 0;
}

extern "C" int mob200_plan_has_offsets(const mob200_Plan* plan)
{
	return plan && plan->have_offsets ? 1 : 0;
}

extern "C" int mob200_plan_export_sidecar(mob200_Plan* plan, size_t stream_index, unsigned int* out, size_t capacity, void* cuda_stream)
{
	if (!plan || stream_index >= plan->n || (!out && capacity))
		return MOB200_ERR_ARGUMENT;
	if (set_device(plan->ctx))
		return MOB200_ERR_CUDA;
	const uint32_t i = plan->sorted_of_caller[stream_index];
	const size_t entries = plan->stream_nblocks[i] ? (size_t)plan->stream_nblocks[i] + 1 : 0;
	if (capacity < entries)
		return MOB200_ERR_ARGUMENT;
	cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
	if (entries)
		CUDA_TRY(cudaMemcpyAsync(out, plan->T.block_offset + plan->stream_block_base[i] + i, entries * 4, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	// a stream whose serial walk failed has kInvalidOffset entries: not a usable sidecar
	for (size_t k = 0; k < entries; ++k)
		if (out[k] == kInvalidOffset)
			return MOB200_ERR_SIDECAR;
	return (int)entries;
}

extern "C" int mob200_plan_run(mob200_Plan* plan, void* cuda_stream)
{
	return mob200_plan_run_ex(plan, cuda_stream, 0);
}

// The decode kernel is persistent and its warp roles wait for each other across CTAs (walker -> producer ->
// decoder, look-back between units): every CTA of a launch must be resident.  One launch alone always is (the
// grid is sized from the occupancy query), two launches that share the device might each hold only part of their
// CTAs and wait for the rest forever.  Launches of this library on one device are therefore chained with an event:
// the kernels of different contexts / CUDA streams run one after the other, copies still overlap them.
struct DeviceSerial
{
	std::mutex mu;
	cudaEvent_t last = nullptr;
	bool recorded = false;
};
static DeviceSerial g_serial[64];

static int launch_serialised(const mob200_Context* ctx, const DevTables& T, uint32_t ctas, cudaStream_t st, cudaEvent_t before, cudaEvent_t after, const DevTables* team_walk)
{
	DeviceSerial& ds = g_serial[ctx->device & 63];
	std::lock_guard<std::mutex> lock(ds.mu);
	if (!ds.last)
		CUDA_TRY(cudaEventCreateWithFlags(&ds.last, cudaEventDisableTiming));
	if (ds.recorded)
		CUDA_TRY(cudaStreamWaitEvent(st, ds.last, 0));
	if (before)
		CUDA_TRY(cudaEventRecord(before, st));
	if (team_walk)
		CUDA_TRY(launch_walk_team(*team_walk, ctx->sm_count, st)); // phase 1 on its own: block offsets and return codes
	CUDA_TRY(launch_decode(T, ctas, st));
	if (after)
		CUDA_TRY(cudaEventRecord(after, st));
	CUDA_TRY(cudaEventRecord(ds.last, st));
	ds.recorded = true;
	return 0;
}

extern "C" int mob200_plan_run_ex(mob200_Plan* plan, void* cuda_stream, int flags)
{
	if (!plan)
		return MOB200_ERR_ARGUMENT;
	if (set_device(plan->ctx))
		return MOB200_ERR_CUDA;
	cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
	const bool block_mode = (flags & MOB200_RUN_BLOCK_PARALLEL) != 0;
	if (block_mode && !plan->have_offsets)
		return MOB200_ERR_ARGUMENT; // no sidecar was imported and no serial walk has run yet
	plan->T.block_mode = block_mode ? 1u : 0u;
	plan->T.wide_walk = block_mode ? 0u : (uint32_t)plan->wide_walk_choice; // block mode: one walker lane per block
	// block mode: the walk no longer limits a batch of few streams, so small-vertex blocks go through the rounds form
	// (run-major order, chained rounds: faster than the plain form down to a thousand blocks -- 16 Ki-vertex normals 52 -> 38 us)
	plan->T.rounds = (block_mode && plan->ctx->rounds_mode == 2) ? (plan->small_blocks_majority ? 1u : 0u) : (uint32_t)plan->rounds_choice;
	if (block_mode && plan->n)
		CUDA_TRY(cudaMemsetAsync(plan->T.status, 0, plan->n * sizeof(int32_t), st)); // block mode reports failures only
	if (!block_mode)
		plan->have_offsets = true; // the serial walk leaves every block's offset in the table (a failed stream is caught when the table is used)
	const bool reuse_tables = plan->T.walker_lead == kDecodeOnly || plan->T.walker_lead == kRewalk;
	if (!reuse_tables || plan->runs == 0) // (decode-only diagnostics: reuse the tables of the first run)
	{
		plan->T.epoch = (plan->T.epoch + 1) & 0x3fffffffu;
		if (plan->T.epoch == 0)
			plan->T.epoch = 1;
	}
	DevTables T = plan->T;
	if (reuse_tables && plan->runs == 0)
		T.walker_lead = 0; // the first run builds the tables
	// Few long streams (the plans that used to take the one-warp-per-stream walker): the offsets-only team walk fills the
	// block-offset table and the status words, then every block is walked and decoded in block mode.
	DevTables Twalk = {};
	const bool two_phase = !block_mode && plan->wide_walk_choice == 1 && plan->ctx->team_walk && plan->T.walker_lead == 0 && plan->n > 0;
	if (two_phase)
	{
		Twalk = T;
		T.block_mode = 1;
		T.wide_walk = 0;
		T.keep_status = 1;
		if (plan->ctx->rounds_mode == 2)
			T.rounds = plan->small_blocks_majority ? 1u : 0u;
	}
	plan->two_phase = two_phase;
	const bool run_major = T.block_mode && plan->tickets_runs && plan->ctx->run_major && (plan->plain_runs ? !T.rounds : T.rounds != 0);
	T.ticket_info = run_major ? plan->tickets_runs : plan->tickets_level;
	T.ticket_shift = run_major ? plan->run_shift : 0u;
	if (two_phase)
		Twalk.ticket_info = plan->tickets_level, Twalk.ticket_shift = 0;

	cudaEvent_t* ev = plan->ev[plan->runs % mob200_Plan::kRing];
	const bool timed = ev[0] != nullptr;
	// unit u runs in CTA u % ctas: a batch with few units still spreads over all SMs
	const uint32_t ctas_max = (uint32_t)(plan->ctx->sm_count * plan->ctx->decode_ctas_per_sm);
	const uint32_t ctas = std::max<uint32_t>((plan->grid + kUnitsPerCta - 1) / kUnitsPerCta, std::min<uint32_t>(ctas_max, plan->grid));
	if (launch_serialised(plan->ctx, T, ctas, st, timed ? ev[0] : nullptr, timed ? ev[1] : nullptr, two_phase ? &Twalk : nullptr))
		return MOB200_ERR_CUDA;
	plan->runs++;
	return 0;
}

extern "C" int mob200_plan_timing_history(mob200_Plan* plan, int max_runs, float* ms)
{
	if (!plan || max_runs < 0 || !plan->ev[0][0])
		return MOB200_ERR_ARGUMENT;
	if (set_device(plan->ctx))
		return MOB200_ERR_CUDA;
	unsigned long long have = plan->runs < (unsigned long long)mob200_Plan::kRing ? plan->runs : (unsigned long long)mob200_Plan::kRing;
	int count = (int)(have < (unsigned long long)max_runs ? have : (unsigned long long)max_runs);
	for (int i = 0; i < count; ++i)
	{
		// i = 0 is the oldest of the `count` most recent runs
		unsigned long long run = plan->runs - count + i;
		cudaEvent_t* ev = plan->ev[run % mob200_Plan::kRing];
		CUDA_TRY(cudaEventSynchronize(ev[1]));
		float a = 0;
		CUDA_TRY(cudaEventElapsedTime(&a, ev[0], ev[1]));
		if (ms)
			ms[i] = a;
	}
	return count;
}

extern "C" int mob200_plan_last_timing(mob200_Plan* plan, float* ms)
{
	int n = mob200_plan_timing_history(plan, 1, ms);
	return n == 1 ? 0 : (n < 0 ? n : MOB200_ERR_ARGUMENT);
}

extern "C" float mob200_plan_create_ms(const mob200_Plan* plan)
{
	return plan ? plan->create_ms : 0.f;
}

extern "C" int mob200_plan_debug_counters(mob200_Plan* plan, unsigned long long* out, int count, int reset)
{
	if (!plan || !out || count < 0 || count > 16)
		return MOB200_ERR_ARGUMENT;
	if (set_device(plan->ctx))
		return MOB200_ERR_CUDA;
	CUDA_TRY(cudaDeviceSynchronize());
	CUDA_TRY(cudaMemcpy(out, plan->T.counters + 16, count * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
	if (reset)
		CUDA_TRY(cudaMemset(plan->T.counters + 16, 0, 16 * sizeof(unsigned long long)));
	return 0;
}

#ifdef MOB200_TRACE
// diagnostics build only: the event trace of unit MOB200_TRACE_UNIT (entry 0 = events recorded); reset afterwards
extern "C" __attribute__((visibility("default"))) int mob200_plan_debug_trace(mob200_Plan* plan, unsigned long long* out, int count)
{
	if (!plan || !out || count < 1 || count > 16384)
		return MOB200_ERR_ARGUMENT;
	if (set_device(plan->ctx))
		return MOB200_ERR_CUDA;
	CUDA_TRY(cudaDeviceSynchronize());
	CUDA_TRY(cudaMemcpy(out, plan->T.counters + 64, (size_t)count * 8, cudaMemcpyDeviceToHost));
	CUDA_TRY(cudaMemset(plan->T.counters + 64, 0, 16384 * 8));
	return 0;
}
#endif

extern "C" int mob200_plan_status(mob200_Plan* plan, int* status, void* cuda_stream)
{
	if (!plan)
		return MOB200_ERR_ARGUMENT;
	if (set_device(plan->ctx))
		return MOB200_ERR_CUDA;
	cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
	std::vector<int> local;
	int* dst = status;
	if (!dst)
	{
		local.resize(plan->n);
		dst = local.data();
	}
	if (plan->n)
		CUDA_TRY(cudaMemcpyAsync(dst, plan->T.status, plan->n * sizeof(int), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	int failed = 0;
	for (size_t i = 0; i < plan->n; ++i)
		failed += dst[i] != 0;
	return failed;
}

extern "C" int mob200_decode_batch_device(mob200_Context* ctx, mob200_Stream* streams, size_t n, void* cuda_stream)
{
	mob200_Plan* plan = nullptr;
	int rc = mob200_plan_create(ctx, streams, n, &plan);
	if (rc)
		return rc;
	rc = mob200_plan_run(plan, cuda_stream);
	std::vector<int> status(n);
	if (rc == 0)
		rc = mob200_plan_status(plan, status.data(), cuda_stream);
	if (rc >= 0)
		for (size_t i = 0; i < n; ++i)
			streams[i].status = status[i];
	mob200_plan_destroy(plan);
	return rc;
}

// ------------------------------------------------------------------------------------------------
// host-pointer batch: H2D -> walk + decode -> D2H, pipelined in chunks over three CUDA streams
// ------------------------------------------------------------------------------------------------

static bool is_pinned(const void* p)
{
	if (!p)
		return false;
	cudaPointerAttributes attr;
	if (cudaPointerGetAttributes(&attr, p) != cudaSuccess)
	{
		cudaGetLastError();
		return false;
	}
	return attr.type == cudaMemoryTypeHost;
}

// A maximal sequence of consecutive streams whose host ranges lie back to back (gaps below 16 bytes, as
// in a glTF buffer or any packed blob): one cudaMemcpyAsync moves the whole sequence, and the device
// copy keeps the host layout (including the 16-byte phase of its first byte).
struct HostRun
{
	size_t first, count; // streams [first, first + count)
	uintptr_t begin;     // host address of the first byte
	size_t bytes;        // length of the host range
	size_t dev_base;     // offset of the run in the device arena (256-byte aligned) ...
	bool pinned;         // ... its first byte sits at dev_base + (begin & 15)
};

// max_gap: bytes that may lie between two merged ranges.  Inputs tolerate up to 15 (the gap is only read);
// outputs must be exactly contiguous (a merged device->host copy would overwrite the gap).
template <typename GetRange>
static void build_runs(size_t n, GetRange range, size_t max_gap, std::vector<HostRun>& runs, std::vector<size_t>& run_of, size_t& arena_bytes)
{
	runs.clear();
	run_of.assign(n, 0);
	arena_bytes = 0;
	for (size_t i = 0; i < n; ++i)
	{
		uintptr_t p;
		size_t len;
		range(i, p, len);
		bool merged = false;
		if (!runs.empty() && len)
		{
			HostRun& r = runs.back();
			uintptr_t end = r.begin + r.bytes;
			if (r.bytes && p >= end && p - end <= max_gap)
			{
				r.bytes = (size_t)(p + len - r.begin);
				r.count = i + 1 - r.first;
				merged = true;
			}
		}
		if (!merged)
		{
			HostRun r;
			r.first = i;
			r.count = 1;
			r.begin = p;
			r.bytes = len;
			r.dev_base = 0;
			r.pinned = false;
			runs.push_back(r);
		}
		run_of[i] = runs.size() - 1;
	}
	for (HostRun& r : runs)
	{
		r.dev_base = arena_bytes;
		arena_bytes += align_up(r.bytes + 32, 256);
		r.pinned = r.bytes ? is_pinned(reinterpret_cast<const void*>(r.begin)) : true;
	}
}

static size_t run_device_offset(const HostRun& r, uintptr_t p)
{
	return r.dev_base + (r.begin & 15) + (size_t)(p - r.begin);
}

static bool stream_args_ok(const mob200_Stream& s)
{
	if (s.vertex_size == 0 || s.vertex_size > 256 || s.vertex_size % 4 != 0) // reference asserts, src/vertexcodec.cpp:1803-1804
		return false;
	if (s.vertex_count >= 0xffffffffull || s.src_size >= 0xffffffffull || !filter_ok(s.filter, s.vertex_size))
		return false;
	if (s.vertex_count && !s.dst)
		return false;
	return true;
}

extern "C" int mob200_decode_batch_host(mob200_Context* ctx, mob200_Stream* streams, size_t n)
{
	return mob200_decode_batch_host_sidecar(ctx, streams, n, nullptr);
}

extern "C" int mob200_decode_batch_host_sidecar(mob200_Context* ctx, mob200_Stream* streams, size_t n, const unsigned int* const* sidecars)
{
	if (!ctx || (n && !streams))
		return MOB200_ERR_ARGUMENT;
	std::lock_guard<std::mutex> lock(ctx->mu);
	if (set_device(ctx))
		return MOB200_ERR_CUDA;
	if (n == 0)
		return 0;

	// Every stream is validated up front: one with illegal arguments gets MOB200_ERR_ARGUMENT as its own status and is
	// replaced by an empty placeholder, the rest of the batch is decoded.
	std::vector<uint8_t> skipped(n, 0);
	std::vector<mob200_Stream> host(streams, streams + n);
	for (size_t i = 0; i < n; ++i)
		if (!stream_args_ok(streams[i]))
		{
			skipped[i] = 1;
			host[i].src = nullptr;
			host[i].src_size = 0;
			host[i].dst = nullptr;
			host[i].vertex_count = 0;
			host[i].vertex_size = 4;
			host[i].filter = MOB200_FILTER_NONE;
		}

	// host ranges -> runs -> device layout
	std::vector<HostRun> in_runs, out_runs;
	std::vector<size_t> in_run_of, out_run_of;
	size_t in_total = 0, out_total = 0;
	build_runs(n, [&](size_t i, uintptr_t& p, size_t& len) {
		p = reinterpret_cast<uintptr_t>(host[i].src);
		len = host[i].src ? host[i].src_size : 0;
	}, 15, in_runs, in_run_of, in_total);
	build_runs(n, [&](size_t i, uintptr_t& p, size_t& len) {
		p = reinterpret_cast<uintptr_t>(host[i].dst);
		len = host[i].vertex_count * host[i].vertex_size;
	}, 0, out_runs, out_run_of, out_total);
	if (ctx->d_in.reserve(in_total + 256) || ctx->d_out.reserve(out_total + 256))
		return MOB200_ERR_CUDA;
	uint8_t* d_in = static_cast<uint8_t*>(ctx->d_in.ptr);
	uint8_t* d_out = static_cast<uint8_t*>(ctx->d_out.ptr);

	std::vector<mob200_Stream> dev(host);
	for (size_t i = 0; i < n; ++i)
	{
		dev[i].src = host[i].src ? d_in + run_device_offset(in_runs[in_run_of[i]], reinterpret_cast<uintptr_t>(host[i].src)) : nullptr;
		dev[i].dst = d_out + run_device_offset(out_runs[out_run_of[i]], reinterpret_cast<uintptr_t>(host[i].dst));
	}

	// chunks of consecutive streams (~kChunkBytes of traffic each) rotate over kSlots in-flight slots
	const size_t kChunkBytes = (size_t)96 << 20;
	const int kSlots = mob200_Context::kHostSlots;
	for (int k = 0; k < kSlots; ++k)
		if (!ctx->slot_stream[k])
		{
			CUDA_TRY(cudaStreamCreateWithFlags(&ctx->slot_stream[k], cudaStreamNonBlocking));
			CUDA_TRY(cudaEventCreateWithFlags(&ctx->slot_done[k], cudaEventDisableTiming));
		}

	std::vector<int> status(n, 0);
	if (ctx->h_status.reserve(n * sizeof(int)))
		return MOB200_ERR_CUDA;
	int* h_status = static_cast<int*>(ctx->h_status.ptr);

	struct SlotWork
	{
		mob200_Plan* plan = nullptr;
		size_t i0 = 0, i1 = 0;
		bool recorded = false;
		std::vector<std::pair<void*, std::pair<const void*, size_t>>> unstage; // dst <- (pinned staging, bytes)
	};
	SlotWork work[mob200_Context::kHostSlots];

	auto retire = [&](int k) -> int {
		SlotWork& w = work[k];
		if (!w.plan)
			return 0;
		int rc = 0;
		if (w.recorded)
		{
			if (cudaEventSynchronize(ctx->slot_done[k]) != cudaSuccess)
				rc = MOB200_ERR_CUDA;
		}
		else if (cudaStreamSynchronize(ctx->slot_stream[k]) != cudaSuccess) // (a chunk that failed half-way: whatever it enqueued must finish)
			rc = MOB200_ERR_CUDA;
		if (rc == 0 && w.recorded)
		{
			for (auto& u : w.unstage)
				memcpy(u.first, u.second.first, u.second.second);
			for (size_t i = w.i0; i < w.i1; ++i)
				status[i] = h_status[i];
		}
		w.unstage.clear();
		mob200_plan_destroy(w.plan);
		w.plan = nullptr;
		w.recorded = false;
		return rc;
	};
	// error path: nothing may stay in flight (DMA into caller memory, kernels, plans) when the call returns
	auto fail = [&](int rc) -> int {
		for (int k = 0; k < kSlots; ++k)
			retire(k);
		cudaGetLastError();
		return rc;
	};
#define SLOT_TRY(expr)                                                                                                  \
	do                                                                                                                  \
	{                                                                                                                   \
		cudaError_t err__ = (expr);                                                                                     \
		if (err__ != cudaSuccess)                                                                                       \
		{                                                                                                               \
			fprintf(stderr, "meshopt_b200: %s failed: %s (%s:%d)\n", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
			return fail(MOB200_ERR_CUDA);                                                                               \
		}                                                                                                               \
	} while (0)

	size_t chunk_index = 0;
	for (size_t i0 = 0; i0 < n; ++chunk_index)
	{
		size_t i1 = i0, bytes = 0;
		while (i1 < n && (i1 == i0 || bytes < kChunkBytes))
		{
			bytes += (host[i1].src ? host[i1].src_size : 0) + host[i1].vertex_count * host[i1].vertex_size;
			++i1;
		}
		const int k = (int)(chunk_index % kSlots);
		if (retire(k))
			return fail(MOB200_ERR_CUDA);
		cudaStream_t st = ctx->slot_stream[k];
		SlotWork& w = work[k];
		w.i0 = i0;
		w.i1 = i1;

		// host -> device: one copy per (run, chunk) intersection
		size_t stage_need_in = 0, stage_need_out = 0;
		for (size_t i = i0; i < i1;)
		{
			const HostRun& r = in_runs[in_run_of[i]];
			size_t last = std::min(i1, r.first + r.count);
			if (!r.pinned)
				for (size_t j = i; j < last; ++j)
					stage_need_in += align_up(host[j].src ? host[j].src_size : 0, 16) + 16;
			i = last;
		}
		for (size_t i = i0; i < i1;)
		{
			const HostRun& r = out_runs[out_run_of[i]];
			size_t last = std::min(i1, r.first + r.count);
			if (!r.pinned)
				for (size_t j = i; j < last; ++j)
					stage_need_out += align_up(host[j].vertex_count * host[j].vertex_size, 16) + 16;
			i = last;
		}
		if ((stage_need_in && ctx->slot_h_in[k].reserve(stage_need_in)) || (stage_need_out && ctx->slot_h_out[k].reserve(stage_need_out)))
			return fail(MOB200_ERR_CUDA);

		size_t cursor = 0;
		for (size_t i = i0; i < i1;)
		{
			const HostRun& r = in_runs[in_run_of[i]];
			size_t last = std::min(i1, r.first + r.count);
			// host range of streams [i, last) inside the run
			uintptr_t h0 = reinterpret_cast<uintptr_t>(host[i].src);
			uintptr_t h1 = reinterpret_cast<uintptr_t>(host[last - 1].src) + (host[last - 1].src ? host[last - 1].src_size : 0);
			if (host[i].src && h1 > h0)
			{
				const void* from = reinterpret_cast<const void*>(h0);
				if (!r.pinned)
				{
					uint8_t* hp = static_cast<uint8_t*>(ctx->slot_h_in[k].ptr) + cursor;
					memcpy(hp, from, h1 - h0);
					from = hp;
					cursor += align_up(h1 - h0, 16) + 16;
				}
				SLOT_TRY(cudaMemcpyAsync(d_in + run_device_offset(r, h0), from, h1 - h0, cudaMemcpyHostToDevice, st));
			}
			i = last;
		}

		// decode the chunk (block mode when every stream of it came with a block-offset sidecar)
		int rc = plan_create_impl(ctx, dev.data() + i0, i1 - i0, &ctx->slot_arena[k], st, false, &w.plan, sidecars ? sidecars + i0 : nullptr);
		if (rc)
			return fail(rc);
		rc = mob200_plan_run_ex(w.plan, st, w.plan->have_offsets ? MOB200_RUN_BLOCK_PARALLEL : 0);
		if (rc)
			return fail(rc);

		// device -> host
		cursor = 0;
		for (size_t i = i0; i < i1;)
		{
			const HostRun& r = out_runs[out_run_of[i]];
			size_t last = std::min(i1, r.first + r.count);
			uintptr_t h0 = reinterpret_cast<uintptr_t>(host[i].dst);
			uintptr_t h1 = reinterpret_cast<uintptr_t>(host[last - 1].dst) + host[last - 1].vertex_count * host[last - 1].vertex_size;
			if (h1 > h0 && r.bytes)
			{
				void* to = reinterpret_cast<void*>(h0);
				if (!r.pinned)
				{
					uint8_t* hp = static_cast<uint8_t*>(ctx->slot_h_out[k].ptr) + cursor;
					w.unstage.push_back({to, {hp, (size_t)(h1 - h0)}});
					to = hp;
					cursor += align_up(h1 - h0, 16) + 16;
				}
				SLOT_TRY(cudaMemcpyAsync(to, d_out + run_device_offset(r, h0), h1 - h0, cudaMemcpyDeviceToHost, st));
			}
			i = last;
		}
		SLOT_TRY(cudaMemcpyAsync(h_status + i0, w.plan->T.status, (i1 - i0) * sizeof(int), cudaMemcpyDeviceToHost, st));
		SLOT_TRY(cudaEventRecord(ctx->slot_done[k], st));
		w.recorded = true;
		i0 = i1;
	}
	for (int k = 0; k < kSlots; ++k)
		if (retire(k))
			return fail(MOB200_ERR_CUDA);
#undef SLOT_TRY

	int failed = 0;
	for (size_t i = 0; i < n; ++i)
	{
		streams[i].status = skipped[i] ? MOB200_ERR_ARGUMENT : status[i];
		failed += streams[i].status != 0;
	}
	return failed;
}

// ------------------------------------------------------------------------------------------------
// several GPUs of one box: streams are independent, so a batch shards with no exchange step
// ------------------------------------------------------------------------------------------------

// Greedy longest-processing-time assignment (the same rule as meshoptimizer_b200/sharding.py): streams in order of
// decreasing cost (ties: lower index first), each to the least loaded rank (ties: lower rank).
extern "C" int mob200_shard_streams(const size_t* costs, size_t n, int world, int* rank_of)
{
	if (world <= 0 || (n && (!costs || !rank_of)))
		return MOB200_ERR_ARGUMENT;
	std::vector<size_t> order(n);
	for (size_t i = 0; i < n; ++i)
		order[i] = i;
	std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return costs[a] > costs[b]; });
	std::vector<unsigned long long> load((size_t)world, 0);
	for (size_t i : order)
	{
		int best = 0;
		for (int r = 1; r < world; ++r)
			if (load[(size_t)r] < load[(size_t)best])
				best = r;
		rank_of[i] = best;
		load[(size_t)best] += costs[i];
	}
	return 0;
}

extern "C" int mob200_decode_batch_multi_host(const int* devices, int n_devices, mob200_Stream* streams, size_t n, const unsigned int* const* sidecars, float* device_ms)
{
	if (!devices || n_devices <= 0 || (n && !streams))
		return MOB200_ERR_ARGUMENT;
	// one context per device, created on first use and kept (their staging arenas are what makes a call cheap)
	static std::mutex mu;
	static mob200_Context* contexts[64] = {};
	{
		std::lock_guard<std::mutex> lock(mu);
		for (int d = 0; d < n_devices; ++d)
		{
			if (devices[d] < 0 || devices[d] >= 64)
				return MOB200_ERR_ARGUMENT;
			if (!contexts[devices[d]] && mob200_context_create(&contexts[devices[d]], devices[d]) != 0)
				return MOB200_ERR_CUDA;
		}
	}

	std::vector<size_t> costs(n);
	for (size_t i = 0; i < n; ++i)
		costs[i] = (streams[i].src ? streams[i].src_size : 0) + streams[i].vertex_count * streams[i].vertex_size; // algorithmic bytes
	std::vector<int> rank_of(n, 0);
	if (mob200_shard_streams(costs.data(), n, n_devices, rank_of.data()))
		return MOB200_ERR_ARGUMENT;

	std::vector<std::vector<size_t>> mine((size_t)n_devices);
	for (size_t i = 0; i < n; ++i)
		mine[(size_t)rank_of[i]].push_back(i);

	std::vector<int> rcs((size_t)n_devices, 0);
	std::vector<std::thread> threads;
	for (int d = 0; d < n_devices; ++d)
		threads.emplace_back([&, d]() {
			const std::vector<size_t>& idx = mine[(size_t)d];
			std::vector<mob200_Stream> part(idx.size());
			std::vector<const unsigned int*> side(idx.size(), nullptr);
			for (size_t k = 0; k < idx.size(); ++k)
			{
				part[k] = streams[idx[k]];
				if (sidecars)
					side[k] = sidecars[idx[k]];
			}
			const auto t0 = std::chrono::steady_clock::now();
			rcs[(size_t)d] = mob200_decode_batch_host_sidecar(contexts[devices[d]], part.data(), part.size(), sidecars ? side.data() : nullptr);
			if (device_ms)
				device_ms[d] = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
			for (size_t k = 0; k < idx.size(); ++k)
				streams[idx[k]].status = part[k].status;
		});
	for (std::thread& t : threads)
		t.join();
	int failed = 0;
	for (int d = 0; d < n_devices; ++d)
	{
		if (rcs[(size_t)d] < 0)
			return rcs[(size_t)d];
		failed += rcs[(size_t)d];
	}
	return failed;
}

// ------------------------------------------------------------------------------------------------
// drop-in symbols
// ------------------------------------------------------------------------------------------------

// pool of contexts for the drop-in symbols: at most kPoolMax per device, created on demand; a caller that finds
// none free waits for one
namespace
{
struct ContextPool
{
	std::mutex mu;
	std::condition_variable cv;
	std::vector<mob200_Context*> idle;
	int created = 0;
};
const int kPoolMax = 8;
ContextPool g_pools[64];
} // namespace

mob200_Context* mob200_pool_acquire()
{
	int device = 0;
	if (cudaGetDevice(&device) != cudaSuccess)
	{
		cudaGetLastError();
		return nullptr;
	}
	ContextPool& pool = g_pools[device & 63];
	std::unique_lock<std::mutex> lock(pool.mu);
	for (;;)
	{
		if (!pool.idle.empty())
		{
			mob200_Context* ctx = pool.idle.back();
			pool.idle.pop_back();
			return ctx;
		}
		if (pool.created < kPoolMax)
		{
			++pool.created;
			lock.unlock();
			mob200_Context* ctx = nullptr;
			if (mob200_context_create(&ctx, device) != 0)
			{
				lock.lock();
				--pool.created;
				pool.cv.notify_one();
				return nullptr;
			}
			return ctx;
		}
		pool.cv.wait(lock);
	}
}

void mob200_pool_release(mob200_Context* ctx)
{
	ContextPool& pool = g_pools[ctx->device & 63];
	{
		std::lock_guard<std::mutex> lock(pool.mu);
		pool.idle.push_back(ctx);
	}
	pool.cv.notify_one();
}

extern "C" int meshopt_decodeVertexVersion(const unsigned char* buffer, size_t buffer_size)
{
	// O(1) framing probe on a host buffer (reference src/vertexcodec.cpp:1782-1797); no decode work
	if (buffer_size < 1)
		return -1;
	unsigned char header = buffer[0];
	if ((header & 0xf0) != kMagic)
		return -1;
	int version = header & 0x0f;
	return version > 1 ? -1 : version;
}

extern "C" int meshopt_decodeVertexBuffer(void* destination, size_t vertex_count, size_t vertex_size, const unsigned char* buffer, size_t buffer_size)
{
	PoolLease lease;
	mob200_Context* ctx = lease.ctx;
	if (!ctx)
		return MOB200_ERR_CUDA;
	mob200_Stream s;
	s.src = buffer;
	s.src_size = buffer ? buffer_size : 0;
	s.dst = destination;
	s.vertex_count = vertex_count;
	s.vertex_size = vertex_size;
	s.filter = MOB200_FILTER_NONE;
	s.status = 0;
	int rc = mob200_decode_batch_host(ctx, &s, 1);
	return rc < 0 ? rc : s.status;
}

static void filter_host(int filter, void* buffer, size_t count, size_t stride)
{
	if (!filter_ok(filter, stride))
	{
		fprintf(stderr, "meshopt_b200: decode filter %d does not accept stride %zu\n", filter, stride);
		abort(); // the reference asserts (src/vertexfilter.cpp:1215,1234,1248,1261)
	}
	if (count == 0)
		return;
	PoolLease lease;
	mob200_Context* ctx = lease.ctx;
	if (!ctx)
	{
		fprintf(stderr, "meshopt_b200: no CUDA device available for meshopt_decodeFilter*\n");
		abort();
	}
	std::lock_guard<std::mutex> lock(ctx->mu);
	size_t bytes = count * stride;
	bool ok = cudaSetDevice(ctx->device) == cudaSuccess && ctx->d_out.reserve(bytes + 16) == 0;
	cudaStream_t st = ctx->stream;
	bool pinned = is_pinned(buffer);
	void* host = buffer;
	if (ok && !pinned)
	{
		ok = ctx->h_out.reserve(bytes) == 0;
		if (ok)
		{
			memcpy(ctx->h_out.ptr, buffer, bytes);
			host = ctx->h_out.ptr;
		}
	}
	ok = ok && cudaMemcpyAsync(ctx->d_out.ptr, host, bytes, cudaMemcpyHostToDevice, st) == cudaSuccess;
	ok = ok && launch_filter(filter, ctx->d_out.ptr, count, stride, ctx->sm_count, st) == cudaSuccess;
	ok = ok && cudaMemcpyAsync(host, ctx->d_out.ptr, bytes, cudaMemcpyDeviceToHost, st) == cudaSuccess;
	ok = ok && cudaStreamSynchronize(st) == cudaSuccess;
	if (!ok)
	{
		fprintf(stderr, "meshopt_b200: CUDA failure in meshopt_decodeFilter*: %s\n", cudaGetErrorString(cudaGetLastError()));
		abort();
	}
	if (!pinned)
		memcpy(buffer, host, bytes);
}

extern "C" void meshopt_decodeFilterOct(void* buffer, size_t count, size_t stride)
{
	filter_host(MOB200_FILTER_OCT, buffer, count, stride);
}

extern "C" void meshopt_decodeFilterQuat(void* buffer, size_t count, size_t stride)
{
	filter_host(MOB200_FILTER_QUAT, buffer, count, stride);
}

extern "C" void meshopt_decodeFilterExp(void* buffer, size_t count, size_t stride)
{
	filter_host(MOB200_FILTER_EXP, buffer, count, stride);
}

extern "C" void meshopt_decodeFilterColor(void* buffer, size_t count, size_t stride)
{
	filter_host(MOB200_FILTER_COLOR, buffer, count, stride);
}

extern "C" int mob200_filter_device(int filter, void* device_buffer, size_t count, size_t stride, void* cuda_stream)
{
	if (filter == MOB200_FILTER_NONE || !filter_ok(filter, stride))
		return MOB200_ERR_ARGUMENT;
	int device = 0, sms = 0;
	CUDA_TRY(cudaGetDevice(&device));
	CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
	CUDA_TRY(launch_filter(filter, device_buffer, count, stride, sms, static_cast<cudaStream_t>(cuda_stream)));
	return 0;
}
