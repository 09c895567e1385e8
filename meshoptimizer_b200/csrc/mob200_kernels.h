// mob200_kernels.h -- launch interface between the host shim (mob200_api.cu) and the kernels.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "mob200_common.h"

// filter ids, identical to enum mob200_Filter in include/meshopt_b200.h
#define MOB200_FILTER_NONE 0
#define MOB200_FILTER_OCT 1
#define MOB200_FILTER_QUAT 2
#define MOB200_FILTER_EXP 3
#define MOB200_FILTER_COLOR 4

namespace mob200
{

// A CTA (one per SM) is kUnitsPerCta independent decode UNITS; a unit is four decoder warps, a producer warp and a
// walker warp with their own slice of shared memory.  The warps of one role are contiguous in the CTA: warp w sits
// on scheduler w % 4, so every scheduler gets the same mix of roles (five decoders, one or two producers, one
// walker).  With five 6-warp CTAs per SM the walkers crowded on two of the four schedulers; a unit's decoder warps
// meet at a barrier every block, so the most loaded scheduler set the pace (measured: fused 1.85 -> 1.70 ms, the
// order of the roles inside the CTA made no difference).
constexpr int kDecodeThreads = 128;  // a unit's decoder warps decode one block at a time: (vertex_size/4) x (vertices/16, rounded up to a power of two) work items
constexpr int kProducerThreads = 32; // a unit's producer warp stages its next blocks (TMA) and resolves their cross-block carry
constexpr int kWalkerThreads = 32;   // a unit's walker warp walks 32 streams, one per lane (or one stream with all lanes)
constexpr int kUnitsPerCta = 5;      // register budget: 65536 / (5 x 192) -> 64 registers per thread
constexpr int kCtaThreads = kUnitsPerCta * (kDecodeThreads + kProducerThreads + kWalkerThreads);
constexpr int kFirstProducerThread = kUnitsPerCta * kDecodeThreads;
constexpr int kFirstWalkerThread = kFirstProducerThread + kUnitsPerCta * kProducerThreads;

constexpr uint32_t kWalkOnly = 0xffffffffu;   // DevTables::walker_lead: decoders off (diagnostic)
constexpr uint32_t kRewalk = 0xfffffffdu;     // DevTables::walker_lead: like kDecodeOnly but the walkers run as well (contention without dependency; diagnostic)
constexpr uint32_t kDecodeOnly = 0xfffffffeu; // DevTables::walker_lead: walkers off, tables of the previous run are reused (diagnostic)

uint32_t decode_smem_bytes();
cudaError_t prepare_decode_kernel();
cudaError_t decode_occupancy(int* ctas_per_sm);

cudaError_t launch_decode(const DevTables& T, uint32_t grid, cudaStream_t stream);
// few long streams: offsets-only walk, one CTA (chain warp + helper warps) per stream; followed by a block-mode launch_decode
cudaError_t launch_walk_team(const DevTables& T, int sm_count, cudaStream_t stream);
cudaError_t launch_filter(int filter, void* data, size_t count, size_t stride, int sm_count, cudaStream_t stream);

} // namespace mob200
