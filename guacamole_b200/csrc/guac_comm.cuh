// guac_comm.cuh — the path's only exchange ("C1"): the per-shard variant records and depth histograms gathered to
// one rank over NCCL (NVLink / NVSwitch), from device memory.  Loci shard naturally (DistributedUtil.scala:537-545), so there
// is no collective on the data path itself; what the reference does with `collect` / `coalesce(1, shuffle = true)`
// (Common.scala:290-293) is:
//   1. one ncclAllGather of six 8-byte counts per rank (compact records, general records, allele bytes, visited loci, tie
//      loci, exact loci);
//   2. one group of ncclSend / ncclRecv: every rank's compact records (8 bytes each, still in HBM in canonical order), its few
//      general records and their allele bytes, into the root's HBM at the offsets the counts give — rank order is locus order;
//   3. one device -> host copy on the root into the pinned block its merged result owns;
//   4. ncclReduce(sum) of the fixed-size depth histogram (k_depth_histogram below).
#pragma once

#include <nccl.h>

#include "guac_host.cuh"
#include "guac_pileup.cuh"
#include "guac_tile.cuh"

#define NCCL_OK(expr)                                                                                        \
  do {                                                                                                       \
    ncclResult_t r_ = (expr);                                                                                \
    if (r_ != ncclSuccess) fail(GUAC_ERR_CUDA, "%s: %s (%s:%d)", #expr, ncclGetErrorString(r_), __FILE__, __LINE__); \
  } while (0)

struct guac_comm {
  guac_ctx* ctx = nullptr;
  int device = 0;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  DevBuf<unsigned long long> d_sizes;      // [6 * world] after the all-gather, + 6 for the local row
  DevBuf<unsigned char> d_stage, d_recv;   // local general records + allele bytes; the root's receive area
  unsigned long long* h_sizes = nullptr;   // pinned mirror
};

namespace guac {

// ---- K_depth_histogram: loci by depth, from the per-locus start / end counts of the difference streams -------------------------
// One warp per granule tile (grid-stride), lane = 32 consecutive loci, exactly the scan of k_call_tile.  Every lane owns a
// private copy of the histogram in shared memory (bin * 32 + lane: conflict-free, no atomics); the copies are summed into
// the global histogram once per warp.
constexpr int kDepthBins = GUAC_DEPTH_BINS;
constexpr int kHistWarps = 4;

__global__ void __launch_bounds__(kHistWarps * 32) k_depth_histogram(DevReads R, const TileDesc* __restrict__ tiles, uint32_t n_tiles,
                                                                     unsigned long long* __restrict__ hist) {
  extern __shared__ __align__(16) uint32_t h_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* H = h_smem + (size_t)warp * kDepthBins * 32;
  for (int i = lane; i < kDepthBins * 32; i += 32) H[i] = 0;
  __syncwarp();
  const uint32_t n_warps = gridDim.x * kHistWarps;
  for (uint32_t tile = blockIdx.x * kHistWarps + warp; tile < n_tiles; tile += n_warps) {
    const TileDesc td = tiles[tile];
    const int tile_lo = td.word0 << 5;
    const uint32_t g = td.gran;
    const GranHdr hdr = R.gs_hdr[g];
    const size_t locus0 = (size_t)g * kGranuleLoci + (size_t)lane * 32;
    const int l0 = tile_lo + (lane << 5);
    const uint32_t in_range = bit_range(td.locus_begin - l0, td.locus_end - l0);
    int run = 0;
    if (R.gs_wide) {
      const uint32_t* d = reinterpret_cast<const uint32_t*>(R.gs_dd) + locus0;
      for (int k = 0; k < 32; ++k) { const uint32_t v = __ldg(d + k); run += (int)(v & 0xFFFFu) - (int)(v >> 16); }
    } else {
      const uint32_t* d = reinterpret_cast<const uint32_t*>(R.gs_dd + locus0);
      for (int j = 0; j < 8; ++j) {
        const uint32_t v = __ldg(d + j);
        run += (int)__dp4a(v & 0x0F0F0F0Fu, 0x01010101u, 0u) - (int)__dp4a((v >> 4) & 0x0F0F0F0Fu, 0x01010101u, 0u);
      }
    }
    int incl = run;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += t;
    }
    int dep = (int)hdr.depth_in + incl - run;
    for (int k = 0; k < 32; ++k) {
      uint32_t s, e;
      if (R.gs_wide) {
        const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(R.gs_dd) + locus0 + k);
        s = v & 0xFFFFu; e = v >> 16;
      } else {
        const uint32_t v = __ldg(R.gs_dd + locus0 + k);
        s = v & 15u; e = v >> 4;
      }
      dep += (int)s - (int)e;
      if ((in_range >> k) & 1u) H[min(dep, kDepthBins - 1) * 32 + lane] += 1u;
    }
  }
  __syncwarp();
  for (int b = lane; b < kDepthBins; b += 32) {
    unsigned long long s = 0;
    for (int l = 0; l < 32; ++l) s += H[b * 32 + ((l + lane) & 31)];
    if (s) atomicAdd(&hist[b], s);
  }
}

}  // namespace guac
