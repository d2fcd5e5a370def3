// usb_usortfull.cuh -- the whole U-sorted candidate list of a query, for exhaustive searches.
//
// k_rank keeps the U counters in shared memory and materialises at most RANK_KCAP candidates, which
// is all a Terminator with -maxaccepts > 0 and -maxrejects > 0 can ever look at.  With -maxaccepts 0
// or -maxrejects 0 (terminator.cpp:23-31) the candidate loop may walk the whole TopOrder list.  For
// that case k_rank also writes U to global memory (RankArgs.u_out) and this kernel turns one U vector
// into the complete list:
//   a5  SetTopBump / SetTopNoBump      udbusortedsearcher.cpp:205-267  (rising MinU, target order)
//   a6  CountSortOrderDesc             countsort.cpp:6-108             (cut NextValue/2, stable, descending)
//
// One warp per (query, strand), three passes:
//   A  32 consecutive targets per step (coalesced).  A step in which no counter exceeds the running
//      maximum cannot change MinU, so its survivors are one comparison and a ballot; the few steps that
//      hold a new maximum are replayed lane by lane in target order with the reference's statements.
//      The survivors (TopTargetIndexes) go to a scratch list in target order; the running maximum
//      before its last rise is CountSortOrderDesc's NextValue, because every strict prefix maximum of
//      U passes the MinU test (MinU < MaxCount after every update).
//   B  histogram of the survivors' counters >= NextValue/2 in shared memory (lanes with equal values
//      are merged with __match_any_sync), turned into descending start offsets by a warp scan.
//   C  stable placement: survivors are revisited in target order, 32 per step; the lanes of one value
//      take consecutive slots in lane order.
// Memory: U, scratch and output are n_seq words per job each; the host bounds jobs x targets.
#pragma once
#include "usb_dev.cuh"

namespace usb {

#define USORTFULL_WARPS 2        // jobs per CTA
#define USORTFULL_HIST 4096      // largest counter value + 1 (unique words of the longest query)

struct UsortFullArgs {
	const uint32_t *u;       // n_jobs x n_seq counters (k_rank, RankArgs.u_out)
	uint32_t *scratch;       // n_jobs x n_seq
	uint32_t *cand_t;        // n_jobs x n_seq: the list, U descending, target ascending within a value
	uint32_t *n_emit;        // list length per job (TopOrder.Size)
	uint32_t n_jobs, n_seq;
	double bump_d;           // BumpPct / 100.0; 0 = SetTopNoBump
	DevCounters *ctr;
};

__global__ void __launch_bounds__(USORTFULL_WARPS * 32) k_usort_full(UsortFullArgs a)
{
	__shared__ uint32_t hist_all[USORTFULL_WARPS][USORTFULL_HIST];
	const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const uint32_t job = blockIdx.x * USORTFULL_WARPS + wib;
	if (job >= a.n_jobs)
		return;
	uint32_t *hist = hist_all[wib];
	const uint32_t N = a.n_seq;
	const uint32_t *U = a.u + (size_t)job * N;
	uint32_t *surv = a.scratch + (size_t)job * N;
	uint32_t *out = a.cand_t + (size_t)job * N;
	const uint32_t lt = (1u << lane) - 1u;

	// ---- pass A: SetTopBump in target order
	uint32_t MinU = 1, MaxCount = 0, NextValue = 0, S = 0;
	for (uint32_t base = 0; base < N; base += 32) {
		const uint32_t t = base + lane;
		const uint32_t u = t < N ? U[t] : 0u;
		bool keep;
		if (!__any_sync(USB_FULL, u > MaxCount))
			keep = u >= MinU;
		else {
			keep = false;
			for (uint32_t l = 0; l < 32; ++l) {
				const uint32_t n = __shfl_sync(USB_FULL, u, l);
				if (n >= MinU) {
					if (n > MaxCount) {
						if (a.bump_d != 0.0) {
							const uint32_t NewMinCount = (uint32_t)(n * a.bump_d);
							if (NewMinCount > MinU && NewMinCount < MaxCount)
								MinU = NewMinCount;
						}
						NextValue = MaxCount;
						MaxCount = n;
					}
					if (lane == l)
						keep = true;
				}
			}
		}
		const uint32_t m = __ballot_sync(USB_FULL, keep);
		if (keep)
			surv[S + __popc(m & lt)] = t;
		S += __popc(m);
	}
	if (MaxCount >= USORTFULL_HIST) {
		if (lane == 0)
			atomicOr(&a.ctr->err, ERR_RECORDS_FULL);
		return;
	}
	const uint32_t MinValue = NextValue / 2;
	__syncwarp();

	// ---- pass B: sizes per value, then descending offsets
	for (uint32_t v = lane; v <= MaxCount; v += 32)
		hist[v] = 0;
	__syncwarp();
	for (uint32_t i = 0; i < S; i += 32) {
		const bool in = i + lane < S;
		const uint32_t u = in ? U[surv[i + lane]] : 0u;
		const bool valid = in && u >= MinValue;
		const uint32_t m = __match_any_sync(USB_FULL, valid ? u : 0xffffffffu);
		if (valid && lane == (uint32_t)__ffs(m) - 1u)
			hist[u] += __popc(m);
		__syncwarp();
	}
	uint32_t total = 0;
	for (uint32_t hi = MaxCount + 1; hi > MinValue; hi -= min(32u, hi - MinValue)) {
		// values hi-1, hi-2, ... one per lane, down to MinValue
		const bool in = hi - MinValue > lane;
		const uint32_t v = in ? hi - 1 - lane : 0u;
		const uint32_t c = in ? hist[v] : 0u;
		uint32_t inc = c;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t x = __shfl_up_sync(USB_FULL, inc, d);
			if (lane >= (uint32_t)d)
				inc += x;
		}
		if (in)
			hist[v] = total + inc - c;
		total += __shfl_sync(USB_FULL, inc, 31);
	}
	__syncwarp();

	// ---- pass C: stable placement
	for (uint32_t i = 0; i < S; i += 32) {
		const bool in = i + lane < S;
		const uint32_t t = in ? surv[i + lane] : 0u;
		const uint32_t u = in ? U[t] : 0u;
		const bool valid = in && u >= MinValue;
		const uint32_t m = __match_any_sync(USB_FULL, valid ? u : 0xffffffffu);
		const uint32_t leader = (uint32_t)__ffs(m) - 1u;
		uint32_t slot = 0;
		if (valid && lane == leader) {
			slot = hist[u];
			hist[u] = slot + __popc(m);
		}
		slot = __shfl_sync(USB_FULL, slot, leader);
		if (valid)
			out[slot + __popc(m & lt)] = t;
		__syncwarp();
	}
	if (lane == 0)
		a.n_emit[job] = total;
}

} // namespace usb
