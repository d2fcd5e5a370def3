// usb_rank.cuh -- kernel K1: UDB posting walk + U-sort candidate ranking, one CTA per
// (query, strand).
//
// Reference behaviour reproduced (results identical, algorithm re-designed for one CTA):
//   a1  query words, bad words dropped      udbsearcher.cpp:128-151, udbparams.cpp:540-555
//   a2  unique words                        udbsearcher.cpp:161-194 (order irrelevant for U)
//   a4  U[t] = #unique query words in t     udbusortedsearcher.cpp:375-410 SetU_NonCoded
//   a5  SetTopBump rising-threshold filter  udbusortedsearcher.cpp:230-267
//   a6  CountSortOrderDesc + NextValue/2    countsort.cpp:6-108
//
// Design: the per-target counters live in shared memory (1 byte per target when the query has
// <= 255 word positions, 2 bytes otherwise), so U never touches HBM.
//   * Word phase: one thread per query position; the threads that discover a unique word fetch
//     its row descriptor (offset, size) at once, so those loads overlap.
//   * Posting walk: warps draw whole rows from a shared cursor; rows start on 16-byte boundaries
//     and are streamed with 128-bit loads that bypass L1 allocation, four in flight per lane,
//     and counted with shared-memory atomics on packed 32-bit words.  Measured on B200 the walk
//     itself runs at the HBM roofline (6.7 TB/s with the L2 hits); what is left is the tail.
//   * Tail: the two order-dependent filters are evaluated exactly from the strict prefix maxima
//     of U ("records").  Segment maxima -> block prefix-max scan ->
//     the few threads whose segment holds records count and place them by a block scan (target
//     order, no sort); one thread replays the SetTopBump threshold evolution over them; every
//     thread then filters its own segment, 16 counters per 128-bit shared load, skipping
//     segments and words whose largest counter is below the counting-sort cut-off (16-bit SIMD
//     max is native on sm_100a, the 8-bit forms are emulated); survivors are placed by a
//     block scan.  The Terminator never looks further than maxaccepts+maxrejects-1 = 32
//     candidates, so the survivors are not sorted: every warp sorts 32 keys in registers and a
//     tree of bitonic merges keeps the 32 smallest (U descending, target ascending = the
//     reference's stable counting-sort order).  More than RANK_KCAP survivors: radix-select.
//   * The tail is a chain of short serial sections (about a third of a query's time with one
//     CTA per SM).  The scratch of the word phase (bitmap, row descriptors) and of the tail
//     (records, survivor keys) share one region, which lets two CTAs of 512 threads live on one
//     SM at 100 000 targets: one CTA's tail overlaps the other's posting walk.
#pragma once
#include "usb_dev.cuh"

namespace usb {

#define RANK_THREADS 1024      // k_rank_big, and k_rank when only one CTA fits per SM
#define RANK_THREADS_2 512     // k_rank when two CTAs fit per SM
#define RANK_VEC 4u            // 128-bit posting loads in flight per lane
#define RANK_CHUNK (128u * RANK_VEC) // postings per warp iteration of the walk
#ifndef RANK_VEC16
#define RANK_VEC16 4u          // same for the 2-byte layout (8 measured no faster)
#endif
#define RANK_KCAP 1024         // max candidates materialised per query
#define RANK_REC_NARROW 256
#define RANK_REC_WIDE 2048

struct RankArgs {
	DevParams P;
	const uint8_t *q;          // concatenated query letters
	const uint64_t *q_off;     // n_q + 1
	uint32_t n_jobs;           // n_q * strands
	uint32_t strands;          // 1 or 2
	IndexView ix;              // CSR segments of the UDB index
	uint32_t n_seq;
	uint32_t k_max;            // <= RANK_KCAP
	uint32_t *aux;             // optional, 4 per job: final SetTopBump MinU, NextValue/2, max U, 0
	uint32_t *cand_t;          // n_jobs * k_max
	uint32_t *cand_u;          // n_jobs * k_max (may be null)
	uint32_t *n_cand;          // TopOrder.Size per job
	uint32_t *n_emit;          // min(TopOrder.Size, k_max) per job
	uint32_t *u_out;           // optional: n_jobs * n_seq
	uint32_t seg_narrow, seg_wide; // targets per thread segment (bank-conflict-free strides)
	uint32_t u_bytes;          // shared bytes reserved for the counters
	uint32_t dedupe_words;     // 32-bit words of the word bitmap (nt) / hash set (aa)
	uint32_t rec_cap;
	double bump_d;             // BumpPct / 100.0 ; 0 = no bump
	uint32_t prof;             // 1 = accumulate phase cycles in ctr->prof (measurement)
	DevCounters *ctr;
};

struct RankShared {
	uint32_t n_rows, row_cur, n_rec, n_chg;
	uint32_t maxv, minv, total, vstar, m_eq, n_sel, bstar, above;
	uint32_t pad3, n_post, n_surv, pad2;
	uint32_t warp_tmp[32];
	long long pt;              // phase timer (measurement)
};

// The scratch region behind the counters, used by two phases in turn.
struct RankScratch {
	// word phase + posting walk
	uint32_t *bitmap;          // dedupe_words
	uint32_t *r_off;           // blockDim.x: row start / 4 (one segment) or the word itself (several segments)
	uint32_t *r_size;          // blockDim.x
	// ranking phase
	unsigned long long *sel;   // RANK_KCAP
	uint32_t *rec_pos, *rec_val, *chg_pos, *chg_minu; // rec_cap each
	uint32_t *hist;            // 256
};

__device__ __forceinline__ RankScratch rank_scratch(uint8_t *X, uint32_t dedupe_words, uint32_t rec_cap)
{
	RankScratch r;
	r.bitmap = (uint32_t *)X;
	r.r_off = r.bitmap + dedupe_words;
	r.r_size = r.r_off + blockDim.x;
	r.sel = (unsigned long long *)X;
	r.rec_pos = (uint32_t *)(r.sel + RANK_KCAP);
	r.rec_val = r.rec_pos + rec_cap;
	r.chg_pos = r.rec_val + rec_cap;
	r.chg_minu = r.chg_pos + rec_cap;
	r.hist = r.chg_minu + rec_cap;
	return r;
}

// exclusive prefix maximum over the CTA in thread order
__device__ __forceinline__ uint32_t block_excl_scan_max(uint32_t v, uint32_t *warp_tmp)
{
	const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	uint32_t inc = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		uint32_t t = __shfl_up_sync(USB_FULL, inc, d);
		if (lane >= (uint32_t)d)
			inc = max(inc, t);
	}
	if (lane == 31)
		warp_tmp[w] = inc;
	__syncthreads();
	if (w == 0) {
		uint32_t xi = warp_tmp[lane];
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			uint32_t t = __shfl_up_sync(USB_FULL, xi, d);
			if (lane >= (uint32_t)d)
				xi = max(xi, t);
		}
		uint32_t ex = __shfl_up_sync(USB_FULL, xi, 1);
		warp_tmp[lane] = lane == 0 ? 0u : ex;
	}
	__syncthreads();
	uint32_t ex = __shfl_up_sync(USB_FULL, inc, 1);
	ex = lane == 0 ? 0u : ex;
	uint32_t r = max(ex, warp_tmp[w]);
	__syncthreads();
	return r;
}

// exclusive prefix sum over the CTA in thread order
__device__ __forceinline__ uint32_t block_excl_scan_sum(uint32_t v, uint32_t *warp_tmp)
{
	const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	uint32_t inc = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		uint32_t t = __shfl_up_sync(USB_FULL, inc, d);
		if (lane >= (uint32_t)d)
			inc += t;
	}
	if (lane == 31)
		warp_tmp[w] = inc;
	__syncthreads();
	if (w == 0) {
		uint32_t x = warp_tmp[lane];
		uint32_t xi = x;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			uint32_t t = __shfl_up_sync(USB_FULL, xi, d);
			if (lane >= (uint32_t)d)
				xi += t;
		}
		warp_tmp[lane] = xi - x;
	}
	__syncthreads();
	uint32_t r = inc - v + warp_tmp[w];
	__syncthreads();
	return r;
}

template <bool WIDE>
__device__ __forceinline__ uint32_t u_get(const uint8_t *U, uint32_t t)
{
	return WIDE ? (uint32_t)((const uint16_t *)U)[t] : (uint32_t)U[t];
}

template <bool WIDE>
__device__ __forceinline__ void u_inc(uint32_t *U32, uint32_t t)
{
	if (WIDE)
		atomicAdd(&U32[t >> 1], 1u << ((t & 1) * 16));
	else
		atomicAdd(&U32[t >> 2], 1u << ((t & 3) * 8));
}

// largest counter of a packed word; 16-bit SIMD max is native on sm_100a, the 8-bit forms are not
template <bool WIDE>
__device__ __forceinline__ uint32_t word_max(uint32_t word)
{
	if (!WIDE)
		word = __vmaxu2(word & 0x00ff00ffu, (word >> 8) & 0x00ff00ffu);
	return max(word & 0xffffu, word >> 16);
}

// f(k, word) for the 32-bit counter words of the 16-byte chunks [c0, c1): one 128-bit shared load
// per four words
template <class F>
__device__ __forceinline__ void for_words(const uint4 *U128, uint32_t c0, uint32_t c1, F f)
{
	for (uint32_t c = c0; c < c1; ++c) {
		const uint4 v = U128[c];
		f(4 * c, v.x);
		f(4 * c + 1, v.y);
		f(4 * c + 2, v.z);
		f(4 * c + 3, v.w);
	}
}

// Survivor test for target t with count u (u > 0): SetTopBump threshold in force at t (changes
// take effect after the record position that caused them) and the counting-sort cut-off.
struct KeepCursor {
	const uint32_t *chg_pos, *chg_minu;
	uint32_t n_chg, minv, cur;
	uint32_t mu, next;   // threshold in force and position of the next change, kept in registers
	__device__ __forceinline__ void reset()
	{
		cur = 0;
		mu = 1;
		next = n_chg ? chg_pos[0] : 0xffffffffu;
	}
	// targets must be visited in ascending order between two reset()s
	__device__ __forceinline__ uint32_t minu(uint32_t t)
	{
		while (next < t) {
			mu = chg_minu[cur];
			++cur;
			next = cur < n_chg ? chg_pos[cur] : 0xffffffffu;
		}
		return mu;
	}
	__device__ __forceinline__ bool keep(uint32_t t, uint32_t u) { return u >= minu(t) && u >= minv; }
};

// Both filters on one packed counter word: SetTopBump threshold in force at each target (KeepCursor)
// and the counting-sort cut-off.  Words must be visited in ascending order after a reset().
template <bool WIDE>
struct SegFilter {
	KeepCursor kc;
	uint32_t floor_thr; // max(NextValue / 2, 1)
	static constexpr uint32_t PER = WIDE ? 2 : 4, LANE_BITS = WIDE ? 16 : 8, LANE_MAX = WIDE ? 0xffffu : 0xffu;
	__device__ __forceinline__ void reset() { kc.reset(); }
	__device__ __forceinline__ static uint32_t lane_of(uint32_t word, uint32_t b) { return (word >> (b * LANE_BITS)) & LANE_MAX; }
	// bit b set = counter b of word k survives
	__device__ __forceinline__ uint32_t mask(uint32_t k, uint32_t word)
	{
		if (word_max<WIDE>(word) < floor_thr)
			return 0u;
		const uint32_t t = k * PER;
		uint32_t thr = max(kc.minu(t), floor_thr);
		uint32_t m = 0;
		if (kc.next >= t + PER - 1) { // no change takes effect inside the word: one threshold
			// x >= thr  <=>  max(x, thr) == x, two 16-bit lanes at a time (native SIMD max)
			const uint32_t thr2 = thr * 0x00010001u;
			if (WIDE) {
				const uint32_t d = __vmaxu2(word, thr2) ^ word;
				m = ((d & 0xffffu) ? 0u : 1u) | ((d >> 16) ? 0u : 2u);
			} else {
				const uint32_t lo = word & 0x00ff00ffu, hi = (word >> 8) & 0x00ff00ffu;
				const uint32_t dl = __vmaxu2(lo, thr2) ^ lo, dh = __vmaxu2(hi, thr2) ^ hi;
				m = ((dl & 0xffffu) ? 0u : 1u) | ((dh & 0xffffu) ? 0u : 2u) | ((dl >> 16) ? 0u : 4u) | ((dh >> 16) ? 0u : 8u);
			}
			return m;
		}
		for (uint32_t b = 0; b < PER; ++b) {
			thr = max(kc.minu(t + b), floor_thr);
			if (lane_of(word, b) >= thr)
				m |= 1u << b;
		}
		return m;
	}
};

__device__ __forceinline__ unsigned long long rank_key(uint32_t u, uint32_t t)
{
	return ((unsigned long long)(0xFFFFu - u) << 32) | t;
}

// ascending bitonic sort of sel[0..n) (n <= RANK_KCAP) by the whole CTA
__device__ void block_sort_keys(unsigned long long *sel, uint32_t n)
{
	const uint32_t tid = threadIdx.x;
	uint32_t P2 = 1;
	while (P2 < n)
		P2 <<= 1;
	for (uint32_t i = n + tid; i < P2; i += blockDim.x)
		sel[i] = ~0ull;
	__syncthreads();
	for (uint32_t k = 2; k <= P2; k <<= 1)
		for (uint32_t j = k >> 1; j > 0; j >>= 1) {
			for (uint32_t i = tid; i < P2; i += blockDim.x) {
				const uint32_t x = i ^ j;
				if (x > i) {
					const unsigned long long A = sel[i], B = sel[x];
					if ((A > B) == ((i & k) == 0)) {
						sel[i] = B;
						sel[x] = A;
					}
				}
			}
			__syncthreads();
		}
}

// ascending bitonic sort of 32 keys, one per lane
__device__ __forceinline__ unsigned long long warp_sort32(unsigned long long key, uint32_t lane)
{
#pragma unroll
	for (uint32_t k = 2; k <= 32; k <<= 1)
#pragma unroll
		for (uint32_t j = k >> 1; j > 0; j >>= 1) {
			const unsigned long long o = __shfl_xor_sync(USB_FULL, key, j);
			const bool up = (lane & k) == 0, lower = (lane & j) == 0;
			key = (lower == up) ? min(key, o) : max(key, o);
		}
	return key;
}

// the 32 smallest keys of two ascending runs, ascending: a[i] vs b[31 - i] leaves a bitonic run
__device__ __forceinline__ unsigned long long warp_merge32(unsigned long long a, unsigned long long b_rev, uint32_t lane)
{
	unsigned long long key = min(a, b_rev);
#pragma unroll
	for (uint32_t j = 16; j > 0; j >>= 1) {
		const unsigned long long o = __shfl_xor_sync(USB_FULL, key, j);
		key = (lane & j) == 0 ? min(key, o) : max(key, o);
	}
	return key;
}

// The 32 smallest keys of sel[0..n) (n <= RANK_KCAP) in ascending order, left in sel[0..32);
// the number of warps must be a power of two.  Every warp sorts its runs of 32 keys in registers
// (shuffles only) and keeps the smallest 32; a tree of bitonic merges over the warps, exchanged
// through shared memory, does the rest.
__device__ void block_top32(unsigned long long *sel, uint32_t n)
{
	const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
	// warp w owns runs w * per .. w * per + per - 1 (consecutive, so that padding runs are last)
	const uint32_t per = (RANK_KCAP / 32) / nw;
	unsigned long long key = ~0ull;
	for (uint32_t r = 0; r < per; ++r) {
		const uint32_t i0 = (w * per + r) * 32;
		if (i0 >= n)
			break;
		unsigned long long k2 = i0 + lane < n ? sel[i0 + lane] : ~0ull;
		k2 = warp_sort32(k2, lane);
		key = r == 0 ? k2 : warp_merge32(key, __shfl_sync(USB_FULL, k2, 31 - lane), lane);
	}
	__syncthreads(); // every key is in a register now: sel becomes the exchange buffer
	for (uint32_t d = 1; d * per * 32 < n; d <<= 1) { // runs starting at or after n are all padding
		const uint32_t role = w & (2 * d - 1);
		if (role == d)
			sel[(w - d) * 32 + lane] = key;
		__syncthreads();
		if (role == 0 && (w + d) * per * 32 < n) // the partner is not all padding
			key = warp_merge32(key, sel[w * 32 + 31 - lane], lane);
		__syncthreads();
	}
	if (w == 0)
		sel[lane] = key;
	__syncthreads();
}

// More than RANK_KCAP survivors (at config 4 every read without a close target: its cut-off is
// low and thousands of targets pass).  Only the first k_max of (U desc, target asc) are wanted, so
// the cut-off is raised by bisection until between k_max and RANK_KCAP survivors remain; a
// counting pass is cheap because segments and words below the trial cut-off are skipped.  When a
// single count value straddles both bounds (many ties), everything above it is taken plus its
// first ties in ascending target order.  Leaves the selection in X.sel[0 .. S.n_sel).
template <bool WIDE>
__device__ void rank_select_bisect(const RankArgs &a, RankShared &S, const RankScratch &X, const uint4 *U128,
  const SegFilter<WIDE> &F0, uint32_t c0, uint32_t c1, uint32_t seg_max)
{
	const uint32_t tid = threadIdx.x, NT = blockDim.x;
	constexpr uint32_t PER = WIDE ? 2 : 4;
	// survivors of this thread's segment with U >= thr (thr >= the original cut-off)
	auto count_at = [&](uint32_t thr) -> uint32_t {
		uint32_t cnt = 0;
		if (seg_max >= thr) {
			SegFilter<WIDE> F = F0;
			F.floor_thr = thr;
			F.reset();
			for_words(U128, c0, c1, [&](uint32_t k, uint32_t word) { cnt += __popc(F.mask(k, word)); });
		}
		return cnt;
	};
	// block total of v, and this thread's exclusive prefix
	auto block_total = [&](uint32_t v, uint32_t &excl) -> uint32_t {
		excl = block_excl_scan_sum(v, S.warp_tmp);
		if (tid == NT - 1)
			S.total = excl + v;
		__syncthreads();
		const uint32_t t = S.total;
		__syncthreads();
		return t;
	};
	uint32_t lo = F0.floor_thr, hi = S.maxv + 1; // count(lo) > RANK_KCAP, count(hi) = 0
	uint32_t thr = 0, my = 0, excl = 0, c_hi = 0, my_hi = 0, excl_hi = 0;
	bool found = false;
	while (hi - lo > 1) {
		const uint32_t mid = (lo + hi) >> 1;
		my = count_at(mid);
		const uint32_t c = block_total(my, excl);
		if (c > RANK_KCAP)
			lo = mid;
		else if (c >= a.k_max) {
			found = true;
			thr = mid;
			break;
		} else {
			hi = mid;
			c_hi = c;
			my_hi = my;
			excl_hi = excl;
		}
	}
	if (!found) { // count(hi) < k_max <= RANK_KCAP < count(hi - 1): take all >= hi, then ties at hi - 1
		thr = hi;
		my = my_hi;
		excl = excl_hi;
	}
	// everything at or above thr, ascending targets
	if (my) {
		SegFilter<WIDE> F = F0;
		F.floor_thr = thr;
		F.reset();
		uint32_t slot = excl;
		for_words(U128, c0, c1, [&](uint32_t k, uint32_t word) {
			uint32_t m = F.mask(k, word);
			while (m) {
				const uint32_t b = (uint32_t)__ffs(m) - 1;
				m &= m - 1;
				X.sel[slot++] = rank_key(SegFilter<WIDE>::lane_of(word, b), k * PER + b);
			}
		});
	}
	if (found) {
		if (tid == 0)
			S.n_sel = S.total; // block_total left count(thr) there
		__syncthreads();
		return;
	}
	// ties: survivors with U == hi - 1, the first k_max - c_hi of them in target order
	const uint32_t tie = hi - 1, need = a.k_max - c_hi;
	uint32_t ties = 0;
	auto for_ties = [&](auto f) {
		if (seg_max < tie)
			return;
		SegFilter<WIDE> F = F0;
		F.floor_thr = max(F0.floor_thr, tie);
		F.reset();
		for_words(U128, c0, c1, [&](uint32_t k, uint32_t word) {
			uint32_t m = F.mask(k, word);
			while (m) {
				const uint32_t b = (uint32_t)__ffs(m) - 1;
				m &= m - 1;
				if (SegFilter<WIDE>::lane_of(word, b) == tie)
					f(k * PER + b);
			}
		});
	};
	for_ties([&](uint32_t) { ++ties; });
	uint32_t rank;
	block_total(ties, rank);
	for_ties([&](uint32_t t) {
		if (rank < need)
			X.sel[c_hi + rank] = rank_key(tie, t);
		++rank;
	});
	if (tid == 0)
		S.n_sel = a.k_max;
	__syncthreads();
}

// Counts vector j of a chunk with `rem` postings left in its row (rows are padded to whole
// vectors, the pad entries of the ragged last vector are skipped).
template <bool WIDE>
__device__ __forceinline__ void vec_count(uint32_t *U32, const uint4 &x, uint32_t j, uint32_t rem, uint32_t lane)
{
	const uint32_t b = 4 * (lane + 32 * j);
	if (b + 3 < rem) {
		u_inc<WIDE>(U32, x.x); u_inc<WIDE>(U32, x.y); u_inc<WIDE>(U32, x.z); u_inc<WIDE>(U32, x.w);
	} else if (b < rem) {
		u_inc<WIDE>(U32, x.x);
		if (b + 1 < rem) u_inc<WIDE>(U32, x.y);
		if (b + 2 < rem) u_inc<WIDE>(U32, x.z);
	}
}

// One vector of eight 2-byte increment descriptors (usb_hostindex.h HostHalf): entry i is the
// index of a 32-bit counter word, its byte class is i / 2.  Narrow counters: one ATOMS per entry
// with a constant increment.  Wide (16-bit) counters: target 4 w + b lives in halfword b % 2 of
// word 2 w + b / 2.
template <bool WIDE>
__device__ __forceinline__ void vec_count16(uint32_t *U32, const uint4 &x)
{
#define USB_INC16(W, B)                                                                  \
	{                                                                                    \
		const uint32_t lo_ = (W) & 0xffffu, hi_ = (W) >> 16;                             \
		if (WIDE) {                                                                      \
			atomicAdd(&U32[2 * lo_ + ((B) >> 1)], 1u << (16 * ((B) & 1)));               \
			atomicAdd(&U32[2 * hi_ + ((B) >> 1)], 1u << (16 * ((B) & 1)));               \
		} else {                                                                         \
			atomicAdd(&U32[lo_], 1u << (8 * (B)));                                       \
			atomicAdd(&U32[hi_], 1u << (8 * (B)));                                       \
		}                                                                                \
	}
	USB_INC16(x.x, 0)
	USB_INC16(x.y, 1)
	USB_INC16(x.z, 2)
	USB_INC16(x.w, 3)
#undef USB_INC16
}

__device__ __forceinline__ uint4 ld_stream128(const uint4 *p)
{
	uint4 x;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "l"(p));
	return x;
}

// Posting rows are streamed once per query: they are not allocated in L1 (most of the unified
// L1/shared memory array is carved out for the counters).
__device__ __forceinline__ void vec_load(uint4 &x, const uint4 *v4, uint32_t j, uint32_t rem, uint32_t lane)
{
	if (4 * (lane + 32 * j) < rem) {
#ifdef USB_RANK_LDG
		x = __ldg(v4 + lane + 32 * j);
#else
		asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
		             : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w)
		             : "l"(v4 + lane + 32 * j));
#endif
	}
}

template <bool WIDE, bool HALF>
__device__ void rank_job(const RankArgs &a, uint32_t job, RankShared &S, uint8_t *U, const RankScratch &X)
{
	const uint32_t tid = threadIdx.x, lane = tid & 31, NT = blockDim.x;
	const uint32_t N = a.n_seq;
	const uint32_t WLEN = a.P.word_length;
	const uint32_t qi = job / a.strands, strand = job % a.strands;
	const uint64_t q0 = a.q_off[qi];
	const uint32_t L = (uint32_t)(a.q_off[qi + 1] - q0);
	const uint8_t *Q = a.q + q0;
	uint32_t *U32 = (uint32_t *)U;
	uint32_t *bitmap = X.bitmap;
	const bool one_seg = a.ix.n_seg == 1;
	const uint32_t PW = NT; // query positions per pass
#define USB_PHASE(i)                                               \
	if (a.prof && tid == 0) {                                      \
		const long long now = clock64();                           \
		atomicAdd(&a.ctr->prof[i], (unsigned long long)(now - S.pt)); \
		S.pt = now;                                                \
	}
	if (a.prof && tid == 0)
		S.pt = clock64();

	// ---- zero shared state (u_bytes is a multiple of 16)
	{
		uint4 *U128 = (uint4 *)U;
		const uint32_t n128 = ((WIDE ? 2 * N : N) + 15) / 16;
		for (uint32_t i = tid; i < n128; i += NT)
			U128[i] = make_uint4(0, 0, 0, 0);
		// nt: bitmap over the 4^w slots; aa: open-addressing hash set of the query's words
		const uint32_t bm0 = a.P.alpha == 4 ? 0u : 0xffffffffu;
		for (uint32_t i = tid; i < a.dedupe_words; i += NT)
			bitmap[i] = bm0;
		if (tid == 0) {
			S.n_rows = 0; S.row_cur = 0; S.n_rec = 0; S.n_chg = 0; S.n_sel = 0; S.n_post = 0; S.n_surv = 0;
		}
	}
	__syncthreads();
	USB_PHASE(0)

	// ---- a1/a2/a4: words -> unique rows -> posting walk
	const uint32_t npos = L >= WLEN ? L - WLEN + 1 : 0;
	for (uint32_t base = 0; base < npos; base += PW) {
		const uint32_t p = base + tid;
		if (tid < PW && p < npos) {
			uint32_t word = 0, bad = 0;
			bool fresh = false;
			if (a.P.alpha == 4) {
				for (uint32_t i = 0; i < WLEN; ++i) {
					uint32_t c = strand ? (uint32_t)c_comp[Q[L - 1 - (p + i)]] : (uint32_t)Q[p + i];
					uint32_t l = udb_letter(c);
					bad |= l >> 2;
					word = (word << 2) | (l & 3);
				}
				if (!bad) {
					uint32_t bit = 1u << (word & 31);
					uint32_t old = atomicOr(&bitmap[word >> 5], bit);
					fresh = !(old & bit);
				}
			} else {
				for (uint32_t i = 0; i < WLEN; ++i) {
					const uint32_t l = c_udb_aa[Q[p + i]];
					bad |= l >> 7;
					word = word * 20u + (l & 31);
				}
				if (!bad) {
					const uint32_t mask = a.P.hash_cap - 1;
					uint32_t h = (word * 2654435761u) >> 12 & mask;
					for (;;) {
						const uint32_t old = atomicCAS(&bitmap[h], 0xffffffffu, word);
						if (old == 0xffffffffu) {
							fresh = true;
							break;
						}
						if (old == word)
							break;
						h = (h + 1) & mask;
					}
				}
			}
			if (fresh && HALF) {
				// row descriptor now: the loads of all unique words of the pass overlap
				const uint32_t size = __ldg(a.ix.seg[0].row_size + word);
				const uint32_t groups = __ldg(a.ix.row_groups + word);
				const uint64_t off = __ldg(a.ix.row_off16 + word);
				if (size) {
					const uint32_t idx = atomicAdd(&S.n_rows, 1u);
					X.r_off[idx] = (uint32_t)(off >> 3);
					X.r_size[idx] = groups;
					atomicAdd(&S.n_post, size);
				}
			} else if (fresh) {
				if (one_seg) {
					// row descriptor now: the loads of all unique words of the pass overlap
					const SegDesc &seg = a.ix.seg[0];
					const uint32_t size = __ldg(seg.row_size + word);
					const uint64_t off = __ldg(seg.row_off + word);
					if (size) {
						const uint32_t idx = atomicAdd(&S.n_rows, 1u);
						X.r_off[idx] = (uint32_t)(off >> 2);
						X.r_size[idx] = size;
					}
				} else
					X.r_off[atomicAdd(&S.n_rows, 1u)] = word;
			}
		}
		__syncthreads();
		const uint32_t n_rows = S.n_rows;
		if (HALF || one_seg) {
			if (!HALF) {
				const uint32_t size = tid < n_rows ? X.r_size[tid] : 0;
				const uint32_t sz = __reduce_add_sync(USB_FULL, size);
				if (lane == 0 && sz)
					atomicAdd(&S.n_post, sz);
			}
			USB_PHASE(1)
			const uint4 *P4 = HALF ? (const uint4 *)a.ix.post16 : (const uint4 *)a.ix.seg[0].postings;
			// whole rows per warp, drawn from a shared cursor
			for (;;) {
				uint32_t r = 0;
				if (lane == 0)
					r = atomicAdd(&S.row_cur, 1u);
				r = __shfl_sync(USB_FULL, r, 0);
				if (r >= n_rows)
					break;
				const uint32_t size = X.r_size[r];
				const uint4 *v4 = P4 + X.r_off[r];
				if (HALF) {
					// size = groups of 32 vectors; every entry is a valid increment (padding
					// points at dummy words), so there is nothing to test
					for (uint32_t g0 = 0; g0 < size; g0 += RANK_VEC16) {
						uint4 x[RANK_VEC16];
#pragma unroll
						for (uint32_t j = 0; j < RANK_VEC16; ++j)
							if (g0 + j < size)
								x[j] = ld_stream128(v4 + (g0 + j) * 32 + lane);
#pragma unroll
						for (uint32_t j = 0; j < RANK_VEC16; ++j)
							if (g0 + j < size)
								vec_count16<WIDE>(U32, x[j]);
					}
					continue;
				}
				for (uint32_t done = 0; done < size; done += RANK_CHUNK) {
					const uint32_t rem = min(size - done, RANK_CHUNK);
					uint4 x[RANK_VEC];
#pragma unroll
					for (uint32_t j = 0; j < RANK_VEC; ++j)
						vec_load(x[j], v4 + done / 4, j, rem, lane);
#pragma unroll
					for (uint32_t j = 0; j < RANK_VEC; ++j)
						vec_count<WIDE>(U32, x[j], j, rem, lane);
				}
			}
		} else {
			// several index segments (a growing cluster_fast database): whole rows per warp
			uint32_t my_post = 0;
			for (;;) {
				uint32_t r = 0;
				if (lane == 0)
					r = atomicAdd(&S.row_cur, 1u);
				r = __shfl_sync(USB_FULL, r, 0);
				if (r >= n_rows)
					break;
				const uint32_t word = X.r_off[r];
				for (uint32_t sg = 0; sg < a.ix.n_seg; ++sg) {
					const SegDesc &seg = a.ix.seg[sg];
					const uint32_t size = seg.row_size[word];
					if (size == 0)
						continue;
					const uint4 *v4 = (const uint4 *)(seg.postings + seg.row_off[word]);
					my_post += size;
					for (uint32_t done = 0; done < size; done += (RANK_CHUNK)) {
						const uint32_t rem = min(size - done, (RANK_CHUNK));
						uint4 x[RANK_VEC];
#pragma unroll
						for (uint32_t j = 0; j < RANK_VEC; ++j)
							vec_load(x[j], v4 + done / 4, j, rem, lane);
#pragma unroll
						for (uint32_t j = 0; j < RANK_VEC; ++j)
							vec_count<WIDE>(U32, x[j], j, rem, lane);
					}
				}
			}
			if (lane == 0 && my_post)
				atomicAdd(&S.n_post, my_post);
		}
		__syncthreads();
		USB_PHASE(2)
		if (tid == 0) {
			S.n_rows = 0;
			S.row_cur = 0;
		}
		__syncthreads();
	}

	if (a.u_out)
		for (uint32_t t = tid; t < N; t += NT)
			a.u_out[(uint64_t)job * N + t] = u_get<WIDE>(U, t);

	// ---- strict prefix maxima ("records") of U in target order; SIMD max over the segment
	uint32_t *rec_pos = X.rec_pos, *rec_val = X.rec_val, *chg_pos = X.chg_pos, *chg_minu = X.chg_minu;
	const uint32_t seg = WIDE ? a.seg_wide : a.seg_narrow;
	const uint32_t t0 = min(N, tid * seg), t1 = min(N, t0 + seg);
	const uint32_t PER = WIDE ? 2 : 4;   // counters per 32-bit word
	// segment start is word aligned; threads past the end own nothing
	const uint32_t w0 = t0 / PER, w1 = t0 < t1 ? (t1 + PER - 1) / PER : w0;
	// the same range in 16-byte chunks (segments are whole chunks; counters past N are zero)
	const uint4 *U128 = (const uint4 *)U;
	const uint32_t c0 = w0 / 4, c1 = t0 < t1 ? (w1 + 3) / 4 : c0;
	uint32_t m = 0;
	{
		uint32_t acc = 0;
		for_words(U128, c0, c1, [&](uint32_t, uint32_t word) {
			if (WIDE)
				acc = __vmaxu2(acc, word);
			else
				acc = __vmaxu2(__vmaxu2(acc, word & 0x00ff00ffu), (word >> 8) & 0x00ff00ffu);
		});
		m = max(acc & 0xffffu, acc >> 16);
	}
	// records are counted per thread and placed by a block scan, which leaves them in target order
	USB_PHASE(9)
	const uint32_t run0 = block_excl_scan_max(m, S.warp_tmp);
	USB_PHASE(10)
	uint32_t nrec = 0;
	if (m > run0) {
		uint32_t run = run0;
		for_words(U128, c0, c1, [&](uint32_t, uint32_t word) {
			if (word_max<WIDE>(word) > run)
				for (uint32_t b = 0; b < PER; ++b) {
					const uint32_t u = (word >> (b * (WIDE ? 16 : 8))) & (WIDE ? 0xffffu : 0xffu);
					if (u > run) {
						++nrec;
						run = u;
					}
				}
		});
	}
	uint32_t rslot = block_excl_scan_sum(nrec, S.warp_tmp);
	if (tid == NT - 1)
		S.n_rec = rslot + nrec;
	if (nrec) {
		uint32_t run = run0;
		for_words(U128, c0, c1, [&](uint32_t k, uint32_t word) {
			if (word_max<WIDE>(word) > run)
				for (uint32_t b = 0; b < PER; ++b) {
					const uint32_t u = (word >> (b * (WIDE ? 16 : 8))) & (WIDE ? 0xffffu : 0xffu);
					if (u > run) {
						if (rslot < a.rec_cap) {
							rec_pos[rslot] = k * PER + b;
							rec_val[rslot] = u;
						}
						++rslot;
						run = u;
					}
				}
		});
	}
	__syncthreads();
	USB_PHASE(11)
	if (tid == 0) {
		uint32_t n = S.n_rec;
		if (n > a.rec_cap) {
			atomicOr(&a.ctr->err, ERR_RECORDS_FULL);
			n = a.rec_cap;
		}
		// SetTopBump replayed over the records (udbusortedsearcher.cpp:247-263)
		uint32_t MinU = 1, MaxCount = 0, nchg = 0;
		for (uint32_t k = 0; k < n; ++k) {
			uint32_t v = rec_val[k];
			if (a.bump_d != 0.0) {
				uint32_t NewMin = (uint32_t)((double)v * a.bump_d);
				if (NewMin > MinU && NewMin < MaxCount) {
					MinU = NewMin;
					chg_pos[nchg] = rec_pos[k];
					chg_minu[nchg] = MinU;
					++nchg;
				}
			}
			MaxCount = v;
		}
		S.n_chg = nchg;
		S.pad2 = MinU; // SetTopBump threshold in force after the last target
		S.maxv = n ? rec_val[n - 1] : 0;
		// countsort.cpp:12-24: NextValue = running max before its last increase
		S.minv = (n >= 2 ? rec_val[n - 2] : 0) / 2;
	}
	__syncthreads();
	USB_PHASE(3)

	// ---- survivors of both filters: counted per thread, placed by a block scan (ascending targets).
	// A segment whose largest counter is below the counting-sort cut-off has none; a word is
	// looked at counter by counter only when its largest counter reaches the threshold.
	SegFilter<WIDE> F{KeepCursor{chg_pos, chg_minu, S.n_chg, S.minv, 0, 1, 0}, max(S.minv, 1u)};
	const uint32_t floor_thr = F.floor_thr;
	constexpr uint32_t LANE_BITS = WIDE ? 16 : 8, LANE_MAX = WIDE ? 0xffffu : 0xffu;
	{
		uint32_t cnt = 0;
		if (m >= floor_thr) {
			F.reset();
			for_words(U128, c0, c1, [&](uint32_t k, uint32_t word) { cnt += __popc(F.mask(k, word)); });
		}
		USB_PHASE(6)
		uint32_t slot = block_excl_scan_sum(cnt, S.warp_tmp);
		USB_PHASE(7)
		if (tid == NT - 1)
			S.n_surv = slot + cnt;
		if (cnt && slot < RANK_KCAP) {
			F.reset();
			for_words(U128, c0, c1, [&](uint32_t k, uint32_t word) {
				uint32_t mask = F.mask(k, word);
				while (mask) {
					const uint32_t b = (uint32_t)__ffs(mask) - 1;
					mask &= mask - 1;
					if (slot < RANK_KCAP)
						X.sel[slot] = rank_key((word >> (b * LANE_BITS)) & LANE_MAX, k * PER + b);
					++slot;
				}
			});
		}
		USB_PHASE(8)
	}
	__syncthreads();
	USB_PHASE(4)
	const uint32_t total = S.n_surv;
	uint32_t nsel;
	const bool top32 = a.k_max <= 32 && (NT & (NT - 1)) == 0 && NT >= 32;
	if (total <= RANK_KCAP) {
		nsel = min(total, a.k_max);
		if (top32)
			block_top32(X.sel, total);
		else
			block_sort_keys(X.sel, total);
	} else {
		rank_select_bisect<WIDE>(a, S, X, U128, F, c0, c1, m);
		nsel = min(S.n_sel, (uint32_t)RANK_KCAP);
		if (top32)
			block_top32(X.sel, nsel);
		else
			block_sort_keys(X.sel, nsel);
		nsel = min(nsel, a.k_max);
	}
	USB_PHASE(12)
	for (uint32_t i = tid; i < nsel; i += NT) {
		const unsigned long long key = X.sel[i];
		a.cand_t[(uint64_t)job * a.k_max + i] = (uint32_t)key;
		if (a.cand_u)
			a.cand_u[(uint64_t)job * a.k_max + i] = 0xFFFFu - (uint32_t)(key >> 32);
	}
	if (tid == 0) {
		a.n_cand[job] = total;
		a.n_emit[job] = nsel;
		if (a.aux) {
			a.aux[4 * job] = S.pad2;
			a.aux[4 * job + 1] = S.minv;
			a.aux[4 * job + 2] = S.maxv;
			a.aux[4 * job + 3] = 0;
		}
		atomicAdd(&a.ctr->postings, (unsigned long long)S.n_post);
	}
	USB_PHASE(5)
#undef USB_PHASE
}

template <bool HALF>
__global__ void __launch_bounds__(RANK_THREADS, 1) k_rank(const RankArgs a)
{
	extern __shared__ __align__(16) uint8_t rank_smem[];
	RankShared &S = *(RankShared *)rank_smem;
	uint8_t *U = rank_smem + ((sizeof(RankShared) + 15) & ~(size_t)15);
	const RankScratch X = rank_scratch(U + a.u_bytes, a.dedupe_words, a.rec_cap);
	const uint32_t job = blockIdx.x;
	if (job >= a.n_jobs)
		return;
	const uint32_t qi = job / a.strands;
	const uint32_t L = (uint32_t)(a.q_off[qi + 1] - a.q_off[qi]);
	const uint32_t npos = L >= a.P.word_length ? L - a.P.word_length + 1 : 0;
	if (npos > 255)
		rank_job<true, HALF>(a, job, S, U, X);
	else
		rank_job<false, HALF>(a, job, S, U, X);
}

// dedupe_bytes: nt = slots / 8 (bitmap), aa = 4 * hash_cap
inline size_t rank_smem_bytes(uint32_t n_seq, bool wide, size_t dedupe_bytes, uint32_t rec_cap, uint32_t threads,
  uint32_t *u_bytes)
{
	// + the dummy words behind the counters that padding entries of the 2-byte layout increment
	// (usb_hostindex.h HostHalf: words dummy0 .. dummy0 + 31, twice as far for 16-bit counters)
	size_t ub = ((((size_t)n_seq + 15) & ~(size_t)15) + 128) * (wide ? 2 : 1);
	*u_bytes = (uint32_t)ub;
	const size_t walk = dedupe_bytes + (size_t)4 * (2 * threads);
	const size_t rank = (size_t)8 * RANK_KCAP + (size_t)16 * rec_cap + 4 * 256;
	return ((sizeof(RankShared) + 15) & ~(size_t)15) + ub + ((std::max(walk, rank) + 15) & ~(size_t)15);
}

// Per-thread segment length (in targets): whole 16-byte chunks of counters, an odd number of them,
// so that the 128-bit loads of the eight lanes of a quarter warp fall into different bank groups.
inline uint32_t rank_segment(uint32_t n_seq, bool wide, uint32_t threads)
{
	uint32_t per = (n_seq + threads - 1) / threads;
	uint32_t unit = wide ? 8 : 16;
	uint32_t m = (per + unit - 1) / unit;
	if (m == 0)
		m = 1;
	if ((m & 1) == 0)
		++m;
	return m * unit;
}

} // namespace usb
