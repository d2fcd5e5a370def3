// usb_rankbig.cuh -- kernel K1b: candidate ranking for big databases (more than -big targets),
// the counterpart of UDBUsortedSearcher::UDBSearchBig (udbusortedsearcherbig.cpp:31-135).
//
// Reference behaviour reproduced:
//   a8  GetWordCountingParams: QueryStep from -id and -stepwords   wordparams.cpp:125-192
//       unique query words in first-occurrence order, every QueryStep-th one is counted
//       candidates = touched targets in first-touch order                udbusortedsearcherbig.cpp:82-100
//       CountSortSubsetDesc: drop U < NextValue/2, stable descending       countsort.cpp:110-191
//
// Design: the counter array is too large for shared memory here (one byte or halfword per target,
// millions of targets), so every resident CTA owns a counter array in global memory that stays in
// L2; the few sampled posting rows (about 11 at -id 0.97 for 250 bp reads) are walked by all
// warps with fire-and-forget packed atomic adds.  The reference's order-dependent parts are
// derived without storing the touch order:
//   * first-touch rank of a target = (first sampled row containing it, target index), because
//     rows are walked in word order and are ascending;
//   * NextValue (running max before its last increase, in first-touch order) = max U over the
//     targets that precede p*, the first-touched target holding the global max; those are exactly
//     the entries of the sampled rows before p*'s row plus the entries of that row below p*.
// Survivors (U >= NextValue/2) are usually few: they get their first-row by binary search in the
// sampled rows and are bitonic-sorted by (U desc, first row, target).  When more than 1024
// survive, the first k_max of that order are selected instead (radix-select on U, ties at the
// cut taken row by row in first-touch order).
#pragma once
#include "usb_rank.cuh"

namespace usb {

#define BIG_MAX_POS 4096   // query word positions supported by the big path
#define BIG_MAX_ROWS BIG_MAX_POS // sampled rows per query (QueryStep can be 1)

struct RankBigArgs {
	DevParams P;
	const uint8_t *q;
	const uint64_t *q_off;
	uint32_t n_jobs, strands;
	IndexView ix;              // CSR segments; a row = its fragments in segment order
	uint32_t n_seq;
	uint32_t k_max;
	uint32_t *aux;             // optional, 4 per job: 0, NextValue/2, max U, QueryStep
	uint32_t *cand_t, *cand_u, *n_cand, *n_emit;
	uint32_t *u_out;           // optional: n_jobs * n_seq
	uint8_t *u_arena;          // gridDim.x counter arrays of u_stride bytes
	uint64_t u_stride;         // >= 2 * n_seq, multiple of 16
	uint32_t stepwords;
	uint32_t *rows_out;        // optional: the sampled words of every job, rows_cap per job
	uint32_t *n_rows_out;      // their number (0xffffffff when more than rows_cap)
	uint32_t rows_cap;
	uint32_t variant;          // measurement knob (USB_BIG_VARIANT)
	DevCounters *ctr;
};

struct RankBigShared {
	uint32_t n_uniq, n_rows, row_cur, n_surv, n_sel, found, gmax, kstar, tstar, nextv, vstar, above, m_eq, taken;
	uint32_t n_post;
	uint32_t warp_tmp[32];
	uint32_t hist[256];
	uint32_t words[BIG_MAX_POS];     // word per position (0xffffffff = bad)
	uint32_t uniq[BIG_MAX_POS];      // unique words, first-occurrence order
	uint32_t rows[BIG_MAX_ROWS];     // sampled words
	unsigned long long sel[RANK_KCAP];
};

__device__ __forceinline__ uint32_t block_reduce_max(uint32_t v, uint32_t *warp_tmp)
{
	const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	v = __reduce_max_sync(USB_FULL, v);
	if (lane == 0)
		warp_tmp[w] = v;
	__syncthreads();
	uint32_t r = warp_tmp[lane];
	r = __reduce_max_sync(USB_FULL, r);
	__syncthreads();
	return r;
}

__device__ __forceinline__ uint32_t block_reduce_min(uint32_t v, uint32_t *warp_tmp)
{
	const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	v = __reduce_min_sync(USB_FULL, v);
	if (lane == 0)
		warp_tmp[w] = v;
	__syncthreads();
	uint32_t r = warp_tmp[lane];
	r = __reduce_min_sync(USB_FULL, r);
	__syncthreads();
	return r;
}

// Counter widths of the big-database path: 4 bits when a query counts with at most 15 words (the usual
// case: about 11 sampled words at -id 0.97 for 250 bp reads; halves the array that is zeroed and read
// per query), 8 bits up to 255 words, 16 bits beyond.
#define BIG_NIB 0
#define BIG_BYTE 1
#define BIG_WIDE 2

template <int MODE>
__device__ __forceinline__ uint32_t ug_get(const uint8_t *U, uint32_t t)
{
	if (MODE == BIG_WIDE)
		return (uint32_t)((const volatile uint16_t *)U)[t];
	if (MODE == BIG_BYTE)
		return (uint32_t)((const volatile uint8_t *)U)[t];
	return ((uint32_t)((const volatile uint8_t *)U)[t >> 1] >> ((t & 1) * 4)) & 0xfu;
}

// all bits of every counter of the packed word that is >= thr (1 <= thr <= counter maximum)
template <int MODE>
__device__ __forceinline__ uint32_t ge_mask(uint32_t w, uint32_t thr)
{
	if (MODE == BIG_WIDE)
		return __vcmpgeu2(w, thr * 0x00010001u);
	if (MODE == BIG_BYTE)
		return __vcmpgeu4(w, thr * 0x01010101u);
	const uint32_t t4 = thr * 0x01010101u;
	return (__vcmpgeu4(w & 0x0f0f0f0fu, t4) & 0x0f0f0f0fu) | (__vcmpgeu4((w >> 4) & 0x0f0f0f0fu, t4) & 0xf0f0f0f0u);
}

// number of counters of the packed word equal to v (v >= 1; BIG_NIB / BIG_BYTE)
template <int MODE>
__device__ __forceinline__ uint32_t eq_count(uint32_t w, uint32_t v)
{
	const uint32_t v4 = v * 0x01010101u;
	if (MODE == BIG_BYTE)
		return __popc(__vcmpeq4(w, v4)) / 8;
	return (__popc(__vcmpeq4(w & 0x0f0f0f0fu, v4)) + __popc(__vcmpeq4((w >> 4) & 0x0f0f0f0fu, v4))) / 8;
}

// index of the first sampled row that contains target t (rows ascending), or n_rows
__device__ __forceinline__ uint32_t first_row_of(const RankBigArgs &a, const RankBigShared &S, uint32_t t,
  uint32_t row_limit)
{
	// only the segment whose target range holds t can contain it
	uint32_t sg = 0;
	while (sg + 1 < a.ix.n_seg && t >= a.ix.seg[sg].base + a.ix.seg[sg].count)
		++sg;
	const SegDesc &seg = a.ix.seg[sg];
	for (uint32_t k = 0; k < row_limit; ++k) {
		const uint32_t word = S.rows[k];
		const uint32_t *row = seg.postings + seg.row_off[word];
		const uint32_t size = seg.row_size[word];
		uint32_t lo = 0, hi = size;
		while (lo < hi) {
			const uint32_t mid = (lo + hi) >> 1;
			if (__ldg(row + mid) < t)
				lo = mid + 1;
			else
				hi = mid;
		}
		if (lo < size && __ldg(row + lo) == t)
			return k;
	}
	return row_limit;
}

__device__ __forceinline__ unsigned long long big_key(uint32_t u, uint32_t k, uint32_t t)
{
	return ((unsigned long long)(0xFFFFu - u) << 48) | ((unsigned long long)(k & 0xFFFFu) << 32) | t;
}

template <int MODE>
__device__ void rank_big_job(const RankBigArgs &a, uint32_t job, RankBigShared &S, uint8_t *U, uint32_t n_rows,
  uint32_t step)
{
	constexpr bool WIDE = MODE == BIG_WIDE;
	uint32_t minv_out = 1;
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	constexpr uint32_t NW = RANK_THREADS / 32;
	const uint32_t N = a.n_seq;
	uint32_t *U32 = (uint32_t *)U;
	constexpr uint32_t PER = MODE == BIG_WIDE ? 2 : MODE == BIG_BYTE ? 4 : 8, BITS = 32 / PER, CMASK = (1u << BITS) - 1;
	const uint32_t n_words32 = (N + PER - 1) / PER;
	const uint32_t n128 = (n_words32 + 3) / 4; // (u_stride leaves room for the rounding)

	// ---- zero the counters (global, L2 resident) and the survivor histogram
	{
		uint4 *U128 = (uint4 *)U;
		for (uint32_t i = tid; i < n128; i += RANK_THREADS)
			__stcg(U128 + i, make_uint4(0, 0, 0, 0));
		for (uint32_t i = tid; i < 256; i += RANK_THREADS)
			S.hist[i] = 0;
	}
	__syncthreads();

	// ---- count: every warp takes a part of a sampled row; four postings in flight per lane.  The
	// adds return the old word: the largest value any add produces is the maximum of U, so no pass
	// over the counters is needed for it.
	uint32_t my_max = 0;
	{
		const uint32_t parts = n_rows >= NW || n_rows == 0 ? 1u : min(8u, NW / n_rows);
		uint32_t my_post = 0;
		for (uint32_t item = warp; item < n_rows * parts; item += NW) {
			const uint32_t r = item / parts, part = item - r * parts;
			const uint32_t word = S.rows[r];
			for (uint32_t sg = 0; sg < a.ix.n_seg; ++sg) {
				const SegDesc &seg = a.ix.seg[sg];
				const uint32_t size = seg.row_size[word];
				if (size == 0)
					continue;
				const uint32_t *row = seg.postings + seg.row_off[word];
				const uint32_t b0 = (uint32_t)((uint64_t)size * part / parts);
				const uint32_t b1 = (uint32_t)((uint64_t)size * (part + 1) / parts);
				if (part == 0)
					my_post += size;
				for (uint32_t i0 = b0; i0 < b1; i0 += 128) {
					uint32_t t[4];
#pragma unroll
					for (uint32_t j = 0; j < 4; ++j) {
						const uint32_t i = i0 + 32 * j + lane;
						t[j] = i < b1 ? __ldg(row + i) : 0xffffffffu;
					}
#pragma unroll
					for (uint32_t j = 0; j < 4; ++j)
						if (t[j] != 0xffffffffu) {
							const uint32_t sh = (t[j] & (PER - 1)) * BITS;
							const uint32_t old = atomicAdd(&U32[t[j] / PER], 1u << sh);
							my_max = max(my_max, ((old >> sh) & CMASK) + 1);
						}
				}
			}
		}
		if (lane == 0 && my_post)
			atomicAdd(&S.n_post, my_post);
	}
	__threadfence();
	const uint32_t gmax = block_reduce_max(my_max, S.warp_tmp);

	if (a.u_out)
		for (uint32_t t = tid; t < N; t += RANK_THREADS)
			a.u_out[(uint64_t)job * N + t] = ug_get<MODE>(U, t);

	uint32_t total = 0, nsel = 0;
	if (gmax > 0) {
		// ---- p* = first-touched target holding gmax: first sampled row with such a target, lowest t
		uint32_t kstar = 0, tstar = 0xffffffffu;
		for (uint32_t k = 0; k < n_rows; ++k) {
			const uint32_t word = S.rows[k];
			uint32_t best = 0xffffffffu;
			for (uint32_t sg = 0; sg < a.ix.n_seg; ++sg) {
				const SegDesc &seg = a.ix.seg[sg];
				const uint32_t size = seg.row_size[word];
				const uint32_t *row = seg.postings + seg.row_off[word];
				for (uint32_t i = tid; i < size; i += RANK_THREADS) {
					const uint32_t t = __ldg(row + i);
					if (ug_get<MODE>(U, t) == gmax)
						best = min(best, t);
				}
			}
			best = block_reduce_min(best, S.warp_tmp);
			if (best != 0xffffffffu) {
				kstar = k;
				tstar = best;
				break;
			}
		}
		// ---- NextValue = max U over everything touched before p*
		uint32_t nv = 0;
		for (uint32_t k = 0; k <= kstar; ++k) {
			const uint32_t word = S.rows[k];
			for (uint32_t sg = 0; sg < a.ix.n_seg; ++sg) {
				const SegDesc &seg = a.ix.seg[sg];
				const uint32_t size = seg.row_size[word];
				const uint32_t *row = seg.postings + seg.row_off[word];
				for (uint32_t i = tid; i < size; i += RANK_THREADS) {
					const uint32_t t = __ldg(row + i);
					if (k == kstar && t >= tstar)
						break; // ascending row: the rest of this thread's stride is beyond p* too
					nv = max(nv, ug_get<MODE>(U, t)); // (no target before p* holds gmax)
				}
			}
		}
		const uint32_t nextv = block_reduce_max(nv, S.warp_tmp);
		const uint32_t minv = max(nextv / 2, 1u);
		minv_out = minv;

		// ---- survivors: one pass of 128-bit loads; slots are handed out per warp; the numbers of
		// survivors with U = minv .. minv + 3 are kept in registers, larger values go to the histogram
		{
			const uint4 *U128 = (const uint4 *)U;
			const bool possible = minv <= CMASK;
			uint32_t h[4] = {0, 0, 0, 0};
			for (uint32_t i0 = 0; possible && i0 < n128; i0 += RANK_THREADS) {
				const uint32_t i = i0 + tid;
				uint4 x = make_uint4(0, 0, 0, 0);
				if (i < n128)
					x = __ldcg(U128 + i);
				const uint32_t w4[4] = {x.x, x.y, x.z, x.w};
				uint32_t m[4], c = 0;
#pragma unroll
				for (uint32_t j = 0; j < 4; ++j) {
					m[j] = w4[j] == 0 ? 0u : ge_mask<MODE>(w4[j], minv);
					c += __popc(m[j]) / BITS;
				}
				if (__ballot_sync(USB_FULL, c != 0) == 0)
					continue;
				uint32_t incl = c;
#pragma unroll
				for (uint32_t d = 1; d < 32; d <<= 1) {
					const uint32_t v = __shfl_up_sync(USB_FULL, incl, d);
					if (lane >= d)
						incl += v;
				}
				uint32_t base = 0;
				if (lane == 31)
					base = atomicAdd(&S.n_surv, incl);
				base = __shfl_sync(USB_FULL, base, 31);
				uint32_t slot = base + incl - c;
				if (!WIDE && a.variant == 0) {
					// histogram without a loop over the bytes: U = minv .. minv + 3 by SIMD compares, the
					// (few) larger ones one by one
#pragma unroll
					for (uint32_t j = 0; j < 4; ++j) {
						if (m[j] == 0)
							continue;
#pragma unroll
						for (uint32_t d = 0; d < 4; ++d)
							if (minv + d <= CMASK)
								h[d] += eq_count<WIDE ? BIG_BYTE : MODE>(w4[j], minv + d);
						uint32_t big = minv + 4 <= CMASK ? ge_mask<MODE>(w4[j], minv + 4) : 0u;
						while (big) {
							const uint32_t b = (uint32_t)(__ffs(big) - 1) / BITS;
							big &= ~(CMASK << (b * BITS));
							atomicAdd(&S.hist[(w4[j] >> (b * BITS)) & CMASK], 1u);
						}
					}
				}
				if (base >= RANK_KCAP && a.variant == 0)
					continue; // (warp-uniform) the slots are used up: only the counts matter from here on
#pragma unroll
				for (uint32_t j = 0; j < 4; ++j) {
					uint32_t mask = m[j];
					while (mask) {
						const uint32_t b = (uint32_t)(__ffs(mask) - 1) / BITS;
						mask &= ~(CMASK << (b * BITS));
						const uint32_t u = (w4[j] >> (b * BITS)) & CMASK;
						const uint32_t t = (i * 4 + j) * PER + b;
						if (slot < RANK_KCAP)
							S.sel[slot] = ((unsigned long long)u << 32) | t;
						++slot;
						if (!WIDE && a.variant != 0) {
							const uint32_t d = u - minv;
							h[0] += d == 0;
							h[1] += d == 1;
							h[2] += d == 2;
							h[3] += d == 3;
							if (d >= 4)
								atomicAdd(&S.hist[u], 1u);
						}
					}
				}
			}
			if (!WIDE) {
#pragma unroll
				for (uint32_t d = 0; d < 4; ++d) {
					const uint32_t sum = __reduce_add_sync(USB_FULL, h[d]);
					if (lane == 0 && sum && minv + d <= CMASK)
						atomicAdd(&S.hist[minv + d], sum);
				}
			}
		}
		__syncthreads();
		total = S.n_surv;
		if (total <= RANK_KCAP) {
			for (uint32_t i = tid; i < total; i += RANK_THREADS) {
				const unsigned long long e = S.sel[i];
				const uint32_t t = (uint32_t)e, u = (uint32_t)(e >> 32);
				S.sel[i] = big_key(u, first_row_of(a, S, t, n_rows), t);
			}
			__syncthreads();
			block_sort_keys(S.sel, total);
			nsel = min(total, a.k_max);
		} else {
			// ---- more survivors than slots (the usual case for a query without a close target:
			// NextValue is 1 or 2 then): select the first k_max of (U desc, first-touch order)
			if (tid == 0) {
				S.n_sel = 0;
				S.taken = 0;
			}
			const volatile uint32_t *V = (const volatile uint32_t *)U32;
			if (WIDE) {
				// radix select on U, two levels (the histogram of the survivor pass is not kept here)
				for (uint32_t i = tid; i < 256; i += RANK_THREADS)
					S.hist[i] = 0;
				__syncthreads();
				for (uint32_t i = tid; i < n_words32; i += RANK_THREADS) {
					const uint32_t word = V[i];
					for (uint32_t b = 0; word && b < PER; ++b) {
						const uint32_t u = (word >> (b * BITS)) & CMASK;
						if (u >= minv)
							atomicAdd(&S.hist[u >> 8], 1u);
					}
				}
			}
			__syncthreads();
			if (tid == 0) {
				uint32_t cum = 0;
				for (int b = 255; b >= 0; --b) {
					cum += S.hist[b];
					if (cum >= a.k_max) {
						S.vstar = (uint32_t)b;
						S.above = cum - S.hist[b];
						break;
					}
				}
				S.m_eq = a.k_max - S.above;
			}
			__syncthreads();
			if (WIDE) {
				const uint32_t bstar = S.vstar;
				for (uint32_t i = tid; i < 256; i += RANK_THREADS)
					S.hist[i] = 0;
				__syncthreads();
				for (uint32_t i = tid; i < n_words32; i += RANK_THREADS) {
					const uint32_t word = V[i];
					for (uint32_t b = 0; word && b < PER; ++b) {
						const uint32_t u = (word >> (b * 16)) & 0xffffu;
						if (u >= minv && (u >> 8) == bstar)
							atomicAdd(&S.hist[u & 255], 1u);
					}
				}
				__syncthreads();
				if (tid == 0) {
					uint32_t cum = S.above;
					for (int b = 255; b >= 0; --b) {
						cum += S.hist[b];
						if (cum >= a.k_max) {
							S.vstar = (bstar << 8) | (uint32_t)b;
							S.above = cum - S.hist[b];
							break;
						}
					}
					S.m_eq = a.k_max - S.above;
				}
				__syncthreads();
			}
			const uint32_t vstar = S.vstar, m_eq = S.m_eq;
			// everything above the cut value (fewer than k_max targets)
			if (S.above) {
				const uint4 *U128 = (const uint4 *)U;
				for (uint32_t i = tid; i < n128; i += RANK_THREADS) {
					const uint4 x = __ldcg(U128 + i);
					const uint32_t w4[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
					for (uint32_t j = 0; j < 4; ++j) {
						if (w4[j] == 0 || vstar >= CMASK)
							continue;
						uint32_t mask = ge_mask<MODE>(w4[j], vstar + 1);
						while (mask) {
							const uint32_t b = (uint32_t)(__ffs(mask) - 1) / BITS;
							mask &= ~(CMASK << (b * BITS));
							const uint32_t u = (w4[j] >> (b * BITS)) & CMASK;
							const uint32_t t = (i * 4 + j) * PER + b;
							const uint32_t slot = atomicAdd(&S.n_sel, 1u);
							if (slot < RANK_KCAP)
								S.sel[slot] = big_key(u, first_row_of(a, S, t, n_rows), t);
						}
					}
				}
			}
			__syncthreads();
			// ties at the cut: row by row in first-touch order until m_eq are taken
			for (uint32_t k = 0; k < n_rows && S.taken < m_eq; ++k) {
				const uint32_t word = S.rows[k];
				for (uint32_t sg = 0; sg < a.ix.n_seg; ++sg) {
				const SegDesc &seg = a.ix.seg[sg];
				const uint32_t size = seg.row_size[word];
				const uint32_t *row = seg.postings + seg.row_off[word];
				for (uint32_t base = 0; base < size && S.taken < m_eq; base += RANK_THREADS) {
					const uint32_t i = base + tid;
					uint32_t t = 0;
					bool hit = false;
					if (i < size) {
						t = __ldg(row + i);
						hit = ug_get<MODE>(U, t) == vstar && first_row_of(a, S, t, k) == k;
					}
					const uint32_t rank = block_excl_scan_sum(hit ? 1u : 0u, S.warp_tmp);
					const uint32_t already = S.taken;
					__syncthreads();
					if (hit && already + rank < m_eq) {
						const uint32_t slot = atomicAdd(&S.n_sel, 1u);
						if (slot < RANK_KCAP)
							S.sel[slot] = big_key(vstar, k, t);
						atomicAdd(&S.taken, 1u);
					}
					__syncthreads();
				}
				}
			}
			__syncthreads();
			nsel = min(S.n_sel, (uint32_t)RANK_KCAP);
			block_sort_keys(S.sel, nsel);
		}
	}
	for (uint32_t i = tid; i < nsel; i += RANK_THREADS) {
		const unsigned long long key = S.sel[i];
		a.cand_t[(uint64_t)job * a.k_max + i] = (uint32_t)key;
		if (a.cand_u)
			a.cand_u[(uint64_t)job * a.k_max + i] = 0xFFFFu - (uint32_t)(key >> 48);
	}
	if (tid == 0) {
		a.n_cand[job] = total;
		a.n_emit[job] = nsel;
		if (a.aux) {
			a.aux[4 * job] = 0;
			a.aux[4 * job + 1] = minv_out;
			a.aux[4 * job + 2] = gmax;
			a.aux[4 * job + 3] = step;
		}
		atomicAdd(&a.ctr->postings, (unsigned long long)S.n_post);
	}
	__syncthreads();
}

__global__ void __launch_bounds__(RANK_THREADS, 1) k_rank_big(const RankBigArgs a)
{
	extern __shared__ __align__(16) uint8_t rankbig_smem[];
	RankBigShared &S = *(RankBigShared *)rankbig_smem;
	const uint32_t tid = threadIdx.x;
	uint8_t *U = a.u_arena + (uint64_t)blockIdx.x * a.u_stride;
	const uint32_t WLEN = a.P.word_length;
	for (;;) {
		__syncthreads();
		if (tid == 0)
			S.found = atomicAdd(&a.ctr->job_rank, 1u);
		__syncthreads();
		const uint32_t job = S.found;
		if (job >= a.n_jobs)
			break;
		const uint32_t qi = job / a.strands, strand = job % a.strands;
		const uint64_t q0 = a.q_off[qi];
		const uint32_t L = (uint32_t)(a.q_off[qi + 1] - q0);
		const uint8_t *Q = a.q + q0;
		const uint32_t npos = L >= WLEN ? min(L - WLEN + 1, (uint32_t)BIG_MAX_POS) : 0;
		// ---- words per position
		for (uint32_t p = tid; p < npos; p += RANK_THREADS) {
			uint32_t word = 0, bad = 0;
			for (uint32_t i = 0; i < WLEN; ++i) {
				uint32_t c = strand ? (uint32_t)c_comp[Q[L - 1 - (p + i)]] : (uint32_t)Q[p + i];
				uint32_t l = udb_letter(c);
				bad |= l >> 2;
				word = (word << 2) | (l & 3);
			}
			S.words[p] = bad ? 0xffffffffu : word;
		}
		if (tid == 0) {
			S.n_uniq = 0; S.row_cur = 0; S.n_surv = 0; S.n_sel = 0; S.n_post = 0;
		}
		__syncthreads();
		// ---- unique words in first-occurrence order (udbsearcher.cpp:161-194): ordered compaction
		uint32_t base_cnt = 0;
		for (uint32_t base = 0; base < npos; base += RANK_THREADS) {
			const uint32_t p = base + tid;
			bool first = false;
			uint32_t w = 0xffffffffu;
			if (p < npos) {
				w = S.words[p];
				first = w != 0xffffffffu;
				for (uint32_t qpos = 0; first && qpos < p; ++qpos)
					first = S.words[qpos] != w;
			}
			const uint32_t rank = block_excl_scan_sum(first ? 1u : 0u, S.warp_tmp);
			if (first)
				S.uniq[base_cnt + rank] = w;
			// total of this chunk = rank + flag of the last thread
			if (tid == RANK_THREADS - 1)
				S.n_uniq = base_cnt + rank + (first ? 1u : 0u);
			__syncthreads();
			base_cnt = S.n_uniq;
		}
		const uint32_t nu = base_cnt;
		// ---- QueryStep (wordparams.cpp:125-192), double arithmetic without fused multiply-add
		uint32_t Thresh = 1;
		{
			const double WordFract = __dsub_rn(1.0, __dmul_rn(__dsub_rn(1.0, a.P.id_d), (double)WLEN));
			if (!(WordFract < 0.0)) {
				const double x = __dmul_rn(WordFract, (double)nu);
				Thresh = x < 1.0 ? 1u : (uint32_t)x;
			}
		}
		uint32_t Step = 1;
		if (a.stepwords != 0) {
			Step = Thresh / a.stepwords;
			if (Step == 0)
				Step = 1;
		}
		const uint32_t n_rows = min((nu + Step - 1) / Step, (uint32_t)BIG_MAX_ROWS);
		for (uint32_t r = tid; r < n_rows; r += RANK_THREADS)
			S.rows[r] = S.uniq[r * Step];
		if (a.rows_out) { // the words this query counts with, for the cluster round's conflict kernel
			for (uint32_t r = tid; r < n_rows && r < a.rows_cap; r += RANK_THREADS)
				a.rows_out[(uint64_t)job * a.rows_cap + r] = S.uniq[r * Step];
			if (tid == 0)
				a.n_rows_out[job] = n_rows <= a.rows_cap ? n_rows : 0xffffffffu;
		}
		__syncthreads();
		if (n_rows > 255)
			rank_big_job<BIG_WIDE>(a, job, S, U, n_rows, Step);
		else if (n_rows > 15 || a.variant == 2)
			rank_big_job<BIG_BYTE>(a, job, S, U, n_rows, Step);
		else
			rank_big_job<BIG_NIB>(a, job, S, U, n_rows, Step);
	}
}

} // namespace usb
