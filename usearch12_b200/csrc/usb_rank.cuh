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
// <= 255 word positions, 2 bytes otherwise), so U never touches HBM.  Posting rows start on
// 16-byte boundaries and are streamed with 128-bit loads, four in flight per lane (64 KB in
// flight per SM), and counted with shared-memory atomics on packed 32-bit words.
// The two order-dependent filters are evaluated exactly from the strict prefix maxima of U
// ("records"): records are few, so one thread replays the
// threshold evolution over them; every thread then filters its own contiguous target segment
// against the threshold in force there, four (two) counters per instruction with the byte
// (halfword) SIMD compares.  Survivors (TopOrder) are usually a few hundred: they are collected
// unordered and bitonic-sorted by (U descending, target ascending), which is the reference's
// stable counting-sort order.  Only when more than RANK_KCAP targets survive does the kernel fall
// back to a radix-select of the first k_max of that order (the Terminator can never look
// further than maxaccepts+maxrejects-1 candidates).
#pragma once
#include "usb_dev.cuh"

namespace usb {

#define RANK_THREADS 1024
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
	uint32_t rec_cap;
	double bump_d;             // BumpPct / 100.0 ; 0 = no bump
	DevCounters *ctr;
};

struct RankShared {
	uint32_t n_rows, row_cur, n_rec, n_chg;
	uint32_t maxv, minv, total, vstar, m_eq, n_sel, bstar, above;
	uint32_t take_all, n_post, n_surv, pad2;
	uint32_t warp_tmp[32];
	uint32_t hist[256];
	uint32_t rows[RANK_THREADS];
	unsigned long long sel[RANK_KCAP];
};

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

// Survivor test for target t with count u (u > 0): SetTopBump threshold in force at t (changes
// take effect after the record position that caused them) and the counting-sort cut-off.
struct KeepCursor {
	const uint32_t *chg_pos, *chg_minu;
	uint32_t n_chg, minv, cur;
	__device__ __forceinline__ uint32_t minu(uint32_t t)
	{
		while (cur < n_chg && chg_pos[cur] < t)
			++cur;
		return cur == 0 ? 1u : chg_minu[cur - 1];
	}
	__device__ __forceinline__ bool keep(uint32_t t, uint32_t u) { return u >= minu(t) && u >= minv; }
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
	for (uint32_t i = n + tid; i < P2; i += RANK_THREADS)
		sel[i] = ~0ull;
	__syncthreads();
	for (uint32_t k = 2; k <= P2; k <<= 1)
		for (uint32_t j = k >> 1; j > 0; j >>= 1) {
			for (uint32_t i = tid; i < P2; i += RANK_THREADS) {
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

// Rare path: more than RANK_KCAP survivors.  Radix-select the k_max-th largest surviving count,
// take ties at the cut in ascending target order (block scan), leave the selection in S.sel.
template <bool WIDE>
__device__ void rank_select_fallback(const RankArgs &a, RankShared &S, const uint8_t *U, KeepCursor kc, uint32_t t0,
  uint32_t t1)
{
	const uint32_t tid = threadIdx.x;
	for (uint32_t i = tid; i < 256; i += RANK_THREADS)
		S.hist[i] = 0;
	if (tid == 0)
		S.n_sel = 0;
	__syncthreads();
	kc.cur = 0;
	for (uint32_t t = t0; t < t1; ++t) {
		const uint32_t u = u_get<WIDE>(U, t);
		if (u && kc.keep(t, u))
			atomicAdd(&S.hist[WIDE ? (u >> 8) : u], 1u);
	}
	__syncthreads();
	if (tid == 0) {
		uint32_t cum = 0;
		for (int b = 255; b >= 0; --b) {
			cum += S.hist[b];
			if (cum >= a.k_max) {
				S.bstar = (uint32_t)b;
				S.above = cum - S.hist[b];
				break;
			}
		}
		if (!WIDE) {
			S.vstar = S.bstar;
			S.m_eq = a.k_max - S.above;
		}
	}
	__syncthreads();
	if (WIDE) {
		for (uint32_t i = tid; i < 256; i += RANK_THREADS)
			S.hist[i] = 0;
		__syncthreads();
		kc.cur = 0;
		const uint32_t bstar = S.bstar;
		for (uint32_t t = t0; t < t1; ++t) {
			const uint32_t u = u_get<WIDE>(U, t);
			if (u && (u >> 8) == bstar && kc.keep(t, u))
				atomicAdd(&S.hist[u & 255], 1u);
		}
		__syncthreads();
		if (tid == 0) {
			uint32_t cum = S.above;
			for (int b = 255; b >= 0; --b) {
				cum += S.hist[b];
				if (cum >= a.k_max) {
					S.vstar = (S.bstar << 8) | (uint32_t)b;
					S.m_eq = a.k_max - (cum - S.hist[b]);
					break;
				}
			}
		}
		__syncthreads();
	}
	const uint32_t vstar = S.vstar, m_eq = S.m_eq;
	uint32_t c_eq = 0;
	kc.cur = 0;
	for (uint32_t t = t0; t < t1; ++t) {
		const uint32_t u = u_get<WIDE>(U, t);
		if (u == vstar && u && kc.keep(t, u))
			++c_eq;
	}
	uint32_t eq_rank = block_excl_scan_sum(c_eq, S.warp_tmp);
	kc.cur = 0;
	for (uint32_t t = t0; t < t1; ++t) {
		const uint32_t u = u_get<WIDE>(U, t);
		if (!u || !kc.keep(t, u))
			continue;
		bool take = u > vstar;
		if (!take && u == vstar)
			take = (eq_rank++ < m_eq);
		if (take) {
			const uint32_t slot = atomicAdd(&S.n_sel, 1u);
			if (slot < RANK_KCAP)
				S.sel[slot] = rank_key(u, t);
		}
	}
	__syncthreads();
}

template <bool WIDE>
__device__ void rank_job(const RankArgs &a, uint32_t job, RankShared &S, uint8_t *U, uint32_t *bitmap,
  uint32_t *rec_pos, uint32_t *rec_val, uint32_t *chg_pos, uint32_t *chg_minu)
{
	const uint32_t tid = threadIdx.x, lane = tid & 31;
	const uint32_t N = a.n_seq;
	const uint32_t WLEN = a.P.word_length;
	const uint32_t qi = job / a.strands, strand = job % a.strands;
	const uint64_t q0 = a.q_off[qi];
	const uint32_t L = (uint32_t)(a.q_off[qi + 1] - q0);
	const uint8_t *Q = a.q + q0;
	uint32_t *U32 = (uint32_t *)U;

	// ---- zero shared state (u_bytes is a multiple of 16)
	{
		uint4 *U128 = (uint4 *)U;
		const uint32_t n128 = ((WIDE ? 2 * N : N) + 15) / 16;
		for (uint32_t i = tid; i < n128; i += RANK_THREADS)
			U128[i] = make_uint4(0, 0, 0, 0);
		// nt: bitmap over the 4^w slots; aa: open-addressing hash set of the query's words
		const uint32_t nbm = a.P.alpha == 4 ? a.P.slots / 32 : a.P.hash_cap;
		const uint32_t bm0 = a.P.alpha == 4 ? 0u : 0xffffffffu;
		for (uint32_t i = tid; i < nbm; i += RANK_THREADS)
			bitmap[i] = bm0;
		if (tid == 0) {
			S.n_rows = 0; S.row_cur = 0; S.n_rec = 0; S.n_chg = 0; S.n_sel = 0; S.n_post = 0; S.n_surv = 0;
		}
	}
	__syncthreads();

	// ---- a1/a2/a4: words -> unique rows -> posting walk
	const uint32_t npos = L >= WLEN ? L - WLEN + 1 : 0;
	for (uint32_t base = 0; base < npos; base += RANK_THREADS) {
		const uint32_t p = base + tid;
		if (p < npos) {
			uint32_t word = 0, bad = 0;
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
					if (!(old & bit))
						S.rows[atomicAdd(&S.n_rows, 1u)] = word;
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
							S.rows[atomicAdd(&S.n_rows, 1u)] = word;
							break;
						}
						if (old == word)
							break;
						h = (h + 1) & mask;
					}
				}
			}
		}
		__syncthreads();
		const uint32_t n_rows = S.n_rows;
		uint32_t my_post = 0;
		for (;;) {
			uint32_t r = 0;
			if (lane == 0)
				r = atomicAdd(&S.row_cur, 1u);
			r = __shfl_sync(USB_FULL, r, 0);
			if (r >= n_rows)
				break;
			const uint32_t word = S.rows[r];
			for (uint32_t sg = 0; sg < a.ix.n_seg; ++sg) {
				const SegDesc &seg = a.ix.seg[sg];
				const uint32_t size = seg.row_size[word];
				if (size == 0)
					continue;
				const uint32_t *row = seg.postings + seg.row_off[word];
				const uint4 *v4 = (const uint4 *)row;
				const uint32_t nvec = size >> 2;
				my_post += size;
				uint32_t i = lane;
				for (; i + 96 < nvec; i += 128) {
					const uint4 x0 = __ldg(v4 + i), x1 = __ldg(v4 + i + 32), x2 = __ldg(v4 + i + 64), x3 = __ldg(v4 + i + 96);
					u_inc<WIDE>(U32, x0.x); u_inc<WIDE>(U32, x0.y); u_inc<WIDE>(U32, x0.z); u_inc<WIDE>(U32, x0.w);
					u_inc<WIDE>(U32, x1.x); u_inc<WIDE>(U32, x1.y); u_inc<WIDE>(U32, x1.z); u_inc<WIDE>(U32, x1.w);
					u_inc<WIDE>(U32, x2.x); u_inc<WIDE>(U32, x2.y); u_inc<WIDE>(U32, x2.z); u_inc<WIDE>(U32, x2.w);
					u_inc<WIDE>(U32, x3.x); u_inc<WIDE>(U32, x3.y); u_inc<WIDE>(U32, x3.z); u_inc<WIDE>(U32, x3.w);
				}
				for (; i < nvec; i += 32) {
					const uint4 x = __ldg(v4 + i);
					u_inc<WIDE>(U32, x.x); u_inc<WIDE>(U32, x.y); u_inc<WIDE>(U32, x.z); u_inc<WIDE>(U32, x.w);
				}
				if (lane < (size & 3))
					u_inc<WIDE>(U32, __ldg(row + 4 * nvec + lane));
			}
		}
		if (lane == 0 && my_post)
			atomicAdd(&S.n_post, my_post);
		__syncthreads();
		if (tid == 0) {
			S.n_rows = 0;
			S.row_cur = 0;
		}
		__syncthreads();
	}

	if (a.u_out)
		for (uint32_t t = tid; t < N; t += RANK_THREADS)
			a.u_out[(uint64_t)job * N + t] = u_get<WIDE>(U, t);

	// ---- strict prefix maxima ("records") of U in target order; SIMD max over the segment
	const uint32_t seg = WIDE ? a.seg_wide : a.seg_narrow;
	const uint32_t t0 = min(N, tid * seg), t1 = min(N, t0 + seg);
	const uint32_t PER = WIDE ? 2 : 4;   // counters per 32-bit word
	// segment start is word aligned; threads past the end own nothing
	const uint32_t w0 = t0 / PER, w1 = t0 < t1 ? (t1 + PER - 1) / PER : w0;
	uint32_t m = 0;
	{
		uint32_t acc = 0;
		for (uint32_t k = w0; k < w1; ++k)
			acc = WIDE ? __vmaxu2(acc, U32[k]) : __vmaxu4(acc, U32[k]);
		if (WIDE)
			m = max(acc & 0xffff, acc >> 16);
		else
			m = max(max(acc & 0xff, (acc >> 8) & 0xff), max((acc >> 16) & 0xff, acc >> 24));
	}
	uint32_t run = block_excl_scan_max(m, S.warp_tmp);
	if (m > run)
		for (uint32_t t = t0; t < t1; ++t) {
			uint32_t u = u_get<WIDE>(U, t);
			if (u > run) {
				uint32_t idx = atomicAdd(&S.n_rec, 1u);
				if (idx < a.rec_cap) {
					rec_pos[idx] = t;
					rec_val[idx] = u;
				}
				run = u;
			}
		}
	__syncthreads();
	if (tid == 0) {
		uint32_t n = S.n_rec;
		if (n > a.rec_cap) {
			atomicOr(&a.ctr->err, ERR_RECORDS_FULL);
			n = a.rec_cap;
		}
		for (uint32_t i = 1; i < n; ++i) { // few records: insertion sort by position
			uint32_t p = rec_pos[i], v = rec_val[i];
			uint32_t j = i;
			while (j > 0 && rec_pos[j - 1] > p) {
				rec_pos[j] = rec_pos[j - 1];
				rec_val[j] = rec_val[j - 1];
				--j;
			}
			rec_pos[j] = p;
			rec_val[j] = v;
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

	// ---- survivors of both filters, collected unordered
	KeepCursor kc{chg_pos, chg_minu, S.n_chg, S.minv, 0};
	{
		const uint32_t mu0 = kc.minu(t0);          // threshold at the segment start
		const uint32_t cur0 = kc.cur;
		// constant over the segment unless a change position lies in [t0, t1 - 1)
		const bool constant = !(cur0 < kc.n_chg && chg_pos[cur0] + 1 < t1);
		if (constant) {
			const uint32_t thr = max(max(mu0, kc.minv), 1u);
			if (thr <= (WIDE ? 0xffffu : 0xffu)) {
				const uint32_t thr_v = WIDE ? thr * 0x00010001u : thr * 0x01010101u;
				for (uint32_t k = w0; k < w1; ++k) {
					const uint32_t word = U32[k];
					uint32_t mask = WIDE ? __vcmpgeu2(word, thr_v) : __vcmpgeu4(word, thr_v);
					while (mask) {
						const uint32_t b = (uint32_t)(__ffs(mask) - 1) / (WIDE ? 16 : 8);
						mask &= ~((WIDE ? 0xffffu : 0xffu) << (b * (WIDE ? 16 : 8)));
						const uint32_t u = (word >> (b * (WIDE ? 16 : 8))) & (WIDE ? 0xffffu : 0xffu);
						const uint32_t t = k * PER + b;
						const uint32_t slot = atomicAdd(&S.n_surv, 1u);
						if (slot < RANK_KCAP)
							S.sel[slot] = rank_key(u, t);
					}
				}
			}
		} else {
			kc.cur = 0;
			for (uint32_t t = t0; t < t1; ++t) {
				const uint32_t u = u_get<WIDE>(U, t);
				if (u && kc.keep(t, u)) {
					const uint32_t slot = atomicAdd(&S.n_surv, 1u);
					if (slot < RANK_KCAP)
						S.sel[slot] = rank_key(u, t);
				}
			}
		}
	}
	__syncthreads();
	const uint32_t total = S.n_surv;
	uint32_t nsel;
	if (total <= RANK_KCAP) {
		nsel = min(total, a.k_max);
		block_sort_keys(S.sel, total);
	} else {
		rank_select_fallback<WIDE>(a, S, U, kc, t0, t1);
		nsel = min(S.n_sel, (uint32_t)RANK_KCAP);
		block_sort_keys(S.sel, nsel);
	}
	for (uint32_t i = tid; i < nsel; i += RANK_THREADS) {
		const unsigned long long key = S.sel[i];
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
}

__global__ void __launch_bounds__(RANK_THREADS, 1) k_rank(const RankArgs a)
{
	extern __shared__ __align__(16) uint8_t rank_smem[];
	RankShared &S = *(RankShared *)rank_smem;
	uint8_t *U = rank_smem + ((sizeof(RankShared) + 15) & ~(size_t)15);
	uint32_t *bitmap = (uint32_t *)(U + a.u_bytes);
	uint32_t *rec_pos = bitmap + (a.P.alpha == 4 ? a.P.slots / 32 : a.P.hash_cap);
	uint32_t *rec_val = rec_pos + a.rec_cap;
	uint32_t *chg_pos = rec_val + a.rec_cap;
	uint32_t *chg_minu = chg_pos + a.rec_cap;
	const uint32_t job = blockIdx.x;
	if (job >= a.n_jobs)
		return;
	const uint32_t qi = job / a.strands;
	const uint32_t L = (uint32_t)(a.q_off[qi + 1] - a.q_off[qi]);
	const uint32_t npos = L >= a.P.word_length ? L - a.P.word_length + 1 : 0;
	if (npos > 255)
		rank_job<true>(a, job, S, U, bitmap, rec_pos, rec_val, chg_pos, chg_minu);
	else
		rank_job<false>(a, job, S, U, bitmap, rec_pos, rec_val, chg_pos, chg_minu);
}

// dedupe_bytes: nt = slots / 8 (bitmap), aa = 4 * hash_cap
inline size_t rank_smem_bytes(uint32_t n_seq, bool wide, size_t dedupe_bytes, uint32_t rec_cap, uint32_t *u_bytes)
{
	size_t ub = (((size_t)n_seq * (wide ? 2 : 1)) + 15) & ~(size_t)15;
	*u_bytes = (uint32_t)ub;
	return ((sizeof(RankShared) + 15) & ~(size_t)15) + ub + dedupe_bytes + (size_t)4 * rec_cap * 4;
}

// Per-thread segment length (in targets) such that consecutive threads start in different
// shared-memory banks: seg * width / 4 must be odd.
inline uint32_t rank_segment(uint32_t n_seq, bool wide)
{
	uint32_t per = (n_seq + RANK_THREADS - 1) / RANK_THREADS;
	uint32_t unit = wide ? 2 : 4;
	uint32_t m = (per + unit - 1) / unit;
	if (m == 0)
		m = 1;
	if ((m & 1) == 0)
		++m;
	return m * unit;
}

} // namespace usb
