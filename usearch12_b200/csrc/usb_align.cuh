// usb_align.cuh -- kernel K2: per-(query,strand) candidate loop, one WARP per job.
//
// A warp walks its query's U-sorted candidates strictly in order, exactly like
// UDBUsortedSearcher::SearchImpl (udbusortedsearcher.cpp:138-151) does, so Accepter/Terminator
// semantics need no speculation; the parallelism inside one attempt is the 32 lanes, the
// parallelism across queries is thousands of resident warps pulling jobs from a global cursor.
//
// Reference behaviour reproduced:
//   a10 HSPFinder::SetA / SeqToWords            hspfinder.cpp:226-270,304-331
//   a11 HSPFinder::UngappedBlast, IsGlobalHSP   ungappedblast.cpp:8-211, hspfinder.cpp:594-636
//   a12 Chain + IsStaggered + HSP %id gate      chainer.cpp:352-500, hsp.h:102-126, getglobalhsps.cpp:22-60
//   a9/a13 GlobalAlign_AllOpts, GetHole, AlignHSPMem   globalalignmem.cpp:25-236
//   a14 ViterbiFastMainDiagMem / ViterbiFastBandMem   viterbifastbandmem.cpp:12-253, diagbox.h:150-171
//   a15 TraceBackBitMem                          tracebackbitmem.cpp:8-73
//   a17 AlignResult::FillLo statistics           arscorer.cpp:201-296
//   a7  Accepter::IsAccept (-id) + Terminator    accepter.cpp:27-38, terminator.cpp:64-100
//
// Arithmetic: the reference DP is float with every default score a multiple of 0.5; here all
// scores are doubled integers, which is exact.
//
// Lane mapping: seed search = one target word position per lane; DP = one band column per lane
// with the horizontal-gap state resolved by a max-plus warp scan; path statistics = one
// alignment column per lane with ballot/popc prefix counts.
#pragma once
#include "usb_dev.cuh"
#include "usb_tables.h"

namespace usb {

// Letter tables of the amino acid path (config 1: amino acid -usearch_global, makedbsearcher.cpp:132-140).
// Sequences are staged as the 6-bit character classes of usb_tables.h; substitution scores come
// from BLOSUM62 (blosum62.cpp:17-96), identities from g_MatchMxAmino (alpha2.cpp:250-279), HSP
// words are numbers of hspw letters in base 20 with wildcards as letter 0 (hspfinder.cpp:226-270).
struct AlignTables {
	int8_t score[USB_NCODE * USB_NCODE];
	unsigned long long match[USB_NCODE];
	uint8_t word_letter[USB_NCODE];
	uint8_t code[256];
};

struct AlignArgs {
	DevParams P;
	const uint8_t *q;
	const uint64_t *q_off;
	uint32_t n_jobs, strands;
	const uint32_t *cand_t;   // n_jobs * k_max, from k_rank
	const uint32_t *n_emit;
	uint32_t k_max;
	const uint32_t *pair_q;   // pairs mode (usb_align_pairs) when non-null: job i = pair i
	const uint32_t *pair_t;
	const uint8_t *db_seq;    // 16-byte aligned targets
	const uint64_t *db_off;
	const uint32_t *db_len;
	usb_hit *hits;
	uint32_t hits_cap;
	uint32_t *runs;
	uint32_t runs_cap;
	usb_qstat *qstat;         // per job
	uint8_t *aligned;         // pairs mode
	uint32_t *hsp_out;        // pairs mode, optional
	uint32_t max_hsp;
	uint8_t *slab;            // per-warp global workspace
	uint64_t slab_stride;
	uint32_t ql_cap, tl_cap;  // padded capacities of the per-warp sequence buffers
	uint32_t hsp_cap;
	uint32_t fast_bytes;      // per-warp bytes of the "fast" arrays
	uint32_t scratch_bytes;   // per-warp bytes of the scratch union inside the fast arrays
	uint32_t fast_in_smem;    // 1: fast arrays in shared memory, 0: in the slab
	const AlignTables *tab;   // amino acid path only
	// label identities and size= annotations of queries and targets (Accepter rules that read labels)
	const uint32_t *q_label, *q_size, *t_label, *t_size;
	const uint32_t *n_cand_all; // TopOrder.Size per job when skipped pairs can exhaust the materialised candidates, else null
	DevCounters *ctr;
};

// Per-warp workspace.  "fast" arrays live in shared memory when they fit, else in the slab
// (generic pointers: the code is the same).
struct WarpWs {
	uint8_t *A, *Ac;          // query letters (strand-adjusted) and their nt codes
	const uint8_t *B;         // current target letters (global memory; only read for wildcards)
	uint8_t *Bc;              // current target nt codes
	uint32_t *A2, *An2;       // query packed 2 bits/base (16 per word) and wildcard flags (bit 0 of each pair)
	uint32_t *B2, *Bn2;       // same for the target
	uint8_t *cnt;             // seed table: words' occurrence counts (capped at 8)
	uint16_t *start, *pos;    // seed table CSR
	uint8_t *scratch;         // union: seed-table fill cursors | seed queues | DP rows (when they fit)
	int *rows_slab;           // DP rows for rectangles too wide for the scratch (global memory)
	// slab-only
	uint8_t *TB;              // trace bytes, (LA+1) x (LB+1)
	char *path, *rev;
	HspRec *ung;
	uint32_t *order, *prev, *chain;
	int *cscore;
	uint32_t LA, LB, nwordsA;
	const AlignTables *T;     // amino acid path: tables in shared memory
	bool seed_dirty;          // a wide DP borrowed the seed table's memory: rebuild before the next seed search
};

inline __host__ __device__ uint32_t pad16(uint32_t x) { return (x + 15u) & ~15u; }

inline __host__ __device__ uint32_t align_fast_bytes(uint32_t ql_cap, uint32_t tl_cap, uint32_t hsp_words)
{
	return 2 * ql_cap + tl_cap + 2 * pad16(ql_cap / 4 + 8) + 2 * pad16(tl_cap / 4 + 8) + pad16(hsp_words) +
	  pad16(2 * hsp_words) + pad16(2 * ql_cap);
}

inline __host__ __device__ uint64_t align_slab_bytes(uint32_t ql_cap, uint32_t tl_cap, uint32_t hsp_cap)
{
	uint64_t tb = ((uint64_t)(ql_cap + 1) * (tl_cap + 1) + 15) & ~(uint64_t)15;
	uint64_t path = pad16(ql_cap + tl_cap + 16);
	return tb + 2 * path + (uint64_t)hsp_cap * (sizeof(HspRec) + 16) + 2 * pad16(4 * (tl_cap + 8));
}

__device__ __forceinline__ void ws_setup(const AlignArgs &a, WarpWs &w, uint8_t *fast, uint8_t *slab)
{
	uint8_t *p = fast;
	w.A = p; p += a.ql_cap;
	w.Ac = p; p += a.ql_cap;
	w.Bc = p; p += a.tl_cap;
	w.A2 = (uint32_t *)p; p += pad16(a.ql_cap / 4 + 8);
	w.An2 = (uint32_t *)p; p += pad16(a.ql_cap / 4 + 8);
	w.B2 = (uint32_t *)p; p += pad16(a.tl_cap / 4 + 8);
	w.Bn2 = (uint32_t *)p; p += pad16(a.tl_cap / 4 + 8);
	w.cnt = p; p += pad16(a.P.hsp_words);
	w.start = (uint16_t *)p; p += pad16(2 * a.P.hsp_words);
	w.pos = (uint16_t *)p; p += pad16(2 * a.ql_cap);
	w.scratch = p;
	uint8_t *s = slab;
	w.TB = s; s += ((uint64_t)(a.ql_cap + 1) * (a.tl_cap + 1) + 15) & ~(uint64_t)15;
	w.path = (char *)s; s += pad16(a.ql_cap + a.tl_cap + 16);
	w.rev = (char *)s; s += pad16(a.ql_cap + a.tl_cap + 16);
	w.ung = (HspRec *)s; s += (uint64_t)a.hsp_cap * sizeof(HspRec);
	w.order = (uint32_t *)s; s += (uint64_t)a.hsp_cap * 4;
	w.prev = (uint32_t *)s; s += (uint64_t)a.hsp_cap * 4;
	w.chain = (uint32_t *)s; s += (uint64_t)a.hsp_cap * 4;
	w.cscore = (int *)s; s += (uint64_t)a.hsp_cap * 4;
	w.rows_slab = (int *)s;
}

// %id identity of query position qp vs target position tp; raw letters are only needed (and the
// target's only fetched from global memory) when a wildcard is involved.
template <bool AA> __device__ __forceinline__ bool pos_match(const WarpWs &w, uint32_t qp, uint32_t tp)
{
	const uint32_t ca = w.Ac[qp], cb = w.Bc[tp];
	if constexpr (AA)
		return (w.T->match[ca] >> cb) & 1ull;
	if ((ca | cb) < 4)
		return ca == cb;
	return chars_match_dev(w.A[qp], w.B[tp], ca, cb);
}

// doubled substitution score of two staged letters
template <bool AA> __device__ __forceinline__ int subst2g(const AlignArgs &a, const WarpWs &w, uint32_t ca, uint32_t cb)
{
	if constexpr (AA)
		return 2 * (int)w.T->score[ca * USB_NCODE + cb];
	else
		return subst2(a.P, ca, cb);
}

// ------------------------------------------------------------------ sequence staging
template <bool AA>
__device__ __forceinline__ void load_query(const AlignArgs &a, WarpWs &w, const uint8_t *Q, uint32_t L, uint32_t strand)
{
	const uint32_t lane = lane_id();
	__syncwarp(); // lanes may still be reading the previous query's arrays
	if constexpr (AA) {
		for (uint32_t i = lane; i < L; i += 32) {
			const uint32_t c = Q[i];
			w.A[i] = (uint8_t)c;
			w.Ac[i] = w.T->code[c];
		}
		w.LA = L;
		__syncwarp();
		return;
	}
	for (uint32_t i = lane; i < L; i += 32) {
		uint32_t c = strand ? (uint32_t)c_comp[Q[L - 1 - i]] : (uint32_t)Q[i];
		w.A[i] = (uint8_t)c;
		w.Ac[i] = (uint8_t)nt_code(c);
	}
	w.LA = L;
	__syncwarp();
	const uint32_t nw16 = (L + 15) / 16;
	for (uint32_t k = lane; k < nw16 + 2; k += 32) {
		uint32_t v = 0, n = 0;
		if (k < nw16)
			for (uint32_t j = 0; j < 16 && 16 * k + j < L; ++j) {
				const uint32_t c = w.Ac[16 * k + j];
				v |= (c & 3) << (2 * j);
				n |= (c >> 2) << (2 * j);
			}
		w.A2[k] = v & ~(n * 3);
		w.An2[k] = n;
	}
	__syncwarp();
}

template <bool AA> __device__ __forceinline__ void load_target(const AlignArgs &a, WarpWs &w, uint32_t t)
{
	const uint32_t lane = lane_id();
	const uint32_t L = a.db_len[t];
	__syncwarp(); // lanes may still be reading the previous target's arrays
	w.B = a.db_seq + a.db_off[t];
	if constexpr (AA) {
		for (uint32_t i = lane; i < L; i += 32)
			w.Bc[i] = w.T->code[w.B[i]];
		w.LB = L;
		__syncwarp();
		return;
	}
	const uint4 *src = (const uint4 *)w.B;
	uint4 *dC = (uint4 *)w.Bc;
	const uint32_t n16 = (L + 15) / 16;
	for (uint32_t i = lane; i < n16 + 2; i += 32) {
		uint32_t p2 = 0, pn = 0;
		if (i < n16) {
			uint4 v = __ldg(src + i);
			uint32_t in[4] = {v.x, v.y, v.z, v.w}, out[4];
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				uint32_t x = in[k];
#pragma unroll
				for (int b = 0; b < 4; ++b) {
					const uint32_t c = nt_code((x >> (8 * b)) & 0xff);
					out[k] = b == 0 ? c : (out[k] | (c << (8 * b)));
					p2 |= (c & 3) << (2 * (4 * k + b));
					pn |= (c >> 2) << (2 * (4 * k + b));
				}
			}
			dC[i] = make_uint4(out[0], out[1], out[2], out[3]);
		}
		w.B2[i] = p2 & ~(pn * 3);
		w.Bn2[i] = pn;
	}
	w.LB = L;
	__syncwarp();
}

// ------------------------------------------------------------------ a10: seed table of the query
// HSP words treat wildcards as letter 0 and are never dropped (hspfinder.cpp:238-240).
// 16 bases starting at position p of a packed array (two aligned word loads + funnel shift).
__device__ __forceinline__ uint32_t ext16(const uint32_t *X, uint32_t p)
{
	const uint32_t k = p >> 4;
	return __funnelshift_r(X[k], X[k + 1], (p & 15) * 2);
}

// The word value is only a table key shared by query and target side (first base in the low
// bits); wildcards are packed as letter 0, so they behave like 'A' exactly as in the reference.
__device__ __forceinline__ uint32_t hsp_word_at(const uint32_t *X2, uint32_t p, uint32_t hsp_words)
{
	return ext16(X2, p) & (hsp_words - 1);
}

// HSP word at position p of the query (which = 0) or the target (1)
template <bool AA> __device__ __forceinline__ uint32_t word_at(const AlignArgs &a, const WarpWs &w, int which, uint32_t p)
{
	if constexpr (AA) {
		const uint8_t *c = (which ? w.Bc : w.Ac) + p;
		uint32_t word = 0;
		for (uint32_t i = 0; i < a.P.hspw; ++i)
			word = word * 20u + w.T->word_letter[c[i]];
		return word;
	} else
		return hsp_word_at(which ? w.B2 : w.A2, p, a.P.hsp_words);
}

// Builds cnt[word] = min(8, occurrences) and, per word, the first 8 query positions in query
// order (hspfinder.cpp:304-323), as a CSR (start, pos).  Query order inside a word is kept by
// processing positions in chunks of 32 and ranking equal words inside a chunk with match_any.
template <bool AA> __device__ void build_seed_table(const AlignArgs &a, WarpWs &w)
{
	const uint32_t lane = lane_id();
	const uint32_t hw = a.P.hspw, HW = a.P.hsp_words;
	uint8_t *fil = w.scratch;
	uint32_t *cnt32 = (uint32_t *)w.cnt, *fil32 = (uint32_t *)fil;
	for (uint32_t i = lane; i < HW / 4; i += 32) {
		cnt32[i] = 0;
		fil32[i] = 0;
	}
	const uint32_t nw = w.LA >= hw ? w.LA - hw + 1 : 0;
	w.nwordsA = nw;
	__syncwarp();
	for (uint32_t base = 0; base < nw; base += 32) {
		const uint32_t p = base + lane;
		const bool valid = p < nw;
		const uint32_t word = valid ? word_at<AA>(a, w, 0, p) : (0x80000000u | lane);
		const uint32_t peers = __match_any_sync(USB_FULL, word);
		if (valid && (peers & lanemask_lt()) == 0) {
			uint32_t c = w.cnt[word] + __popc(peers);
			w.cnt[word] = (uint8_t)min(c, 8u);
		}
		__syncwarp();
	}
	// exclusive scan of cnt -> start
	{
		const uint32_t per = HW / 32;
		uint32_t sum = 0;
		for (uint32_t i = 0; i < per; ++i)
			sum += w.cnt[lane * per + i];
		uint32_t inc = sum;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			uint32_t t = __shfl_up_sync(USB_FULL, inc, d);
			if (lane >= (uint32_t)d)
				inc += t;
		}
		uint32_t run = inc - sum;
		for (uint32_t i = 0; i < per; ++i) {
			w.start[lane * per + i] = (uint16_t)run;
			run += w.cnt[lane * per + i];
		}
	}
	__syncwarp();
	for (uint32_t base = 0; base < nw; base += 32) {
		const uint32_t p = base + lane;
		const bool valid = p < nw;
		const uint32_t word = valid ? word_at<AA>(a, w, 0, p) : (0x80000000u | lane);
		const uint32_t peers = __match_any_sync(USB_FULL, word);
		uint32_t before = 0;
		if (valid)
			before = fil[word];
		__syncwarp();
		if (valid) {
			uint32_t slot = before + __popc(peers & lanemask_lt());
			if (slot < 8)
				w.pos[w.start[word] + slot] = (uint16_t)p;
			if ((peers & lanemask_lt()) == 0)
				fil[word] = (uint8_t)min(before + __popc(peers), 8u);
		}
		__syncwarp();
	}
}

// ------------------------------------------------------------------ a11: ungapped seeds + X-drop
// hspfinder.cpp:594-636
__device__ __forceinline__ bool is_global_hsp(uint32_t ALo, uint32_t BLo, uint32_t LA, uint32_t LB)
{
	const uint32_t AR = LA - ALo, BR = LB - BLo;
	if (LA <= LB) {
		const uint32_t MaxGap = LA / 4 + 1;
		if (ALo > BLo && ALo - BLo > MaxGap)
			return false;
		if (AR > BR && AR - BR > MaxGap)
			return false;
	} else {
		const uint32_t MaxGap = LB / 4 + 1;
		if (BLo > ALo && BLo - ALo > MaxGap)
			return false;
		if (BR > AR && BR - AR > MaxGap)
			return false;
	}
	return true;
}

// One lane extends one seed letter by letter (ungappedblast.cpp:76-193): right from the seed's
// last letter, then left from its first letter starting at the best right-extended score.
// Scores are doubled.  Generic form, used when the score signs do not allow the packed walk.
template <bool AA>
__device__ __forceinline__ void extend_seed_bytes(const AlignArgs &a, const WarpWs &w, uint32_t APos, uint32_t BPos,
  int &best_out, uint32_t &best_lo_out, uint32_t &best_hi_out)
{
	const DevParams &P = a.P;
	const uint32_t LA = w.LA, LB = w.LB, hw = P.hspw;
	const uint8_t *Ac = w.Ac, *Bc = w.Bc;
	int score = 0;
	for (uint32_t j = 0; j < hw; ++j)
		score += subst2g<AA>(a, w, Ac[APos + j], Bc[BPos + j]);
	int best = score;
	uint32_t bp = BPos + hw - 1, ap = APos + hw - 1, best_hi = bp;
	for (;;) {
		++bp;
		if (bp >= LB)
			break;
		++ap;
		if (ap >= LA)
			break;
		score += subst2g<AA>(a, w, Ac[ap], Bc[bp]);
		if (score > best) {
			best = score;
			best_hi = bp;
		} else if ((float)(best - score) > P.xdrop2)
			break;
	}
	bp = BPos;
	ap = APos;
	uint32_t best_lo = BPos;
	score = best;
	for (;;) {
		if (bp == 0 || ap == 0)
			break;
		--bp;
		--ap;
		score += subst2g<AA>(a, w, Ac[ap], Bc[bp]);
		if (score > best) {
			best = score;
			best_lo = bp;
		} else if ((float)(best - score) > P.xdrop2)
			break;
	}
	best_out = best;
	best_lo_out = best_lo;
	best_hi_out = best_hi;
}

#define M55 0x55555555u

// Same walk on the 2-bit packed sequences, 16 letters per step (requires match > 0 > mismatch;
// wildcard pairs score 0).  Inside a run of matches the score only rises, so the run can be
// applied at once: it cannot trigger the X-drop test and, if it ends above the best score, its
// last letter is the new best end.  A wildcard pair changes nothing.  A mismatch is the only
// step that can terminate.  The result is identical to extend_seed_bytes.
template <bool WILD>
__device__ __forceinline__ void extend_seed_packed(const AlignArgs &a, const WarpWs &w, uint32_t APos, uint32_t BPos,
  int seed2, int &best_out, uint32_t &best_lo_out, uint32_t &best_hi_out)
{
	const DevParams &P = a.P;
	const uint32_t LA = w.LA, LB = w.LB, hw = P.hspw;
	int score = seed2, best = seed2;
	uint32_t best_hi = BPos + hw - 1;
	{
		uint32_t pa = APos + hw, pb = BPos + hw;
		for (;;) {
			const uint32_t k = min(16u, min(LA - pa, LB - pb));
			if (k == 0)
				break;
			const uint32_t x = ext16(w.A2, pa) ^ ext16(w.B2, pb);
			const uint32_t wl = WILD ? (ext16(w.An2, pa) | ext16(w.Bn2, pb)) & M55 : 0u;
			uint32_t stop = ((x | (x >> 1)) & M55) | wl; // pairs that are not a definite match
			uint32_t done = 0;
			bool term = false;
			while (done < k) {
				const uint32_t nxt = min(k, stop ? (uint32_t)(__ffs(stop) - 1) >> 1 : 16u);
				if (nxt > done) {
					score += (int)(nxt - done) * P.match2;
					if (score > best) {
						best = score;
						best_hi = pb + nxt - 1;
					}
				}
				if (nxt >= k)
					break;
				if (!((wl >> (2 * nxt)) & 1)) {
					score += P.mismatch2;
					if ((float)(best - score) > P.xdrop2) {
						term = true;
						break;
					}
				}
				done = nxt + 1;
				stop &= stop - 1;
			}
			if (term || k < 16)
				break;
			pa += 16;
			pb += 16;
		}
	}
	uint32_t best_lo = BPos;
	score = best;
	{
		uint32_t pa = APos, pb = BPos; // letters available to the left
		for (;;) {
			const uint32_t k = min(16u, min(pa, pb));
			if (k == 0)
				break;
			const uint32_t sh = 32 - 2 * k;
			const uint32_t x = ext16(w.A2, pa - k) ^ ext16(w.B2, pb - k);
			const uint32_t wl = WILD ? ((ext16(w.An2, pa - k) | ext16(w.Bn2, pb - k)) & M55) << sh : 0u;
			uint32_t stop = (((x | (x >> 1)) & M55) << sh) | wl; // nearest letter at bit 30
			uint32_t done = 0;
			bool term = false;
			while (done < k) {
				const uint32_t nxt = min(k, (uint32_t)__clz(stop) >> 1);
				if (nxt > done) {
					score += (int)(nxt - done) * P.match2;
					if (score > best) {
						best = score;
						best_lo = pb - nxt;
					}
				}
				if (nxt >= k)
					break;
				const uint32_t bit = 0x40000000u >> (2 * nxt);
				if (!(wl & bit)) {
					score += P.mismatch2;
					if ((float)(best - score) > P.xdrop2) {
						term = true;
						break;
					}
				}
				done = nxt + 1;
				stop &= ~bit;
			}
			if (term || k < 16)
				break;
			pa -= 16;
			pb -= 16;
		}
	}
	best_out = best;
	best_lo_out = best_lo;
	best_hi_out = best_hi;
}

// Cheap exact rejection of a seed before walking it.  (1) IsGlobalHSP only depends on the
// diagonal, which the extension never leaves.  (2) Look at the 16 letter pairs on each side of
// the seed: if a side has so few possible matches that the X-drop walk must stop inside the
// window (or the sequence ends there), the gain on that side is at most its match count; when
// both sides are bounded like that and seed + bounds < MinGlobalHSPScore the seed can never be
// accepted.  About 85 % of random seeds end here.
template <bool WILD>
__device__ __forceinline__ bool seed_may_pass_p(const AlignArgs &a, const WarpWs &w, uint32_t ap, uint32_t bp, int &seed2)
{
	const DevParams &P = a.P;
	const uint32_t LA = w.LA, LB = w.LB, hw = P.hspw;
	if (!is_global_hsp(ap, bp, LA, LB))
		return false;
	const uint32_t seedmask = M55 & (P.hsp_words - 1);
	const uint32_t wseed = WILD ? (ext16(w.An2, ap) | ext16(w.Bn2, bp)) & seedmask : 0u;
	seed2 = P.match2 * (int)(hw - __popc(wseed));
	const uint32_t ra = ap + hw, rb = bp + hw;
	const uint32_t kR = min(16u, min(LA - ra, LB - rb));
	uint32_t mR = 0;
	if (kR) {
		const uint32_t x = ext16(w.A2, ra) ^ ext16(w.B2, rb);
		const uint32_t wl = WILD ? ext16(w.An2, ra) | ext16(w.Bn2, rb) : 0u;
		const uint32_t mis = ((x | (x >> 1)) & ~wl & M55) & (kR == 16 ? M55 : ((1u << (2 * kR)) - 1));
		mR = kR - __popc(mis);
	}
	const uint32_t kL = min(16u, min(ap, bp));
	uint32_t mL = 0;
	if (kL) {
		const uint32_t x = ext16(w.A2, ap - kL) ^ ext16(w.B2, bp - kL);
		const uint32_t wl = WILD ? ext16(w.An2, ap - kL) | ext16(w.Bn2, bp - kL) : 0u;
		const uint32_t mis = ((x | (x >> 1)) & ~wl & M55) & (kL == 16 ? M55 : ((1u << (2 * kL)) - 1));
		mL = kL - __popc(mis);
	}
	const bool deadR = kR < 16 || (float)(-((int)mR * P.match2 + (int)(16 - mR) * P.mismatch2)) > P.xdrop2;
	const bool deadL = kL < 16 || (float)(-((int)mL * P.match2 + (int)(16 - mL) * P.mismatch2)) > P.xdrop2;
	if (deadR && deadL && (float)(seed2 + (int)(mR + mL) * P.match2) < P.minscore2)
		return false;
	return true;
}

__device__ __forceinline__ bool seed_may_pass(const AlignArgs &a, const WarpWs &w, uint32_t ap, uint32_t bp, int &seed2)
{
	return seed_may_pass_p<true>(a, w, ap, bp, seed2);
}

// Acceptance test of a seed extended on the packed sequences (ungappedblast.cpp:172-193).
template <bool WILD>
__device__ __forceinline__ bool extend_seed_p(const AlignArgs &a, const WarpWs &w, uint32_t APos, uint32_t BPos, int seed2,
  uint32_t MinLength, HspRec &out, uint32_t &Bhi_out)
{
	int best;
	uint32_t best_lo, best_hi;
	extend_seed_packed<WILD>(a, w, APos, BPos, seed2, best, best_lo, best_hi);
	const uint32_t Length = best_hi - best_lo + 1;
	const uint32_t Alo = best_lo - (BPos - APos); // same diagonal; wrap-around arithmetic is exact
	if (Length < MinLength || (float)best < a.P.minscore2)
		return false;
	if (!is_global_hsp(Alo, best_lo, w.LA, w.LB))
		return false;
	out.Loi = Alo;
	out.Loj = best_lo;
	out.Len = Length;
	out.score2 = best;
	Bhi_out = best_hi;
	return true;
}

// Acceptance test of an extended seed (ungappedblast.cpp:172-193).
template <bool AA>
__device__ __forceinline__ bool extend_seed(const AlignArgs &a, const WarpWs &w, uint32_t APos, uint32_t BPos,
  int seed2, bool packed, uint32_t MinLength, HspRec &out, uint32_t &Bhi_out)
{
	int best;
	uint32_t best_lo, best_hi;
	if (!AA && packed)
		extend_seed_packed<true>(a, w, APos, BPos, seed2, best, best_lo, best_hi);
	else
		extend_seed_bytes<AA>(a, w, APos, BPos, best, best_lo, best_hi);
	const uint32_t Length = best_hi - best_lo + 1;
	const uint32_t Alo = best_lo - (BPos - APos); // same diagonal; wrap-around arithmetic is exact
	if (Length < MinLength || (float)best < a.P.minscore2)
		return false;
	if (!is_global_hsp(Alo, best_lo, w.LA, w.LB))
		return false;
	out.Loi = Alo;
	out.Loj = best_lo;
	out.Len = Length;
	out.score2 = best;
	Bhi_out = best_hi;
	return true;
}

// Seed queues inside the per-warp scratch: q1 = target word positions that have seeds, q2 =
// seeds that survived the pre-filter, both in reference scan order.
#define SEEDQ1 384
#define SEEDQ2 512
#define SEED_SCRATCH_BYTES (8 * SEEDQ1 + 6 * SEEDQ2)

// Extends the queued survivors 32 at a time and accepts in order (see ungapped_blast).
// (Measured and dropped: letting only the first seed of every diagonal walk and the others wait
// -- the walks of a batch run in lockstep, so a batch costs its longest walk whether one or twenty
// lanes walk that diagonal, and every waiting seed that is needed after all costs a second round:
// k_align 347 -> 473 ms.)
template <bool AA>
__device__ __forceinline__ void extend_queued(const AlignArgs &a, WarpWs &w, const uint32_t *q2b, const uint16_t *q2a,
  uint32_t n2, bool packed, uint32_t MinLength, uint32_t &cur, uint32_t &nung)
{
	const uint32_t lane = lane_id();
	const DevParams &P = a.P;
	for (uint32_t s0 = 0; s0 < n2; s0 += 32) {
		const uint32_t s = s0 + lane;
		HspRec h;
		h.Loi = h.Loj = h.Len = 0;
		h.score2 = 0;
		uint32_t bhi = 0, bp = 0;
		bool ok = false;
		if (s < n2) {
			bp = q2b[s];
			const uint32_t ap = q2a[s];
			if (bp >= cur) {
				int seed2 = 0;
				if (!AA && packed) {
					const uint32_t seedmask = M55 & (P.hsp_words - 1);
					seed2 = P.match2 * (int)(P.hspw - __popc((ext16(w.An2, ap) | ext16(w.Bn2, bp)) & seedmask));
				}
				ok = extend_seed<AA>(a, w, ap, bp, seed2, packed, MinLength, h, bhi);
			}
		}
		uint32_t okmask = __ballot_sync(USB_FULL, ok);
		while (okmask) {
			const int src = __ffs(okmask) - 1;
			okmask &= okmask - 1;
			const uint32_t sbp = __shfl_sync(USB_FULL, bp, src);
			if (sbp < cur)
				continue; // inside an HSP accepted a moment ago: the scan never sees this seed
			HspRec g;
			g.Loi = __shfl_sync(USB_FULL, h.Loi, src);
			g.Loj = __shfl_sync(USB_FULL, h.Loj, src);
			g.Len = __shfl_sync(USB_FULL, h.Len, src);
			g.score2 = __shfl_sync(USB_FULL, h.score2, src);
			const uint32_t sbhi = __shfl_sync(USB_FULL, bhi, src);
			if (nung < a.hsp_cap) {
				if (lane == 0)
					w.ung[nung] = g;
				++nung;
			} else if (lane == 0)
				atomicOr(&a.ctr->err, ERR_HSP_FULL);
			cur = sbhi + 1;
		}
	}
	__syncwarp();
}

// Returns the number of ungapped HSPs written to w.ung (ungappedblast.cpp:45-210).  Reference
// order: target word positions ascending; at each, the query positions of that word in query
// order; the first seed that yields an acceptable HSP wins and the scan jumps past that HSP.
// Here: (1) the target positions whose word occurs in the query are queued in order (one ballot
// per 32 positions), (2) every lane takes one queued position and pre-filters its seeds; the
// survivors are queued in the same (position, query order) order, (3) survivors are extended 32
// at a time.  An extension does not depend on scan history, so the sequential result is: walk
// the survivors in order, accept the first acceptable one at or after the current scan
// position, move the scan position past it, continue.
template <bool AA> __device__ uint32_t ungapped_blast(const AlignArgs &a, WarpWs &w, uint32_t MinLength)
{
	const uint32_t lane = lane_id();
	const DevParams &P = a.P;
	const uint32_t hw = P.hspw, HW = P.hsp_words;
	const uint32_t LB = w.LB;
	if (LB < 2 * hw)
		return 0;
	const bool packed = !AA && P.match2 > 0 && P.mismatch2 < 0;
	const uint32_t nwordsB = LB - hw + 1;
	uint32_t *q1b = (uint32_t *)w.scratch;
	uint32_t *q1s = q1b + SEEDQ1;
	uint32_t *q2b = q1s + SEEDQ1;
	uint16_t *q2a = (uint16_t *)(q2b + SEEDQ2);
	uint32_t nung = 0;
	uint32_t scan = 0; // next target word position to queue
	uint32_t cur = 0;  // sequential scan position (seeds before it are never examined)
	while (scan < nwordsB) {
		uint32_t n1 = 0;
		while (scan < nwordsB && n1 + 32 <= SEEDQ1) {
			const uint32_t bpos = scan + lane;
			uint32_t na = 0, word = 0;
			if (bpos < nwordsB) {
				word = word_at<AA>(a, w, 1, bpos);
				na = w.cnt[word];
			}
			const uint32_t m = __ballot_sync(USB_FULL, na != 0);
			if (na) {
				const uint32_t d = n1 + __popc(m & lanemask_lt());
				q1b[d] = bpos;
				q1s[d] = (uint32_t)w.start[word] | (na << 16);
			}
			n1 += __popc(m);
			scan += 32;
		}
		__syncwarp();
		uint32_t n2 = 0;
		for (uint32_t e0 = 0; e0 < n1; e0 += 32) {
			if (n2 + 256 > SEEDQ2) {
				extend_queued<AA>(a, w, q2b, q2a, n2, packed, MinLength, cur, nung);
				n2 = 0;
			}
			const uint32_t e = e0 + lane;
			uint32_t bp = 0, st = 0, na = 0;
			if (e < n1) {
				bp = q1b[e];
				const uint32_t v = q1s[e];
				st = v & 0xffff;
				na = bp >= cur ? v >> 16 : 0;
			}
			uint32_t keepbits = 0;
			for (uint32_t i = 0; i < na; ++i) {
				int seed2;
				if (!packed || seed_may_pass(a, w, w.pos[st + i], bp, seed2))
					keepbits |= 1u << i;
			}
			// exclusive prefix of the per-lane survivor counts (0..8) from four ballots
			const uint32_t c = __popc(keepbits), lt = lanemask_lt();
			const uint32_t b0 = __ballot_sync(USB_FULL, c & 1), b1 = __ballot_sync(USB_FULL, c & 2),
			               b2 = __ballot_sync(USB_FULL, c & 4), b3 = __ballot_sync(USB_FULL, c & 8);
			uint32_t d = n2 + __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt) + 8 * __popc(b3 & lt);
			while (keepbits) {
				const uint32_t i = __ffs(keepbits) - 1;
				keepbits &= keepbits - 1;
				q2b[d] = bp;
				q2a[d] = w.pos[st + i];
				++d;
			}
			n2 += __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2) + 8 * __popc(b3);
		}
		__syncwarp();
		extend_queued<AA>(a, w, q2b, q2a, n2, packed, MinLength, cur, nung);
		__syncwarp(); // the queue is refilled next: every lane is done reading it
		scan = max(scan, cur);
	}
	__syncwarp();
	return nung;
}

// ------------------------------------------------------------------ a12: chaining
// hsp.h:102-126 (three of the four terminal-gap terms are clamped, as in the reference)
__device__ __forceinline__ bool is_staggered(const HspRec &h, uint32_t LA, uint32_t LB)
{
	int Hii = (int)(h.Loi + h.Len - 1), Hij = (int)(h.Loj + h.Len - 1);
	int gla = (int)h.Loi - (int)h.Loj;
	int glb = (int)h.Loj - (int)h.Loi;
	int gra = ((int)LA - Hii - 1) - ((int)LB - Hij - 1);
	int grb = ((int)LB - Hij - 1) - ((int)LA - Hii - 1);
	if (gla < 0) gla = 0;
	if (glb < 0) glb = 0;
	if (grb < 0) grb = 0;
	int GapA = gla + gra, GapB = glb + grb;
	if (GapA == 0 || GapB == 0)
		return false;
	double r = LA < LB ? (double)GapA / (double)LA : (double)GapB / (double)LB;
	return r > 0.5;
}

// Best colinear chain (chainer.cpp:352-500; the "dominated chain" pruning there never fires, so
// this is the plain O(K^2) recurrence over HSPs ordered by query start, stable).  K is small
// (mean 2.9): lane 0 does it.  Returns the chain length; indices into w.ung in w.chain.
__device__ uint32_t chain_hsps(const AlignArgs &a, WarpWs &w, uint32_t n)
{
	uint32_t len = 0;
	const uint32_t lane = lane_id();
	if (n > 0 && n <= 32) {
		// Up to 32 HSPs (always, in practice): lane k holds HSP k in registers and the recurrence
		// runs on shuffles and ballots -- the arrays of the serial version live in global memory
		// and every step of it was a dependent load.
		const bool have = lane < n;
		HspRec h = have ? w.ung[lane] : HspRec{0, 0, 0, 0};
		// stable order by query start: rank = HSPs that come before this one
		uint32_t rank = 0;
		for (uint32_t j = 0; j < n; ++j) {
			const uint32_t lj = __shfl_sync(USB_FULL, h.Loi, j);
			rank += (lj < h.Loi || (lj == h.Loi && j < lane)) ? 1u : 0u;
		}
		int cs = 0;
		uint32_t prev = 0xffffffffu;
		for (uint32_t oi = 0; oi < n; ++oi) {
			const uint32_t k = (uint32_t)__ffs(__ballot_sync(USB_FULL, have && rank == oi)) - 1;
			const uint32_t kLoi = __shfl_sync(USB_FULL, h.Loi, k), kLoj = __shfl_sync(USB_FULL, h.Loj, k);
			const int kscore = __shfl_sync(USB_FULL, h.score2, k);
			// best predecessor: largest chain score among the compatible HSPs earlier in the order,
			// the earliest of them on ties (the serial scan only replaces on a strictly larger score)
			const bool ok = have && rank < oi && h.Loi + h.Len - 1 < kLoi && h.Loj + h.Len - 1 < kLoj;
			const uint32_t key = ok ? (((uint32_t)cs << 6) | (63u - rank)) : 0u;
			const uint32_t bestkey = __reduce_max_sync(USB_FULL, key);
			uint32_t bestc = 0xffffffffu;
			int best = 0;
			if (bestkey) {
				const uint32_t brank = 63u - (bestkey & 63u);
				bestc = (uint32_t)__ffs(__ballot_sync(USB_FULL, have && rank == brank)) - 1;
				best = (int)(bestkey >> 6);
			}
			if (lane == k) {
				prev = bestc;
				cs = best + kscore;
			}
		}
		// best chain end: largest score, lowest index on ties
		const uint32_t okey = have ? (((uint32_t)cs << 6) | (63u - lane)) : 0u;
		const uint32_t opt = 63u - (__reduce_max_sync(USB_FULL, okey) & 63u);
		uint32_t members = 0; // lanes on the chain
		for (uint32_t k = opt; k != 0xffffffffu; k = __shfl_sync(USB_FULL, prev, k)) {
			members |= 1u << k;
			++len;
		}
		// chain in order of query start == ascending rank; position = chain members ranked lower
		{
			uint32_t pos = 0;
			for (uint32_t j = 0; j < n; ++j) {
				const uint32_t rj = __shfl_sync(USB_FULL, rank, j);
				pos += (((members >> j) & 1u) && rj < rank) ? 1u : 0u;
			}
			if ((members >> lane) & 1u)
				w.chain[pos] = lane;
		}
		const bool stag = ((members >> lane) & 1u) && is_staggered(h, w.LA, w.LB); // hspfinder.cpp:537-553
		if (__any_sync(USB_FULL, stag))
			len = 0;
		__syncwarp();
		return len;
	}
	if (lane == 0 && n > 0) {
		const HspRec *H = w.ung;
		uint32_t *order = w.order, *prev = w.prev;
		int *cs = w.cscore;
		for (uint32_t i = 0; i < n; ++i) {
			uint32_t j = i;
			const uint32_t key = H[i].Loi;
			while (j > 0 && H[order[j - 1]].Loi > key) {
				order[j] = order[j - 1];
				--j;
			}
			order[j] = i;
		}
		for (uint32_t oi = 0; oi < n; ++oi) {
			const uint32_t k = order[oi];
			const HspRec h = H[k];
			int best = 0;
			uint32_t bestc = 0xffffffffu;
			for (uint32_t oj = 0; oj < oi; ++oj) {
				const uint32_t c = order[oj];
				const HspRec g = H[c];
				if (g.Loi + g.Len - 1 < h.Loi && g.Loj + g.Len - 1 < h.Loj && (bestc == 0xffffffffu || cs[c] > best)) {
					bestc = c;
					best = cs[c];
				}
			}
			prev[k] = bestc;
			cs[k] = bestc == 0xffffffffu ? h.score2 : cs[bestc] + h.score2;
		}
		uint32_t opt = 0;
		for (uint32_t k = 1; k < n; ++k)
			if (cs[k] > cs[opt])
				opt = k;
		for (uint32_t k = opt; k != 0xffffffffu; k = prev[k])
			++len;
		uint32_t i = len;
		for (uint32_t k = opt; k != 0xffffffffu; k = prev[k])
			w.chain[--i] = k;
		for (uint32_t c = 0; c < len; ++c) // hspfinder.cpp:537-553
			if (is_staggered(H[w.chain[c]], w.LA, w.LB)) {
				len = 0;
				break;
			}
	}
	len = __shfl_sync(USB_FULL, len, 0);
	__syncwarp();
	return len;
}

// ------------------------------------------------------------------ a14/a15: banded Viterbi
struct GapCosts { // doubled; "L"/"R" = left/right terminal variants (alnparams.cpp:100-152)
	int OpenA, ExtA, OpenB, ExtB, LOpenA, LExtA, LOpenB, LExtB, ROpenA, RExtA, ROpenB, RExtB;
};

__device__ __forceinline__ GapCosts hole_costs(const DevParams &P, bool leftA, bool leftB, bool rightA, bool rightB)
{
	GapCosts g;
	g.OpenA = g.OpenB = P.open2;
	g.ExtA = g.ExtB = P.ext2;
	g.LOpenA = leftA ? P.topen2 : P.open2;   g.LExtA = leftA ? P.text2 : P.ext2;
	g.LOpenB = leftB ? P.topen2 : P.open2;   g.LExtB = leftB ? P.text2 : P.ext2;
	g.ROpenA = rightA ? P.topen2 : P.open2;  g.RExtA = rightA ? P.text2 : P.ext2;
	g.ROpenB = rightB ? P.topen2 : P.open2;  g.RExtB = rightB ? P.text2 : P.ext2;
	return g;
}

// diagbox.h:150-171
__device__ __forceinline__ void band_range(uint32_t LA, uint32_t LB, uint32_t dlo, uint32_t dhi, uint32_t i,
  uint32_t &sj, uint32_t &ej)
{
	sj = (dlo + i >= LA) ? dlo + i - LA : 0;
	if (sj >= LB)
		sj = LB - 1;
	ej = (dhi + i + 1 >= LA) ? dhi + i + 1 - LA : 0;
	if (ej > LB)
		ej = LB;
}

// max-plus inclusive warp scan: R[l] = max_{k<=l} (v[k] + (l-k)*ext)
__device__ __forceinline__ int scan_gap(int v, int ext)
{
	const uint32_t lane = lane_id();
	int R = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		int t = __shfl_up_sync(USB_FULL, R, d);
		if (lane >= (uint32_t)d)
			R = max(R, t + d * ext);
	}
	return R;
}

// Global alignment of Ac[0..LA) x Bc[0..LB) restricted to the main-diagonal band
// (viterbifastbandmem.cpp:232-253), LA, LB >= 1.  Appends the path at out[0..) and returns its
// length; *score2 receives the doubled score.  Rows are swept in order; within a row each lane
// owns one column: M and D depend only on the previous row, the I state is a prefix max-plus scan.
template <bool AA>
__device__ uint32_t viterbi_band(const AlignArgs &a, WarpWs &w, const uint8_t *Ac, uint32_t LA, const uint8_t *Bc,
  uint32_t LB, const GapCosts &G, char *out, int *score2, uint32_t *cells)
{
	const uint32_t lane = lane_id();
	const DevParams &P = a.P;
	uint32_t dlo = min(LA, LB), dhi = max(LA, LB);
	dlo = dlo > P.band ? dlo - P.band : 1;
	dhi += P.band;
	dhi = min(dhi, LA + LB - 1);
	if (P.band == 0) { // -band 0: ViterbiFastMem (viterbifastmem.cpp:9-170) == a band over every diagonal
		dlo = 1;
		dhi = LA + LB - 1;
	}
	// trace rows start on 4-byte boundaries and the two DP rows on 16-byte boundaries: wide rows are
	// swept four columns per lane with vector loads and stores
	const uint64_t W = ((uint64_t)LB + 4) & ~(uint64_t)3;
	// rows in shared memory when the rectangle is narrow enough, else in the global slab
	const bool rows_fit = 8u * (LB + 12) <= a.scratch_bytes;
	// next choice: the query's seed table (cnt, start, pos) lies right in front of the scratch and
	// is not needed any more for this candidate; it is rebuilt if another candidate follows
	const uint32_t big_bytes = (uint32_t)((w.scratch + a.scratch_bytes) - w.cnt);
	const bool rows_big = !rows_fit && a.fast_in_smem && 8u * (LB + 12) <= big_bytes;
	if (rows_big)
		w.seed_dirty = true;
	int *Mrow = (rows_fit ? (int *)w.scratch : rows_big ? (int *)w.cnt : w.rows_slab) + 4;
	int *Drow = Mrow + ((LB + 4 + 3) & ~3u);
	uint8_t *TB = w.TB;
	for (uint32_t j = lane; j <= LB + 1; j += 32) {
		Mrow[(int)j - 1] = USB_NEG;
		if (j <= LB)
			Drow[j] = USB_NEG;
	}
	__syncwarp();
	int OpenA = G.LOpenA, ExtA = G.LExtA;
	uint32_t ncell = 0;
	for (uint32_t i = 0; i < LA; ++i) {
		uint32_t sj, ej;
		band_range(LA, LB, dlo, dhi, i, sj, ej);
		if (ej == 0)
			continue;
		ncell += ej - sj;
		const uint32_t ca = Ac[i];
		uint8_t *TBrow = TB + (uint64_t)i * W;
		int Mcarry = (i == 0) ? 0 : (sj == 0 ? USB_NEG : Mrow[(int)sj - 1]);
		if (sj > 0 && lane == 0)
			TBrow[sj - 1] = TB_IM;
		int Icarry = USB_NEG;
		int lastMold = Mcarry;
		if (ej - sj >= 96) {
			// Wide row: four consecutive columns per lane, 128 per step.  Inside a lane the insert
			// state runs serially over its four columns; across lanes it is the same max-plus scan
			// as below on the lanes' aggregates (what leaves the lane if nothing enters it), with a
			// step cost of four extensions.  One scan, three shuffles and vector accesses serve 128
			// cells.  Columns left of sj (the chunk starts on a multiple of four) are dead cells;
			// column sj - 1 carries the row's start value so that column sj sees it as `saved`.
			const uint32_t b0 = sj & ~3u;
			int McW = (b0 == sj) ? Mcarry : USB_NEG;
			for (uint32_t base = b0; base < ej; base += 128) {
				const uint32_t j0 = base + 4 * lane;
				int mo[4], dol[4];
				if (j0 >= sj && j0 + 3 < ej) {
					const int4 m4 = *(const int4 *)(Mrow + j0), d4 = *(const int4 *)(Drow + j0);
					mo[0] = m4.x; mo[1] = m4.y; mo[2] = m4.z; mo[3] = m4.w;
					dol[0] = d4.x; dol[1] = d4.y; dol[2] = d4.z; dol[3] = d4.w;
				} else {
#pragma unroll
					for (int c = 0; c < 4; ++c) {
						const uint32_t j = j0 + c;
						const bool act = j >= sj && j < ej;
						mo[c] = act ? Mrow[j] : (j + 1 == sj ? Mcarry : USB_NEG);
						dol[c] = act ? Drow[j] : USB_NEG;
					}
				}
				int sv[4];
				sv[0] = __shfl_up_sync(USB_FULL, mo[3], 1);
				if (lane == 0)
					sv[0] = McW;
				sv[1] = mo[0]; sv[2] = mo[1]; sv[3] = mo[2];
				const int g0 = sv[0] + OpenA, g1 = sv[1] + OpenA, g2 = sv[2] + OpenA, g3 = sv[3] + OpenA;
				const int L = max(max(g0 + 3 * ExtA, g1 + 2 * ExtA), max(g2 + ExtA, g3));
				const int R = scan_gap(L, 4 * ExtA);
				const int Iout = max(R, Icarry + (int)(lane + 1) * 4 * ExtA);
				int Ie = __shfl_up_sync(USB_FULL, Iout, 1);
				if (lane == 0)
					Ie = Icarry;
				const int gg[4] = {g0, g1, g2, g3};
				int mn[4], dn[4];
				uint32_t tb = 0;
				uint32_t b4 = 0;
				if (j0 < ej) {
					// four target letters; Bc + j0 is not word aligned in general (a hole starts anywhere
					// in the target): two aligned words and a funnel shift (the buffer is padded)
					const uintptr_t pb = (uintptr_t)(Bc + j0);
					const uint32_t *pw = (const uint32_t *)(pb & ~(uintptr_t)3);
					b4 = __funnelshift_r(pw[0], pw[1], 8 * (uint32_t)(pb & 3));
				}
#pragma unroll
				for (int c = 0; c < 4; ++c) {
					const uint32_t j = j0 + c;
					const bool act = j >= sj && j < ej;
					uint32_t bits = 0;
					int xM = sv[c];
					if (dol[c] > xM) {
						xM = dol[c];
						bits = TB_DM;
					}
					if (Ie > xM) {
						xM = Ie;
						bits = TB_IM;
					}
					mn[c] = xM + (act ? subst2g<AA>(a, w, ca, (b4 >> (8 * c)) & 0xffu) : 0);
					const int ob = (j == 0) ? G.LOpenB : G.OpenB, eb = (j == 0) ? G.LExtB : G.ExtB;
					const int md = sv[c] + ob;
					int dnew = dol[c] + eb;
					if (md >= dnew) {
						dnew = md;
						bits |= TB_MD;
					}
					if (gg[c] >= Ie + ExtA)
						bits |= TB_MI;
					dn[c] = dnew;
					tb |= bits << (8 * c);
					Ie = max(gg[c], Ie + ExtA);
				}
				// carries: old M of the row's last column (right edge), of this step's last column, I
				const uint32_t last = min(127u, ej - 1 - base);
				const uint32_t lc = last & 3;
				const int msel = lc == 0 ? mo[0] : lc == 1 ? mo[1] : lc == 2 ? mo[2] : mo[3];
				lastMold = __shfl_sync(USB_FULL, msel, last >> 2);
				McW = __shfl_sync(USB_FULL, mo[3], 31);
				Icarry = __shfl_sync(USB_FULL, Iout, 31);
				if (j0 >= sj && j0 + 3 < ej) {
					*(int4 *)(Mrow + j0) = make_int4(mn[0], mn[1], mn[2], mn[3]);
					*(int4 *)(Drow + j0) = make_int4(dn[0], dn[1], dn[2], dn[3]);
					*(uint32_t *)(TBrow + j0) = tb;
				} else {
#pragma unroll
					for (int c = 0; c < 4; ++c) {
						const uint32_t j = j0 + c;
						if (j >= sj && j < ej) {
							Mrow[j] = mn[c];
							Drow[j] = dn[c];
							TBrow[j] = (uint8_t)(tb >> (8 * c));
						}
					}
				}
			}
		} else
		for (uint32_t base = sj; base < ej; base += 32) {
			const uint32_t j = base + lane;
			const bool act = j < ej;
			const int mold = act ? Mrow[j] : USB_NEG;
			const int dold = act ? Drow[j] : USB_NEG;
			int saved = __shfl_up_sync(USB_FULL, mold, 1);
			if (lane == 0)
				saved = Mcarry;
			// I state: I(j+1) = max(saved(j) + OpenA, I(j) + ExtA)
			const int g = saved + OpenA;
			const int R = scan_gap(g, ExtA);
			const int Iincl = max(R, Icarry + (int)(lane + 1) * ExtA);
			int Iexcl = __shfl_up_sync(USB_FULL, Iincl, 1);
			if (lane == 0)
				Iexcl = Icarry;
			uint32_t bits = 0;
			int xM = saved;
			if (dold > xM) {
				xM = dold;
				bits = TB_DM;
			}
			if (Iexcl > xM) {
				xM = Iexcl;
				bits = TB_IM;
			}
			const int mnew = xM + (act ? subst2g<AA>(a, w, ca, Bc[j]) : 0);
			const int ob = (j == 0) ? G.LOpenB : G.OpenB, eb = (j == 0) ? G.LExtB : G.ExtB;
			const int md = saved + ob;
			int dnew = dold + eb;
			if (md >= dnew) {
				dnew = md;
				bits |= TB_MD;
			}
			if (g >= Iexcl + ExtA)
				bits |= TB_MI;
			const uint32_t last_lane = min(31u, ej - 1 - base);
			lastMold = __shfl_sync(USB_FULL, mold, last_lane);
			Mcarry = __shfl_sync(USB_FULL, mold, 31);
			Icarry = __shfl_sync(USB_FULL, Iincl, 31);
			if (act) {
				Mrow[j] = mnew;
				Drow[j] = dnew;
				TBrow[j] = (uint8_t)bits;
			}
		}
		__syncwarp();
		if (lane == 0) { // right-edge column, every row (viterbifastbandmem.cpp:165-179)
			uint32_t tb = 0;
			const int md = lastMold + G.ROpenB;
			int d = Drow[LB] + G.RExtB;
			if (md >= d) {
				d = md;
				tb = TB_MD;
			}
			Drow[LB] = d;
			TBrow[LB] = (uint8_t)tb;
		}
		__syncwarp();
		OpenA = G.OpenA;
		ExtA = G.ExtA;
	}
	// bottom row: horizontal gaps after the last query letter (:186-206), strict '>'
	uint32_t sj, ej;
	band_range(LA, LB, dlo, dhi, LA - 1, sj, ej);
	int FinalI = USB_NEG;
	{
		uint8_t *TBrow = TB + (uint64_t)LA * W;
		if (lane == 0)
			Mrow[(int)sj - 1] = USB_NEG;
		__syncwarp();
		int Icarry = USB_NEG;
		for (uint32_t base = sj; base < ej; base += 32) {
			const uint32_t j = base + lane;
			const bool act = j < ej;
			const int g = (act ? Mrow[(int)j - 1] : USB_NEG) + G.ROpenA;
			const int R = scan_gap(g, G.RExtA);
			const int Iincl = max(R, Icarry + (int)(lane + 1) * G.RExtA);
			int Iexcl = __shfl_up_sync(USB_FULL, Iincl, 1);
			if (lane == 0)
				Iexcl = Icarry;
			if (act)
				TBrow[j] = (g > Iexcl + G.RExtA) ? TB_MI : 0;
			const uint32_t last_lane = min(31u, ej - 1 - base);
			FinalI = __shfl_sync(USB_FULL, Iincl, last_lane);
			Icarry = __shfl_sync(USB_FULL, Iincl, 31);
		}
	}
	__syncwarp();
	const int FinalM = Mrow[(int)LB - 1], FinalD = Drow[LB];
	int Score = FinalM;
	char State = 'M';
	if (FinalD > Score) {
		Score = FinalD;
		State = 'D';
	}
	if (FinalI > Score) {
		Score = FinalI;
		State = 'I';
	}
	if (score2)
		*score2 = Score;
	if (cells)
		*cells += ncell;

	// traceback (tracebackbitmem.cpp:8-73), reversed into w.rev.  The walk is serial, one dependent
	// trace-byte load per step; here the warp looks 32 steps ahead along the current state's run
	// (diagonal for M, column for D, row for I): lane l loads the byte step l would read if the
	// state did not change, the first lane whose byte ends the run (or runs off the matrix) is
	// found with a ballot, and the whole run is emitted at once.
	uint32_t n = 0;
	{
		uint32_t i = LA, j = LB;
		char *rev = w.rev;
		const uint32_t limit = LA + LB;
		bool bad = false;
		while ((i != 0 || j != 0) && n < limit) {
			bool valid;
			uint32_t ci, cj;
			if (State == 'M') {
				valid = i > lane && j > lane;
				ci = i - 1 - lane;
				cj = j - 1 - lane;
			} else if (State == 'D') {
				valid = i > lane;
				ci = i - 1 - lane;
				cj = j;
			} else {
				valid = j > lane;
				ci = i;
				cj = j - 1 - lane;
			}
			const uint32_t t = valid ? (uint32_t)TB[(uint64_t)ci * W + cj] : 0u;
			const bool leave = State == 'M' ? (t & (TB_DM | TB_IM)) != 0 : State == 'D' ? (t & TB_MD) != 0 : (t & TB_MI) != 0;
			const uint32_t vmask = __ballot_sync(USB_FULL, valid);
			const uint32_t stop = __ballot_sync(USB_FULL, !valid || leave);
			const uint32_t first = stop ? (uint32_t)__ffs(stop) - 1 : 32u;
			const bool first_valid = first < 32 && ((vmask >> first) & 1u);
			uint32_t cnt = first < 32 ? first + (first_valid ? 1u : 0u) : 32u;
			const bool cut = cnt > limit - n;
			cnt = min(cnt, limit - n);
			if (cnt == 0) { // the serial walk would step off the matrix here
				bad = true;
				break;
			}
			if (lane < cnt)
				rev[n + lane] = State;
			n += cnt;
			const char prev = State;
			if (first_valid && !cut) {
				const uint32_t tf = __shfl_sync(USB_FULL, t, first);
				State = prev == 'M' ? ((tf & TB_DM) ? 'D' : 'I') : 'M';
			}
			if (prev == 'M') {
				i -= cnt;
				j -= cnt;
			} else if (prev == 'D')
				i -= cnt;
			else
				j -= cnt;
		}
		if ((bad || i != 0 || j != 0) && lane == 0)
			atomicOr(&a.ctr->err, ERR_TRACE);
	}
	__syncwarp();
	for (uint32_t k = lane; k < n; k += 32)
		out[k] = w.rev[n - 1 - k];
	__syncwarp();
	return n;
}

// ------------------------------------------------------------------ a9/a13: whole pair
__device__ __forceinline__ uint32_t fill_run(char *out, char c, uint32_t n)
{
	for (uint32_t k = lane_id(); k < n; k += 32)
		out[k] = c;
	return n;
}

// globalalignmem.cpp:70-112 AlignHSPMem
template <bool AA>
__device__ uint32_t align_hole(const AlignArgs &a, WarpWs &w, uint32_t Loi, uint32_t Loj, uint32_t Leni, uint32_t Lenj,
  char *out, usb_qstat &st)
{
	if (Leni == 0)
		return fill_run(out, 'I', Lenj);
	if (Lenj == 0)
		return fill_run(out, 'D', Leni);
	const GapCosts G = hole_costs(a.P, Loi == 0, Loj == 0, Loi + Leni == w.LA, Loj + Lenj == w.LB);
	++st.n_dp;
	return viterbi_band<AA>(a, w, w.Ac + Loi, Leni, w.Bc + Loj, Lenj, G, out, nullptr, &st.dp_cells);
}

// GlobalAlign_AllOpts (globalalignmem.cpp:129-236).  Returns path length, 0 = rejected (no AR).
// The seed table of the query must be current.
template <bool AA>
__device__ uint32_t global_align(const AlignArgs &a, WarpWs &w, usb_qstat &st, uint32_t *n_chain_out)
{
	const DevParams &P = a.P;
	const uint32_t lane = lane_id();
	const uint32_t LA = w.LA, LB = w.LB;
	if (P.fulldp) { // globalalignmem.cpp:153-157 FullDPAlways: no HSPs, no identity gate
		if (n_chain_out)
			*n_chain_out = 0;
		if (LA == 0 || LB == 0)
			return 0;
		const GapCosts G = hole_costs(P, true, true, true, true);
		++st.n_dp;
		return viterbi_band<AA>(a, w, w.Ac, LA, w.Bc, LB, G, w.path, nullptr, &st.dp_cells);
	}
	uint32_t MinHSPLength = P.min_hsp_len == 0 ? 32 : P.min_hsp_len;
	MinHSPLength = min(MinHSPLength, LA / 4);
	MinHSPLength = max(MinHSPLength, 16u);
	const uint32_t nung = ungapped_blast<AA>(a, w, MinHSPLength);
	const uint32_t nchain = chain_hsps(a, w, nung);
	if (n_chain_out)
		*n_chain_out = nchain;
	// HSP identity gate (getglobalhsps.cpp:22-60, globalalignmem.cpp:171)
	uint32_t TotalLength = 0, TotalSame = 0;
	for (uint32_t c = 0; c < nchain; ++c) {
		const HspRec h = w.ung[w.chain[c]];
		TotalLength += h.Len;
		for (uint32_t base = 0; base < h.Len; base += 32) {
			const uint32_t k = base + lane;
			bool same = false;
			if (k < h.Len)
				same = pos_match<AA>(w, h.Loi + k, h.Loj + k);
			TotalSame += __popc(__ballot_sync(USB_FULL, same));
		}
	}
	const float HSPFractId = TotalLength == 0 ? 0.0f : (float)TotalSame / (float)TotalLength;
	if (HSPFractId < P.min_hsp_fract_id)
		return 0;
	char *path = w.path;
	uint32_t n = 0;
	if (nchain == 0) {
		if (P.min_hsp_len > 0 && LA > 64)
			return 0;
		if (LA == 0 || LB == 0)
			return 0;
		const GapCosts G = hole_costs(P, true, true, true, true);
		++st.n_dp;
		return viterbi_band<AA>(a, w, w.Ac, LA, w.Bc, LB, G, path, nullptr, &st.dp_cells);
	}
	uint32_t Loi = 0, Loj = 0;
	for (uint32_t c = 0; c < nchain; ++c) {
		const HspRec h = w.ung[w.chain[c]];
		n += align_hole<AA>(a, w, Loi, Loj, h.Loi - Loi, h.Loj - Loj, path + n, st);
		n += fill_run(path + n, 'M', h.Len);
		Loi = h.Loi + h.Len;
		Loj = h.Loj + h.Len;
	}
	n += align_hole<AA>(a, w, Loi, Loj, LA - Loi, LB - Loj, path + n, st);
	__syncwarp();
	return n;
}

// ------------------------------------------------------------------ a17: statistics + hit record
// arscorer.cpp:201-296 FillLo over the path between the first and last M column.
template <bool AA> __device__ bool path_stats(const AlignArgs &a, const WarpWs &w, uint32_t n, usb_hit &h)
{
	const uint32_t lane = lane_id();
	const char *path = w.path;
	uint32_t first = 0xffffffffu, last = 0, qpos = 0, tpos = 0;
	// first M column and the letters consumed before it
	for (uint32_t base = 0; base < n && first == 0xffffffffu; base += 32) {
		const uint32_t c = base + lane;
		const char ch = c < n ? path[c] : 0;
		const uint32_t mm = __ballot_sync(USB_FULL, ch == 'M');
		const uint32_t dm = __ballot_sync(USB_FULL, ch == 'D');
		const uint32_t im = __ballot_sync(USB_FULL, ch == 'I');
		if (mm) {
			const uint32_t f = __ffs(mm) - 1;
			first = base + f;
			qpos += __popc(dm & ((1u << f) - 1));
			tpos += __popc(im & ((1u << f) - 1));
		} else {
			qpos += __popc(dm);
			tpos += __popc(im);
		}
	}
	if (first == 0xffffffffu)
		return false;
	for (uint32_t base = (n - 1) & ~31u;; base -= 32) {
		const uint32_t c = base + lane;
		const char ch = c < n ? path[c] : 0;
		const uint32_t mm = __ballot_sync(USB_FULL, ch == 'M');
		if (mm) {
			last = base + 31 - __clz(mm);
			break;
		}
		if (base == 0)
			break;
	}
	h.first_mq = qpos;
	h.first_mt = tpos;
	h.first_mcol = first;
	h.alnlen = last - first + 1;
	uint32_t ids = 0, mism = 0, gaps = 0, opens = 0;
	uint32_t prev_carry = 'M';
	for (uint32_t base = first; base <= last; base += 32) {
		const uint32_t c = base + lane;
		const bool in = c <= last;
		const uint32_t ch = in ? (uint32_t)path[c] : 0u;
		const bool isM = ch == 'M', isD = ch == 'D', isI = ch == 'I';
		const uint32_t qm = __ballot_sync(USB_FULL, isM || isD);
		const uint32_t tm = __ballot_sync(USB_FULL, isM || isI);
		const uint32_t qp = qpos + __popc(qm & lanemask_lt());
		const uint32_t tp = tpos + __popc(tm & lanemask_lt());
		bool same = false;
		if (isM)
			same = pos_match<AA>(w, qp, tp);
		uint32_t prev = __shfl_up_sync(USB_FULL, ch, 1);
		if (lane == 0)
			prev = prev_carry;
		ids += __popc(__ballot_sync(USB_FULL, same));
		mism += __popc(__ballot_sync(USB_FULL, isM && !same));
		gaps += __popc(__ballot_sync(USB_FULL, isD || isI));
		opens += __popc(__ballot_sync(USB_FULL, (isD || isI) && prev == 'M'));
		prev_carry = __shfl_sync(USB_FULL, ch, 31);
		qpos += __popc(qm);
		tpos += __popc(tm);
	}
	h.ids = ids;
	h.mism = mism;
	h.intgaps = gaps;
	h.opens = opens;
	h.last_mq = qpos - 1;
	h.last_mt = tpos - 1;
	return true;
}

// Appends the path as runs (length << 2 | op) to the run arena; returns false when full.
__device__ bool emit_runs(const AlignArgs &a, const WarpWs &w, uint32_t n, usb_hit &h)
{
	const uint32_t lane = lane_id();
	const char *path = w.path;
	uint32_t nruns = 0;
	uint32_t carry = 0;
	for (uint32_t base = 0; base < n; base += 32) {
		const uint32_t c = base + lane;
		const uint32_t ch = c < n ? (uint32_t)path[c] : 0u;
		uint32_t prev = __shfl_up_sync(USB_FULL, ch, 1);
		if (lane == 0)
			prev = carry;
		nruns += __popc(__ballot_sync(USB_FULL, c < n && ch != prev));
		carry = __shfl_sync(USB_FULL, ch, 31);
	}
	uint32_t off = 0;
	if (lane == 0)
		off = atomicAdd(&a.ctr->n_runs, nruns);
	off = __shfl_sync(USB_FULL, off, 0);
	if ((uint64_t)off + nruns > a.runs_cap) {
		if (lane == 0)
			atomicOr(&a.ctr->err, ERR_RUNS_FULL);
		return false;
	}
	uint32_t *runs = a.runs + off;
	uint32_t k = 0;
	carry = 0;
	for (uint32_t base = 0; base < n; base += 32) {
		const uint32_t c = base + lane;
		const uint32_t ch = c < n ? (uint32_t)path[c] : 0u;
		uint32_t prev = __shfl_up_sync(USB_FULL, ch, 1);
		if (lane == 0)
			prev = carry;
		const bool startrun = c < n && ch != prev;
		const uint32_t sm = __ballot_sync(USB_FULL, startrun);
		if (startrun) {
			const uint32_t op = ch == 'M' ? 0u : ch == 'D' ? 1u : 2u;
			runs[k + __popc(sm & lanemask_lt())] = (c << 2) | op; // start column for now
		}
		k += __popc(sm);
		carry = __shfl_sync(USB_FULL, ch, 31);
	}
	__syncwarp();
	// start columns -> lengths
	for (uint32_t base = 0; base < nruns; base += 32) {
		const uint32_t r = base + lane;
		uint32_t v = 0, nxt = 0;
		if (r < nruns) {
			v = runs[r];
			nxt = (r + 1 < nruns) ? (runs[r + 1] >> 2) : n;
		}
		__syncwarp();
		if (r < nruns)
			runs[r] = ((nxt - (v >> 2)) << 2) | (v & 3);
		__syncwarp();
	}
	h.run_off = off;
	h.run_cnt = nruns;
	return true;
}

// ------------------------------------------------------------------ Accepter rules (all candidate loops)
#define ACC_PAIR_FLAGS (USB_ACC_SELF | USB_ACC_NOTSELF | USB_ACC_SELFID | USB_ACC_MIN_SIZERATIO | USB_ACC_MINQT | \
                        USB_ACC_MAXQT | USB_ACC_MINSL | USB_ACC_MAXSL)
#define ERR_KCAP 1024u // skipped pairs used up the materialised candidates before the Terminator fired

// Accepter::RejectPair (accepter.cpp:145-197): rules on the pair itself, before any alignment.
// Q = raw query letters, L its length; the target's raw (masked) letters come from the database.
template <class Args>
__device__ bool reject_pair(const Args &a, uint32_t qi, uint32_t strand, uint32_t t, const uint8_t *Q, uint32_t L)
{
	const DevParams &P = a.P;
	const uint32_t f = P.accept_flags;
	if ((f & USB_ACC_SELF) && a.q_label[qi] == a.t_label[t])
		return true;
	if ((f & USB_ACC_NOTSELF) && a.q_label[qi] != a.t_label[t])
		return true;
	const uint32_t TL = a.db_len[t];
	if ((f & USB_ACC_SELFID) && TL == L) { // same length and the same letters, byte for byte
		const uint8_t *T = a.db_seq + a.db_off[t];
		bool diff = false;
		for (uint32_t i = lane_id(); i < L; i += 32) {
			const uint32_t c = strand ? (uint32_t)c_comp[Q[L - 1 - i]] : (uint32_t)Q[i];
			diff |= c != (uint32_t)T[i];
		}
		if (!__any_sync(USB_FULL, diff))
			return true;
	}
	if (f & USB_ACC_MIN_SIZERATIO) {
		const double Ratio = (double)a.t_size[t] / (double)a.q_size[qi];
		if (Ratio < P.min_sizeratio_d)
			return true;
	}
	if (f & (USB_ACC_MINQT | USB_ACC_MAXQT | USB_ACC_MINSL | USB_ACC_MAXSL)) {
		const double q = (double)L, tt = (double)TL;
		const double s = (double)min(L, TL), l = (double)max(L, TL);
		const double qt = q / tt, sl = s / l;
		if ((f & USB_ACC_MINQT) && qt < P.minqt_d)
			return true;
		if ((f & USB_ACC_MAXQT) && qt > P.maxqt_d)
			return true;
		if ((f & USB_ACC_MINSL) && sl < P.minsl_d)
			return true;
		if ((f & USB_ACC_MAXSL) && sl > P.maxsl_d)
			return true;
	}
	return false;
}

// Accepter::IsAcceptLo (accepter.cpp:41-94) on the statistics of an alignment.
template <class Args>
__device__ __forceinline__ bool accept_hit(const Args &a, const usb_hit &h, uint32_t qi, bool local = false)
{
	const DevParams &P = a.P;
	const uint32_t f = P.accept_flags;
	const double fid = h.alnlen == 0 ? 0.0 : (double)h.ids / (double)h.alnlen;
	if (fid < P.id_d)
		return false;
	if ((f & USB_ACC_MAXID) && fid > P.maxid_d)
		return false;
	if ((f & USB_ACC_MINCOLS) && h.alnlen < P.mincols)
		return false;
	if ((f & USB_ACC_MAXGAPS) && h.intgaps > P.maxgaps)
		return false;
	if (f & (USB_ACC_QUERY_COV | USB_ACC_MAX_QUERY_COV)) {
		const double cov = (double)(h.last_mq - h.first_mq + 1) / (double)h.ql; // arscorer.cpp:122-137
		if ((f & USB_ACC_QUERY_COV) && cov < P.query_cov_d)
			return false;
		if ((f & USB_ACC_MAX_QUERY_COV) && cov > P.max_query_cov_d)
			return false;
	}
	if (f & (USB_ACC_TARGET_COV | USB_ACC_MAX_TARGET_COV)) {
		// arscorer.cpp:139-154: letter pairs over TL for a global alignment, the segment length for a local one
		const double cov = (double)(local ? h.last_mt - h.first_mt + 1 : h.ids + h.mism) / (double)h.tl;
		if ((f & USB_ACC_TARGET_COV) && cov < P.target_cov_d)
			return false;
		if ((f & USB_ACC_MAX_TARGET_COV) && cov > P.max_target_cov_d)
			return false;
	}
	if ((f & USB_ACC_MAXDIFFS) && h.mism + h.intgaps > P.maxdiffs)
		return false;
	if ((f & USB_ACC_MINDIFFS) && h.mism + h.intgaps < P.mindiffs)
		return false;
	if ((f & USB_ACC_ABSKEW) && (double)a.t_size[h.target] / (double)a.q_size[qi] < P.abskew_d)
		return false;
	return true;
}


// ------------------------------------------------------------------ the job loop
template <bool AA> __device__ void align_job(const AlignArgs &a, WarpWs &w, uint32_t job)
{
	const uint32_t lane = lane_id();
	const bool pairs = a.pair_q != nullptr;
	const uint32_t qi = pairs ? a.pair_q[job] : job / a.strands;
	const uint32_t strand = pairs ? 0u : job % a.strands;
	const uint32_t ncand = pairs ? 1u : a.n_emit[job];
	usb_qstat st;
	st.n_cand = 0; st.n_tried = 0; st.n_hspfail = 0; st.n_dp = 0; st.dp_cells = 0; st.n_accept = 0; st.seq_bytes = 0;
	if (ncand != 0) {
		const uint64_t q0 = a.q_off[qi];
		const uint32_t L = (uint32_t)(a.q_off[qi + 1] - q0);
		load_query<AA>(a, w, a.q + q0, L, strand);
		build_seed_table<AA>(a, w);
		w.seed_dirty = false;
		uint32_t acc = 0, rej = 0;
		bool stopped = pairs;
		for (uint32_t k = 0; k < ncand; ++k) {
			const uint32_t t = pairs ? a.pair_t[job] : a.cand_t[(uint64_t)job * a.k_max + k];
			if (w.seed_dirty) {
				build_seed_table<AA>(a, w);
				w.seed_dirty = false;
			}
			// Accepter::RejectPair before any alignment (searcher.cpp:63-67): skipped without a Terminator
			// call on the small-database path, a reject on the big one (udbusortedsearcherbig.cpp:119-128)
			if (!pairs && (a.P.accept_flags & ACC_PAIR_FLAGS) &&
			    reject_pair(a, qi, strand, t, a.q + q0, L)) {
				if (a.P.reject_pair_counts && a.P.maxrejects > 0 && ++rej == a.P.maxrejects) {
					stopped = true;
					break;
				}
				continue;
			}
			load_target<AA>(a, w, t);
			++st.n_tried;
			st.seq_bytes += w.LA + w.LB;
			uint32_t nchain = 0;
			const uint32_t n = global_align<AA>(a, w, st, &nchain);
			if (pairs && a.hsp_out) {
				uint32_t *ho = a.hsp_out + (uint64_t)job * (1 + 4 * a.max_hsp);
				if (lane == 0) {
					ho[0] = nchain;
					for (uint32_t c = 0; c < nchain && c < a.max_hsp; ++c) {
						const HspRec h = w.ung[w.chain[c]];
						ho[1 + 4 * c] = h.Loi;
						ho[2 + 4 * c] = h.Loj;
						ho[3 + 4 * c] = h.Len;
						ho[4 + 4 * c] = (uint32_t)h.score2;
					}
				}
			}
			bool accept = false;
			if (n == 0)
				++st.n_hspfail;
			else {
				usb_hit h;
				h.query = qi; h.target = t; h.strand = strand; h.rank = pairs ? job : k;
				h.ql = w.LA; h.tl = w.LB; h.run_off = 0; h.run_cnt = 0; h.raw = 0; h.sub = 0;
				if (!path_stats<AA>(a, w, n, h)) {
					if (lane == 0)
						atomicOr(&a.ctr->err, ERR_NO_M);
				} else {
					// accepter.cpp:27-38: reject iff double(ids)/double(cols) < (double)(float)id
					accept = pairs || accept_hit(a, h, qi);
					if (accept) {
						if (emit_runs(a, w, n, h)) {
							uint32_t slot = 0;
							if (lane == 0)
								slot = atomicAdd(&a.ctr->n_hits, 1u);
							slot = __shfl_sync(USB_FULL, slot, 0);
							if (slot < a.hits_cap) {
								if (lane == 0)
									a.hits[slot] = h;
							} else if (lane == 0)
								atomicOr(&a.ctr->err, ERR_HITS_FULL);
						}
						++st.n_accept;
					}
				}
			}
			if (pairs) {
				if (lane == 0)
					a.aligned[job] = n != 0;
				break;
			}
			// terminator.cpp:64-100
			if (accept)
				++acc;
			else
				++rej;
			if ((a.P.maxaccepts > 0 && acc == a.P.maxaccepts) || (a.P.maxrejects > 0 && rej == a.P.maxrejects)) {
				stopped = true;
				break;
			}
		}
		// every materialised candidate was looked at; with skipped pairs the reference may go on
		if (!stopped && a.n_cand_all && a.n_cand_all[job] > ncand && lane == 0)
			atomicOr(&a.ctr->err, ERR_KCAP);
	}
	if (lane == 0 && a.qstat) {
		usb_qstat *o = a.qstat + job;
		*o = st; // n_cand is filled in by the host from k_rank's output
	}
}

#define ALIGN_MAX_WARPS 16
// AA = true: amino acid alphabet; the letter tables sit in front of the per-warp arrays in shared memory.
#define ALIGN_TAB_BYTES ((uint32_t)((sizeof(AlignTables) + 15) & ~(size_t)15))
template <bool AA> __global__ void __launch_bounds__(ALIGN_MAX_WARPS * 32, 1) k_align(const AlignArgs a)
{
	extern __shared__ __align__(16) uint8_t align_smem[];
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + warp;
	uint8_t *slab = a.slab + (uint64_t)gw * a.slab_stride;
	uint8_t *smem = align_smem;
	WarpWs w;
	w.T = nullptr;
	if constexpr (AA) {
		for (uint32_t i = threadIdx.x; i < sizeof(AlignTables) / 4; i += blockDim.x)
			((uint32_t *)smem)[i] = ((const uint32_t *)a.tab)[i];
		__syncthreads();
		w.T = (const AlignTables *)smem;
		smem += ALIGN_TAB_BYTES;
	}
	uint8_t *fast = a.fast_in_smem ? smem + (size_t)warp * a.fast_bytes : slab;
	if (!a.fast_in_smem)
		slab += a.fast_bytes;
	ws_setup(a, w, fast, slab);
	for (;;) {
		uint32_t job = 0;
		if (lane == 0)
			job = atomicAdd(&a.ctr->job, 1u);
		job = __shfl_sync(USB_FULL, job, 0);
		if (job >= a.n_jobs)
			break;
		align_job<AA>(a, w, job);
	}
}

// ------------------------------------------------------------------ stage kernel: Viterbi only
struct ViterbiArgs {
	AlignArgs base;          // P, slab, caps, ctr
	const uint8_t *a_seq; const uint64_t *a_off;
	const uint8_t *b_seq; const uint64_t *b_off;
	const uint8_t *flags;
	uint32_t n;
	char *paths; const uint64_t *path_off;
	int *score2;
};

__global__ void __launch_bounds__(ALIGN_MAX_WARPS * 32, 1) k_viterbi(const ViterbiArgs v)
{
	extern __shared__ __align__(16) uint8_t align_smem[];
	const AlignArgs &a = v.base;
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + warp;
	uint8_t *slab = a.slab + (uint64_t)gw * a.slab_stride;
	uint8_t *fast = a.fast_in_smem ? align_smem + (size_t)warp * a.fast_bytes : slab;
	if (!a.fast_in_smem)
		slab += a.fast_bytes;
	WarpWs w;
	w.T = nullptr;
	ws_setup(a, w, fast, slab);
	for (;;) {
		uint32_t job = 0;
		if (lane == 0)
			job = atomicAdd(&a.ctr->job, 1u);
		job = __shfl_sync(USB_FULL, job, 0);
		if (job >= v.n)
			break;
		const uint32_t LA = (uint32_t)(v.a_off[job + 1] - v.a_off[job]);
		const uint32_t LB = (uint32_t)(v.b_off[job + 1] - v.b_off[job]);
		const uint8_t *A = v.a_seq + v.a_off[job], *B = v.b_seq + v.b_off[job];
		for (uint32_t i = lane; i < LA; i += 32)
			w.Ac[i] = (uint8_t)nt_code(A[i]);
		for (uint32_t i = lane; i < LB; i += 32)
			w.Bc[i] = (uint8_t)nt_code(B[i]);
		w.LA = LA;
		w.LB = LB;
		__syncwarp();
		const uint32_t f = v.flags[job];
		const GapCosts G = hole_costs(a.P, f & 1, f & 2, f & 4, f & 8);
		char *out = v.paths + v.path_off[job];
		int sc = 0;
		uint32_t cells = 0;
		const uint32_t n = viterbi_band<false>(a, w, w.Ac, LA, w.Bc, LB, G, out, &sc, &cells);
		if (lane == 0) {
			out[n] = 0;
			v.score2[job] = sc;
		}
		__syncwarp();
	}
}

} // namespace usb
