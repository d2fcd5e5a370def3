// usb_stage.cuh -- the candidate loop of -usearch_global as a staged pipeline (nucleotide path).
//
// The reference walks a query's U-sorted candidates one after the other
// (UDBUsortedSearcher::SearchImpl, udbusortedsearcher.cpp:138-151): SetTarget, Align, Accepter,
// Terminator.  91 % of all attempts at the BASELINE workload end at the HSP identity gate of
// GlobalAlign_AllOpts (globalalignmem.cpp:171) and never reach the DP, and the gate of a
// (query, target) pair has no state: it does not depend on what happened to earlier candidates.
// So the loop is cut into stages over candidate ranges [ka, kb):
//
//   k_stage_prep   list of the jobs the Terminator has not stopped (the work items of the stage)
//   k_gate         per job: packed query, seed table, and for each candidate of the stage the packed
//                  target -> ungapped HSPs -> chain -> HSP identity gate.  Lean: 2-bit letters only,
//                  ~5.7 KB of shared memory per warp, <= 64 registers => 32 warps per SM.
//                  Survivors leave a record (job, k, chained HSPs).
//   k_dp           per record: holes by banded Viterbi, path, FillLo statistics, -id test; the hit
//                  is staged next to the record.
//   k_commit       per job, candidates in order: Terminator counts (terminator.cpp:64-100); hits
//                  of candidates examined before the Terminator fired are copied to the output.
//
// Stage 0 is candidate 0 of every job (70 % of the reads are accepted there), stage 1 candidates
// 1..64, and so on.  Work done for candidates behind the point where the Terminator fires is
// wasted but never visible: results are exactly those of the sequential loop.
//
// Reference behaviour reproduced by the stages: see usb_align.cuh (same device functions).
#pragma once
#include "usb_align.cuh"

namespace usb {

#define ERR_REC_FULL 256u
#define ERR_HSPARENA_FULL 512u

#define STAGE_MAX 20          // stages per batch: [0,1), then ranges of STAGE_WIDTH candidates
#define STAGE_WIDTH 64
#define GATE_MAX_WARPS 32
#define GATE_Q1 256           // flat seed queue (ring, power of two): < 32 waiting + at most 8 x 16 per slow scan step
#define GATE_Q2 64            // survivors of the pre-filter waiting for a full batch of walks
#define DP_MAX_WARPS 16

struct PassRec {
	uint32_t job, k, target, nchain, hsp_off;
	uint32_t hit, n_dp, dp_cells; // filled in by k_dp: hit = 1 + index of the staged hit, 0 = rejected by -id
};

struct StageCounters {
	uint32_t n_recs, n_hsp_words, n_staged, pad;
	unsigned long long dp_cells, dp_seq_bytes; // totals over the records (measurement)
	struct {
		uint32_t n_items, gate_cursor, rec_begin, dp_cursor;
	} st[STAGE_MAX];
};

struct StageArgs {
	AlignArgs A;
	const uint32_t *db2, *dbn;     // packed targets (k_pack_targets)
	const uint8_t *db_wild;        // per target: any letter outside ACGTU
	uint32_t ka, kb, stage;
	uint32_t *job_state;           // done << 31 | accepts << 16 | rejects
	float2 *job_ids;               // lowest / highest accepted identity per job (-termid, -termidd), else null
	uint32_t *verdict;             // n_jobs x k_max: 1 = failed the gate, rec << 2 | 2 = record
	uint32_t *items;               // jobs of this stage; null: every job (stage 0 and pairs mode)
	PassRec *recs;
	uint32_t recs_cap;
	uint32_t *hsp_arena;           // chained HSPs of the records: Loi, Loj, Len
	uint32_t hsp_arena_cap;        // words
	usb_hit *hits_stage;           // hits of the records that passed -id, until k_commit selects them
	uint32_t stage_cap;
	StageCounters *sc;
	// gate geometry (per warp)
	uint32_t g_qw, g_tw;           // u32 words of a packed query / target incl. 2 words of padding
	uint32_t g_start_bytes, g_q1_bytes, g_bytes;
};

// ------------------------------------------------------------------ packed database
// word offset of target t in db2 / dbn: one word per 16 letters (targets are padded to 16 letters)
// plus two zero words behind every target for ext16's look-ahead, rounded up to 16 bytes so that a
// packed target can be fetched with one bulk copy (cp.async.bulk needs 16-byte aligned addresses).
// Target t occupies [off, off + n16 + 2); off(t + 1) >= off(t) + n16 + 3.
__host__ __device__ inline uint64_t pack_off(const uint64_t *db_off, uint32_t t)
{
	return (db_off[t] / 16 + 6ull * t + 3) & ~3ull;
}
// words of db2 / dbn that hold n targets ending at letter offset end_off
__host__ __device__ inline uint64_t pack_words(uint64_t end_off, uint64_t n) { return end_off / 16 + 6 * n + 16; }

// One warp per target: 2-bit letters (wildcards as 0) and wildcard flags (bit 0 of each pair).
__global__ void k_pack_targets(const uint8_t *db_seq, const uint64_t *db_off, const uint32_t *db_len, uint32_t t0, uint32_t n,
  uint32_t *db2, uint32_t *dbn, uint8_t *db_wild)
{
	const uint32_t lane = threadIdx.x & 31;
	for (uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += gridDim.x * (blockDim.x >> 5)) {
		const uint32_t t = t0 + i;
		const uint32_t L = db_len[t];
		const uint4 *src = (const uint4 *)(db_seq + db_off[t]);
		const uint64_t o = pack_off(db_off, t);
		const uint32_t n16 = (L + 15) / 16;
		bool wild = false;
		for (uint32_t k = lane; k < n16 + 2; k += 32) {
			uint32_t p2 = 0, pn = 0;
			if (k < n16) {
				const uint4 v = __ldg(src + k);
				const uint32_t in[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
				for (int q = 0; q < 4; ++q)
#pragma unroll
					for (int b = 0; b < 4; ++b) {
						const uint32_t pos = 16 * k + 4 * q + b;
						const uint32_t c = pos < L ? nt_code((in[q] >> (8 * b)) & 0xff) : 0u;
						p2 |= (c & 3) << (2 * (4 * q + b));
						pn |= (c >> 2) << (2 * (4 * q + b));
					}
			}
			db2[o + k] = p2 & ~(pn * 3);
			dbn[o + k] = pn;
			wild |= pn != 0;
		}
		wild = __any_sync(USB_FULL, wild);
		if (lane == 0)
			db_wild[t] = wild ? 1 : 0;
	}
}

// ------------------------------------------------------------------ stage preparation
// Work items of stage [ka, kb): the jobs that are still running and have candidates in the range.
__global__ void k_stage_prep(const StageArgs S)
{
	const uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x;
	if (i0 == 0)
		S.sc->st[S.stage].rec_begin = S.sc->n_recs;
	if (!S.items)
		return;
	const uint32_t lane = threadIdx.x & 31;
	for (uint32_t j0 = blockIdx.x * blockDim.x; j0 < S.A.n_jobs; j0 += gridDim.x * blockDim.x) {
		const uint32_t job = j0 + threadIdx.x;
		const bool act = job < S.A.n_jobs && !(S.job_state[job] >> 31) && min(S.A.n_emit[job], S.kb) > S.ka;
		const uint32_t m = __ballot_sync(USB_FULL, act);
		uint32_t base = 0;
		if (lane == 0 && m)
			base = atomicAdd(&S.sc->st[S.stage].n_items, __popc(m));
		base = __shfl_sync(USB_FULL, base, 0);
		if (act)
			S.items[base + __popc(m & ((1u << lane) - 1))] = job;
	}
}

// ------------------------------------------------------------------ gate kernel
struct GateWs {
	WarpWs w;              // A2, An2, B2, Bn2, LA, LB, ung/order/prev/chain/cscore (slab) are set
	uint16_t *start, *pos; // CSR seed table of the query: first 8 positions of every word, query order
	uint32_t *q1b;
	uint16_t *q1a;
	uint32_t *q2b;
	uint16_t *q2a;
	uint32_t *B2base;      // two packed-target buffers filled by bulk copies: B2base + b * B2words
	uint32_t B2words;
	uint64_t *bar;         // their mbarriers
	uint32_t phase;        // bit b: parity barrier b completes next
	const uint8_t *Qraw;   // raw letters of the query (global), for wildcard identities
	uint32_t strand;
	bool wild;             // query or target holds a letter outside ACGTU
};

// Packs the query (strand-adjusted) two bits per letter; returns whether it holds a wildcard.
__device__ __forceinline__ bool gate_load_query(GateWs &g, const uint8_t *Q, uint32_t L, uint32_t strand)
{
	const uint32_t lane = lane_id();
	const uint32_t nw16 = (L + 15) / 16;
	bool wild = false;
	for (uint32_t k = lane; k < nw16 + 2; k += 32) {
		uint32_t v = 0, n = 0;
		if (k < nw16)
			for (uint32_t j = 0; j < 16 && 16 * k + j < L; ++j) {
				const uint32_t i = 16 * k + j;
				const uint32_t ch = strand ? (uint32_t)c_comp[Q[L - 1 - i]] : (uint32_t)Q[i];
				const uint32_t c = nt_code(ch);
				v |= (c & 3) << (2 * j);
				n |= (c >> 2) << (2 * j);
			}
		g.w.A2[k] = v & ~(n * 3);
		g.w.An2[k] = n;
		wild |= n != 0;
	}
	g.w.LA = L;
	g.Qraw = Q;
	g.strand = strand;
	__syncwarp();
	return __any_sync(USB_FULL, wild);
}

// ---- TMA staging of the packed targets: one lane arms the warp's mbarrier with the byte count
// and issues a bulk copy global -> shared (cp.async.bulk, SASS UBLKCP); the warp waits on the
// barrier's phase before it reads the buffer.  Two buffers per warp: the next candidate's letters
// are in flight while the current one is searched.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
	             "l"(src), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	asm volatile("{\n"
	             ".reg .pred p;\n"
	             "WAIT_%=:\n"
	             "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	             "@p bra DONE_%=;\n"
	             "bra WAIT_%=;\n"
	             "DONE_%=:\n"
	             "}" ::"r"(smem_u32(bar)),
	             "r"(parity)
	             : "memory");
}

// starts the copy of target t's packed letters into buffer b (lane 0 issues it)
__device__ __forceinline__ void gate_prefetch_target(const StageArgs &S, GateWs &g, uint32_t t, uint32_t b)
{
	if (lane_id() == 0) {
		const uint32_t nw = ((S.A.db_len[t] + 15) / 16 + 2 + 3) & ~3u; // whole 16-byte units (the arrays are padded)
		tma_load_1d(g.B2base + b * g.B2words, S.db2 + pack_off(S.A.db_off, t), 4 * nw, g.bar + b);
	}
}

// waits for buffer b (filled by gate_prefetch_target for target t) and makes it the current target
__device__ __forceinline__ void gate_take_target(const StageArgs &S, GateWs &g, uint32_t t, uint32_t b, bool twild)
{
	const uint32_t lane = lane_id();
	const uint32_t L = S.A.db_len[t];
	mbar_wait(g.bar + b, (g.phase >> b) & 1u);
	g.phase ^= 1u << b;
	g.w.B2 = g.B2base + b * g.B2words;
	g.w.B = S.A.db_seq + S.A.db_off[t];
	if (twild) {
		const uint64_t o = pack_off(S.A.db_off, t);
		const uint32_t nw = (L + 15) / 16 + 2;
		for (uint32_t k = lane; k < nw; k += 32)
			g.w.Bn2[k] = __ldg(S.dbn + o + k);
	}
	g.w.LB = L;
	__syncwarp();
}

// hspfinder.cpp:304-323 SetA as a CSR: start[word] .. start[word + 1] index pos[], at most 8 entries
__device__ void gate_seed_table(const StageArgs &S, GateWs &g)
{
	const uint32_t lane = lane_id();
	const uint32_t hw = S.A.P.hspw, HW = S.A.P.hsp_words;
	uint16_t *cnt = g.start;           // counts first, scanned in place
	uint8_t *fil = (uint8_t *)g.q1b;   // fill cursors (the queue is idle while the table is built)
	for (uint32_t i = lane; i < HW / 2 + 1; i += 32)
		((uint32_t *)cnt)[i] = 0;
	for (uint32_t i = lane; i < HW / 4; i += 32)
		((uint32_t *)fil)[i] = 0;
	const uint32_t nw = g.w.LA >= hw ? g.w.LA - hw + 1 : 0;
	g.w.nwordsA = nw;
	__syncwarp();
	for (uint32_t base = 0; base < nw; base += 32) {
		const uint32_t p = base + lane;
		const bool valid = p < nw;
		const uint32_t word = valid ? hsp_word_at(g.w.A2, p, HW) : (0x80000000u | lane);
		const uint32_t peers = __match_any_sync(USB_FULL, word);
		if (valid && (peers & lanemask_lt()) == 0)
			cnt[word] = (uint16_t)min((uint32_t)cnt[word] + __popc(peers), 8u);
		__syncwarp();
	}
	{
		const uint32_t per = HW / 32;
		uint32_t sum = 0;
		for (uint32_t i = 0; i < per; ++i)
			sum += cnt[lane * per + i];
		uint32_t inc = sum;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t t = __shfl_up_sync(USB_FULL, inc, d);
			if (lane >= (uint32_t)d)
				inc += t;
		}
		uint32_t run = inc - sum;
		for (uint32_t i = 0; i < per; ++i) {
			const uint32_t c = cnt[lane * per + i];
			cnt[lane * per + i] = (uint16_t)run;
			run += c;
		}
		if (lane == 31)
			cnt[HW] = (uint16_t)run;
	}
	__syncwarp();
	for (uint32_t base = 0; base < nw; base += 32) {
		const uint32_t p = base + lane;
		const bool valid = p < nw;
		const uint32_t word = valid ? hsp_word_at(g.w.A2, p, HW) : (0x80000000u | lane);
		const uint32_t peers = __match_any_sync(USB_FULL, word);
		uint32_t before = 0;
		if (valid)
			before = fil[word];
		__syncwarp();
		if (valid) {
			const uint32_t slot = before + __popc(peers & lanemask_lt());
			if (slot < 8)
				g.pos[g.start[word] + slot] = (uint16_t)p;
			if ((peers & lanemask_lt()) == 0)
				fil[word] = (uint8_t)min(before + __popc(peers), 8u);
		}
		__syncwarp();
	}
}

// One batch of queued survivors (at most 32, in scan order): every live seed walks, the results
// are committed in order (extend_queued of usb_align.cuh on the gate's queues).
template <bool WILD>
__device__ __forceinline__ void gate_extend(const StageArgs &S, GateWs &g, uint32_t n2, uint32_t MinLength, uint32_t &cur,
  uint32_t &nung)
{
	const uint32_t lane = lane_id();
	const DevParams &P = S.A.P;
	WarpWs &w = g.w;
	HspRec h;
	h.Loi = h.Loj = h.Len = 0;
	h.score2 = 0;
	uint32_t bhi = 0, bp = 0;
	bool ok = false;
	if (lane < n2) {
		bp = g.q2b[lane];
		const uint32_t ap = g.q2a[lane];
		if (bp >= cur) {
			const uint32_t seedmask = M55 & (P.hsp_words - 1);
			const uint32_t wseed = WILD ? (ext16(w.An2, ap) | ext16(w.Bn2, bp)) & seedmask : 0u;
			const int seed2 = P.match2 * (int)(P.hspw - __popc(wseed));
			ok = extend_seed_p<WILD>(S.A, w, ap, bp, seed2, MinLength, h, bhi);
		}
	}
	uint32_t okmask = __ballot_sync(USB_FULL, ok);
	while (okmask) {
		const int src = __ffs(okmask) - 1;
		okmask &= okmask - 1;
		const uint32_t sbp = __shfl_sync(USB_FULL, bp, src);
		if (sbp < cur)
			continue; // inside an HSP accepted a moment ago: the scan never sees this seed
		HspRec r;
		r.Loi = __shfl_sync(USB_FULL, h.Loi, src);
		r.Loj = __shfl_sync(USB_FULL, h.Loj, src);
		r.Len = __shfl_sync(USB_FULL, h.Len, src);
		r.score2 = __shfl_sync(USB_FULL, h.score2, src);
		const uint32_t sbhi = __shfl_sync(USB_FULL, bhi, src);
		if (nung < S.A.hsp_cap) {
			if (lane == 0)
				w.ung[nung] = r;
			++nung;
		} else if (lane == 0)
			atomicOr(&S.A.ctr->err, ERR_HSP_FULL);
		cur = sbhi + 1;
	}
}

// ungappedblast.cpp:45-210 as a stream: scan 128 target word positions per step (four consecutive
// positions per lane, their words cut from one 16-letter window) -> flat seed queue (a ring, scan
// order: position, then query order) -> pre-filter one seed per lane -> survivor queue -> walks in
// full batches of 32.  Seeds behind the sequential scan position `cur` are dropped wherever they are.
template <bool WILD> __device__ uint32_t gate_ungapped(const StageArgs &S, GateWs &g, uint32_t MinLength)
{
	const uint32_t lane = lane_id();
	const DevParams &P = S.A.P;
	WarpWs &w = g.w;
	const uint32_t hw = P.hspw, HW = P.hsp_words;
	const uint32_t LB = w.LB;
	if (LB < 2 * hw)
		return 0;
	const uint32_t nwordsB = LB - hw + 1;
	constexpr uint32_t QM = GATE_Q1 - 1;
	uint32_t nung = 0, scan = 0, cur = 0, n2 = 0;
	uint32_t head = 0, tail = 0; // ring of queued seeds [head, tail)
	// pre-filter: full batches of 32 queued seeds (all = true: the rest too)
	auto drain = [&](bool all) {
		while (tail - head >= 32 || (all && tail != head)) {
			const uint32_t cnt = min(32u, tail - head);
			bool keep = false;
			uint32_t bp = 0, ap = 0;
			if (lane < cnt) {
				bp = g.q1b[(head + lane) & QM];
				ap = g.q1a[(head + lane) & QM];
				int seed2;
				keep = bp >= cur && seed_may_pass_p<WILD>(S.A, w, ap, bp, seed2);
			}
			head += cnt;
			const uint32_t km = __ballot_sync(USB_FULL, keep);
			if (keep) {
				const uint32_t d = n2 + __popc(km & lanemask_lt());
				g.q2b[d] = bp;
				g.q2a[d] = (uint16_t)ap;
			}
			n2 += __popc(km);
			__syncwarp();
			if (n2 >= 32) {
				gate_extend<WILD>(S, g, 32, MinLength, cur, nung);
				const uint32_t rest = n2 - 32; // move the rest of the survivor queue to its front
				uint32_t mb = 0, ma = 0;
				if (lane < rest) {
					mb = g.q2b[32 + lane];
					ma = g.q2a[32 + lane];
				}
				__syncwarp();
				if (lane < rest) {
					g.q2b[lane] = mb;
					g.q2a[lane] = (uint16_t)ma;
				}
				n2 = rest;
				__syncwarp();
			}
		}
	};
	while (scan < nwordsB) {
		const uint32_t bpos0 = scan + 4 * lane;
		uint32_t st[4], na[4];
		uint32_t c = 0;
		{
			const uint32_t x = bpos0 < nwordsB ? ext16(w.B2, bpos0) : 0u;
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				st[i] = 0;
				na[i] = 0;
				if (bpos0 + i < nwordsB) {
					const uint32_t word = (x >> (2 * i)) & (HW - 1);
					st[i] = g.start[word];
					na[i] = (uint32_t)g.start[word + 1] - st[i];
				}
				c += na[i];
			}
		}
		uint32_t incl = c;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t t = __shfl_up_sync(USB_FULL, incl, d);
			if (lane >= (uint32_t)d)
				incl += t;
		}
		const uint32_t T = __shfl_sync(USB_FULL, incl, 31);
		if (tail - head + T <= GATE_Q1) {
			uint32_t d = tail + incl - c;
#pragma unroll
			for (int i = 0; i < 4; ++i)
				for (uint32_t j = 0; j < na[i]; ++j, ++d) {
					g.q1b[d & QM] = bpos0 + i;
					g.q1a[d & QM] = g.pos[st[i] + j];
				}
			tail += T;
			__syncwarp();
			drain(false);
		} else {
			// a low-complexity stretch (up to 8 seeds per position): 16 positions at a time
			for (uint32_t r = 0; r < 8; ++r) {
				const uint32_t bpos = scan + 16 * r + lane;
				uint32_t n1 = 0, s1 = 0;
				if (lane < 16 && bpos < nwordsB) {
					const uint32_t word = hsp_word_at(w.B2, bpos, HW);
					s1 = g.start[word];
					n1 = (uint32_t)g.start[word + 1] - s1;
				}
				uint32_t in2 = n1;
#pragma unroll
				for (int d = 1; d < 32; d <<= 1) {
					const uint32_t t = __shfl_up_sync(USB_FULL, in2, d);
					if (lane >= (uint32_t)d)
						in2 += t;
				}
				uint32_t d = tail + in2 - n1;
				for (uint32_t j = 0; j < n1; ++j, ++d) {
					g.q1b[d & QM] = bpos;
					g.q1a[d & QM] = g.pos[s1 + j];
				}
				tail += __shfl_sync(USB_FULL, in2, 31);
				__syncwarp();
				drain(false);
			}
		}
		scan += 128;
		if (cur > scan)
			scan = cur;
	}
	drain(true);
	if (n2)
		gate_extend<WILD>(S, g, n2, MinLength, cur, nung);
	__syncwarp();
	return nung;
}

// number of identical letter pairs of an ungapped HSP (hspfinder.cpp:561-579 GetHSPIdCount)
template <bool WILD> __device__ __forceinline__ uint32_t gate_hsp_ids(const GateWs &g, const HspRec &h)
{
	const uint32_t lane = lane_id();
	const WarpWs &w = g.w;
	uint32_t same = 0;
	for (uint32_t base = 0; base < h.Len; base += 512) {
		const uint32_t off = base + 16 * lane;
		if (off < h.Len) {
			const uint32_t k = min(16u, h.Len - off);
			const uint32_t mask = k == 16 ? M55 : ((1u << (2 * k)) - 1) & M55;
			const uint32_t x = ext16(w.A2, h.Loi + off) ^ ext16(w.B2, h.Loj + off);
			const uint32_t mis = (x | (x >> 1)) & M55;
			uint32_t wl = 0;
			if (WILD)
				wl = (ext16(w.An2, h.Loi + off) | ext16(w.Bn2, h.Loj + off)) & mask;
			same += __popc(~mis & ~wl & mask);
			while (wl) { // a wildcard pair: IUPAC-aware identity of the raw letters (alpha2.cpp:220-264)
				const uint32_t j = (uint32_t)(__ffs(wl) - 1) >> 1;
				wl &= wl - 1;
				const uint32_t qp = h.Loi + off + j, tp = h.Loj + off + j;
				const uint32_t ca = g.strand ? (uint32_t)c_comp[g.Qraw[w.LA - 1 - qp]] : (uint32_t)g.Qraw[qp];
				const uint32_t cb = w.B[tp];
				same += chars_match_dev(ca, cb, nt_code(ca), nt_code(cb)) ? 1u : 0u;
			}
		}
	}
#pragma unroll
	for (int d = 16; d; d >>= 1)
		same += __shfl_xor_sync(USB_FULL, same, d);
	return same;
}

// GlobalAlign_AllOpts up to the gate (globalalignmem.cpp:129-176).  Returns 0 = no alignment,
// 1 = goes on to the DP with nchain chained HSPs (w.chain / w.ung).
template <bool WILD> __device__ uint32_t gate_pair(const StageArgs &S, GateWs &g, uint32_t &nchain)
{
	const DevParams &P = S.A.P;
	WarpWs &w = g.w;
	const uint32_t LA = w.LA, LB = w.LB;
	nchain = 0;
	if (P.fulldp) // globalalignmem.cpp:153-157 FullDPAlways: no HSPs, no identity gate
		return (LA == 0 || LB == 0) ? 0u : 1u;
	uint32_t MinHSPLength = P.min_hsp_len == 0 ? 32 : P.min_hsp_len;
	MinHSPLength = min(MinHSPLength, LA / 4);
	MinHSPLength = max(MinHSPLength, 16u);
	const uint32_t nung = gate_ungapped<WILD>(S, g, MinHSPLength);
	nchain = chain_hsps(S.A, w, nung);
	uint32_t TotalLength = 0, TotalSame = 0;
	for (uint32_t c = 0; c < nchain; ++c) {
		const HspRec h = w.ung[w.chain[c]];
		TotalLength += h.Len;
		TotalSame += gate_hsp_ids<WILD>(g, h);
	}
	const float HSPFractId = TotalLength == 0 ? 0.0f : (float)TotalSame / (float)TotalLength;
	if (HSPFractId < P.min_hsp_fract_id)
		return 0;
	if (nchain == 0) {
		if (P.min_hsp_len > 0 && LA > 64)
			return 0;
		if (LA == 0 || LB == 0)
			return 0;
	}
	return 1;
}

__global__ void __launch_bounds__(GATE_MAX_WARPS * 32, 1) k_gate(const StageArgs S)
{
	extern __shared__ __align__(16) uint8_t stage_smem[];
	const AlignArgs &a = S.A;
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + warp;
	GateWs g;
	{
		uint8_t *p = stage_smem + (size_t)warp * S.g_bytes;
		g.w.A2 = (uint32_t *)p; p += 4 * S.g_qw;
		g.w.An2 = (uint32_t *)p; p += 4 * S.g_qw;
		g.B2base = (uint32_t *)p; p += 8 * S.g_tw;
		g.B2words = S.g_tw;
		g.w.B2 = g.B2base;
		g.w.Bn2 = (uint32_t *)p; p += 4 * S.g_tw;
		g.bar = (uint64_t *)p; p += 16;
		g.phase = 0;
		g.start = (uint16_t *)p; p += S.g_start_bytes;
		g.pos = (uint16_t *)p; p += pad16(2 * a.ql_cap);
		g.q1b = (uint32_t *)p; p += S.g_q1_bytes;
		g.q1a = (uint16_t *)p; p += pad16(2 * GATE_Q1);
		g.q2b = (uint32_t *)p; p += 4 * GATE_Q2;
		g.q2a = (uint16_t *)p;
		uint8_t *s = a.slab + (uint64_t)gw * a.slab_stride;
		g.w.ung = (HspRec *)s; s += (uint64_t)a.hsp_cap * sizeof(HspRec);
		g.w.order = (uint32_t *)s; s += (uint64_t)a.hsp_cap * 4;
		g.w.prev = (uint32_t *)s; s += (uint64_t)a.hsp_cap * 4;
		g.w.chain = (uint32_t *)s; s += (uint64_t)a.hsp_cap * 4;
		g.w.cscore = (int *)s;
		g.w.T = nullptr;
		// the wildcard arrays are only read when a wildcard is present; keep them defined anyway
		for (uint32_t k = lane; k < S.g_tw; k += 32)
			g.w.Bn2[k] = 0;
		if (lane == 0) {
			mbar_init(g.bar, 1);
			mbar_init(g.bar + 1, 1);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		__syncwarp();
	}
	const bool pairs = a.pair_q != nullptr;
	const uint32_t n_items = S.items ? S.sc->st[S.stage].n_items : a.n_jobs;
	uint32_t cur_job = 0xffffffffu;
	bool qwild = false;
	for (;;) {
		uint32_t it = 0;
		if (lane == 0)
			it = atomicAdd(&S.sc->st[S.stage].gate_cursor, 1u);
		it = __shfl_sync(USB_FULL, it, 0);
		if (it >= n_items)
			break;
		const uint32_t job = S.items ? S.items[it] : it;
		const uint32_t k0 = S.ka, k1 = pairs ? 1u : min(S.kb, a.n_emit[job]);
		if (k0 >= k1)
			continue;
		const uint32_t qi = pairs ? a.pair_q[job] : job / a.strands;
		const uint32_t strand = pairs ? 0u : job % a.strands;
		if (job != cur_job || pairs) {
			const uint64_t q0 = a.q_off[qi];
			const uint32_t L = (uint32_t)(a.q_off[qi + 1] - q0);
			qwild = gate_load_query(g, a.q + q0, L, strand);
			if (!a.P.fulldp)
				gate_seed_table(S, g);
			cur_job = job;
		}
		// the packed letters of candidate k + 1 are fetched while candidate k is searched; a
		// candidate whose pair is rejected beforehand (Accepter::RejectPair) is fetched all the same
		auto target_of = [&](uint32_t k) { return pairs ? a.pair_t[job] : a.cand_t[(uint64_t)job * a.k_max + k]; };
		const bool stage_targets = !a.P.fulldp;
		uint32_t buf = 0;
		if (stage_targets)
			gate_prefetch_target(S, g, target_of(k0), buf);
		for (uint32_t k = k0; k < k1; ++k) {
			const uint32_t t = target_of(k);
			const uint32_t cur = buf;
			if (stage_targets) {
				buf ^= 1u;
				__syncwarp(); // every lane is done with the other buffer (the candidate before this one)
				if (k + 1 < k1)
					gate_prefetch_target(S, g, target_of(k + 1), buf);
			}
			if (a.P.accept_flags & ACC_PAIR_FLAGS) {
				const uint64_t q0 = a.q_off[qi];
				if (reject_pair(a, qi, strand, t, a.q + q0, (uint32_t)(a.q_off[qi + 1] - q0))) {
					// skipped without a Terminator call, or counted as a reject on the big-database path
					if (lane == 0)
						S.verdict[(uint64_t)job * a.k_max + (pairs ? 0u : k)] = a.P.reject_pair_counts ? 1u : 3u;
					if (stage_targets) { // consume the buffer's phase
						mbar_wait(g.bar + cur, (g.phase >> cur) & 1u);
						g.phase ^= 1u << cur;
					}
					continue;
				}
			}
			const bool twild = S.db_wild[t] != 0;
			uint32_t nchain = 0, pass;
			if (a.P.fulldp) {
				g.w.LB = a.db_len[t];
				pass = (g.w.LA == 0 || g.w.LB == 0) ? 0u : 1u;
			} else {
				gate_take_target(S, g, t, cur, twild);
				g.wild = qwild || twild;
				pass = g.wild ? gate_pair<true>(S, g, nchain) : gate_pair<false>(S, g, nchain);
				if (twild) { // leave the wildcard array clean for the next target
					__syncwarp(); // (every lane is done reading it)
					const uint32_t nw = (g.w.LB + 15) / 16 + 2;
					for (uint32_t j = lane; j < nw; j += 32)
						g.w.Bn2[j] = 0;
					__syncwarp();
				}
			}
			if (pairs && a.hsp_out) {
				uint32_t *ho = a.hsp_out + (uint64_t)job * (1 + 4 * a.max_hsp);
				if (lane == 0) {
					ho[0] = nchain;
					for (uint32_t c = 0; c < nchain && c < a.max_hsp; ++c) {
						const HspRec h = g.w.ung[g.w.chain[c]];
						ho[1 + 4 * c] = h.Loi;
						ho[2 + 4 * c] = h.Loj;
						ho[3 + 4 * c] = h.Len;
						ho[4 + 4 * c] = (uint32_t)h.score2;
					}
				}
			}
			uint32_t v = 1;
			if (pass) {
				uint32_t rec = 0, ho = 0;
				if (lane == 0) {
					rec = atomicAdd(&S.sc->n_recs, 1u);
					ho = atomicAdd(&S.sc->n_hsp_words, 3 * nchain);
				}
				rec = __shfl_sync(USB_FULL, rec, 0);
				ho = __shfl_sync(USB_FULL, ho, 0);
				if (rec >= S.recs_cap || (uint64_t)ho + 3 * nchain > S.hsp_arena_cap) {
					if (lane == 0)
						atomicOr(&a.ctr->err, rec >= S.recs_cap ? ERR_REC_FULL : ERR_HSPARENA_FULL);
				} else {
					for (uint32_t c = lane; c < nchain; c += 32) {
						const HspRec h = g.w.ung[g.w.chain[c]];
						S.hsp_arena[ho + 3 * c] = h.Loi;
						S.hsp_arena[ho + 3 * c + 1] = h.Loj;
						S.hsp_arena[ho + 3 * c + 2] = h.Len;
					}
					if (lane == 0) {
						PassRec r;
						r.job = job; r.k = k; r.target = t; r.nchain = nchain; r.hsp_off = ho;
						r.hit = 0; r.n_dp = 0; r.dp_cells = 0;
						S.recs[rec] = r;
					}
					v = (rec << 2) | 2u;
				}
			}
			if (lane == 0)
				S.verdict[(uint64_t)job * a.k_max + (pairs ? 0u : k)] = v;
		}
	}
}

// ------------------------------------------------------------------ DP kernel
// Byte codes of the query (strand-adjusted) for the DP and the statistics.
__device__ __forceinline__ void dp_load_query(WarpWs &w, const uint8_t *Q, uint32_t L, uint32_t strand)
{
	__syncwarp(); // lanes may still be reading the previous record's arrays
	for (uint32_t i = lane_id(); i < L; i += 32) {
		const uint32_t c = strand ? (uint32_t)c_comp[Q[L - 1 - i]] : (uint32_t)Q[i];
		w.A[i] = (uint8_t)c;
		w.Ac[i] = (uint8_t)nt_code(c);
	}
	w.LA = L;
	__syncwarp();
}

// Byte codes of the target from its packed form: 16 letters per lane and step.
__device__ __forceinline__ void dp_load_target(const StageArgs &S, WarpWs &w, uint32_t t)
{
	const uint32_t L = S.A.db_len[t];
	const uint64_t o = pack_off(S.A.db_off, t);
	w.B = S.A.db_seq + S.A.db_off[t];
	const bool twild = S.db_wild[t] != 0;
	uint4 *dC = (uint4 *)w.Bc;
	const uint32_t n16 = (L + 15) / 16;
	__syncwarp(); // lanes may still be reading the previous record's arrays
	for (uint32_t k = lane_id(); k < n16; k += 32) {
		const uint32_t p2 = __ldg(S.db2 + o + k);
		const uint32_t pn = twild ? __ldg(S.dbn + o + k) : 0u;
		uint32_t out[4];
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			uint32_t v = 0;
#pragma unroll
			for (int b = 0; b < 4; ++b) {
				const uint32_t sh = 2 * (4 * q + b);
				v |= (((p2 >> sh) & 3u) | (((pn >> sh) & 1u) << 2)) << (8 * b);
			}
			out[q] = v;
		}
		dC[k] = make_uint4(out[0], out[1], out[2], out[3]);
	}
	w.LB = L;
	__syncwarp();
}

__global__ void __launch_bounds__(DP_MAX_WARPS * 32, 1) k_dp(const StageArgs S)
{
	extern __shared__ __align__(16) uint8_t stage_smem[];
	const AlignArgs &a = S.A;
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + warp;
	WarpWs w;
	{
		uint8_t *p = stage_smem + (size_t)warp * a.fast_bytes;
		w.A = p; p += a.ql_cap;
		w.Ac = p; p += a.ql_cap;
		w.Bc = p; p += a.tl_cap;
		w.scratch = p;
		w.cnt = p; // no seed table here: the "borrow the seed table" rows of viterbi_band never apply
		w.A2 = w.An2 = w.B2 = w.Bn2 = nullptr;
		w.start = w.pos = nullptr;
		uint8_t *s = a.slab + (uint64_t)gw * a.slab_stride;
		w.TB = s; s += ((uint64_t)(a.ql_cap + 1) * (a.tl_cap + 1) + 15) & ~(uint64_t)15;
		w.path = (char *)s; s += pad16(a.ql_cap + a.tl_cap + 16);
		w.rev = (char *)s; s += pad16(a.ql_cap + a.tl_cap + 16);
		w.rows_slab = (int *)s;
		w.ung = nullptr;
		w.order = w.prev = w.chain = nullptr;
		w.cscore = nullptr;
		w.T = nullptr;
		w.seed_dirty = false;
	}
	const bool pairs = a.pair_q != nullptr;
	const uint32_t rec_begin = S.sc->st[S.stage].rec_begin;
	const uint32_t rec_end = min(S.sc->n_recs, S.recs_cap);
	uint32_t cur_job = 0xffffffffu;
	for (;;) {
		uint32_t c = 0;
		if (lane == 0)
			c = atomicAdd(&S.sc->st[S.stage].dp_cursor, 1u);
		c = __shfl_sync(USB_FULL, c, 0);
		const uint32_t ri = rec_begin + c;
		if (ri >= rec_end)
			break;
		const PassRec r = S.recs[ri];
		const uint32_t job = r.job;
		const uint32_t qi = pairs ? a.pair_q[job] : job / a.strands;
		const uint32_t strand = pairs ? 0u : job % a.strands;
		if (job != cur_job || pairs) {
			const uint64_t q0 = a.q_off[qi];
			dp_load_query(w, a.q + q0, (uint32_t)(a.q_off[qi + 1] - q0), strand);
			cur_job = job;
		}
		dp_load_target(S, w, r.target);
		usb_qstat st;
		st.n_dp = 0;
		st.dp_cells = 0;
		const uint32_t LA = w.LA, LB = w.LB;
		char *path = w.path;
		uint32_t n = 0;
		if (r.nchain == 0) { // -fulldp, or no HSP on a short query (globalalignmem.cpp:153-157,173-179)
			const GapCosts G = hole_costs(a.P, true, true, true, true);
			++st.n_dp;
			n = viterbi_band<false>(a, w, w.Ac, LA, w.Bc, LB, G, path, nullptr, &st.dp_cells);
		} else {
			uint32_t Loi = 0, Loj = 0;
			for (uint32_t h = 0; h < r.nchain; ++h) {
				const uint32_t hLoi = S.hsp_arena[r.hsp_off + 3 * h], hLoj = S.hsp_arena[r.hsp_off + 3 * h + 1],
				               hLen = S.hsp_arena[r.hsp_off + 3 * h + 2];
				n += align_hole<false>(a, w, Loi, Loj, hLoi - Loi, hLoj - Loj, path + n, st);
				n += fill_run(path + n, 'M', hLen);
				Loi = hLoi + hLen;
				Loj = hLoj + hLen;
			}
			n += align_hole<false>(a, w, Loi, Loj, LA - Loi, LB - Loj, path + n, st);
		}
		__syncwarp();
		usb_hit h;
		h.query = qi; h.target = r.target; h.strand = strand; h.rank = pairs ? job : r.k;
		h.ql = LA; h.tl = LB; h.run_off = 0; h.run_cnt = 0; h.raw = 0; h.sub = 0;
		uint32_t hit = 0;
		if (!path_stats<false>(a, w, n, h)) {
			if (lane == 0)
				atomicOr(&a.ctr->err, ERR_NO_M);
		} else {
			// accepter.cpp:27-94: reject iff double(ids)/double(cols) < (double)(float)id, then the other rules
			if ((pairs || accept_hit(a, h, qi)) && emit_runs(a, w, n, h)) {
				if (lane == 0)
					hit = atomicAdd(&S.sc->n_staged, 1u);
				hit = __shfl_sync(USB_FULL, hit, 0);
				if (hit < S.stage_cap) {
					if (lane == 0)
						S.hits_stage[hit] = h;
					++hit;
				} else {
					if (lane == 0)
						atomicOr(&a.ctr->err, ERR_HITS_FULL);
					hit = 0;
				}
			}
		}
		if (lane == 0) {
			atomicAdd(&S.sc->dp_cells, (unsigned long long)st.dp_cells);
			atomicAdd(&S.sc->dp_seq_bytes, (unsigned long long)(LA + LB));
			S.recs[ri].hit = hit;
			S.recs[ri].n_dp = st.n_dp;
			S.recs[ri].dp_cells = st.dp_cells;
		}
	}
}

// ------------------------------------------------------------------ commit kernel
// One thread per job: the candidates of the stage in order, Terminator counts, hits to the output.
__global__ void k_commit(const StageArgs S)
{
	const AlignArgs &a = S.A;
	const bool pairs = a.pair_q != nullptr;
	for (uint32_t job = blockIdx.x * blockDim.x + threadIdx.x; job < a.n_jobs; job += gridDim.x * blockDim.x) {
		uint32_t state = S.stage == 0 ? 0u : S.job_state[job];
		if (state >> 31)
			continue;
		usb_qstat st;
		if (S.stage == 0) {
			st.n_cand = 0; st.n_tried = 0; st.n_hspfail = 0; st.n_dp = 0; st.dp_cells = 0; st.n_accept = 0; st.seq_bytes = 0;
		} else
			st = a.qstat[job];
		uint32_t acc = (state >> 16) & 0x7fff, rej = state & 0xffff;
		// HitMgr::GetMinFractId / GetMaxFractId (hitmgr.cpp:508-532) of the hits accepted so far
		float minid = 1.0f, maxid = 0.0f;
		if (S.job_ids && S.stage != 0) {
			minid = S.job_ids[job].x;
			maxid = S.job_ids[job].y;
		}
		const uint32_t ncand = pairs ? 1u : a.n_emit[job];
		const uint32_t n = min(ncand, S.kb);
		bool done = false;
		uint32_t LA = 0;
		if (n > S.ka) {
			const uint32_t qi = pairs ? a.pair_q[job] : job / a.strands;
			LA = (uint32_t)(a.q_off[qi + 1] - a.q_off[qi]);
		}
		for (uint32_t k = S.ka; k < n; ++k) {
			const uint32_t v = S.verdict[(uint64_t)job * a.k_max + k];
			const uint32_t t = pairs ? a.pair_t[job] : a.cand_t[(uint64_t)job * a.k_max + k];
			if (v == 3u) // RejectPair: no alignment, no Terminator call (searcher.cpp:63-67)
				continue;
			++st.n_tried;
			st.seq_bytes += LA + a.db_len[t];
			bool accept = false;
			if ((v & 3u) != 2u) {
				++st.n_hspfail;
				if (pairs)
					a.aligned[job] = 0;
			} else {
				const uint32_t ri = v >> 2;
				const PassRec r = S.recs[ri];
				st.n_dp += r.n_dp;
				st.dp_cells += r.dp_cells;
				accept = r.hit != 0;
				if (pairs)
					a.aligned[job] = 1;
				if (accept) {
					const usb_hit hh = S.hits_stage[r.hit - 1];
					const uint32_t slot = atomicAdd(&a.ctr->n_hits, 1u);
					if (slot < a.hits_cap)
						a.hits[slot] = hh;
					else
						atomicOr(&a.ctr->err, ERR_HITS_FULL);
					++st.n_accept;
					const float fid = (float)(hh.alnlen == 0 ? 0.0 : (double)hh.ids / (double)hh.alnlen);
					minid = fminf(minid, fid);
					maxid = fmaxf(maxid, fid);
				}
			}
			if (pairs)
				break;
			// terminator.cpp:64-100: -termid / -termidd look at the hits accepted so far (this one included)
			if ((a.P.accept_flags & USB_ACC_TERMID) && acc + (accept ? 1u : 0u) > 0 && (double)minid <= a.P.termid_d) {
				done = true;
				break;
			}
			if ((a.P.accept_flags & USB_ACC_TERMIDD) && acc + (accept ? 1u : 0u) > 0 &&
			    (double)(maxid - minid) > a.P.termidd_d) {
				done = true;
				break;
			}
			if (accept)
				++acc;
			else
				++rej;
			if ((a.P.maxaccepts > 0 && acc == a.P.maxaccepts) || (a.P.maxrejects > 0 && rej == a.P.maxrejects)) {
				done = true;
				break;
			}
		}
		if (S.kb >= ncand) {
			// every materialised candidate was looked at; with skipped pairs the reference may go on
			if (!done && a.n_cand_all && a.n_cand_all[job] > ncand)
				atomicOr(&a.ctr->err, ERR_KCAP);
			done = true;
		}
		S.job_state[job] = (done ? 0x80000000u : 0u) | (acc << 16) | rej;
		if (S.job_ids)
			S.job_ids[job] = make_float2(minid, maxid);
		if (a.qstat)
			a.qstat[job] = st;
	}
}

} // namespace usb
