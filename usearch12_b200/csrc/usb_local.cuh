// usb_local.cuh -- kernel K4: the -usearch_local candidate loop, one warp per (query, strand).
//
// Reference behaviour reproduced (results identical, algorithm re-designed for a warp):
//   a18 LocalAligner2::SetQueryImpl        localaligner2.cpp:64-155   query word -> positions
//       LocalAligner2::AlignMulti          localmulti.cpp:9-118       seeds in (target pos, query pos) order,
//                                                                     KeepAR overlap filter, skip past HSP end
//       LocalAligner::AlignPos             localaligner.cpp:101-211   ungapped X-drop both ways, anchor
//       GetAnchor                          localaligner.cpp:11-64
//   a19 XDropAlignMemMaxL2                 xdropalignmem.cpp:26-216   backward + forward gapped extension
//       XDropFwdFastMem / TraceBack        xdropfwdmem.cpp:271-749
//       XDropBwdFastMem                    xdropbwdmem.cpp:23-70
//   a20 EStats gates                       estats.cpp:65-96           (thresholds per query from the host)
//   a7  Searcher::Align (AlignMulti arm)   searcher.cpp:31-49, Terminator terminator.cpp:64-100
//   a17 AlignResult::FillLo                arscorer.cpp:201-296
//
// Lane mapping.  Seeds: one target word position per lane looks its word up in the query's sorted
// (word, position) list; the seeds are queued in reference order and their ungapped extensions
// (state-free) are evaluated 32 at a time, one per lane; the few that pass are then handled one
// after the other in order, which is all the order-dependent state (skip position, overlap list)
// needs.  Gapped X-drop DP: rows are sequential, one window column per lane; the insert state is a
// max-plus warp scan, the running best score a prefix maximum, and the data-dependent growth of
// the row to the right is resolved per 32-column chunk from a ballot of the growth tests
// (speculative cells to the right of the first cell that does not grow are discarded).
// All scores are integers (BLOSUM62 / integer nt scores, local gaps -10/-1); the float -9e9
// sentinel of the reference, which absorbs every addition, is the exact integer USB_NEG here
// (every dead value is clamped back to it), so even dead cells carry the reference's trace bits.
#pragma once
#include "usb_align.cuh" // Accepter rules shared by all candidate loops
#include "usb_dev.cuh"
#include "usb_tables.h"

namespace usb {

#define ERR_TB_FULL 64u
#define ERR_AR_FULL 128u
#define LOCAL_MAX_AR 32
#define LOCAL_SEEDQ 64
#define LOCAL_MAX_WARPS 16
#define LOCAL_MAXL 4096   // xdpmem.h:6 g_MaxL: longer extensions are split (XDropFwdSplit / XDropBwdSplit)
#define LOCAL_LONG_MAX 65000 // longest sequence of the long mode (positions are packed in 16 bits)

struct LocalDevTables {
	int8_t score[USB_NCODE * USB_NCODE];
	unsigned long long match[USB_NCODE];
	uint8_t word_letter[USB_NCODE];
	uint8_t code[256];
};

struct LocalArgs {
	DevParams P;
	const uint8_t *q;
	const uint64_t *q_off;
	uint32_t n_jobs, strands;
	const uint32_t *cand_t;    // n_jobs * k_max, from k_rank
	const uint32_t *n_emit;
	uint32_t k_max;
	const uint32_t *pair_q;    // pairs mode (usb_local_pairs) when non-null: job i = pair i, every AR is a hit
	const uint32_t *pair_t;
	const uint8_t *db_seq;
	const uint64_t *db_off;
	const uint32_t *db_len;
	const uint32_t *q_label, *q_size, *t_label, *t_size; // Accepter rules on labels / size= annotations (or null)
	const uint32_t *n_cand_all; // TopOrder.Size per job when skipped pairs can exhaust the materialised candidates
	const float *min_ungapped; // per query: (float) EStats::GetMinUngappedRawScore(QL)
	const int *min_gapped;     // per query: smallest raw score with RawScoreToEvalue(score) <= -evalue
	const LocalDevTables *tab;
	usb_hit *hits;
	uint32_t hits_cap;
	uint32_t *runs;
	uint32_t runs_cap;
	usb_qstat *qstat;
	uint8_t *slab;
	uint64_t slab_stride;
	uint32_t ql_cap, tl_cap, qk_cap; // padded capacities (letters, letters, sorted word keys = power of two)
	uint32_t fast_bytes;
	uint32_t long_mode;              // sequences above LOCAL_MAXL: letters and word keys live in the slab
	uint32_t row_cap;                // columns of the DP rows (tl_cap, or LOCAL_MAXL + 16 in long mode)
	uint32_t tb_cap;                 // trace bytes per warp
	float xdrop_u, xdrop_g;          // -xdrop_u, -xdrop_g
	float abs_open_f, abs_ext_f;
	int open, ext;                   // local gap open / extend (negative)
	uint32_t w, alpha, alpha_hi;     // LocalAligner2 word length, alphabet size, alpha^(w-1)
	DevCounters *ctr;
};

struct LocalShared {
	int8_t score[USB_NCODE * USB_NCODE];
	unsigned long long match[USB_NCODE];
	uint8_t word_letter[USB_NCODE];
};

struct LocalWs {
	uint8_t *A, *B;        // letter codes of the query (strand-adjusted) and of the current target
	uint32_t *qk;          // sorted (word << 16 | position) keys of the query
	int *Mrow, *Drow;      // X-drop DP rows, absolute column index
	uint32_t *seedq;       // queued seeds (qpos << 16 | tpos), LOCAL_SEEDQ entries
	// slab
	uint8_t *tb;           // trace bytes, rows stored back to back
	uint32_t *rowlo, *rowhi, *rowoff;
	char *path, *tmp;
	uint32_t LA, LB, nq;
};

inline __host__ __device__ uint32_t lpad16(uint32_t x) { return (x + 15u) & ~15u; }

// Long mode (a sequence above LOCAL_MAXL letters): the DP rows stay in shared memory -- an extension never
// spans more than LOCAL_MAXL columns, longer ones are split -- but the letters and the sorted word keys of
// whole sequences move to the slab.
inline __host__ __device__ uint32_t local_fast_bytes(uint32_t ql_cap, uint32_t tl_cap, uint32_t qk_cap, uint32_t row_cap,
  bool long_mode)
{
	const uint32_t rows = 2 * lpad16(4 * (row_cap + 8)) + 4 * LOCAL_SEEDQ;
	return long_mode ? rows : ql_cap + tl_cap + 4 * qk_cap + rows;
}

inline __host__ __device__ uint64_t local_slab_bytes(uint32_t ql_cap, uint32_t tl_cap, uint32_t tb_cap, uint32_t qk_cap,
  bool long_mode)
{
	return (uint64_t)lpad16(tb_cap) + 3ull * lpad16(4 * (ql_cap + 8)) + 2ull * lpad16(ql_cap + tl_cap + 16) +
	       (long_mode ? (uint64_t)ql_cap + tl_cap + 4ull * qk_cap : 0ull);
}

__device__ __forceinline__ void local_ws_setup(const LocalArgs &a, LocalWs &w, uint8_t *fast, uint8_t *slab)
{
	uint8_t *p = fast;
	if (!a.long_mode) {
		w.A = p; p += a.ql_cap;
		w.B = p; p += a.tl_cap;
		w.qk = (uint32_t *)p; p += 4 * a.qk_cap;
	}
	w.Mrow = (int *)p; p += lpad16(4 * (a.row_cap + 8));
	w.Drow = (int *)p; p += lpad16(4 * (a.row_cap + 8));
	w.seedq = (uint32_t *)p;
	uint8_t *s = slab;
	w.tb = s; s += lpad16(a.tb_cap);
	w.rowlo = (uint32_t *)s; s += lpad16(4 * (a.ql_cap + 8));
	w.rowhi = (uint32_t *)s; s += lpad16(4 * (a.ql_cap + 8));
	w.rowoff = (uint32_t *)s; s += lpad16(4 * (a.ql_cap + 8));
	w.path = (char *)s; s += lpad16(a.ql_cap + a.tl_cap + 16);
	w.tmp = (char *)s; s += lpad16(a.ql_cap + a.tl_cap + 16);
	if (a.long_mode) {
		w.A = s; s += a.ql_cap;
		w.B = s; s += a.tl_cap;
		w.qk = (uint32_t *)s;
	}
}

__device__ __forceinline__ int clampneg(int x) { return x < USB_NEG / 2 ? USB_NEG : x; }

// word of the LocalAligner2 alphabet starting at position p (wildcards are letter 0)
__device__ __forceinline__ uint32_t local_word(const LocalArgs &a, const LocalShared &S, const uint8_t *codes, uint32_t p)
{
	uint32_t word = 0;
	for (uint32_t i = 0; i < a.w; ++i)
		word = word * a.alpha + S.word_letter[codes[p + i]];
	return word;
}

// ascending bitonic sort of n <= cap keys by one warp (cap = power of two, padded with ~0)
__device__ void warp_sort_keys(uint32_t *k, uint32_t n, uint32_t cap)
{
	const uint32_t lane = lane_id();
	uint32_t P2 = 32;
	while (P2 < n)
		P2 <<= 1;
	if (P2 > cap)
		P2 = cap;
	for (uint32_t i = n + lane; i < P2; i += 32)
		k[i] = 0xffffffffu;
	__syncwarp();
	for (uint32_t kk = 2; kk <= P2; kk <<= 1)
		for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
			for (uint32_t i = lane; i < P2; i += 32) {
				const uint32_t x = i ^ j;
				if (x > i) {
					const uint32_t A = k[i], B = k[x];
					if ((A > B) == ((i & kk) == 0)) {
						k[i] = B;
						k[x] = A;
					}
				}
			}
			__syncwarp();
		}
}

// LocalAligner2::SetQueryImpl: the per-word position lists of the reference (ascending query
// position inside a word) are the runs of equal words in the sorted key array.
__device__ void local_load_query(const LocalArgs &a, const LocalShared &S, LocalWs &w, const uint8_t *Q, uint32_t L,
  uint32_t strand)
{
	const uint32_t lane = lane_id();
	for (uint32_t i = lane; i < L; i += 32) {
		const uint32_t c = strand ? (uint32_t)c_comp[Q[L - 1 - i]] : (uint32_t)Q[i];
		w.A[i] = a.tab->code[c];
	}
	w.LA = L;
	__syncwarp();
	const uint32_t nq = L > a.w ? L - a.w + 1 : 0; // QL <= WordLength: no words (localaligner2.cpp:81-82)
	for (uint32_t p = lane; p < nq; p += 32)
		w.qk[p] = (local_word(a, S, w.A, p) << 16) | p;
	w.nq = nq;
	__syncwarp();
	if (nq)
		warp_sort_keys(w.qk, nq, a.qk_cap);
}

__device__ __forceinline__ void local_load_target(const LocalArgs &a, LocalWs &w, uint32_t t)
{
	const uint32_t lane = lane_id();
	const uint32_t L = a.db_len[t];
	const uint8_t *src = a.db_seq + a.db_off[t];
	const uint4 *s4 = (const uint4 *)src;
	uint32_t *d32 = (uint32_t *)w.B;
	const uint32_t n16 = (L + 15) / 16;
	for (uint32_t i = lane; i < n16; i += 32) {
		const uint4 v = __ldg(s4 + i);
		const uint32_t in[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			uint32_t o = 0;
#pragma unroll
			for (int b = 0; b < 4; ++b)
				o |= (uint32_t)a.tab->code[(in[k] >> (8 * b)) & 0xff] << (8 * b);
			d32[4 * i + k] = o;
		}
	}
	w.LB = L;
	__syncwarp();
}

// first index in the sorted keys with key >= v
__device__ __forceinline__ uint32_t key_lower_bound(const uint32_t *k, uint32_t n, uint32_t v)
{
	uint32_t lo = 0, hi = n;
	while (lo < hi) {
		const uint32_t mid = (lo + hi) >> 1;
		if (k[mid] < v)
			lo = mid + 1;
		else
			hi = mid;
	}
	return lo;
}

// ------------------------------------------------------------------ a19: gapped X-drop extension
// One direction of XDropAlignMemMaxL2.  A'(i) = Ab[i * dA], B'(j) = Bb[j * dB] (dA = dB = -1 for the
// backward extension, which the reference runs on reversed copies).  Returns the best score and
// (Besti, Bestj); trace bytes go to w.tb with per-row windows in w.rowlo/rowhi/rowoff.
__device__ int xdrop_dp(const LocalArgs &a, const LocalShared &S, LocalWs &w, const uint8_t *Ab, int dA, uint32_t LA,
  const uint8_t *Bb, int dB, uint32_t LB, uint32_t &Besti, uint32_t &Bestj, uint32_t &rows, uint32_t &cells)
{
	const uint32_t lane = lane_id();
	int *Mrow = w.Mrow, *Drow = w.Drow;
	int best = S.score[Ab[0] * USB_NCODE + Bb[0]];
	Besti = 0;
	Bestj = 0;
	uint32_t pjlo = 0, pjhi = 0, jlo = 1, jhi = 1;
	if (lane == 0) {
		Mrow[0] = best;
		Drow[0] = USB_NEG;
		Drow[1] = USB_NEG;
	}
	__syncwarp();
	uint32_t tboff = 0;
	const float X = a.xdrop_g;
	uint32_t i = 1;
	bool overflow = false;
	for (; i < LA; ++i) {
		const int8_t *srow = S.score + (uint32_t)Ab[(int)i * dA] * USB_NCODE;
		uint32_t next_jlo = 0xffffffffu, njhi = 0xffffffffu;
		// carries across 32-column chunks
		int carry_m = (jlo == pjlo) ? USB_NEG : Mrow[jlo - 1]; // M0 of the chunk's first cell
		int carry_i = USB_NEG;                                  // I0 of the chunk's first cell
		int run_best = best;
		uint32_t row_bestj = 0xffffffffu;
		uint32_t jhi_cur = jhi;
		int last_m_old = USB_NEG;
		const uint32_t row_tb = tboff;
		if ((uint64_t)row_tb + (LB - jlo) + 2 > a.tb_cap) {
			overflow = true;
			break;
		}
		uint32_t jend = jhi; // final jhi of this row
		for (uint32_t base = jlo;; base += 32) {
			const uint32_t j = base + lane;
			const bool inb = j < LB;
			// previous-row values (columns beyond the previous window are dead)
			int m_old = USB_NEG, d_old = USB_NEG;
			if (inb) {
				if (j <= pjhi)
					m_old = Mrow[j];
				if (j <= pjhi + 1 && !(j == jlo && jlo == pjlo))
					d_old = Drow[j];
			}
			int m0 = __shfl_up_sync(USB_FULL, m_old, 1);
			if (lane == 0)
				m0 = carry_m;
			// insert state: I0(j+1) = max(M0(j) + open, I0(j) + ext), I0(first) = carry_i
			const int mi = clampneg(m0 + a.open);
			// y = mi - lane*ext  (ext < 0);  exclusive prefix max of y, then + (lane-1)*ext
			int y = mi <= USB_NEG ? USB_NEG : mi - (int)lane * a.ext;
			int inc = y;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const int t = __shfl_up_sync(USB_FULL, inc, d);
				if (lane >= (uint32_t)d)
					inc = max(inc, t);
			}
			int exc = __shfl_up_sync(USB_FULL, inc, 1);
			int i0 = USB_NEG;
			if (lane > 0 && exc > USB_NEG)
				i0 = exc + ((int)lane - 1) * a.ext;
			if (carry_i > USB_NEG)
				i0 = max(i0, carry_i + (int)lane * a.ext);
			i0 = clampneg(i0);
			// MATCH
			uint32_t bits = 0;
			int xm = m0;
			if (d_old > xm) {
				xm = d_old;
				bits = TB_DM;
			}
			if (i0 > xm) {
				xm = i0;
				bits = TB_IM;
			}
			const int sc = inb ? clampneg(xm + (int)srow[Bb[(int)j * dB]]) : USB_NEG;
			// running best before / after this cell (prefix maximum in column order)
			int pm = sc;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const int t = __shfl_up_sync(USB_FULL, pm, d);
				if (lane >= (uint32_t)d)
					pm = max(pm, t);
			}
			int bb = __shfl_up_sync(USB_FULL, pm, 1);
			bb = lane == 0 ? run_best : max(bb, run_best);
			const int ba = max(bb, sc);
			const float hM = (float)(sc - bb) + X;
			// DELETE (not for the first column of the row)
			int d_new = d_old;
			float hD = -1.0f;
			if (j != jlo) {
				const int md = clampneg(m0 + a.open);
				const int de = clampneg(d_old + a.ext);
				if (md >= de) {
					d_new = md;
					bits |= TB_MD;
				} else
					d_new = de;
				hD = (float)(d_new - ba) + X;
			}
			// INSERT
			const int ie = clampneg(i0 + a.ext);
			int i_next;
			if (mi >= ie) {
				i_next = mi;
				bits |= TB_MI;
			} else
				i_next = ie;
			const float hI = (float)(i_next - ba) + X;
			// which speculative cells exist: the row grows past jhi while the last cell passes the
			// growth test (xdropfwdmem.cpp:533-556,618-636)
			const bool grow = inb && j + 1 < LB && (hM > a.abs_ext_f || hI > a.abs_ext_f);
			const uint32_t growmask = __ballot_sync(USB_FULL, grow);
			uint32_t n_exist; // number of existing lanes in this chunk
			bool more;        // row continues into the next chunk
			if (jhi_cur >= base + 32) {
				n_exist = 32;
				more = true;
			} else {
				const uint32_t l0 = jhi_cur - base; // lane of the current last column
				const uint32_t stopm = ~growmask & ~((1u << l0) - 1);
				if (stopm) {
					n_exist = (uint32_t)__ffs(stopm); // lanes 0 .. ffs-1
					more = false;
				} else {
					n_exist = 32;
					more = true;
					jhi_cur = base + 32; // the next chunk's first column exists
				}
			}
			const bool ex = lane < n_exist && inb;
			const uint32_t exmask = __ballot_sync(USB_FULL, ex);
			// window of the next row
			uint32_t cand = 0xffffffffu;
			if (ex) {
				if (hM > 0.0f)
					cand = j + 1;
				if (hM > a.abs_open_f)
					cand = j;
				if (hD > 0.0f)
					cand = min(cand, j - 1);
				if (hI > 0.0f)
					cand = min(cand, j + 1);
			}
			next_jlo = min(next_jlo, __reduce_min_sync(USB_FULL, cand));
			const uint32_t am = __ballot_sync(USB_FULL, ex && (hM > 0.0f || hI > 0.0f));
			const uint32_t dm = __ballot_sync(USB_FULL, ex && hD > 0.0f);
			if (am) {
				const uint32_t la = 31 - __clz(am);
				njhi = base + la + 1;
				const uint32_t dab = la == 31 ? 0u : (dm & ~((2u << la) - 1));
				if (dab)
					njhi = max(njhi, base + (31 - __clz(dab)) - 1);
			} else if (dm)
				njhi = max(njhi, base + (31 - __clz(dm)) - 1);
			// best cell: later cells win ties (xdropfwdmem.cpp:558-566)
			{
				const int cm = ex ? sc : USB_NEG;
				const int rmax = __reduce_max_sync(USB_FULL, cm);
				if (exmask && rmax >= run_best) {
					const uint32_t eq = __ballot_sync(USB_FULL, ex && sc == rmax);
					run_best = rmax;
					row_bestj = base + (31 - __clz(eq));
				}
			}
			// stores
			if (ex) {
				Mrow[j] = sc;
				if (j != jlo)
					Drow[j] = d_new;
				w.tb[row_tb + (j - jlo)] = (uint8_t)bits;
			}
			const uint32_t last = n_exist - 1;
			carry_m = __shfl_sync(USB_FULL, m_old, 31);
			carry_i = __shfl_sync(USB_FULL, i_next, 31);
			if (!more) {
				jend = base + last;
				last_m_old = __shfl_sync(USB_FULL, m_old, last);
				break;
			}
			__syncwarp();
		}
		__syncwarp();
		// end of the D row (xdropfwdmem.cpp:649-671): column jend + 1 <= LB
		if (lane == 0) {
			const uint32_t j1 = jend + 1;
			int d_old = USB_NEG;
			if (j1 <= pjhi + 1 && !(j1 == jlo && jlo == pjlo))
				d_old = Drow[j1];
			const int md = clampneg(last_m_old + a.open);
			const int de = clampneg(d_old + a.ext);
			uint8_t b = 0;
			int dn = de;
			if (md >= de) {
				dn = md;
				b = TB_MD;
			}
			Drow[j1] = dn;
			w.tb[row_tb + (j1 - jlo)] = b;
			w.rowlo[i] = jlo;
			w.rowhi[i] = jend;
			w.rowoff[i] = row_tb;
		}
		tboff = row_tb + (jend - jlo) + 2;
		cells += jend - jlo + 1;
		if (row_bestj != 0xffffffffu) {
			best = run_best;
			Besti = i;
			Bestj = row_bestj;
		}
		__syncwarp();
		if (next_jlo == 0xffffffffu) {
			++i;
			break;
		}
		pjlo = jlo;
		pjhi = jend;
		jlo = min(next_jlo, LB - 1);
		jhi = min(njhi, LB - 1);
	}
	rows = i; // rows 1 .. rows-1 hold trace bytes
	if (overflow) {
		if (lane == 0)
			atomicOr(&a.ctr->err, ERR_TB_FULL);
		return 0;
	}
	return best;
}

// XDropFwdTraceBackBitMem (xdropfwdmem.cpp:271-342) by lane 0: raw traceback order (best cell
// first) into out; returns the number of columns, 0 on a trace error.
__device__ uint32_t xdrop_traceback(const LocalArgs &a, const LocalWs &w, uint32_t Besti, uint32_t Bestj, uint32_t rows,
  char *out, uint32_t cap)
{
	uint32_t n = 0;
	if (lane_id() == 0) {
		uint32_t i = Besti, j = Bestj;
		char st = 'M';
		bool bad = false;
		for (;;) {
			if (n >= cap) {
				bad = true;
				break;
			}
			out[n++] = st;
			if (i == 0 && j == 0)
				break;
			uint32_t ri, cj;
			if (st == 'M') {
				ri = i; cj = j;
			} else if (st == 'D') {
				ri = i; cj = j + 1;
			} else {
				ri = i + 1; cj = j;
			}
			if (ri < 1 || ri >= rows || cj < w.rowlo[ri] || cj > w.rowhi[ri] + 1) {
				bad = true;
				break;
			}
			const uint8_t c = w.tb[w.rowoff[ri] + (cj - w.rowlo[ri])];
			char nx;
			if (st == 'M') {
				nx = (c & TB_DM) ? 'D' : (c & TB_IM) ? 'I' : 'M';
				if (i == 0 || j == 0) { bad = true; break; }
				--i; --j;
			} else if (st == 'D') {
				nx = (c & TB_MD) ? 'M' : 'D';
				if (i == 0) { bad = true; break; }
				--i;
			} else {
				nx = (c & TB_MI) ? 'M' : 'I';
				if (j == 0) { bad = true; break; }
				--j;
			}
			st = nx;
		}
		if (bad) {
			atomicOr(&a.ctr->err, ERR_TRACE);
			n = 0;
		}
	}
	return __shfl_sync(USB_FULL, n, 0);
}

struct LocalHsp {
	uint32_t Loi, Loj, Leni, Lenj;
	int score;
};

// XDropAlignMemMaxL2 (xdropalignmem.cpp:26-216): w.path receives the whole path; returns its length
// (0 = no alignment).
// xdropfwdsplit.cpp:15-22 GetSubL
__device__ __forceinline__ uint32_t local_sub_l(uint32_t L)
{
	if (L <= LOCAL_MAXL)
		return L;
	if (L < 2 * LOCAL_MAXL)
		return L / 2;
	return LOCAL_MAXL;
}

// One extension of XDropAlignMemMaxL2 in one direction (dir = -1 backward from (i0, j0) over la x lb
// letters, +1 forward), the piece XDropFwdFastMem / XDropBwdFastMem compute: score, lengths and the path
// in w.tmp -- in raw traceback order, which is the forward order of the sequences for the backward
// extension and the reverse of it for the forward one.  *m = path length; false = trace arena full.
__device__ bool xdrop_piece(const LocalArgs &a, const LocalShared &S, LocalWs &w, uint32_t i0, uint32_t j0, int dir, uint32_t la,
  uint32_t lb, int &score, uint32_t &leni, uint32_t &lenj, uint32_t &m, usb_qstat &st)
{
	const uint32_t lane = lane_id();
	if (la == 1 || lb == 1) {
		score = S.score[w.A[i0] * USB_NCODE + w.B[j0]];
		leni = 1;
		lenj = 1;
		if (lane == 0)
			w.tmp[0] = 'M';
		m = 1;
		__syncwarp();
		return true;
	}
	uint32_t bi, bj, rows = 0, cells = 0;
	score = xdrop_dp(a, S, w, w.A + i0, dir, la, w.B + j0, dir, lb, bi, bj, rows, cells);
	++st.n_dp;
	st.dp_cells += cells;
	if (score <= 0) {
		score = 0;
		leni = 0;
		lenj = 0;
		m = 0;
		return true;
	}
	m = xdrop_traceback(a, w, bi, bj, rows, w.tmp, a.ql_cap + a.tl_cap);
	if (m == 0)
		return false;
	leni = bi + 1;
	lenj = bj + 1;
	__syncwarp();
	return true;
}

__device__ uint32_t xdrop_align(const LocalArgs &a, const LocalShared &S, LocalWs &w, uint32_t AncLoi, uint32_t AncLoj,
  uint32_t AncLen, LocalHsp &H, usb_qstat &st)
{
	const uint32_t lane = lane_id();
	H.score = 0;
	if (AncLen <= 1)
		return 0;
	const uint32_t LA = w.LA, LB = w.LB;
	const uint32_t AncHii = AncLoi + AncLen - 1, AncHij = AncLoj + AncLen - 1;
	uint32_t n = 0;
	// backward: prefixes A[0..AncLoi], B[0..AncLoj] reversed; above g_MaxL letters in pieces, each piece
	// in front of the ones before it (XDropBwdSplit, xdropbwdsplit.cpp:15-79: PrependPath)
	uint32_t BwdLeni = 0, BwdLenj = 0;
	int BwdScore = 0;
	{
		const uint32_t la = AncLoi + 1, lb = AncLoj + 1;
		const bool split = AncLoi > LOCAL_MAXL || AncLoj > LOCAL_MAXL; // xdropalignmem.cpp:87
		uint32_t doneA = 0, doneB = 0;
		for (;;) {
			if (split && (doneA == la || doneB == lb))
				break;
			const uint32_t sla = split ? local_sub_l(la - doneA) : la, slb = split ? local_sub_l(lb - doneB) : lb;
			int sc;
			uint32_t li, lj, m;
			if (!xdrop_piece(a, S, w, AncLoi - doneA, AncLoj - doneB, -1, sla, slb, sc, li, lj, m, st))
				return 0;
			if (split && sc == 0)
				break;
			BwdScore += sc;
			BwdLeni += li;
			BwdLenj += lj;
			if (m) {
				// shift what is there to the right by m (from the end, 32 letters at a time), piece in front
				for (uint32_t k = n; k > 0;) {
					const uint32_t c = min(32u, k);
					char ch = 0;
					if (lane < c)
						ch = w.path[k - c + lane];
					__syncwarp();
					if (lane < c)
						w.path[k - c + lane + m] = ch;
					__syncwarp();
					k -= c;
				}
				for (uint32_t k = lane; k < m; k += 32)
					w.path[k] = w.tmp[k];
				n += m;
				__syncwarp();
			}
			if (!split || (li < sla && lj < slb))
				break;
			doneA += li;
			doneB += lj;
		}
	}
	// the anchor without its first and last column (they belong to the two extensions)
	for (uint32_t k = lane; k + 2 < AncLen; k += 32)
		w.path[n + k] = 'M';
	n += AncLen - 2;
	// forward: suffixes from the last anchor column (XDropFwdSplit, xdropfwdsplit.cpp:24-91: AppendPath)
	uint32_t FwdLeni = 0, FwdLenj = 0;
	int FwdScore = 0;
	{
		const uint32_t la = LA - AncHii, lb = LB - AncHij;
		const bool split = la > LOCAL_MAXL || lb > LOCAL_MAXL; // xdropalignmem.cpp:120
		for (;;) {
			if (split && (FwdLeni == la || FwdLenj == lb))
				break;
			const uint32_t sla = split ? local_sub_l(la - FwdLeni) : la, slb = split ? local_sub_l(lb - FwdLenj) : lb;
			int sc;
			uint32_t li, lj, m;
			if (!xdrop_piece(a, S, w, AncHii + FwdLeni, AncHij + FwdLenj, 1, sla, slb, sc, li, lj, m, st))
				return 0;
			if (split && sc == 0)
				break;
			FwdScore += sc;
			FwdLeni += li;
			FwdLenj += lj;
			if (m == 1 && (sla == 1 || slb == 1)) {
				if (lane == 0)
					w.path[n] = 'M';
			} else
				for (uint32_t k = lane; k < m; k += 32)
					w.path[n + k] = w.tmp[m - 1 - k];
			n += m;
			__syncwarp();
			if (!split || (li < sla && lj < slb))
				break;
		}
	}
	__syncwarp();
	// total = Bwd + Fwd + Anchor - the two duplicated end columns (xdropalignmem.cpp:162-189)
	int anc = 0;
	for (uint32_t k = lane; k < AncLen; k += 32)
		anc += S.score[w.A[AncLoi + k] * USB_NCODE + w.B[AncLoj + k]];
	anc = __reduce_add_sync(USB_FULL, anc);
	const int dupe = S.score[w.A[AncLoi] * USB_NCODE + w.B[AncLoj]] + S.score[w.A[AncHii] * USB_NCODE + w.B[AncHij]];
	H.score = BwdScore + FwdScore + anc - dupe;
	H.Loi = AncLoi + 1 - BwdLeni;
	H.Loj = AncLoj + 1 - BwdLenj;
	H.Leni = BwdLeni + FwdLeni + AncLen - 2;
	H.Lenj = BwdLenj + FwdLenj + AncLen - 2;
	return n;
}

// GetAnchor (localaligner.cpp:11-64) by lane 0: best run of strictly positive pair scores.
__device__ int local_anchor(const LocalShared &S, const LocalWs &w, uint32_t Loi, uint32_t Loj, uint32_t L,
  uint32_t &AncLoi, uint32_t &AncLoj, uint32_t &AncLen)
{
	int bestscore = 0;
	uint32_t beststart = 0, bestlen = 0;
	if (lane_id() == 0) {
		uint32_t startk = 0xffffffffu;
		int anc = 0;
		for (uint32_t k = 0; k < L; ++k) {
			const int sc = S.score[w.A[Loi + k] * USB_NCODE + w.B[Loj + k]];
			if (sc > 0) {
				if (startk == 0xffffffffu) {
					startk = k;
					anc = sc;
				} else
					anc += sc;
			} else {
				if (anc > bestscore) {
					bestscore = anc;
					beststart = startk;
					bestlen = k - startk;
				}
				startk = 0xffffffffu;
			}
		}
		if (anc > bestscore) {
			bestscore = anc;
			beststart = startk;
			bestlen = L - startk;
		}
	}
	bestscore = __shfl_sync(USB_FULL, bestscore, 0);
	AncLoi = Loi + __shfl_sync(USB_FULL, beststart, 0);
	AncLoj = Loj + __shfl_sync(USB_FULL, beststart, 0);
	AncLen = __shfl_sync(USB_FULL, bestlen, 0);
	return bestscore;
}

// ------------------------------------------------------------------ a17: statistics of a local hit
// FillLo (arscorer.cpp:201-296) over the whole path (a local path starts and ends with M).
__device__ void local_path_stats(const LocalShared &S, const LocalWs &w, uint32_t n, const LocalHsp &H, usb_hit &h)
{
	const uint32_t lane = lane_id();
	const char *path = w.path;
	uint32_t qpos = H.Loi, tpos = H.Loj;
	uint32_t ids = 0, mism = 0, gaps = 0, opens = 0;
	uint32_t prev_carry = 'M';
	for (uint32_t base = 0; base < n; base += 32) {
		const uint32_t c = base + lane;
		const bool in = c < n;
		const uint32_t ch = in ? (uint32_t)path[c] : 0u;
		const bool isM = ch == 'M', isD = ch == 'D', isI = ch == 'I';
		const uint32_t qm = __ballot_sync(USB_FULL, isM || isD);
		const uint32_t tm = __ballot_sync(USB_FULL, isM || isI);
		const uint32_t qp = qpos + __popc(qm & lanemask_lt());
		const uint32_t tp = tpos + __popc(tm & lanemask_lt());
		bool same = false;
		if (isM)
			same = (S.match[w.A[qp]] >> w.B[tp]) & 1ull;
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
	h.first_mq = H.Loi;
	h.first_mt = H.Loj;
	h.first_mcol = 0;
	h.alnlen = n;
	h.ids = ids;
	h.mism = mism;
	h.intgaps = gaps;
	h.opens = opens;
	h.last_mq = qpos - 1;
	h.last_mt = tpos - 1;
}

// Appends the path as runs (length << 2 | op) to the run arena; returns false when full.
__device__ bool local_emit_runs(const LocalArgs &a, const LocalWs &w, uint32_t n, usb_hit &h)
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
			runs[k + __popc(sm & lanemask_lt())] = (c << 2) | op;
		}
		k += __popc(sm);
		carry = __shfl_sync(USB_FULL, ch, 31);
	}
	__syncwarp();
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

// HSPData::OverlapFract > 0.5 (hsp.h:74-89, localaligner2.cpp:252-258); x = new AR, y = kept AR
__device__ __forceinline__ bool local_large_overlap(const LocalHsp &x, const LocalHsp &y)
{
	if (x.Leni == 0 || x.Lenj == 0)
		return false;
	const uint32_t xHii = x.Loi + x.Leni - 1, xHij = x.Loj + x.Lenj - 1;
	const uint32_t yHii = y.Loi + y.Leni - 1, yHij = y.Loj + y.Lenj - 1;
	const uint32_t MaxLoi = max(x.Loi, y.Loi), MaxLoj = max(x.Loj, y.Loj);
	const uint32_t MinHii = min(xHii, yHii), MinHij = min(xHij, yHij);
	const uint32_t Ovi = MinHii < MaxLoi ? 0 : MinHii - MaxLoi;
	const uint32_t Ovj = MinHij < MaxLoj ? 0 : MinHij - MaxLoj;
	const double f = (double)(Ovi * Ovj) / (double)(x.Leni * x.Lenj);
	return f > 0.5;
}

// ------------------------------------------------------------------ one candidate target
struct LocalTargetState {
	LocalHsp kept;      // lane k holds the k-th kept AR of this target
	uint32_t n_kept;
	uint32_t next_tpos; // target positions below this are skipped (localmulti.cpp:105-111)
	bool any_accept;
};

// Ungapped extension of one seed (localaligner.cpp:109-165), one seed per lane.
__device__ __forceinline__ int local_ungapped(const LocalArgs &a, const LocalShared &S, const LocalWs &w, uint32_t qpos,
  uint32_t tpos, uint32_t &LeftLength, uint32_t &RightLength)
{
	int best = 0, tot = 0;
	uint32_t len = 0, k = 0;
	int i = (int)qpos, j = (int)tpos;
	while (i >= 0 && j >= 0) {
		++k;
		tot += S.score[w.A[i] * USB_NCODE + w.B[j]];
		if (tot > best) {
			best = tot;
			len = k;
		} else if ((float)(best - tot) > a.xdrop_u)
			break;
		--i;
		--j;
	}
	LeftLength = len;
	int rbest = 0;
	tot = 0;
	len = 0;
	k = 0;
	i = (int)qpos + 1;
	j = (int)tpos + 1;
	while (i < (int)w.LA && j < (int)w.LB) {
		++k;
		tot += S.score[w.A[i] * USB_NCODE + w.B[j]];
		if (tot > rbest) {
			rbest = tot;
			len = k;
		} else if ((float)(rbest - tot) > a.xdrop_u)
			break;
		++i;
		++j;
	}
	RightLength = len;
	return best + rbest;
}

// Evaluates the first n (<= 32) queued seeds; handles the passing ones in order.
__device__ void local_process_seeds(const LocalArgs &a, const LocalShared &S, LocalWs &w, uint32_t n, uint32_t qi,
  uint32_t t, uint32_t strand, uint32_t rank, LocalTargetState &ts, usb_qstat &st, bool pairs)
{
	const uint32_t lane = lane_id();
	uint32_t qpos = 0, tpos = 0, ll = 0, rl = 0;
	bool pass = false;
	if (lane < n) {
		const uint32_t sd = w.seedq[lane];
		qpos = sd >> 16;
		tpos = sd & 0xffff;
		if (tpos >= ts.next_tpos) {
			const int sc = local_ungapped(a, S, w, qpos, tpos, ll, rl);
			pass = !((float)sc < a.min_ungapped[qi]);
		}
	}
	uint32_t pm = __ballot_sync(USB_FULL, pass);
	while (pm) {
		const uint32_t l = (uint32_t)__ffs(pm) - 1;
		pm &= pm - 1;
		const uint32_t sq = __shfl_sync(USB_FULL, qpos, l), stp = __shfl_sync(USB_FULL, tpos, l);
		const uint32_t sl = __shfl_sync(USB_FULL, ll, l), sr = __shfl_sync(USB_FULL, rl, l);
		if (stp < ts.next_tpos)
			continue;
		const uint32_t Loi = sq + 1 - sl, Loj = stp + 1 - sl, SegLength = sl + sr;
		uint32_t AncLoi, AncLoj, AncLen;
		const int anc = local_anchor(S, w, Loi, Loj, SegLength, AncLoi, AncLoj, AncLen);
		if (anc <= 0)
			continue;
		LocalHsp H;
		const uint32_t np = xdrop_align(a, S, w, AncLoi, AncLoj, AncLen, H, st);
		if (np == 0 || H.score <= 0)
			continue;
		if (H.score < a.min_gapped[qi]) // E-value gate (localaligner.cpp:198-203)
			continue;
		// KeepAR (localaligner2.cpp:239-250)
		const bool ov = lane < ts.n_kept && local_large_overlap(H, ts.kept);
		if (__any_sync(USB_FULL, ov))
			continue;
		if (ts.n_kept >= LOCAL_MAX_AR) {
			if (lane == 0)
				atomicOr(&a.ctr->err, ERR_AR_FULL);
			continue;
		}
		if (lane == ts.n_kept)
			ts.kept = H;
		const uint32_t sub = ts.n_kept++;
		const uint32_t nt = H.Loj + H.Lenj; // Hij + 1
		ts.next_tpos = nt > stp ? nt : stp + 1;
		// Searcher::Align: Accepter on every AR of the target (searcher.cpp:36-47)
		usb_hit h;
		h.query = qi; h.target = t; h.strand = strand; h.rank = rank;
		h.ql = w.LA; h.tl = w.LB; h.run_off = 0; h.run_cnt = 0;
		h.raw = H.score;
		h.sub = sub;
		local_path_stats(S, w, np, H, h);
		// Accepter::IsAcceptLo (accepter.cpp:41-94); -evalue was applied above
		const bool accept = pairs || accept_hit(a, h, qi, true);
		if (accept) {
			if (local_emit_runs(a, w, np, h)) {
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
			ts.any_accept = true;
			++st.n_accept;
		}
	}
}

// LocalAligner2::AlignMulti for the staged query and target.  Returns AnyAccepts.
__device__ bool local_align_multi(const LocalArgs &a, const LocalShared &S, LocalWs &w, uint32_t qi, uint32_t t,
  uint32_t strand, uint32_t rank, usb_qstat &st, bool pairs)
{
	const uint32_t lane = lane_id();
	LocalTargetState ts;
	ts.n_kept = 0;
	ts.next_tpos = 0;
	ts.any_accept = false;
	ts.kept.Loi = ts.kept.Loj = ts.kept.Leni = ts.kept.Lenj = 0;
	ts.kept.score = 0;
	const uint32_t TL = w.LB;
	if (TL < 2 * a.w || w.nq == 0)
		return false;
	const uint32_t nwords = TL - a.w + 1;
	uint32_t nqd = 0; // queued seeds
	for (uint32_t base = 0; base < nwords;) {
		if (base < ts.next_tpos) { // everything below next_tpos is skipped anyway
			base = ts.next_tpos;
			continue;
		}
		const uint32_t tp = base + lane;
		uint32_t lo = 0, cnt = 0;
		if (tp < nwords) {
			const uint32_t word = local_word(a, S, w.B, tp);
			lo = key_lower_bound(w.qk, w.nq, word << 16);
			uint32_t hi = lo;
			while (hi < w.nq && (w.qk[hi] >> 16) == word)
				++hi;
			cnt = hi - lo;
		}
		// seeds of this chunk in (target position, query position) order
		uint32_t inc = cnt;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t x = __shfl_up_sync(USB_FULL, inc, d);
			if (lane >= (uint32_t)d)
				inc += x;
		}
		const uint32_t total = __shfl_sync(USB_FULL, inc, 31);
		const uint32_t excl = inc - cnt;
		uint32_t done = 0; // seeds of this chunk already queued
		while (done < total) {
			const uint32_t room = LOCAL_SEEDQ - nqd;
			const uint32_t take = min(room, total - done);
			// lane's seeds with chunk index in [done, done + take)
			for (uint32_t k = 0; k < cnt; ++k) {
				const uint32_t g = excl + k;
				if (g >= done && g < done + take)
					w.seedq[nqd + (g - done)] = ((w.qk[lo + k] & 0xffff) << 16) | tp;
			}
			nqd += take;
			done += take;
			__syncwarp();
			while (nqd >= 32) {
				local_process_seeds(a, S, w, 32, qi, t, strand, rank, ts, st, pairs);
				__syncwarp();
				// shift the rest down
				const uint32_t rest = nqd - 32;
				uint32_t v0 = 0;
				if (lane < rest)
					v0 = w.seedq[32 + lane];
				__syncwarp();
				if (lane < rest)
					w.seedq[lane] = v0;
				nqd = rest;
				__syncwarp();
			}
		}
		base += 32;
	}
	if (nqd) {
		__syncwarp();
		local_process_seeds(a, S, w, nqd, qi, t, strand, rank, ts, st, pairs);
	}
	return ts.any_accept;
}

// ------------------------------------------------------------------ the job loop
__device__ void local_job(const LocalArgs &a, const LocalShared &S, LocalWs &w, uint32_t job)
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
		local_load_query(a, S, w, a.q + q0, L, strand);
		uint32_t acc = 0, rej = 0;
		bool stopped = pairs;
		for (uint32_t k = 0; k < ncand; ++k) {
			const uint32_t t = pairs ? a.pair_t[job] : a.cand_t[(uint64_t)job * a.k_max + k];
			// (Accepter::RejectPair rules are refused on the host for local searches: the reference applies them
			// per AR inside IsAccept, searcher.cpp:26-49, and crashes when one rejects a pair)
			local_load_target(a, w, t);
			++st.n_tried;
			st.seq_bytes += w.LA + w.LB;
			const bool any = local_align_multi(a, S, w, qi, t, strand, pairs ? job : k, st, pairs);
			if (pairs)
				break;
			if (!any)
				++st.n_hspfail;
			// terminator.cpp:64-100, one count per target (searcher.cpp:48)
			if (any)
				++acc;
			else
				++rej;
			if ((a.P.maxaccepts > 0 && acc == a.P.maxaccepts) || (a.P.maxrejects > 0 && rej == a.P.maxrejects)) {
				stopped = true;
				break;
			}
		}
		if (!stopped && a.n_cand_all && a.n_cand_all[job] > ncand && lane == 0)
			atomicOr(&a.ctr->err, ERR_KCAP);
	}
	if (lane == 0 && a.qstat)
		a.qstat[job] = st;
}

__global__ void __launch_bounds__(LOCAL_MAX_WARPS * 32, 1) k_local(const LocalArgs a)
{
	extern __shared__ __align__(16) uint8_t local_smem[];
	LocalShared &S = *(LocalShared *)local_smem;
	for (uint32_t i = threadIdx.x; i < USB_NCODE * USB_NCODE; i += blockDim.x)
		S.score[i] = a.tab->score[i];
	for (uint32_t i = threadIdx.x; i < USB_NCODE; i += blockDim.x) {
		S.match[i] = a.tab->match[i];
		S.word_letter[i] = a.tab->word_letter[i];
	}
	__syncthreads();
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + warp;
	uint8_t *fast = local_smem + ((sizeof(LocalShared) + 15) & ~(size_t)15) + (size_t)warp * a.fast_bytes;
	uint8_t *slab = a.slab + (uint64_t)gw * a.slab_stride;
	LocalWs w;
	local_ws_setup(a, w, fast, slab);
	for (;;) {
		uint32_t job = 0;
		if (lane == 0)
			job = atomicAdd(&a.ctr->job, 1u);
		job = __shfl_sync(USB_FULL, job, 0);
		if (job >= a.n_jobs)
			break;
		local_job(a, S, w, job);
	}
}

} // namespace usb
