// usb_dev.cuh -- device-side constants, parameter block and small helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/usb200.h"

namespace usb {

#define USB_FULL 0xffffffffu
// Dead-state sentinel of the integer DP.  The reference uses float -9e9 (mx.h:12), which absorbs
// additions; here every reachable score is > -2^27 so NEG + (anything reachable) stays below all
// live values and trace bits of live states are identical.
#define USB_NEG (-(1 << 28))

// trace bits, same meaning as tracebit.h:4-7
#define TB_DM 1
#define TB_IM 2
#define TB_MD 4
#define TB_MI 8

// error bits reported through DevCounters.err
#define ERR_HITS_FULL 1u
#define ERR_RUNS_FULL 2u
#define ERR_HSP_FULL 4u
#define ERR_TRACE 8u
#define ERR_RECORDS_FULL 16u
#define ERR_NO_M 32u

struct DevParams {
	int match2, mismatch2;               // 2 x substitution scores (setnucmx.cpp:33-87)
	int open2, ext2, topen2, text2;      // 2 x gap penalties (alnparams.cpp:378-384)
	float xdrop2;                        // 2 x XDropGlobalHSP (alnheuristics.cpp:31)
	float minscore2;                     // 2 x MinGlobalHSPScore (alnheuristics.cpp:40-44)
	float min_hsp_fract_id;              // MinGlobalHSPFractId = max(id, 0.75) (alnheuristics.cpp:39)
	uint32_t min_hsp_len;                // MinGlobalHSPLength (-minhsp)
	uint32_t band;                       // BandRadius
	uint32_t hspw, hsp_words, hsp_hi;    // HSP finder word length, 4^w, 4^(w-1)
	uint32_t word_length, slots;         // UDB
	uint32_t maxaccepts, maxrejects;
	uint32_t bump;
	uint32_t fulldp;                     // -fulldp (FullDPAlways)
	uint32_t alpha;                      // UDB alphabet size: 4 (nt) or 20 (aa), udbparams.cpp:235-261
	uint32_t hash_cap;                   // aa: entries of the per-CTA word hash set (power of two), else 0
	double id_d;                         // (double)(float)id (accepter.cpp:36-38)
	// Accepter / Terminator options beyond -id (usb200.h USB_ACC_*), widened like oget_flt does
	uint32_t accept_flags;
	uint32_t mincols, maxgaps, maxdiffs, mindiffs;
	uint32_t reject_pair_counts;         // 1: a RejectPair-ed target counts as a reject (big-database path,
	                                     // udbusortedsearcherbig.cpp:119-128), 0: it is skipped (searcher.cpp:63-67)
	double maxid_d, query_cov_d, max_query_cov_d, target_cov_d, max_target_cov_d, abskew_d, min_sizeratio_d;
	double minqt_d, maxqt_d, minsl_d, maxsl_d, termid_d, termidd_d;
};

struct DevCounters {
	uint32_t job;      // dynamic job cursor
	uint32_t n_hits;
	uint32_t n_runs;
	uint32_t err;
	unsigned long long postings; // UDB postings walked by k_rank
	uint32_t job_rank;           // job cursor of k_rank_big
	uint32_t pad;
	unsigned long long prof[16];  // k_rank phase cycles, summed over CTAs (only with RankArgs.prof)
};

// The UDB index as the kernels see it: a few CSR segments over consecutive target ranges
// (usb_hostindex.h).  A word's posting row is the concatenation of its fragments in segment order,
// which is ascending target order.
#define USB_MAX_SEG 24
struct SegDesc {
	const uint64_t *row_off;   // slots + 1, multiples of 4
	const uint32_t *row_size;  // slots
	const uint32_t *postings;  // global target indexes
	uint32_t base, count;
};
struct IndexView {
	SegDesc seg[USB_MAX_SEG];
	uint32_t n_seg;
	uint32_t n_seq;
	// 2-byte increment descriptors (usb_hostindex.h HostHalf): one static segment; then
	// seg[0].postings / row_off are null and only k_rank may walk the index
	const uint16_t *post16;
	const uint64_t *row_off16;   // slots + 1, entries (multiples of 256)
	const uint32_t *row_groups;  // slots: groups of 256 entries
};

struct HspRec {
	uint32_t Loi, Loj, Len;
	int score2;
};

__constant__ uint16_t c_cls[256];
__constant__ uint8_t c_upper[256];
__constant__ uint8_t c_comp[256];
__constant__ uint8_t c_udb_aa[256];   // amino UDB letter of a raw character, 0xff = bad (udbparams.cpp:546-552)

// 0..3 for ACGTU in either case, 4 for everything else (alpha.cpp g_CharToLetterNucleo).
__device__ __forceinline__ uint32_t nt_code(uint32_t c)
{
	uint32_t u = c & 0xDFu;
	uint32_t r = 4;
	r = (u == 'A') ? 0u : r;
	r = (u == 'C') ? 1u : r;
	r = (u == 'G') ? 2u : r;
	r = (u == 'T' || u == 'U') ? 3u : r;
	return ((c | 0x20u) >= 'a' && (c | 0x20u) <= 'z') ? r : 4u;
}

// UDB letter (udbparams.cpp:540-555): lower-case or non-ACGTU kills the word.
__device__ __forceinline__ uint32_t udb_letter(uint32_t c)
{
	uint32_t r = 4;
	r = (c == 'A') ? 0u : r;
	r = (c == 'C') ? 1u : r;
	r = (c == 'G') ? 2u : r;
	r = (c == 'T' || c == 'U') ? 3u : r;
	return r;
}

// %id identity of two raw characters (alpha2.cpp:220-264); ca/cb are their nt_code()s.
__device__ __forceinline__ bool chars_match_dev(uint32_t a, uint32_t b, uint32_t ca, uint32_t cb)
{
	if ((ca | cb) < 4)
		return ca == cb;
	uint32_t xa = c_cls[a], xb = c_cls[b];
	if (!(xa & 0x100) || !(xb & 0x100))
		return (xa & 0x200) && (xb & 0x200);
	if (c_upper[a] == c_upper[b])
		return true;
	return ((xa & 0xf) & (xb >> 4)) || ((xb & 0xf) & (xa >> 4));
}

__device__ __forceinline__ int subst2(const DevParams &P, uint32_t ca, uint32_t cb)
{
	return ((ca | cb) & 4) ? 0 : (ca == cb ? P.match2 : P.mismatch2);
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1; }

} // namespace usb
