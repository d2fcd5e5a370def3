// usb_api.cu -- C ABI (include/usb200.h) over the CUDA kernels.  No CPU fallback: every compute
// entry point needs a CUDA device and fails with USB_ECUDA otherwise.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <string>
#include <unordered_map>
#include <thread>
#include <vector>

#include "usb_align.cuh"
#include "usb_stage.cuh"
#include "usb_local.cuh"
#include "usb_hostindex.h"
#include "usb_rank.cuh"
#include "usb_rankbig.cuh"
#include "usb_usortfull.cuh"

using namespace usb;

// ------------------------------------------------------------------ errors
static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...)
{
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	g_err = buf;
	return code;
}

// same for the host-only translation units of the library (usb_udbfile.cpp)
namespace usb {
int fail_msg(int code, const char *msg)
{
	g_err = msg;
	return code;
}
}

#define CK(call)                                                                                     \
	do {                                                                                             \
		cudaError_t e_ = (call);                                                                     \
		if (e_ != cudaSuccess)                                                                       \
			return fail(USB_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)

extern "C" const char *usb_last_error(void) { return g_err.c_str(); }

// cudaFuncSetAttribute state is per function and per device, shared by every searcher and host
// thread of the process: it is tracked here under one mutex and the dynamic shared-memory limit is
// only ever raised.  The mutex is held across "set attribute + launch" so that a launch never sees
// a limit or carve-out another thread chose.
enum { FN_RANK_F, FN_RANK_T, FN_RANK_BIG, FN_ALIGN_NT, FN_ALIGN_AA, FN_VITERBI, FN_LOCAL, FN_GATE, FN_DP, FN_CONFLICT, FN_COUNT };
static std::mutex g_attr_mu;
static size_t g_attr_smem[64][FN_COUNT];
static int g_attr_carve[64][FN_COUNT];

// call with g_attr_mu held
static cudaError_t func_smem_locked(const void *fn, int id, size_t bytes, int carve = -2)
{
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess)
		return e;
	if (dev < 0 || dev >= 64)
		return cudaErrorInvalidDevice;
	if (bytes > g_attr_smem[dev][id]) {
		if ((e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)) != cudaSuccess)
			return e;
		g_attr_smem[dev][id] = bytes;
	}
	if (carve != -2 && carve != g_attr_carve[dev][id] - 1000) {
		if ((e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, carve)) != cudaSuccess)
			return e;
		g_attr_carve[dev][id] = carve + 1000; // 0 = never set
	}
	return cudaSuccess;
}

extern "C" int usb_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

extern "C" void usb_default_params(usb_params *p, int cluster_fast)
{
	memset(p, 0, sizeof *p);
	p->struct_size = sizeof(usb_params);
	p->is_nucleo = 1;
	p->id = 0.97f;
	p->maxaccepts = 1;
	p->maxrejects = cluster_fast ? 8 : 32;
	p->strand_both = 0;
	p->word_length = 8;
	p->big = 100000;
	p->bump = 50;
	p->stepwords = 8;
	p->band = 16;
	p->minhsp = 16;
	p->hspw = 5;
	p->xdrop_nw = 8.0f;
	p->match = 1.0f;
	p->mismatch = -2.0f;
	p->gap_open = -10.0f;
	p->gap_ext = -1.0f;
	p->term_gap_open = -0.5f;
	p->term_gap_ext = -0.5f;
	p->dbmask = cluster_fast ? 0 : 1;
	p->cluster_mode = cluster_fast;
	p->local = 0;
	p->evalue = 10.0f;
	p->xdrop_u = 16.0f;
	p->xdrop_g = 32.0f;
	p->lopen = -10.0f;
	p->lext = -1.0f;
	p->ka_dbsize = 1e9f;
	p->accept_flags = USB_ACC_MAXID; // o_defaults.inc:7: -maxid has a default, which counts as set
	p->maxid = 1.0f;
}

extern "C" void usb_set_local(usb_params *p, int nucleo, float evalue)
{
	p->local = 1;
	p->evalue = evalue;
	p->is_nucleo = nucleo ? 1 : 0;
	if (nucleo) {
		p->word_length = 8;
		p->hspw = 5;
	} else {
		p->word_length = 5;
		p->hspw = 3;
	}
}

// Amino acid database searched globally (makedbsearcher.cpp:132-140 builds a GlobalAligner for either
// alphabet): UDB words of 5 letters over 20 (udbparams.cpp:251-257), HSP words of 3
// (alnheuristics.cpp:37), BLOSUM62, gap open -17 / extend -1 (alnparams.cpp:381-384), no strands.
extern "C" void usb_set_amino(usb_params *p)
{
	p->is_nucleo = 0;
	p->strand_both = 0;
	p->word_length = 5;
	p->hspw = 3;
	p->gap_open = -17.0f;
	p->gap_ext = -1.0f;
}

#define CLUSTER_ROWS_CAP 32 // sampled words per query the cluster round's conflict kernel handles

// ------------------------------------------------------------------ objects
template <class T> struct DevBuf {
	T *p = nullptr;
	size_t cap = 0;
	int reserve(size_t n)
	{
		if (n <= cap)
			return 0;
		if (p)
			cudaFree(p);
		p = nullptr;
		cap = 0;
		size_t want = n + n / 8 + 64;
		cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
		if (e != cudaSuccess) {
			cudaGetLastError();
			return fail(USB_ENOMEM, "cudaMalloc of %zu bytes failed: %s", want * sizeof(T), cudaGetErrorString(e));
		}
		cap = want;
		return 0;
	}
	// grows to hold n elements, keeping the first `used` (device-to-device copy), doubling
	int grow_keep(size_t n, size_t used)
	{
		if (n <= cap)
			return 0;
		size_t want = std::max(n + n / 8 + 64, cap * 2);
		T *q = nullptr;
		cudaError_t e = cudaMalloc((void **)&q, want * sizeof(T));
		if (e != cudaSuccess) {
			cudaGetLastError();
			return fail(USB_ENOMEM, "cudaMalloc of %zu bytes failed: %s", want * sizeof(T), cudaGetErrorString(e));
		}
		if (p && used) {
			e = cudaMemcpy(q, p, used * sizeof(T), cudaMemcpyDeviceToDevice);
			if (e != cudaSuccess) {
				cudaGetLastError();
				cudaFree(q);
				return fail(USB_ECUDA, "device copy of %zu bytes failed: %s", used * sizeof(T), cudaGetErrorString(e));
			}
		}
		if (p)
			cudaFree(p);
		p = q;
		cap = want;
		return 0;
	}
	void release()
	{
		if (p)
			cudaFree(p);
		p = nullptr;
		cap = 0;
	}
};

// One immutable CSR segment of the index on the host and on the device.  On the device the
// postings are either 4-byte ascending rows (any kernel) or the 2-byte bank-aware layout of
// HostHalf (k_rank only), never both.
struct IndexSegment {
	HostCSR H;
	DevBuf<uint64_t> d_row_off;
	DevBuf<uint32_t> d_row_size, d_postings, d_row_groups;
	DevBuf<uint16_t> d_post16;
	bool on_dev = false, half = false;
	int upload(bool want_half, uint32_t n_targets)
	{
		if (on_dev && half == want_half)
			return 0;
		int rc;
		if (want_half) {
			HostHalf hh;
			make_half(H, n_targets, 0, hh);
			d_postings.release();
			if ((rc = d_row_off.reserve(hh.row_off.size())) || (rc = d_row_size.reserve(H.row_size.size())) ||
			    (rc = d_row_groups.reserve(hh.row_groups.size())) || (rc = d_post16.reserve(hh.postings.size())))
				return rc;
			CK(cudaMemcpy(d_row_off.p, hh.row_off.data(), hh.row_off.size() * 8, cudaMemcpyHostToDevice));
			CK(cudaMemcpy(d_row_size.p, H.row_size.data(), H.row_size.size() * 4, cudaMemcpyHostToDevice));
			CK(cudaMemcpy(d_row_groups.p, hh.row_groups.data(), hh.row_groups.size() * 4, cudaMemcpyHostToDevice));
			CK(cudaMemcpy(d_post16.p, hh.postings.data(), hh.postings.size() * 2, cudaMemcpyHostToDevice));
		} else {
			d_post16.release();
			d_row_groups.release();
			if ((rc = d_row_off.reserve(H.row_off.size())) || (rc = d_row_size.reserve(H.row_size.size())) ||
			    (rc = d_postings.reserve(H.postings.size())))
				return rc;
			CK(cudaMemcpy(d_row_off.p, H.row_off.data(), H.row_off.size() * 8, cudaMemcpyHostToDevice));
			CK(cudaMemcpy(d_row_size.p, H.row_size.data(), H.row_size.size() * 4, cudaMemcpyHostToDevice));
			CK(cudaMemcpy(d_postings.p, H.postings.data(), H.postings.size() * 4, cudaMemcpyHostToDevice));
		}
		on_dev = true;
		half = want_half;
		return 0;
	}
	void release()
	{
		d_row_off.release();
		d_row_size.release();
		d_postings.release();
		d_post16.release();
		d_row_groups.release();
		on_dev = false;
	}
};

// USB_TIMING=1: wall-clock phases of the append path, printed at exit (measurement aid)
struct AppendTimers {
	double t[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
	uint64_t calls = 0;
	bool on = getenv("USB_TIMING") != nullptr;
	double tk = 0;
	static double now()
	{
		struct timespec ts;
		clock_gettime(CLOCK_MONOTONIC, &ts);
		return ts.tv_sec + 1e-9 * ts.tv_nsec;
	}
	void start()
	{
		if (on)
			tk = now();
	}
	void lap(int i)
	{
		if (on) {
			cudaDeviceSynchronize();
			const double x = now();
			t[i] += x - tk;
			tk = x;
		}
	}
	~AppendTimers()
	{
		if (on && calls)
			fprintf(stderr, "usb_index_append: %llu calls; host copy %.3fs letters up %.3fs pack %.3fs | count+scan %.3fs sizes down %.3fs "
			                "plan %.3fs pool grow %.3fs move %.3fs fill %.3fs tables up %.3fs | layout %.3fs\n",
			  (unsigned long long)calls, t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8], t[9], t[10]);
	}
};
static AppendTimers g_at;

#include "usb_dynseg.inc"

struct usb_index {
	int device = 0;
	usb_params P;
	HostSeqs S;                          // masked SeqDB (host copy)
	std::vector<IndexSegment *> segs;    // ascending target ranges covering [0, dyn ? dyn->base : S.n())
	DynSegment *dyn = nullptr;           // growable tail segment for small appends (cluster_fast)
	uint32_t n_dev = 0;                  // targets whose letters are on the device
	uint64_t n_postings = 0;
	bool no_half = false;                // a kernel other than k_rank needed the 4-byte postings
	DevBuf<uint8_t> d_seqs;
	DevBuf<uint64_t> d_seq_off;
	DevBuf<uint32_t> d_seq_len;
	// nucleotide targets packed two bits per letter + wildcard flags (k_pack_targets, usb_stage.cuh)
	DevBuf<uint32_t> d_db2, d_dbn;
	DevBuf<uint8_t> d_wild;
	// label identities / size= annotations of the targets (usb_index_set_attrs); n_attr = targets covered
	DevBuf<uint32_t> d_t_label, d_t_size;
	uint32_t n_label = 0, n_size = 0;
	std::vector<uint32_t> row_tmp;       // usb_index_row scratch
};

#include "usb_ixbuild.inc"
#include "usb_derep.inc"

// 2-byte postings: one static segment from target 0, at most USB_HALF_MAX_TARGETS, and no use of the
// big-database or cluster kernels (they walk 4-byte rows).
static bool index_wants_half(const usb_index *ix)
{
	return !ix->no_half && !ix->P.cluster_mode && !ix->dyn && ix->segs.size() == 1 && ix->segs[0]->H.base == 0 &&
	       ix->S.n() <= USB_HALF_MAX_TARGETS && ix->S.n() <= ix->P.big && !getenv("USB_NO_HALF");
}

// Brings every segment to the layout the index wants now (no-op when nothing changed).
static int index_sync_layout(usb_index *ix)
{
	const bool half = index_wants_half(ix);
	for (IndexSegment *g : ix->segs) {
		int rc = g->upload(half, ix->S.n());
		if (rc)
			return rc;
	}
	return 0;
}

static void fill_index_view(const usb_index *ix, IndexView &v)
{
	memset(&v, 0, sizeof v);
	v.n_seg = (uint32_t)ix->segs.size();
	v.n_seq = ix->S.n();
	for (uint32_t i = 0; i < v.n_seg; ++i) {
		const IndexSegment *g = ix->segs[i];
		v.seg[i].row_off = g->half ? nullptr : g->d_row_off.p;
		v.seg[i].row_size = g->d_row_size.p;
		v.seg[i].postings = g->d_postings.p;
		if (g->half) {
			v.post16 = g->d_post16.p;
			v.row_off16 = g->d_row_off.p;
			v.row_groups = g->d_row_groups.p;
		}
		v.seg[i].base = g->H.base;
		v.seg[i].count = g->H.count;
	}
	if (ix->dyn) {
		SegDesc &d = v.seg[v.n_seg++];
		d.row_off = ix->dyn->d_row_off.p;
		d.row_size = ix->dyn->d_row_size.p;
		d.postings = ix->dyn->d_pool.p;
		d.base = ix->dyn->base;
		d.count = ix->dyn->count;
	}
}

// Result arrays are not value-initialised on resize (the download overwrites them); they are plain
// pageable memory: page-locking a fresh 100 MB result costs more (about 1.3 ms per MB here) than
// copying it out of the searcher's one page-locked staging buffer.
template <class T> struct NoInitAlloc {
	using value_type = T;
	NoInitAlloc() = default;
	template <class U> NoInitAlloc(const NoInitAlloc<U> &) {}
	T *allocate(size_t n)
	{
		void *p = aligned_alloc(64, (n * sizeof(T) + 63) & ~(size_t)63);
		if (!p)
			throw std::bad_alloc();
		return (T *)p;
	}
	void deallocate(T *q, size_t) { free(q); }
	template <class U> void construct(U *p) noexcept { ::new ((void *)p) U; }
	template <class U, class... A> void construct(U *p, A &&...a) { ::new ((void *)p) U(std::forward<A>(a)...); }
	template <class U> bool operator==(const NoInitAlloc<U> &) const { return true; }
	template <class U> bool operator!=(const NoInitAlloc<U> &) const { return false; }
};
template <class T> using NoInitVec = std::vector<T, NoInitAlloc<T>>;

struct usb_result {
	NoInitVec<usb_hit> hits;
	NoInitVec<uint32_t> runs;
	NoInitVec<uint64_t> qoff;
	NoInitVec<usb_qstat> qstat;
};

struct usb_searcher {
	usb_index *ix = nullptr;
	usb_params P;
	DevParams D;
	int num_sms = 0;
	size_t smem_optin = 0, smem_per_sm = 0, l2_persist_bytes = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
	uint64_t launches = 0;
	// current batch
	uint32_t n_q = 0, n_jobs = 0, strands = 1, k_max = 0, max_ql = 0;
	bool ran = false;
	uint32_t last_hits = 0, last_runs = 0;
	uint64_t last_postings = 0;
	DevBuf<uint32_t> d_grp_cnt;   // download_result: hits per group, then scatter cursors
	DevBuf<uint64_t> d_grp_off;
	DevBuf<usb_hit> d_hits_grp;
	DevBuf<uint8_t> d_q;
	DevBuf<uint64_t> d_qoff;
	DevBuf<uint32_t> d_cand_t, d_cand_u, d_ncand, d_nemit, d_runs, d_uout, d_aux;
	DevBuf<uint32_t> d_cand_full, d_scratch_full; // exhaustive searches: whole candidate lists (usb_usortfull.cuh)
	DevBuf<usb_hit> d_hits;
	DevBuf<usb_qstat> d_qstat;
	DevBuf<DevCounters> d_ctr;
	DevBuf<uint8_t> d_slab, d_uarena;
	// pinned host staging for the raw (unordered) hit records of a batch
	void *h_stage = nullptr;
	size_t h_stage_cap = 0;
	bool big = false;       // UDBSearchBig path (sticky, udbusortedsearcher.cpp:39-58)
	bool kcap_retry = false; // the batch is being repeated with whole candidate lists (ERR_KCAP)
	// -usearch_local
	DevBuf<LocalDevTables> d_ltab;
	DevBuf<float> d_min_ungapped;
	DevBuf<int> d_min_gapped;
	DevBuf<AlignTables> d_atab;          // amino acid usearch_global: letter tables of k_align
	// staged candidate loop (usb_stage.cuh)
	DevBuf<uint32_t> d_job_state, d_verdict, d_items, d_hsp_arena;
	DevBuf<PassRec> d_recs;
	DevBuf<usb_hit> d_hits_stage;
	DevBuf<StageCounters> d_sc;
	DevBuf<uint8_t> d_gslab;
	DevBuf<float2> d_job_ids;
	// cluster rounds: the block's own sequences laid out as targets (usb_cluster.inc)
	DevBuf<uint8_t> d_tseq, d_twild, d_pal;
	DevBuf<uint64_t> d_toff;
	DevBuf<uint32_t> d_tlen, d_tdb2, d_tdbn, d_ppq, d_ppt;
	DevBuf<uint32_t> d_rows_out, d_nrows_out;          // sampled words per job (k_rank_big)
	DevBuf<uint32_t> d_tcnt, d_trow_size, d_tpost, d_lim, d_theta, d_conf; // conflict kernel
	DevBuf<uint64_t> d_trow_off;
	DevBuf<uint32_t> d_q_label, d_q_size; // usb_batch_set_query_attrs
	uint32_t n_q_label = 0, n_q_size = 0;
	cudaEvent_t ev_st[4 * STAGE_MAX + 1];
	int n_ev_st = 0;
	float ms_gate = 0, ms_dp = 0, ms_misc = 0, ms_rank = 0;
	uint32_t last_recs = 0, last_hsp_words = 0;
	uint64_t last_dp_cells = 0, last_dp_seq_bytes = 0;
	std::vector<float> es_min_ungapped; // per query length, NAN = not computed yet
	std::vector<int> es_min_gapped;     // per query length, INT_MIN = not computed yet
};

// exhaustive searches: (query-strand, target) pairs per batch -- 16 bytes of lists each
static const uint64_t EXHAUSTIVE_MAX_CELLS = 1ull << 29;

static bool is_int2(float x) { return std::floor(2.0 * (double)x) == 2.0 * (double)x; }

// Validates the option snapshot and derives the device constants (alnheuristics.cpp:26-69).
static int make_dev_params(const usb_params *p, DevParams &D)
{
	if (!p || p->struct_size != sizeof(usb_params))
		return fail(USB_EINVAL, "usb_params.struct_size mismatch (header/library version skew)");
	if (!p->is_nucleo && !p->local && (p->cluster_mode || p->hspw != 3))
		return fail(USB_EINVAL, "amino acid usearch_global: cluster_mode is not supported and hspw must be 3 (got %u)", p->hspw);
	if (p->is_nucleo) {
		if (p->word_length < 2 || p->word_length > 8)
			return fail(USB_EINVAL, "word_length %u unsupported (2..8)", p->word_length);
	} else if (p->word_length < 2 || p->word_length > 5)
		return fail(USB_EINVAL, "amino word_length %u unsupported (2..5)", p->word_length);
	if (p->local) {
		if (p->hspw < 2 || (p->is_nucleo ? p->hspw > 8 : p->hspw > 3))
			return fail(USB_EINVAL, "local seed word length (hspw) %u unsupported (nt 2..8, aa 2..3)", p->hspw);
		if (p->cluster_mode || p->fulldp)
			return fail(USB_EINVAL, "local = 1 cannot be combined with cluster_mode / fulldp");
		const float iv[] = {p->lopen, p->lext, p->match, p->mismatch};
		for (float v : iv)
			if (std::floor((double)v) != (double)v || std::fabs(v) > 100)
				return fail(USB_EINVAL, "local alignment needs integer scores (got %g)", (double)v);
		if (!(p->lopen < 0.0f) || !(p->lext < 0.0f))
			return fail(USB_EINVAL, "local gap penalties must be negative (lopen %g, lext %g)", (double)p->lopen, (double)p->lext);
		if (!(p->evalue > 0.0f) || !(p->ka_dbsize > 0.0f) || !(p->xdrop_u >= 0.0f) || !(p->xdrop_g >= 0.0f))
			return fail(USB_EINVAL, "local: evalue, ka_dbsize must be > 0 and xdrop_u, xdrop_g >= 0");
	} else if (p->is_nucleo && (p->hspw < 3 || p->hspw > 6))
		return fail(USB_EINVAL, "hspw %u unsupported (3..6)", p->hspw);
	const float sc[] = {p->match, p->mismatch, p->gap_open, p->gap_ext, p->term_gap_open, p->term_gap_ext};
	for (float v : sc)
		if (!is_int2(v) || std::fabs(v) > 1000)
			return fail(USB_EINVAL, "score %g is not a multiple of 0.5: the integer DP cannot reproduce it", (double)v);
	if (!(p->id >= 0.0f && p->id <= 1.0f))
		return fail(USB_EINVAL, "-id %g out of range", (double)p->id);
	D.match2 = (int)std::lround(2.0 * p->match);
	D.mismatch2 = (int)std::lround(2.0 * p->mismatch);
	D.open2 = (int)std::lround(2.0 * p->gap_open);
	D.ext2 = (int)std::lround(2.0 * p->gap_ext);
	D.topen2 = (int)std::lround(2.0 * p->term_gap_open);
	D.text2 = (int)std::lround(2.0 * p->term_gap_ext);
	D.xdrop2 = 2.0f * p->xdrop_nw;
	D.min_hsp_len = p->minhsp;
	float minscore;
	if (p->is_nucleo) {
		// MinGlobalHSPScore = FractId * Length * match, evaluated in float like the reference
		D.min_hsp_fract_id = p->id > 0.75f ? p->id : 0.75f;
		minscore = D.min_hsp_fract_id * (float)p->minhsp * p->match;
	} else {
		// alnheuristics.cpp:40-58: amino acids gate at max(id, 0.5) and scale the score threshold by
		// the smallest diagonal entry of the matrix over the 20 letters (BLOSUM62: 4)
		float min_diag = 9e9f;
		for (int i = 0; i < 24; ++i)
			if (kBlosumOrder[i] != 'B' && kBlosumOrder[i] != 'Z' && kBlosumOrder[i] != 'X' && kBlosumOrder[i] != '*')
				min_diag = std::min(min_diag, (float)kBlosum62[i][i]);
		D.min_hsp_fract_id = p->id > 0.5f ? p->id : 0.5f;
		minscore = D.min_hsp_fract_id * min_diag * (float)p->minhsp;
	}
	D.minscore2 = 2.0f * minscore;
	D.band = p->band;
	D.hspw = p->hspw;
	D.hsp_words = p->local ? 0u : p->is_nucleo ? 1u << (2 * p->hspw) : 8000u;
	D.hsp_hi = D.hsp_words / (p->is_nucleo ? 4 : 20);
	D.word_length = p->word_length;
	D.alpha = p->is_nucleo ? 4 : 20;
	D.slots = udb_slots(D.alpha, p->word_length);
	D.hash_cap = 0;
	D.maxaccepts = p->maxaccepts;
	D.maxrejects = p->maxrejects;
	D.bump = p->bump;
	D.fulldp = p->fulldp != 0;
	D.id_d = (double)p->id;
	// Accepter / Terminator options (accepter.cpp:41-94,145-197; terminator.cpp:66-86)
	D.accept_flags = p->accept_flags;
	if ((p->accept_flags & USB_ACC_MAXID) && !(p->maxid < 1.0f))
		D.accept_flags &= ~USB_ACC_MAXID; // identities never exceed 1: the default -maxid 1.0 rejects nothing
	const uint32_t extra = D.accept_flags;
	if (extra && p->cluster_mode)
		return fail(USB_EINVAL, "Accepter / Terminator options beyond -id (accept_flags 0x%x) are not supported in cluster mode",
		  extra);
	if ((extra & (USB_ACC_TERMID | USB_ACC_TERMIDD)) && p->strand_both)
		return fail(USB_EINVAL, "-termid / -termidd with -strand both: the reference carries the accepted hits of the plus "
		                        "strand into the minus-strand search; not supported");
	D.mincols = p->mincols;
	D.maxgaps = p->maxgaps;
	D.maxdiffs = p->maxdiffs;
	D.mindiffs = p->mindiffs;
	D.reject_pair_counts = 0;
	D.maxid_d = (double)p->maxid;
	D.query_cov_d = (double)p->query_cov;
	D.max_query_cov_d = (double)p->max_query_cov;
	D.target_cov_d = (double)p->target_cov;
	D.max_target_cov_d = (double)p->max_target_cov;
	D.abskew_d = (double)p->abskew;
	D.min_sizeratio_d = (double)p->min_sizeratio;
	D.minqt_d = (double)p->minqt;
	D.maxqt_d = (double)p->maxqt;
	D.minsl_d = (double)p->minsl;
	D.maxsl_d = (double)p->maxsl;
	D.termid_d = (double)p->termid;
	D.termidd_d = (double)p->termidd;
	return 0;
}

static bool g_tables_uploaded[64];

static int upload_tables(int device)
{
	if (device < 0 || device >= 64)
		return fail(USB_EINVAL, "device %d out of range", device);
	if (g_tables_uploaded[device])
		return 0;
	CharTables T;
	build_char_tables(T);
	CK(cudaMemcpyToSymbol(c_cls, T.cls, sizeof T.cls));
	CK(cudaMemcpyToSymbol(c_upper, T.upper, sizeof T.upper));
	CK(cudaMemcpyToSymbol(c_comp, T.comp, sizeof T.comp));
	CK(cudaMemcpyToSymbol(c_udb_aa, udb_letters(20), 256));
	g_tables_uploaded[device] = true;
	return 0;
}

// ------------------------------------------------------------------ index
extern "C" int usb_index_create(int device, const usb_params *p, const uint8_t *seqs, const uint64_t *seq_off,
  uint32_t n_seq, usb_index **out)
{
	if (!out || (!seqs && n_seq) || !seq_off)
		return fail(USB_EINVAL, "usb_index_create: null argument");
	DevParams D;
	int rc = make_dev_params(p, D);
	if (rc)
		return rc;
	if (usb_device_count() <= device)
		return fail(USB_ECUDA, "CUDA device %d not available (no CPU fallback exists)", device);
	CK(cudaSetDevice(device));
	rc = upload_tables(device);
	if (rc)
		return rc;
	usb_index *ix = new usb_index;
	ix->device = device;
	ix->P = *p;
	*out = ix;
	if (n_seq && (rc = usb_index_append(ix, seqs, seq_off, n_seq))) {
		usb_index_free(ix);
		*out = nullptr;
		return rc;
	}
	return 0;
}

// Appends targets [N, N+n): UDBData::AddSIToDB_CopyData (udbbuild.cpp:286) for a whole block.
// The new targets form a CSR segment; neighbouring segments of similar size are concatenated
// (log-structured merge) so that the number of row fragments per word stays logarithmic.
static int index_append_impl(usb_index *ix, const uint8_t *seqs, const uint64_t *seq_off, uint32_t n);

// Capacity hint for an index that will grow by appends (cluster_fast knows its uniques up front):
// host and device arrays for n_seqs targets with n_letters letters in all.  Appends beyond it still work.
extern "C" int usb_index_reserve(usb_index *ix, uint32_t n_seqs, uint64_t n_letters)
{
	if (!ix)
		return fail(USB_EINVAL, "usb_index_reserve: null argument");
	CK(cudaSetDevice(ix->device));
	const uint64_t bytes = n_letters + 16ull * n_seqs + 64; // every target is padded to 16 bytes
	const uint32_t n0 = ix->S.n();
	try {
		ix->S.seqs.reserve(bytes);
		ix->S.seq_off.reserve((size_t)n_seqs + 1);
		ix->S.seq_len.reserve(n_seqs);
	} catch (const std::bad_alloc &) {
		return fail(USB_ENOMEM, "usb_index_reserve: out of host memory (%llu bytes)", (unsigned long long)bytes);
	}
	const uint64_t b0 = ix->S.seq_off[n0];
	int rc;
	if ((rc = ix->d_seqs.grow_keep(bytes + 16, b0)) || (rc = ix->d_seq_off.grow_keep((size_t)n_seqs + 2, (size_t)n0 + 1)) ||
	    (rc = ix->d_seq_len.grow_keep((size_t)n_seqs + 2, n0)))
		return rc;
	if (ix->P.is_nucleo) {
		const size_t w0 = n0 ? (size_t)pack_words(b0, n0) : 0, w1 = (size_t)pack_words(bytes, n_seqs);
		if ((rc = ix->d_db2.grow_keep(w1 + 4, w0)) || (rc = ix->d_dbn.grow_keep(w1 + 4, w0)) ||
		    (rc = ix->d_wild.grow_keep((size_t)n_seqs + 2, n0)))
			return rc;
	}
	return 0;
}

// A failed append (device memory, a CUDA error) must not leave targets on the host side that have
// no postings: the host copy and the device count are rolled back to the state before the call.
extern "C" int usb_index_append(usb_index *ix, const uint8_t *seqs, const uint64_t *seq_off, uint32_t n)
{
	if (!ix)
		return fail(USB_EINVAL, "usb_index_append: null argument");
	const uint32_t n0 = ix->S.n(), dev0 = ix->n_dev, max0 = ix->S.max_len;
	const size_t nseg0 = ix->segs.size();
	const uint64_t post0 = ix->n_postings;
	const int rc = index_append_impl(ix, seqs, seq_off, n);
	if (rc && ix->S.n() != n0 && ix->segs.size() == nseg0 && !(ix->dyn && ix->dyn->base + ix->dyn->count > n0)) {
		ix->S.seq_len.resize(n0);
		ix->S.seq_off.resize((size_t)n0 + 1);
		ix->S.seqs.resize(ix->S.seq_off.back() + 16);
		ix->S.max_len = max0;
		ix->n_dev = dev0;
		ix->n_postings = post0;
	}
	return rc;
}

static int index_append_impl(usb_index *ix, const uint8_t *seqs, const uint64_t *seq_off, uint32_t n)
{
	if (!ix || !seq_off || (!seqs && n))
		return fail(USB_EINVAL, "usb_index_append: null argument");
	if (n == 0)
		return 0;
	CK(cudaSetDevice(ix->device));
	const uint32_t n0 = ix->S.n();
	if ((uint64_t)n0 + n > 0xfffffff0ull)
		return fail(USB_ELIMIT, "too many targets");
	for (uint32_t i = 0; i < n; ++i)
		if (seq_off[i + 1] < seq_off[i] || seq_off[i + 1] - seq_off[i] > (1u << 24))
			return fail(USB_EINVAL, "usb_index_append: bad offsets at target %u", i);
	HostSeqs &S = ix->S;
	g_at.start();
	++g_at.calls;
	S.append(seqs, seq_off, n, ix->P.dbmask && !ix->P.cluster_mode, 0);
	g_at.lap(0);
	int rc;
	// letters and lengths of the new targets
	const uint64_t b0 = S.seq_off[n0], b1 = S.seq_off[n0 + n];
	if ((rc = ix->d_seqs.grow_keep(b1 + 16, b0)) || (rc = ix->d_seq_off.grow_keep((size_t)n0 + n + 1, (size_t)n0 + 1)) ||
	    (rc = ix->d_seq_len.grow_keep((size_t)n0 + n + 1, n0)))
		return rc;
	CK(cudaMemcpy(ix->d_seqs.p + b0, S.seqs.data() + b0, b1 - b0 + 16, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(ix->d_seq_off.p + n0, S.seq_off.data() + n0, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(ix->d_seq_len.p + n0, S.seq_len.data() + n0, (size_t)n * 4, cudaMemcpyHostToDevice));
	ix->n_dev = n0 + n;
	g_at.lap(1);
	if (ix->P.is_nucleo) {
		const size_t w0 = n0 ? (size_t)pack_words(b0, n0) : 0, w1 = (size_t)pack_words(b1, (uint64_t)n0 + n);
		if ((rc = ix->d_db2.grow_keep(w1 + 4, w0)) || (rc = ix->d_dbn.grow_keep(w1 + 4, w0)) ||
		    (rc = ix->d_wild.grow_keep((size_t)n0 + n + 1, n0)))
			return rc;
		k_pack_targets<<<std::min<uint32_t>((n + 7) / 8, 4096), 256>>>(ix->d_seqs.p, ix->d_seq_off.p, ix->d_seq_len.p, n0, n,
		  ix->d_db2.p, ix->d_dbn.p, ix->d_wild.p);
		CK(cudaGetLastError());
		CK(cudaDeviceSynchronize());
	}
	g_at.lap(2);
	// small appends (and everything after the first one) go to the growable tail segment; the
	// initial targets of a search database always form a static segment (2-byte layout)
	// (cluster_fast: every append goes to the tail segment, large ones through the device builder)
	if (ix->dyn || ix->P.cluster_mode || (n < 8192 && n0 > 0)) {
		if (!ix->dyn) {
			ix->dyn = new DynSegment;
			if ((rc = ix->dyn->init(n0, ix->P.word_length, ix->P.is_nucleo ? 4 : 20)))
				return rc;
		}
		const uint64_t before = ix->dyn->n_postings;
		if (ix->P.is_nucleo && n >= 256 && !getenv("USB_HOST_INDEX")) {
			int num_sms = 0;
			CK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, ix->device));
			rc = ix->dyn->append_device(ix->d_seqs.p, ix->d_seq_off.p, ix->d_seq_len.p, n0, n, ix->P.word_length, num_sms);
		} else
			rc = ix->dyn->append(S, n0, n, ix->P.word_length);
		if (rc)
			return rc;
		ix->n_postings += ix->dyn->n_postings - before;
		g_at.start();
		rc = index_sync_layout(ix); // a tail segment ends the 2-byte layout of a static one
		g_at.lap(10);
		return rc;
	}
	// new segment, then merge while the last two are of similar size (or the list is full)
	IndexSegment *g = new IndexSegment;
	// nucleotide indexes are built on the device from the letters uploaded above (USB_HOST_INDEX=1:
	// host builder, the one the tests compare with)
	if (ix->P.is_nucleo && n >= 2048 && !getenv("USB_HOST_INDEX")) {
		if ((rc = build_csr_device(ix, n0, n, g->H))) {
			delete g;
			return rc;
		}
	} else
		build_csr(S, n0, n, ix->P.word_length, ix->P.is_nucleo ? 4 : 20, 0, g->H);
	ix->n_postings += g->H.n_postings;
	ix->segs.push_back(g);
	bool dirty = true;
	while (ix->segs.size() >= 2) {
		IndexSegment *a = ix->segs[ix->segs.size() - 2], *b = ix->segs.back();
		if (!(b->H.count * 2 >= a->H.count || ix->segs.size() > USB_MAX_SEG - 2))
			break;
		IndexSegment *m = new IndexSegment;
		merge_csr(a->H, b->H, m->H);
		a->release();
		b->release();
		delete a;
		delete b;
		ix->segs.pop_back();
		ix->segs.back() = m;
	}
	(void)dirty;
	return index_sync_layout(ix);
}

extern "C" void usb_index_free(usb_index *ix)
{
	if (!ix)
		return;
	cudaSetDevice(ix->device);
	ix->d_seqs.release();
	ix->d_seq_off.release();
	ix->d_seq_len.release();
	ix->d_db2.release();
	ix->d_dbn.release();
	ix->d_wild.release();
	ix->d_t_label.release();
	ix->d_t_size.release();
	for (IndexSegment *g : ix->segs) {
		g->release();
		delete g;
	}
	if (ix->dyn) {
		ix->dyn->release();
		delete ix->dyn;
	}
	delete ix;
}

extern "C" uint32_t usb_index_seq_count(const usb_index *ix) { return ix ? ix->S.n() : 0; }
extern "C" uint64_t usb_index_posting_count(const usb_index *ix) { return ix ? ix->n_postings : 0; }
extern "C" uint32_t usb_index_posting_width(const usb_index *ix)
{
	return ix && ix->segs.size() == 1 && !ix->dyn && ix->segs[0]->half ? 2u : 4u;
}

extern "C" int usb_index_row(const usb_index *ix, uint32_t word, const uint32_t **row, uint32_t *size)
{
	if (!ix || word >= udb_slots(ix->P.is_nucleo ? 4 : 20, ix->P.word_length))
		return fail(USB_EINVAL, "usb_index_row: bad word %u", word);
	std::vector<uint32_t> &tmp = const_cast<usb_index *>(ix)->row_tmp;
	tmp.clear();
	for (const IndexSegment *g : ix->segs)
		tmp.insert(tmp.end(), g->H.postings.begin() + g->H.row_off[word],
		  g->H.postings.begin() + g->H.row_off[word] + g->H.row_size[word]);
	if (ix->dyn && ix->dyn->row_size[word]) {
		const size_t n0 = tmp.size(), n = ix->dyn->row_size[word];
		tmp.resize(n0 + n);
		cudaSetDevice(ix->device);
		CK(cudaMemcpy(tmp.data() + n0, ix->dyn->d_pool.p + ix->dyn->row_off[word], n * 4, cudaMemcpyDeviceToHost));
	}
	*row = tmp.data();
	*size = (uint32_t)tmp.size();
	return 0;
}

extern "C" int usb_index_seq(const usb_index *ix, uint32_t target, const uint8_t **seq, uint32_t *len)
{
	if (!ix || target >= ix->S.n())
		return fail(USB_EINVAL, "usb_index_seq: bad target %u", target);
	*seq = ix->S.seqs.data() + ix->S.seq_off[target];
	*len = ix->S.seq_len[target];
	return 0;
}

// ------------------------------------------------------------------ EStats (estats.cpp:25-101)
// Karlin-Altschul statistics in double, exactly the reference's expressions (same libm): the DB
// size is -ka_dbsize (its default counts as set, makedbsearcher.cpp:92-96) and reaches the
// constructor as a float, like the maximum E-value.
struct EStats {
	double gl, ul, gk, uk, loggk, loguk, dbsize, maxe, log2;
	void init(const usb_params &P)
	{
		if (P.is_nucleo) {
			gl = 1.280; ul = 1.330; gk = 0.460; uk = 0.621;
		} else {
			gl = 0.267; ul = 0.311; gk = 0.0410; uk = 0.128;
		}
		loggk = log(gk);
		loguk = log(uk);
		dbsize = (double)(float)P.ka_dbsize;
		maxe = (double)(float)P.evalue;
		log2 = log(2.0);
	}
	// The expressions are written the way the reference binary evaluates them under its own build
	// flags (-O3 -ffast-math, src/Makefile:11-14: x/Log2 -> x*(1/Log2), NM/pow(2,B) -> NM*exp2(-B),
	// one fused multiply-add), so that printed E-values and bit scores agree to the last digit.
	double min_ungapped_raw(unsigned QL) const // estats.cpp:65-71
	{
		return ((log((double)QL * dbsize) + loguk) - log(maxe)) / ul;
	}
	double raw_to_bits(double raw) const { return std::fma(raw, gl, -loggk) * (1.0 / log2); } // estats.cpp:79-85, gapped
	double raw_to_evalue(double raw, unsigned QL) const                                       // estats.cpp:73-96
	{
		const double x = (loggk - raw * gl) * (1.0 / log2);
		return (double)QL * (exp2(x) * dbsize);
	}
	// smallest integer raw score >= 1 that passes `Evalue > -evalue -> reject` (localaligner.cpp:198-203)
	int min_gapped_raw(unsigned QL, float evalue_opt) const
	{
		const double E = (double)evalue_opt;
		auto pass = [&](int sc) { return !(raw_to_evalue((double)sc, QL) > E); };
		double guess = ((log((double)QL * dbsize) - log(E)) + loggk) / gl;
		int sc = guess > 1e9 ? 1000000000 : guess < 1 ? 1 : (int)guess;
		while (sc > 1 && pass(sc - 1))
			--sc;
		while (sc < 2000000000 && !pass(sc))
			++sc;
		return sc;
	}
};

// ------------------------------------------------------------------ searcher
extern "C" int usb_searcher_create(usb_index *ix, const usb_params *p, usb_searcher **out)
{
	if (!ix || !out)
		return fail(USB_EINVAL, "usb_searcher_create: null argument");
	DevParams D;
	int rc = make_dev_params(p, D);
	if (rc)
		return rc;
	if (p->word_length != ix->P.word_length)
		return fail(USB_EINVAL, "searcher word_length %u != index word_length %u", p->word_length, ix->P.word_length);
	CK(cudaSetDevice(ix->device));
	usb_searcher *s = new usb_searcher;
	s->ix = ix;
	s->big = ix->S.n() > p->big;
	s->P = *p;
	s->D = D;
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, ix->device));
	s->num_sms = prop.multiProcessorCount;
	s->smem_optin = prop.sharedMemPerBlockOptin;
	s->smem_per_sm = prop.sharedMemPerMultiprocessor;
	{
		// share of L2 that may hold persisting lines (USB_RANK_L2_MB overrides; 0 disables)
		size_t want = (size_t)prop.persistingL2CacheMaxSize;
		if (const char *e = getenv("USB_RANK_L2_MB"))
			want = std::min<size_t>(want, (size_t)atol(e) << 20);
		want = std::min<size_t>(want, (size_t)prop.accessPolicyMaxWindowSize);
		if (want && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess)
			s->l2_persist_bytes = want;
		else
			cudaGetLastError();
	}
	CK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
	for (auto &e : s->ev)
		CK(cudaEventCreate(&e));
	for (auto &e : s->ev_st) {
		e = nullptr;
		CK(cudaEventCreate(&e));
	}
	if ((rc = s->d_ctr.reserve(1))) {
		usb_searcher_free(s);
		return rc;
	}
	if (p->local) {
		LocalTables T;
		build_local_tables(p->is_nucleo != 0, (int)p->match, (int)p->mismatch, T);
		LocalDevTables *H = new LocalDevTables;
		for (int a = 0; a < USB_NCODE; ++a) {
			for (int b = 0; b < USB_NCODE; ++b)
				H->score[a * USB_NCODE + b] = T.score[a][b];
			H->match[a] = T.match[a];
			H->word_letter[a] = T.word_letter[a];
		}
		memcpy(H->code, T.code, 256);
		rc = s->d_ltab.reserve(1);
		cudaError_t e = rc ? cudaSuccess : cudaMemcpy(s->d_ltab.p, H, sizeof *H, cudaMemcpyHostToDevice);
		delete H;
		if (rc || e != cudaSuccess) {
			usb_searcher_free(s);
			return rc ? rc : fail(USB_ECUDA, "upload of the local tables failed: %s", cudaGetErrorString(e));
		}
	}
	if (!p->is_nucleo && !p->local) {
		LocalTables T;
		build_local_tables(false, 0, 0, T);
		AlignTables *H = new AlignTables;
		for (int a = 0; a < USB_NCODE; ++a) {
			for (int b = 0; b < USB_NCODE; ++b)
				H->score[a * USB_NCODE + b] = T.score[a][b];
			H->match[a] = T.match[a];
			H->word_letter[a] = T.word_letter[a];
		}
		memcpy(H->code, T.code, 256);
		rc = s->d_atab.reserve(1);
		cudaError_t e = rc ? cudaSuccess : cudaMemcpy(s->d_atab.p, H, sizeof *H, cudaMemcpyHostToDevice);
		delete H;
		if (rc || e != cudaSuccess) {
			usb_searcher_free(s);
			return rc ? rc : fail(USB_ECUDA, "upload of the amino acid tables failed: %s", cudaGetErrorString(e));
		}
	}
	*out = s;
	return 0;
}

extern "C" void usb_searcher_free(usb_searcher *s)
{
	if (!s)
		return;
	cudaSetDevice(s->ix->device);
	if (s->stream)
		cudaStreamSynchronize(s->stream);
	s->d_q.release(); s->d_qoff.release(); s->d_cand_t.release(); s->d_cand_u.release();
	s->d_ncand.release(); s->d_nemit.release(); s->d_runs.release(); s->d_uout.release(); s->d_aux.release();
	s->d_cand_full.release(); s->d_scratch_full.release();
	s->d_grp_cnt.release(); s->d_grp_off.release(); s->d_hits_grp.release();
	s->d_hits.release(); s->d_qstat.release(); s->d_ctr.release(); s->d_slab.release(); s->d_uarena.release();
	s->d_ltab.release(); s->d_min_ungapped.release(); s->d_min_gapped.release(); s->d_atab.release();
	if (s->h_stage)
		cudaFreeHost(s->h_stage);
	for (auto &e : s->ev)
		if (e)
			cudaEventDestroy(e);
	for (auto &e : s->ev_st)
		if (e)
			cudaEventDestroy(e);
	s->d_job_state.release(); s->d_verdict.release(); s->d_items.release(); s->d_hsp_arena.release();
	s->d_recs.release(); s->d_hits_stage.release(); s->d_sc.release(); s->d_gslab.release();
	s->d_job_ids.release(); s->d_q_label.release(); s->d_q_size.release();
	s->d_tseq.release(); s->d_twild.release(); s->d_pal.release(); s->d_toff.release(); s->d_tlen.release();
	s->d_tdb2.release(); s->d_tdbn.release(); s->d_ppq.release(); s->d_ppt.release();
	s->d_rows_out.release(); s->d_nrows_out.release(); s->d_tcnt.release(); s->d_trow_size.release(); s->d_tpost.release();
	s->d_lim.release(); s->d_theta.release(); s->d_conf.release(); s->d_trow_off.release();
	if (s->stream)
		cudaStreamDestroy(s->stream);
	delete s;
}

extern "C" uint64_t usb_searcher_launch_count(const usb_searcher *s) { return s ? s->launches : 0; }

extern "C" int usb_batch_counters(const usb_searcher *s, uint64_t out[4])
{
	if (!s || !out || !s->ran)
		return fail(USB_EINVAL, "usb_batch_counters: no completed batch");
	out[0] = s->last_postings;
	out[1] = s->last_hits;
	out[2] = s->last_runs;
	out[3] = s->n_jobs;
	return 0;
}

extern "C" int usb_index_set_attrs(usb_index *ix, uint32_t first, uint32_t n, const uint32_t *label_id, const uint32_t *size)
{
	if (!ix || (uint64_t)first + n > ix->S.n())
		return fail(USB_EINVAL, "usb_index_set_attrs: targets [%u, %u) out of range", first, first + n);
	CK(cudaSetDevice(ix->device));
	int rc;
	if (label_id) {
		if ((rc = ix->d_t_label.grow_keep((size_t)first + n, ix->n_label)))
			return rc;
		CK(cudaMemcpy(ix->d_t_label.p + first, label_id, (size_t)n * 4, cudaMemcpyHostToDevice));
		ix->n_label = std::max(ix->n_label, first + n);
	}
	if (size) {
		for (uint32_t i = 0; i < n; ++i)
			if (size[i] == 0)
				return fail(USB_EINVAL, "usb_index_set_attrs: size of target %u is 0", first + i);
		if ((rc = ix->d_t_size.grow_keep((size_t)first + n, ix->n_size)))
			return rc;
		CK(cudaMemcpy(ix->d_t_size.p + first, size, (size_t)n * 4, cudaMemcpyHostToDevice));
		ix->n_size = std::max(ix->n_size, first + n);
	}
	return 0;
}

extern "C" int usb_batch_set_query_attrs(usb_searcher *s, uint32_t n_q, const uint32_t *label_id, const uint32_t *size)
{
	if (!s)
		return fail(USB_EINVAL, "usb_batch_set_query_attrs: null searcher");
	CK(cudaSetDevice(s->ix->device));
	int rc;
	s->n_q_label = s->n_q_size = 0;
	if (label_id && n_q) {
		if ((rc = s->d_q_label.reserve(n_q)))
			return rc;
		CK(cudaMemcpy(s->d_q_label.p, label_id, (size_t)n_q * 4, cudaMemcpyHostToDevice));
		s->n_q_label = n_q;
	}
	if (size && n_q) {
		for (uint32_t i = 0; i < n_q; ++i)
			if (size[i] == 0)
				return fail(USB_EINVAL, "usb_batch_set_query_attrs: size of query %u is 0", i);
		if ((rc = s->d_q_size.reserve(n_q)))
			return rc;
		CK(cudaMemcpy(s->d_q_size.p, size, (size_t)n_q * 4, cudaMemcpyHostToDevice));
		s->n_q_size = n_q;
	}
	return 0;
}

extern "C" int usb_batch_kernel_ms(const usb_searcher *s, double out[8])
{
	if (!s || !out || !s->ran)
		return fail(USB_EINVAL, "usb_batch_kernel_ms: no completed batch");
	out[0] = s->ms_rank;
	out[1] = s->ms_gate;
	out[2] = s->ms_dp;
	out[3] = s->ms_misc;
	out[4] = (double)s->last_recs;
	out[5] = (double)s->last_dp_cells;
	out[6] = (double)s->last_dp_seq_bytes;
	out[7] = (double)s->last_hsp_words;
	return 0;
}

// ------------------------------------------------------------------ launches
static int upload_queries(usb_searcher *s, const uint8_t *qseqs, const uint64_t *q_off, uint32_t n_q)
{
	if (!s || !q_off || (!qseqs && n_q))
		return fail(USB_EINVAL, "null query buffers");
	CK(cudaSetDevice(s->ix->device));
	uint32_t max_ql = 0;
	for (uint32_t i = 0; i < n_q; ++i) {
		if (q_off[i + 1] < q_off[i])
			return fail(USB_EINVAL, "q_off not ascending at %u", i);
		uint64_t L = q_off[i + 1] - q_off[i];
		if (L > 65000)
			return fail(USB_ELIMIT, "query %u has %llu letters; the seed table supports up to 65000", i,
			  (unsigned long long)L);
		max_ql = std::max<uint32_t>(max_ql, (uint32_t)L);
	}
	const uint64_t base = n_q ? q_off[0] : 0, total = n_q ? q_off[n_q] - base : 0;
	int rc;
	if ((rc = s->d_q.reserve(total + 16)) || (rc = s->d_qoff.reserve((size_t)n_q + 1)))
		return rc;
	if (total)
		CK(cudaMemcpyAsync(s->d_q.p, qseqs + base, total, cudaMemcpyHostToDevice, s->stream));
	if (base == 0)
		CK(cudaMemcpyAsync(s->d_qoff.p, q_off, ((size_t)n_q + 1) * 8, cudaMemcpyHostToDevice, s->stream));
	else {
		std::vector<uint64_t> rel((size_t)n_q + 1);
		for (uint32_t i = 0; i <= n_q; ++i)
			rel[i] = q_off[i] - base;
		CK(cudaMemcpyAsync(s->d_qoff.p, rel.data(), rel.size() * 8, cudaMemcpyHostToDevice, s->stream));
		CK(cudaStreamSynchronize(s->stream));
	}
	if (s->P.local) {
		// per-query gates of LocalAligner::AlignPos (localaligner.cpp:167-171,198-203), cached by length
		EStats es;
		es.init(s->P);
		if (s->es_min_gapped.size() <= max_ql) {
			s->es_min_gapped.resize((size_t)max_ql + 1, INT32_MIN);
			s->es_min_ungapped.resize((size_t)max_ql + 1, 0.0f);
		}
		std::vector<float> mu(std::max(1u, n_q));
		std::vector<int> mg(std::max(1u, n_q));
		for (uint32_t i = 0; i < n_q; ++i) {
			const uint32_t L = (uint32_t)(q_off[i + 1] - q_off[i]);
			if (s->es_min_gapped[L] == INT32_MIN) {
				s->es_min_ungapped[L] = (float)es.min_ungapped_raw(L);
				s->es_min_gapped[L] = L ? es.min_gapped_raw(L, s->P.evalue) : 1;
			}
			mu[i] = s->es_min_ungapped[L];
			mg[i] = s->es_min_gapped[L];
		}
		if ((rc = s->d_min_ungapped.reserve(mu.size())) || (rc = s->d_min_gapped.reserve(mg.size())))
			return rc;
		CK(cudaMemcpyAsync(s->d_min_ungapped.p, mu.data(), mu.size() * 4, cudaMemcpyHostToDevice, s->stream));
		CK(cudaMemcpyAsync(s->d_min_gapped.p, mg.data(), mg.size() * 4, cudaMemcpyHostToDevice, s->stream));
		CK(cudaStreamSynchronize(s->stream));
	}
	s->n_q = n_q;
	s->max_ql = max_ql;
	s->ran = false;
	return 0;
}

// K1b launch: persistent CTAs, each with its own counter array in global memory.
static int launch_rank_big(usb_searcher *s, uint32_t n_jobs, uint32_t strands, uint32_t k_max, bool want_u)
{
	if (!s->ix->no_half) { // this kernel walks 4-byte rows
		s->ix->no_half = true;
		int rc = index_sync_layout(s->ix);
		if (rc)
			return rc;
	}
	const usb_index *ix = s->ix;
	const uint32_t N = ix->S.n();
	if (s->max_ql >= s->D.word_length && s->max_ql - s->D.word_length + 1 > BIG_MAX_POS)
		return fail(USB_ELIMIT, "big-database path supports queries up to %u letters (got %u)",
		  BIG_MAX_POS + s->D.word_length - 1, s->max_ql);
	RankBigArgs a;
	int rc0;
	memset(&a, 0, sizeof a);
	a.P = s->D;
	a.q = s->d_q.p;
	a.q_off = s->d_qoff.p;
	a.n_jobs = n_jobs;
	a.strands = strands;
	fill_index_view(ix, a.ix);
	a.n_seq = N;
	a.k_max = k_max;
	a.cand_t = s->d_cand_t.p;
	a.cand_u = s->d_cand_u.p;
	a.n_cand = s->d_ncand.p;
	a.n_emit = s->d_nemit.p;
	a.u_out = want_u ? s->d_uout.p : nullptr;
	a.aux = s->d_aux.p;
	a.stepwords = s->P.stepwords;
	{
		static const uint32_t variant = getenv("USB_BIG_VARIANT") ? (uint32_t)atoi(getenv("USB_BIG_VARIANT")) : 0u;
		a.variant = variant;
	}
	a.rows_out = nullptr;
	if (s->P.cluster_mode) { // usb_cluster_round reads the sampled words back on the device
		if ((rc0 = s->d_rows_out.reserve((size_t)n_jobs * CLUSTER_ROWS_CAP)) || (rc0 = s->d_nrows_out.reserve(n_jobs)))
			return rc0;
		a.rows_out = s->d_rows_out.p;
		a.n_rows_out = s->d_nrows_out.p;
		a.rows_cap = CLUSTER_ROWS_CAP;
	}
	a.ctr = s->d_ctr.p;
	a.u_stride = (((uint64_t)N * 2 + 64) + 255) & ~(uint64_t)255;
	const uint32_t grid = std::min<uint32_t>(n_jobs, (uint32_t)s->num_sms * 2);
	int rc = s->d_uarena.reserve((size_t)grid * a.u_stride);
	if (rc)
		return rc;
	a.u_arena = s->d_uarena.p;
	const size_t smem = sizeof(RankBigShared);
	{
		std::lock_guard<std::mutex> lk(g_attr_mu);
		CK(func_smem_locked((const void *)k_rank_big, FN_RANK_BIG, smem));
		k_rank_big<<<grid, RANK_THREADS, smem, s->stream>>>(a);
		CK(cudaGetLastError());
	}
	++s->launches;
	return 0;
}

static int launch_rank(usb_searcher *s, uint32_t n_jobs, uint32_t strands, uint32_t k_max, bool want_u)
{
	const usb_index *ix = s->ix;
	const uint32_t N = ix->S.n();
	int rc;
	if ((rc = s->d_cand_t.reserve((size_t)n_jobs * k_max)) || (rc = s->d_cand_u.reserve((size_t)n_jobs * k_max)) ||
	    (rc = s->d_ncand.reserve(n_jobs)) || (rc = s->d_nemit.reserve(n_jobs)) || (rc = s->d_aux.reserve((size_t)n_jobs * 4 + 4)))
		return rc;
	if (want_u && (rc = s->d_uout.reserve((size_t)n_jobs * N)))
		return rc;
	if (n_jobs == 0)
		return 0;
	if (!s->big && N > s->P.big)
		s->big = true; // sticky, like UDBUsortedSearcher::SetQueryImpl (udbusortedsearcher.cpp:39-58)
	if (s->big && !s->P.is_nucleo)
		return fail(USB_ELIMIT, "amino acid databases larger than -big (%u targets) are not supported", s->P.big);
	if (s->big)
		return launch_rank_big(s, n_jobs, strands, k_max, want_u);
	RankArgs a;
	memset(&a, 0, sizeof a);
	a.P = s->D;
	size_t dedupe_bytes = s->D.slots / 8;
	if (!s->P.is_nucleo) {
		// hash set of the query's words: at least twice the number of word positions
		uint32_t cap = 1024;
		while (cap < 2 * (s->max_ql + 1))
			cap <<= 1;
		a.P.hash_cap = cap;
		dedupe_bytes = (size_t)cap * 4;
	}
	a.q = s->d_q.p;
	a.q_off = s->d_qoff.p;
	a.n_jobs = n_jobs;
	a.strands = strands;
	fill_index_view(ix, a.ix);
	a.n_seq = N;
	a.k_max = k_max;
	a.cand_t = s->d_cand_t.p;
	a.cand_u = s->d_cand_u.p;
	a.n_cand = s->d_ncand.p;
	a.n_emit = s->d_nemit.p;
	a.u_out = want_u ? s->d_uout.p : nullptr;
	a.aux = s->d_aux.p;
	const bool wide = s->max_ql >= s->D.word_length && s->max_ql - s->D.word_length + 1 > 255;
	a.rec_cap = wide ? RANK_REC_WIDE : RANK_REC_NARROW;
	a.dedupe_words = (uint32_t)(dedupe_bytes / 4);
	a.bump_d = s->P.bump / 100.0;
	a.ctr = s->d_ctr.p;
	// Two CTAs of 512 threads per SM when their shared memory fits twice (each CTA also costs
	// 1 KB of system shared memory): the serial scan/sort tail of one query then overlaps the
	// posting walk of another.  Else one CTA of 1024 threads.
	uint32_t threads = RANK_THREADS_2;
	size_t smem = rank_smem_bytes(N, wide, dedupe_bytes, a.rec_cap, threads, &a.u_bytes);
	const bool two = 2 * (smem + 1024) <= s->smem_per_sm && !getenv("USB_RANK_ONE_CTA");
	if (!two) {
		threads = RANK_THREADS;
		smem = rank_smem_bytes(N, wide, dedupe_bytes, a.rec_cap, threads, &a.u_bytes);
	}
	a.seg_narrow = rank_segment(N, false, threads);
	a.seg_wide = rank_segment(N, true, threads);
	if (smem > s->smem_optin)
		return fail(USB_ELIMIT,
		  "U-sort needs %zu bytes of shared memory for %u targets (%s counters) > %zu available; tiled U-sort is not built yet",
		  smem, N, wide ? "2-byte" : "1-byte", s->smem_optin);
	a.prof = getenv("USB_RANK_PROF") ? 1 : 0; // measurement knob: phase cycles to stderr
	// Every query streams ~243 random rows of an index that is larger than L2, so a row's reuse
	// distance is the whole index and an LRU L2 keeps nothing (ncu: 6 % hits).  A persisting access
	// window over the first part of the postings turns that part into L2 hits; the rest streams.
	const bool l2win = a.ix.post16 && s->l2_persist_bytes > 0;
	if (l2win) {
		cudaStreamAttrValue av;
		memset(&av, 0, sizeof av);
		const size_t bytes = std::min<size_t>(s->l2_persist_bytes, (size_t)ix->segs[0]->d_post16.cap * 2);
		av.accessPolicyWindow.base_ptr = (void *)a.ix.post16;
		av.accessPolicyWindow.num_bytes = bytes;
		av.accessPolicyWindow.hitRatio = 1.0f;
		av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
		av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
		CK(cudaStreamSetAttribute(s->stream, cudaStreamAttributeAccessPolicyWindow, &av));
	}
	{
		// two CTAs only fit with the whole L1/shared array carved out as shared memory
		const int carve = two ? (int)cudaSharedmemCarveoutMaxShared : (int)cudaSharedmemCarveoutDefault;
		std::lock_guard<std::mutex> lk(g_attr_mu);
		if (a.ix.post16) {
			CK(func_smem_locked((const void *)k_rank<true>, FN_RANK_T, smem, carve));
			k_rank<true><<<n_jobs, threads, smem, s->stream>>>(a);
		} else {
			CK(func_smem_locked((const void *)k_rank<false>, FN_RANK_F, smem, carve));
			k_rank<false><<<n_jobs, threads, smem, s->stream>>>(a);
		}
		CK(cudaGetLastError());
	}
	if (l2win) { // the window only serves the U-sort kernel
		cudaStreamAttrValue av;
		memset(&av, 0, sizeof av);
		CK(cudaStreamSetAttribute(s->stream, cudaStreamAttributeAccessPolicyWindow, &av));
	}
	++s->launches;
	return 0;
}

struct AlignGeom {
	int wpb;
	uint32_t ql_cap, tl_cap, fast_bytes, scratch_bytes, fast_in_smem;
	size_t smem;
	uint64_t slab_stride;
	uint32_t n_warps, grid;
};

static int align_geometry(usb_searcher *s, uint32_t max_ql, uint32_t max_tl, uint32_t hsp_cap, AlignGeom &g)
{
	g.ql_cap = pad16(max_ql + 16);
	g.tl_cap = pad16(max_tl + 16);
	// fixed per-warp arrays + a scratch union (seed-table fill cursors | seed queues | DP rows).
	// The scratch gets whatever is left of a 1/16 share of shared memory, never less than the
	// seed queues need; rectangles whose DP rows do not fit use the global slab.
	const uint32_t fixed = align_fast_bytes(g.ql_cap, g.tl_cap, s->D.hsp_words);
	const size_t tab_bytes = (!s->P.is_nucleo && !s->P.local) ? ALIGN_TAB_BYTES : 0;
	const size_t budget = s->smem_optin > 1024 + tab_bytes ? s->smem_optin - 1024 - tab_bytes : 0;
	const uint32_t min_scratch = std::max<uint32_t>(SEED_SCRATCH_BYTES, pad16(s->D.hsp_words));
	uint32_t share = (uint32_t)((budget / ALIGN_MAX_WARPS) & ~(size_t)15);
	g.scratch_bytes = share > fixed + min_scratch ? share - fixed : min_scratch;
	g.scratch_bytes = std::min<uint32_t>(g.scratch_bytes, 8u * (g.tl_cap + 8));
	g.scratch_bytes = std::max<uint32_t>(g.scratch_bytes, min_scratch);
	g.fast_bytes = fixed + g.scratch_bytes;
	g.wpb = (int)std::min<size_t>(ALIGN_MAX_WARPS, budget / g.fast_bytes);
	g.fast_in_smem = g.wpb != 0;
	if (!g.fast_in_smem)
		g.wpb = 8;
	g.smem = (g.fast_in_smem ? (size_t)g.wpb * g.fast_bytes : 0) + tab_bytes;
	g.slab_stride = align_slab_bytes(g.ql_cap, g.tl_cap, hsp_cap) + (g.fast_in_smem ? 0 : g.fast_bytes);
	g.slab_stride = (g.slab_stride + 255) & ~(uint64_t)255;
	g.grid = (uint32_t)s->num_sms;
	g.n_warps = g.grid * g.wpb;
	// keep the workspace under ~1/3 of device memory
	static thread_local size_t total_b = 0; // (same device class for every searcher of a process)
	if (total_b == 0) {
		size_t free_b = 0;
		CK(cudaMemGetInfo(&free_b, &total_b));
	}
	const uint64_t limit = std::max<uint64_t>(total_b / 3, (uint64_t)256 << 20);
	while ((uint64_t)g.n_warps * g.slab_stride > limit && g.grid > 1) {
		g.grid = std::max(1u, g.grid / 2);
		g.n_warps = g.grid * g.wpb;
	}
	if ((uint64_t)g.n_warps * g.slab_stride > limit)
		return fail(USB_ELIMIT, "alignment workspace of %llu bytes per warp (query %u x target %u letters) does not fit",
		  (unsigned long long)g.slab_stride, max_ql, max_tl);
	return 0;
}

static cudaError_t launch_align(const AlignArgs &a, const AlignGeom &g, cudaStream_t st)
{
	std::lock_guard<std::mutex> lk(g_attr_mu);
	cudaError_t e;
	if (a.tab) {
		if ((e = func_smem_locked((const void *)k_align<true>, FN_ALIGN_AA, g.smem)) != cudaSuccess)
			return e;
		k_align<true><<<g.grid, g.wpb * 32, g.smem, st>>>(a);
	} else {
		if ((e = func_smem_locked((const void *)k_align<false>, FN_ALIGN_NT, g.smem)) != cudaSuccess)
			return e;
		k_align<false><<<g.grid, g.wpb * 32, g.smem, st>>>(a);
	}
	return cudaGetLastError();
}

static cudaError_t launch_viterbi(const ViterbiArgs &v, const AlignGeom &g, cudaStream_t st)
{
	std::lock_guard<std::mutex> lk(g_attr_mu);
	cudaError_t e = func_smem_locked((const void *)k_viterbi, FN_VITERBI, g.smem);
	if (e != cudaSuccess)
		return e;
	k_viterbi<<<g.grid, g.wpb * 32, g.smem, st>>>(v);
	return cudaGetLastError();
}

static void fill_align_args(usb_searcher *s, const AlignGeom &g, uint32_t hsp_cap, AlignArgs &a)
{
	memset(&a, 0, sizeof a);
	const usb_index *ix = s->ix;
	a.P = s->D;
	a.q = s->d_q.p;
	a.q_off = s->d_qoff.p;
	a.db_seq = ix->d_seqs.p;
	a.db_off = ix->d_seq_off.p;
	a.db_len = ix->d_seq_len.p;
	a.slab = s->d_slab.p;
	a.slab_stride = g.slab_stride;
	a.ql_cap = g.ql_cap;
	a.tl_cap = g.tl_cap;
	a.hsp_cap = hsp_cap;
	a.fast_bytes = g.fast_bytes;
	a.scratch_bytes = g.scratch_bytes;
	a.fast_in_smem = g.fast_in_smem;
	a.tab = (!s->P.is_nucleo && !s->P.local) ? s->d_atab.p : nullptr;
	a.ctr = s->d_ctr.p;
}

struct LocalGeom {
	int wpb;
	bool long_mode;
	uint32_t ql_cap, tl_cap, qk_cap, row_cap, fast_bytes, tb_cap;
	size_t smem;
	uint64_t slab_stride;
	uint32_t n_warps, grid;
};

static int local_geometry(usb_searcher *s, uint32_t max_ql, uint32_t max_tl, LocalGeom &g)
{
	// Sequences above g_MaxL letters (xdpmem.h:6): extensions are split into pieces of at most g_MaxL
	// columns (xdropfwdsplit.cpp, xdropbwdsplit.cpp), so the DP rows keep their size; letters and word
	// keys of the whole sequences move from shared memory to the per-warp slab.
	g.long_mode = max_ql > LOCAL_MAXL || max_tl > LOCAL_MAXL;
	if (max_ql > LOCAL_LONG_MAX || max_tl > LOCAL_LONG_MAX)
		return fail(USB_ELIMIT, "usearch_local supports sequences up to %u letters (query %u, target %u)", LOCAL_LONG_MAX, max_ql,
		  max_tl);
	g.ql_cap = lpad16(max_ql + 16);
	g.tl_cap = lpad16(max_tl + 16);
	g.row_cap = g.long_mode ? std::min<uint32_t>(g.tl_cap, lpad16(LOCAL_MAXL + 16)) : g.tl_cap;
	g.qk_cap = 32;
	while (g.qk_cap < max_ql + 1)
		g.qk_cap <<= 1;
	g.fast_bytes = lpad16(local_fast_bytes(g.ql_cap, g.tl_cap, g.qk_cap, g.row_cap, g.long_mode));
	const size_t fixed = (sizeof(LocalShared) + 15) & ~(size_t)15;
	const size_t budget = s->smem_optin > fixed + 1024 ? s->smem_optin - fixed - 1024 : 0;
	g.wpb = (int)std::min<size_t>(LOCAL_MAX_WARPS, budget / g.fast_bytes);
	if (g.wpb < 1)
		return fail(USB_ELIMIT, "local alignment workspace of %u bytes per warp does not fit in shared memory", g.fast_bytes);
	g.smem = fixed + (size_t)g.wpb * g.fast_bytes;
	const uint64_t full = ((uint64_t)max_ql + 1) * ((uint64_t)max_tl + 2) + 64;
	g.tb_cap = (uint32_t)std::min<uint64_t>(full, (uint64_t)32 << 20);
	g.slab_stride = (local_slab_bytes(g.ql_cap, g.tl_cap, g.tb_cap, g.qk_cap, g.long_mode) + 255) & ~(uint64_t)255;
	g.grid = (uint32_t)s->num_sms;
	g.n_warps = g.grid * g.wpb;
	size_t free_b = 0, total_b = 0;
	CK(cudaMemGetInfo(&free_b, &total_b));
	const uint64_t limit = std::max<uint64_t>(total_b / 3, (uint64_t)256 << 20);
	while ((uint64_t)g.n_warps * g.slab_stride > limit && g.wpb > 1) {
		--g.wpb;
		g.n_warps = g.grid * g.wpb;
		g.smem = fixed + (size_t)g.wpb * g.fast_bytes;
	}
	while ((uint64_t)g.n_warps * g.slab_stride > limit && g.grid > 1) {
		g.grid = std::max(1u, g.grid / 2);
		g.n_warps = g.grid * g.wpb;
	}
	if ((uint64_t)g.n_warps * g.slab_stride > limit)
		return fail(USB_ELIMIT, "local alignment workspace of %llu bytes per warp does not fit",
		  (unsigned long long)g.slab_stride);
	return 0;
}

static void fill_local_args(usb_searcher *s, const LocalGeom &g, LocalArgs &a)
{
	memset(&a, 0, sizeof a);
	const usb_index *ix = s->ix;
	a.P = s->D;
	a.q = s->d_q.p;
	a.q_off = s->d_qoff.p;
	a.db_seq = ix->d_seqs.p;
	a.db_off = ix->d_seq_off.p;
	a.db_len = ix->d_seq_len.p;
	a.min_ungapped = s->d_min_ungapped.p;
	a.min_gapped = s->d_min_gapped.p;
	a.tab = s->d_ltab.p;
	a.slab = s->d_slab.p;
	a.slab_stride = g.slab_stride;
	a.ql_cap = g.ql_cap;
	a.tl_cap = g.tl_cap;
	a.qk_cap = g.qk_cap;
	a.fast_bytes = g.fast_bytes;
	a.long_mode = g.long_mode ? 1u : 0u;
	a.row_cap = g.row_cap;
	a.tb_cap = g.tb_cap;
	a.xdrop_u = s->P.xdrop_u;
	a.xdrop_g = s->P.xdrop_g;
	a.open = (int)s->P.lopen;
	a.ext = (int)s->P.lext;
	a.abs_open_f = -s->P.lopen;
	a.abs_ext_f = -s->P.lext;
	a.w = s->P.hspw;
	a.alpha = s->P.is_nucleo ? 4 : 20;
	a.alpha_hi = 1;
	for (uint32_t i = 1; i < a.w; ++i)
		a.alpha_hi *= a.alpha;
	a.ctr = s->d_ctr.p;
}

static int launch_local(usb_searcher *s, const LocalArgs &a, const LocalGeom &g)
{
	{
		std::lock_guard<std::mutex> lk(g_attr_mu);
		CK(func_smem_locked((const void *)k_local, FN_LOCAL, g.smem));
		k_local<<<g.grid, g.wpb * 32, g.smem, s->stream>>>(a);
		CK(cudaGetLastError());
	}
	++s->launches;
	return 0;
}

// ------------------------------------------------------------------ staged candidate loop
struct StageGeom {
	bool ok;
	// gate
	uint32_t g_qw, g_tw, g_start_bytes, g_q1_bytes, g_bytes, g_wpb;
	size_t g_smem;
	uint64_t gslab_stride;
	// dp
	uint32_t ql_cap, tl_cap, d_fast_bytes, d_scratch_bytes, d_wpb;
	size_t d_smem;
	uint64_t dslab_stride;
	uint32_t grid;
};

static uint32_t up4(uint32_t x) { return (x + 3u) & ~3u; }

// The staged pipeline serves nucleotide global searches whose scores allow the packed X-drop walk
// (match > 0 > mismatch) and whose per-warp arrays fit in shared memory; everything else (amino
// acids, unusual scores, very long sequences) takes the one-kernel candidate loop k_align.
static bool stage_geometry(usb_searcher *s, uint32_t max_ql, uint32_t max_tl, uint32_t hsp_cap, StageGeom &g)
{
	memset(&g, 0, sizeof g);
	if (!s->P.is_nucleo || s->P.local || !(s->D.match2 > 0 && s->D.mismatch2 < 0) || getenv("USB_ONE_KERNEL_ALIGN"))
		return false;
	if (max_ql > 65000 || s->smem_optin < 65536)
		return false;
	g.ql_cap = pad16(max_ql + 16);
	g.tl_cap = pad16(max_tl + 16);
	const size_t budget = s->smem_optin - 1024;
	g.g_qw = up4(g.ql_cap / 16 + 2);
	g.g_tw = up4(g.tl_cap / 16 + 2);
	g.g_start_bytes = pad16(2 * (s->D.hsp_words + 2));
	g.g_q1_bytes = std::max<uint32_t>(pad16(4 * GATE_Q1), pad16(s->D.hsp_words));
	g.g_bytes = 8 * g.g_qw + 12 * g.g_tw + 16 + g.g_start_bytes + pad16(2 * g.ql_cap) + g.g_q1_bytes + pad16(2 * GATE_Q1) +
	            4 * GATE_Q2 + pad16(2 * GATE_Q2);
	g.g_wpb = (uint32_t)std::min<size_t>(GATE_MAX_WARPS, budget / g.g_bytes);
	if (g.g_wpb < 4)
		return false;
	g.g_smem = (size_t)g.g_wpb * g.g_bytes;
	g.gslab_stride = ((uint64_t)hsp_cap * (sizeof(HspRec) + 16) + 255) & ~(uint64_t)255;
	// dp: letters + DP rows (as wide as fits; wider rectangles use rows in the global slab)
	const uint32_t fixed = 2 * g.ql_cap + g.tl_cap;
	const uint32_t want_rows = pad16(8u * (g.tl_cap + 8));
	g.d_wpb = 0;
	for (uint32_t wpb = DP_MAX_WARPS; wpb >= 1; --wpb) {
		const size_t share = (budget / wpb) & ~(size_t)15;
		if (share >= (size_t)fixed + 2048) {
			g.d_wpb = wpb;
			g.d_scratch_bytes = (uint32_t)std::min<size_t>(want_rows, share - fixed);
			break;
		}
	}
	if (!g.d_wpb)
		return false;
	g.d_fast_bytes = fixed + g.d_scratch_bytes;
	g.d_smem = (size_t)g.d_wpb * g.d_fast_bytes;
	g.dslab_stride = (align_slab_bytes(g.ql_cap, g.tl_cap, 0) + 255) & ~(uint64_t)255;
	g.grid = (uint32_t)s->num_sms;
	size_t free_b = 0, total_b = 0;
	if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess)
		return false;
	const uint64_t limit = std::max<uint64_t>(total_b / 3, (uint64_t)256 << 20);
	while ((uint64_t)g.grid * g.d_wpb * g.dslab_stride > limit && g.grid > 1)
		g.grid = std::max(1u, g.grid / 2);
	if ((uint64_t)g.grid * g.d_wpb * g.dslab_stride > limit)
		return false;
	g.ok = true;
	return true;
}

struct StageCaps {
	uint64_t recs, hsp_words, staged;
};

// Launches the stages of one batch on the searcher's stream (no host synchronisation inside).
// a: AlignArgs with the batch fields (n_jobs, cand_t / pairs, hits, runs, qstat ...) filled in.
// packed 2-bit letters of the targets the staged kernels read (default: the index's)
struct PackedTargets {
	const uint32_t *db2, *dbn;
	const uint8_t *wild;
};

static int run_staged(usb_searcher *s, const StageGeom &g, AlignArgs a, uint32_t hsp_cap, const StageCaps &caps,
  const PackedTargets *PT = nullptr)
{
	const usb_index *ix = s->ix;
	const bool pairs = a.pair_q != nullptr;
	const uint32_t n_jobs = a.n_jobs, k_max = pairs ? 1u : a.k_max;
	int rc;
	if ((rc = s->d_job_state.reserve(std::max(1u, n_jobs))) || (rc = s->d_verdict.reserve((size_t)std::max(1u, n_jobs) * k_max)) ||
	    (rc = s->d_items.reserve(std::max(1u, n_jobs))) || (rc = s->d_recs.reserve(caps.recs)) ||
	    (rc = s->d_hsp_arena.reserve(caps.hsp_words)) || (rc = s->d_hits_stage.reserve(caps.staged)) ||
	    (rc = s->d_sc.reserve(1)) || (rc = s->d_gslab.reserve((size_t)g.grid * g.g_wpb * g.gslab_stride)) ||
	    (rc = s->d_slab.reserve((size_t)g.grid * g.d_wpb * g.dslab_stride)))
		return rc;
	CK(cudaMemsetAsync(s->d_sc.p, 0, sizeof(StageCounters), s->stream));
	StageArgs S;
	memset(&S, 0, sizeof S);
	S.A = a;
	S.A.k_max = k_max;
	S.A.ql_cap = g.ql_cap;
	S.A.tl_cap = g.tl_cap;
	S.A.hsp_cap = hsp_cap;
	S.A.fast_in_smem = 1;
	S.A.tab = nullptr;
	S.db2 = PT ? PT->db2 : ix->d_db2.p;
	S.dbn = PT ? PT->dbn : ix->d_dbn.p;
	S.db_wild = PT ? PT->wild : ix->d_wild.p;
	S.job_state = s->d_job_state.p;
	S.job_ids = nullptr;
	if (a.P.accept_flags & (USB_ACC_TERMID | USB_ACC_TERMIDD)) {
		if ((rc = s->d_job_ids.reserve(std::max(1u, n_jobs))))
			return rc;
		S.job_ids = s->d_job_ids.p;
	}
	S.verdict = s->d_verdict.p;
	S.recs = s->d_recs.p;
	S.recs_cap = (uint32_t)std::min<uint64_t>(caps.recs, 0x3ffffff0ull);
	S.hsp_arena = s->d_hsp_arena.p;
	S.hsp_arena_cap = (uint32_t)std::min<uint64_t>(caps.hsp_words, 0xfffffff0ull);
	S.hits_stage = s->d_hits_stage.p;
	S.stage_cap = (uint32_t)std::min<uint64_t>(caps.staged, 0xfffffff0ull);
	S.sc = s->d_sc.p;
	S.g_qw = g.g_qw;
	S.g_tw = g.g_tw;
	S.g_start_bytes = g.g_start_bytes;
	S.g_q1_bytes = g.g_q1_bytes;
	S.g_bytes = g.g_bytes;
	s->n_ev_st = 0;
	auto mark = [&]() { return cudaEventRecord(s->ev_st[s->n_ev_st++], s->stream); };
	CK(mark());
	uint32_t ka = 0, stage = 0;
	while (ka < k_max && stage < STAGE_MAX) {
		const uint32_t kb = stage == 0 ? 1u : (stage == STAGE_MAX - 1 ? k_max : std::min(k_max, ka + STAGE_WIDTH));
		S.ka = ka;
		S.kb = kb;
		S.stage = stage;
		S.items = stage == 0 ? nullptr : s->d_items.p;
		const uint32_t tgrid = std::min<uint32_t>((n_jobs + 255) / 256, (uint32_t)s->num_sms * 8);
		k_stage_prep<<<std::max(1u, tgrid), 256, 0, s->stream>>>(S);
		CK(cudaGetLastError());
		{
			// gate: the HSP scratch lives in its own slab
			StageArgs G = S;
			G.A.slab = s->d_gslab.p;
			G.A.slab_stride = g.gslab_stride;
			std::lock_guard<std::mutex> lk(g_attr_mu);
			CK(func_smem_locked((const void *)k_gate, FN_GATE, g.g_smem));
			k_gate<<<g.grid, g.g_wpb * 32, g.g_smem, s->stream>>>(G);
			CK(cudaGetLastError());
		}
		CK(mark());
		{
			StageArgs D = S;
			D.A.slab = s->d_slab.p;
			D.A.slab_stride = g.dslab_stride;
			D.A.fast_bytes = g.d_fast_bytes;
			D.A.scratch_bytes = g.d_scratch_bytes;
			std::lock_guard<std::mutex> lk(g_attr_mu);
			CK(func_smem_locked((const void *)k_dp, FN_DP, g.d_smem));
			k_dp<<<g.grid, g.d_wpb * 32, g.d_smem, s->stream>>>(D);
			CK(cudaGetLastError());
		}
		CK(mark());
		k_commit<<<std::max(1u, tgrid), 256, 0, s->stream>>>(S);
		CK(cudaGetLastError());
		CK(mark());
		s->launches += 4;
		ka = kb;
		++stage;
	}
	return 0;
}

// per-kernel times of the last staged batch (after the stream was synchronised)
static void stage_times(usb_searcher *s)
{
	s->ms_gate = s->ms_dp = s->ms_misc = 0;
	for (int i = 0; i + 3 < s->n_ev_st; i += 3) {
		float a = 0, b = 0, c = 0;
		cudaEventElapsedTime(&a, s->ev_st[i], s->ev_st[i + 1]);
		cudaEventElapsedTime(&b, s->ev_st[i + 1], s->ev_st[i + 2]);
		cudaEventElapsedTime(&c, s->ev_st[i + 2], s->ev_st[i + 3]);
		s->ms_gate += a;
		s->ms_dp += b;
		s->ms_misc += c;
	}
}

static const char *err_text(uint32_t e)
{
	static thread_local char buf[400];
	if (e & ERR_KCAP) {
		snprintf(buf, sizeof buf, "a query ran through all %u materialised candidates without terminating because pairs were "
		                          "skipped (-self / -notself / -selfid / size and length ratio rules), and the whole lists do not fit: "
		                          "search smaller batches (at most 2^29 query-strand x target pairs), or a database below -big",
		         (unsigned)RANK_KCAP);
		return buf;
	}
	snprintf(buf, sizeof buf, "device error flags 0x%x:%s%s%s%s%s%s%s%s%s%s", e, e & ERR_REC_FULL ? " gate record list full" : "",
	  e & ERR_HSPARENA_FULL ? " HSP arena full" : "", e & ERR_HITS_FULL ? " hit buffer full" : "",
	  e & ERR_RUNS_FULL ? " run arena full" : "", e & ERR_HSP_FULL ? " HSP list full" : "",
	  e & ERR_TRACE ? " traceback left the band" : "", e & ERR_RECORDS_FULL ? " U-sort record list full" : "",
	  e & ERR_NO_M ? " alignment path without M" : "", e & ERR_TB_FULL ? " X-drop trace arena full" : "",
	  e & ERR_AR_FULL ? " more than 32 local alignments for one target" : "");
	return buf;
}

extern "C" int usb_batch_upload(usb_searcher *s, const uint8_t *qseqs, const uint64_t *q_off, uint32_t n_q)
{
	int rc = upload_queries(s, qseqs, q_off, n_q);
	if (rc)
		return rc;
	CK(cudaStreamSynchronize(s->stream));
	return 0;
}

extern "C" int usb_batch_run(usb_searcher *s, float *ms)
{
	if (!s)
		return fail(USB_EINVAL, "null searcher");
	CK(cudaSetDevice(s->ix->device));
	const usb_index *ix = s->ix;
	const uint32_t N = ix->S.n();
	s->strands = s->P.strand_both ? 2 : 1;
	s->n_jobs = s->n_q * s->strands;
	uint32_t k_max = N;
	if (s->P.maxaccepts > 0 && s->P.maxrejects > 0)
		k_max = std::min<uint64_t>(N, (uint64_t)s->P.maxaccepts + s->P.maxrejects - 1);
	// RejectPair-ed candidates are skipped without a Terminator call on the small-database path
	// (searcher.cpp:63-67), so the loop can reach any candidate of the list: materialise as many as
	// the build allows; a query that runs out of them unterminated is reported (ERR_KCAP)
	// (a local search never skips: its rejected pairs are rejects, searcher.cpp:26-49)
	const bool pair_skips = (s->D.accept_flags & ACC_PAIR_FLAGS) != 0 && !(s->big || N > s->P.big) && !s->P.local;
	// (after such a batch ran out of candidates it is repeated with the whole lists, see below)
	const bool pair_skips_all = pair_skips && s->kcap_retry;
	s->kcap_retry = false;
	if (pair_skips)
		k_max = pair_skips_all ? N : std::min<uint32_t>(N, RANK_KCAP);
	if (k_max == 0)
		k_max = 1;
	// -maxaccepts 0 or -maxrejects 0 on more than RANK_KCAP targets: the candidate loop may walk the whole U-sorted
	// list.  k_rank then also writes its counters to global memory and k_usort_full (usb_usortfull.cuh) makes the
	// complete list from them; the loop kernels take it with a stride of N candidates per job.
	const bool exhaustive = k_max > RANK_KCAP;
	if (exhaustive) {
		if (s->big || N > s->P.big)
			return fail(USB_ELIMIT, "exhaustive searches (-maxaccepts 0 / -maxrejects 0) are not built for databases above -big "
			                        "(%u targets): set both options > 0", s->P.big);
		if (s->P.cluster_mode)
			return fail(USB_EINVAL, "cluster rounds need -maxaccepts 1 and -maxrejects > 0");
		// counters, scratch list, candidate list and verdicts: 16 bytes per (query-strand, target)
		if ((uint64_t)s->n_jobs * N > EXHAUSTIVE_MAX_CELLS)
			return fail(USB_ELIMIT, "exhaustive search: %u query-strands x %u targets exceed %llu list entries per batch; "
			                        "search batches of at most %llu queries", s->n_jobs, N, (unsigned long long)EXHAUSTIVE_MAX_CELLS,
			            (unsigned long long)std::max<uint64_t>(1, EXHAUSTIVE_MAX_CELLS / N / s->strands));
	}
	const uint32_t k_rank_max = exhaustive ? (uint32_t)RANK_KCAP : k_max;
	s->k_max = k_max;
	const uint32_t hsp_cap = std::max<uint32_t>(64, ix->S.max_len / 8 + 16);
	const bool local = s->P.local != 0;
	AlignGeom g;
	LocalGeom lg;
	StageGeom sg;
	const bool staged = stage_geometry(s, s->max_ql, ix->S.max_len, hsp_cap, sg);
	int rc = local ? local_geometry(s, s->max_ql, ix->S.max_len, lg)
	               : staged ? 0 : align_geometry(s, s->max_ql, ix->S.max_len, hsp_cap, g);
	if (rc)
		return rc;
	StageCaps caps;
	caps.recs = s->D.fulldp ? (uint64_t)s->n_jobs * k_max + 1024 : (uint64_t)s->n_jobs * 2 + 4096;
	caps.recs = std::max<uint64_t>(caps.recs, s->d_recs.cap);
	caps.hsp_words = std::max<uint64_t>(caps.recs * 12, s->d_hsp_arena.cap);
	// (an exhaustive search starts with room for 64 hits per job; a full buffer is grown and the batch repeated)
	const uint64_t per_job_hits = s->P.maxaccepts > 0 ? s->P.maxaccepts : exhaustive ? 64 : k_max;
	// a local target can contribute several ARs (localmulti.cpp): start with room for two per
	// accepted target and grow on demand
	uint64_t hits_cap = std::max<uint64_t>(1, (uint64_t)s->n_jobs * per_job_hits * (local ? 2 : 1) + (local ? 1024 : 0));
	if (hits_cap > 0xfffffff0ull)
		return fail(USB_ELIMIT, "hit buffer of %llu records too large; use smaller batches", (unsigned long long)hits_cap);
	uint64_t runs_cap = std::max<uint64_t>(s->d_runs.cap, std::max<uint64_t>((uint64_t)1 << 20, hits_cap * 24));
	if ((rc = s->d_qstat.reserve(std::max(1u, s->n_jobs))) ||
	    (!staged && (rc = s->d_slab.reserve(local ? (size_t)lg.n_warps * lg.slab_stride : (size_t)g.n_warps * g.slab_stride))))
		return rc;
	for (int attempt = 0;; ++attempt) {
		if ((rc = s->d_runs.reserve(runs_cap)) || (rc = s->d_hits.reserve(hits_cap)))
			return rc;
		CK(cudaMemsetAsync(s->d_ctr.p, 0, sizeof(DevCounters), s->stream));
		CK(cudaEventRecord(s->ev[0], s->stream));
		if ((rc = launch_rank(s, s->n_jobs, s->strands, k_rank_max, exhaustive)))
			return rc;
		const uint32_t *cand_list = s->d_cand_t.p;
		if (exhaustive && s->n_jobs) {
			if ((rc = s->d_cand_full.reserve((size_t)s->n_jobs * N)) || (rc = s->d_scratch_full.reserve((size_t)s->n_jobs * N)))
				return rc;
			UsortFullArgs f;
			f.u = s->d_uout.p;
			f.scratch = s->d_scratch_full.p;
			f.cand_t = s->d_cand_full.p;
			f.n_emit = s->d_nemit.p;
			f.n_jobs = s->n_jobs;
			f.n_seq = N;
			f.bump_d = s->P.bump / 100.0;
			f.ctr = s->d_ctr.p;
			k_usort_full<<<(s->n_jobs + USORTFULL_WARPS - 1) / USORTFULL_WARPS, USORTFULL_WARPS * 32, 0, s->stream>>>(f);
			CK(cudaGetLastError());
			++s->launches;
			cand_list = s->d_cand_full.p;
		}
		CK(cudaEventRecord(s->ev[1], s->stream));
		if (s->n_jobs && local) {
			LocalArgs a;
			fill_local_args(s, lg, a);
			a.n_jobs = s->n_jobs;
			a.strands = s->strands;
			a.cand_t = cand_list;
			a.n_emit = s->d_nemit.p;
			a.k_max = k_max;
			a.hits = s->d_hits.p;
			a.hits_cap = (uint32_t)hits_cap;
			a.runs = s->d_runs.p;
			a.runs_cap = (uint32_t)std::min<uint64_t>(s->d_runs.cap, 0xfffffff0ull);
			a.qstat = s->d_qstat.p;
			if (s->D.accept_flags) {
				if (s->D.accept_flags & (USB_ACC_TERMID | USB_ACC_TERMIDD))
					return fail(USB_EINVAL, "-termid / -termidd are not supported with -usearch_local");
				// In a local search the reference applies RejectPair only inside IsAccept (searcher.cpp:26-49) and
				// dies with SIGSEGV -- or goes on with corrupted results -- when one of these rules rejects a
				// pair: there is nothing to be identical to
				if (s->D.accept_flags & ACC_PAIR_FLAGS)
					return fail(USB_EINVAL, "-self, -notself, -selfid, -min_sizeratio, -minqt/-maxqt and -minsl/-maxsl are not "
					                        "supported with -usearch_local (the reference crashes on them)");
				if ((s->D.accept_flags & USB_ACC_NEEDS_LABELS) && (ix->n_label < N || s->n_q_label < s->n_q))
					return fail(USB_EINVAL, "-self / -notself need label identities: call usb_index_set_attrs for every target "
					                        "and usb_batch_set_query_attrs for this batch");
				if ((s->D.accept_flags & USB_ACC_NEEDS_SIZES) && (ix->n_size < N || s->n_q_size < s->n_q))
					return fail(USB_EINVAL, "-abskew / -min_sizeratio need size= annotations: call usb_index_set_attrs for every "
					                        "target and usb_batch_set_query_attrs for this batch");
				a.q_label = s->d_q_label.p;
				a.q_size = s->d_q_size.p;
				a.t_label = ix->d_t_label.p;
				a.t_size = ix->d_t_size.p;
				a.P.reject_pair_counts = s->big ? 1u : 0u;
				a.n_cand_all = pair_skips ? s->d_ncand.p : nullptr;
			}
			if ((rc = launch_local(s, a, lg)))
				return rc;
		} else if (s->n_jobs) {
			AlignArgs a;
			if (staged) {
				memset(&a, 0, sizeof a);
				a.P = s->D;
				a.q = s->d_q.p;
				a.q_off = s->d_qoff.p;
				a.db_seq = ix->d_seqs.p;
				a.db_off = ix->d_seq_off.p;
				a.db_len = ix->d_seq_len.p;
				a.ctr = s->d_ctr.p;
			} else
				fill_align_args(s, g, hsp_cap, a);
			a.n_jobs = s->n_jobs;
			a.strands = s->strands;
			a.cand_t = cand_list;
			a.n_emit = s->d_nemit.p;
			a.k_max = k_max;
			a.hits = s->d_hits.p;
			a.hits_cap = (uint32_t)hits_cap;
			a.runs = s->d_runs.p;
			a.runs_cap = (uint32_t)std::min<uint64_t>(s->d_runs.cap, 0xfffffff0ull);
			a.qstat = s->d_qstat.p;
			if (s->D.accept_flags) {
				if (!staged && (s->D.accept_flags & (USB_ACC_TERMID | USB_ACC_TERMIDD)))
					return fail(USB_EINVAL, "-termid / -termidd need the staged candidate loop (nucleotide scores with "
					                        "match > 0 > mismatch, sequences that fit in shared memory)");
				if ((s->D.accept_flags & USB_ACC_NEEDS_LABELS) && (ix->n_label < N || s->n_q_label < s->n_q))
					return fail(USB_EINVAL, "-self / -notself need label identities: call usb_index_set_attrs for every target "
					                        "and usb_batch_set_query_attrs for this batch");
				if ((s->D.accept_flags & USB_ACC_NEEDS_SIZES) && (ix->n_size < N || s->n_q_size < s->n_q))
					return fail(USB_EINVAL, "-abskew / -min_sizeratio need size= annotations: call usb_index_set_attrs for every "
					                        "target and usb_batch_set_query_attrs for this batch");
				a.q_label = s->d_q_label.p;
				a.q_size = s->d_q_size.p;
				a.t_label = ix->d_t_label.p;
				a.t_size = ix->d_t_size.p;
				a.P.reject_pair_counts = s->big ? 1u : 0u;
				a.n_cand_all = pair_skips ? s->d_ncand.p : nullptr;
			}
			if (staged) {
				caps.staged = hits_cap + 1024;
				if ((rc = run_staged(s, sg, a, hsp_cap, caps)))
					return rc;
			} else {
				CK(launch_align(a, g, s->stream));
				++s->launches;
			}
		}
		CK(cudaEventRecord(s->ev[2], s->stream));
		DevCounters c;
		CK(cudaMemcpyAsync(&c, s->d_ctr.p, sizeof c, cudaMemcpyDeviceToHost, s->stream));
		CK(cudaStreamSynchronize(s->stream));
		if (c.err && !(c.err & ~(ERR_RUNS_FULL | ERR_HITS_FULL | ERR_REC_FULL | ERR_HSPARENA_FULL)) && attempt < 6) {
			if (c.err & ERR_RUNS_FULL)
				runs_cap = std::max<uint64_t>(runs_cap * 4, (uint64_t)c.n_runs + 1024);
			if (c.err & ERR_HITS_FULL)
				hits_cap = std::min<uint64_t>(0xfffffff0ull, std::max<uint64_t>(hits_cap * 2, (uint64_t)c.n_hits + 1024));
			if (c.err & ERR_REC_FULL)
				caps.recs *= 4;
			if (c.err & (ERR_REC_FULL | ERR_HSPARENA_FULL))
				caps.hsp_words = std::max<uint64_t>(caps.hsp_words * 4, caps.recs * 12);
			continue;
		}
		if ((c.err & ERR_KCAP) && !(c.err & ~(ERR_KCAP | ERR_RUNS_FULL | ERR_HITS_FULL | ERR_REC_FULL | ERR_HSPARENA_FULL)) && !exhaustive &&
		    N > RANK_KCAP && (uint64_t)s->n_jobs * N <= EXHAUSTIVE_MAX_CELLS) {
			// skipped pairs used up the 1 024 materialised candidates: once more with the whole lists (k_usort_full)
			s->kcap_retry = true;
			return usb_batch_run(s, ms);
		}
		if (c.err)
			return fail(USB_ELIMIT, "%s", err_text(c.err));
		s->last_hits = c.n_hits;
		s->last_runs = c.n_runs;
		s->last_postings = c.postings;
		if (getenv("USB_RANK_PROF") && s->n_jobs) {
			const char *nm[13] = {"zero", "words", "walk", "replay", "filter_tail", "emit", "f_count", "f_scan", "f_write",
			  "r_max", "r_scan", "r_collect", "sort"};
			fprintf(stderr, "k_rank phase cycles per job:");
			for (int i = 0; i < 13; ++i)
				fprintf(stderr, " %s %.0f", nm[i], (double)c.prof[i] / s->n_jobs);
			fprintf(stderr, "\n");
		}
		break;
	}
	s->last_recs = s->last_hsp_words = 0;
	s->last_dp_cells = s->last_dp_seq_bytes = 0;
	if (staged && s->n_jobs && !local) {
		stage_times(s);
		StageCounters sc;
		CK(cudaMemcpy(&sc, s->d_sc.p, sizeof sc, cudaMemcpyDeviceToHost));
		s->last_recs = sc.n_recs;
		s->last_hsp_words = sc.n_hsp_words;
		s->last_dp_cells = sc.dp_cells;
		s->last_dp_seq_bytes = sc.dp_seq_bytes;
	} else
		s->ms_gate = s->ms_dp = s->ms_misc = 0;
	CK(cudaEventElapsedTime(&s->ms_rank, s->ev[0], s->ev[1]));
	if (ms) {
		CK(cudaEventElapsedTime(&ms[0], s->ev[0], s->ev[1]));
		CK(cudaEventElapsedTime(&ms[1], s->ev[1], s->ev[2]));
		CK(cudaEventElapsedTime(&ms[2], s->ev[0], s->ev[2]));
	}
	s->ran = true;
	return 0;
}

// HitMgr::Sort (hitmgr.cpp:477) orders a query's hits with the reference's own quicksort on
// float scores (sort.h:63-102): middle pivot, Hoare partition, descending, not stable.
static void quicksort_desc(const float *v, uint32_t *ord, int lo, int hi)
{
	int i = lo, j = hi;
	const float pivot = v[ord[(lo + hi) / 2]];
	while (i <= j) {
		while (v[ord[i]] > pivot)
			++i;
		while (v[ord[j]] < pivot)
			--j;
		if (i <= j) {
			std::swap(ord[i], ord[j]);
			++i;
			--j;
		}
	}
	if (lo < j)
		quicksort_desc(v, ord, lo, j);
	if (i < hi)
		quicksort_desc(v, ord, i, hi);
}

static void order_hits_like_hitmgr(NoInitVec<usb_hit> &hits, const NoInitVec<uint64_t> &qoff, bool local)
{
	std::vector<float> sc;
	std::vector<uint32_t> ord;
	std::vector<usb_hit> tmp;
	for (size_t q = 0; q + 1 < qoff.size(); ++q) {
		const uint64_t b = qoff[q], e = qoff[q + 1];
		const uint32_t n = (uint32_t)(e - b);
		if (n < 2)
			continue;
		// Searcher::Search appends plus-strand hits, then minus-strand hits (searcher.cpp:144-158)
		// and, for local searches, the ARs of one target in AlignMulti order (searcher.cpp:36-47)
		std::sort(hits.begin() + b, hits.begin() + e, [](const usb_hit &x, const usb_hit &y) {
			return x.strand != y.strand ? x.strand < y.strand : x.rank != y.rank ? x.rank < y.rank : x.sub < y.sub;
		});
		sc.resize(n);
		ord.resize(n);
		tmp.assign(hits.begin() + b, hits.begin() + e);
		for (uint32_t i = 0; i < n; ++i) {
			const usb_hit &h = tmp[i];
			// arscorer.cpp:818-824: local = raw score, global = fractional identity
			sc[i] = local ? (float)h.raw : (float)(h.alnlen == 0 ? 0.0 : (double)h.ids / (double)h.alnlen);
			ord[i] = i;
		}
		quicksort_desc(sc.data(), ord.data(), 0, (int)n - 1);
		for (uint32_t i = 0; i < n; ++i)
			hits[b + i] = tmp[ord[i]];
	}
}

// Result objects are recycled (two at most): their vectors keep their capacity, so a steady stream
// of batches does not page-fault through ~200 MB of fresh allocations per call.
static std::mutex g_pool_mu;
static std::vector<usb_result *> g_pool;

static usb_result *result_new()
{
	std::lock_guard<std::mutex> lk(g_pool_mu);
	if (g_pool.empty())
		return new usb_result;
	usb_result *r = g_pool.back();
	g_pool.pop_back();
	return r;
}

static void result_recycle(usb_result *r)
{
	if (!r)
		return;
	{
		std::lock_guard<std::mutex> lk(g_pool_mu);
		if (g_pool.size() < 8) {
			r->hits.clear();
			r->runs.clear();
			r->qoff.clear();
			r->qstat.clear();
			g_pool.push_back(r);
			return;
		}
	}
	delete r;
}

// USB_TIMING=1: wall-clock phases of usb_search_batch, printed at exit (measurement aid)
struct BatchTimers {
	double t[6] = {0, 0, 0, 0, 0, 0};
	uint64_t calls = 0;
	bool on = getenv("USB_TIMING") != nullptr;
	~BatchTimers()
	{
		if (on && calls)
			fprintf(stderr, "usb_search_batch: %llu calls; upload %.3fs run %.3fs copy back %.3fs group %.3fs order %.3fs\n",
			  (unsigned long long)calls, t[0], t[1], t[2], t[3], t[4]);
	}
};
static BatchTimers g_bt;

static int download_result(usb_searcher *s, uint32_t n_q, bool group, usb_result **out)
{
	double tk = g_bt.on ? AppendTimers::now() : 0;
	usb_result *r = result_new();
	const uint32_t nh = s->last_hits, nr = s->last_runs;
	// The hits are grouped on the device (count per group, offsets, scatter); everything comes back
	// through the searcher's page-locked staging buffer and is copied into the result by a few threads.
	const bool want_q = s->n_jobs && group;
	const size_t b_hits = ((size_t)nh * sizeof(usb_hit) + 63) & ~(size_t)63, b_off = (((size_t)n_q + 1) * 8 + 63) & ~(size_t)63,
	             b_runs = ((size_t)nr * 4 + 63) & ~(size_t)63,
	             b_qstat = want_q ? ((size_t)s->n_jobs * sizeof(usb_qstat) + 63) & ~(size_t)63 : 0,
	             b_nc = want_q ? ((size_t)s->n_jobs * 4 + 63) & ~(size_t)63 : 0;
	const size_t stage_bytes = b_hits + b_off + b_runs + b_qstat + b_nc + 64;
	if (stage_bytes > s->h_stage_cap) {
		if (s->h_stage)
			cudaFreeHost(s->h_stage);
		s->h_stage = nullptr;
		s->h_stage_cap = 0;
		const size_t want = stage_bytes * 5 / 4;
		if (cudaHostAlloc(&s->h_stage, want, cudaHostAllocDefault) != cudaSuccess) {
			cudaGetLastError();
			result_recycle(r);
			return fail(USB_ENOMEM, "cudaHostAlloc of %zu bytes failed", want);
		}
		s->h_stage_cap = want;
	}
	uint8_t *st_hits = (uint8_t *)s->h_stage, *st_off = st_hits + b_hits, *st_runs = st_off + b_off, *st_qstat = st_runs + b_runs,
	        *st_nc = st_qstat + b_qstat;
	const uint32_t *ncand = (const uint32_t *)st_nc;
	try {
		r->runs.resize(nr);
		r->qstat.resize(s->n_jobs);
		r->hits.resize(nh);
		r->qoff.resize((size_t)n_q + 1);
	} catch (const std::bad_alloc &) {
		result_recycle(r);
		return fail(USB_ENOMEM, "out of host memory for the result (%u hits)", nh);
	}
	int rc;
	if ((rc = s->d_grp_cnt.reserve((size_t)n_q + 1)) || (rc = s->d_grp_off.reserve((size_t)n_q + 1)) ||
	    (rc = s->d_hits_grp.reserve((size_t)nh + 1))) {
		result_recycle(r);
		return rc;
	}
	cudaError_t e = cudaMemsetAsync(s->d_grp_cnt.p, 0, ((size_t)n_q + 1) * 4, s->stream);
	if (e == cudaSuccess && nh) {
		const uint32_t grid = std::min<uint32_t>((nh + 255) / 256, (uint32_t)s->num_sms * 8);
		k_hit_count<<<grid, 256, 0, s->stream>>>(s->d_hits.p, nh, group ? 1 : 0, s->d_grp_cnt.p);
		k_row_offsets<<<1, 1024, 0, s->stream>>>(s->d_grp_cnt.p, n_q, s->d_grp_off.p);
		e = cudaMemsetAsync(s->d_grp_cnt.p, 0, ((size_t)n_q + 1) * 4, s->stream);
		k_hit_scatter<<<grid, 256, 0, s->stream>>>(s->d_hits.p, nh, group ? 1 : 0, s->d_grp_off.p, s->d_grp_cnt.p, s->d_hits_grp.p);
		if (e == cudaSuccess)
			e = cudaGetLastError();
		if (e == cudaSuccess)
			e = cudaMemcpyAsync(st_hits, s->d_hits_grp.p, (size_t)nh * sizeof(usb_hit), cudaMemcpyDeviceToHost, s->stream);
		if (e == cudaSuccess)
			e = cudaMemcpyAsync(st_off, s->d_grp_off.p, ((size_t)n_q + 1) * 8, cudaMemcpyDeviceToHost, s->stream);
		s->launches += 3;
	} else
		memset(st_off, 0, ((size_t)n_q + 1) * 8);
	if (e == cudaSuccess && nr)
		e = cudaMemcpyAsync(st_runs, s->d_runs.p, (size_t)nr * 4, cudaMemcpyDeviceToHost, s->stream);
	if (e == cudaSuccess && want_q) {
		e = cudaMemcpyAsync(st_qstat, s->d_qstat.p, (size_t)s->n_jobs * sizeof(usb_qstat), cudaMemcpyDeviceToHost, s->stream);
		if (e == cudaSuccess)
			e = cudaMemcpyAsync(st_nc, s->d_ncand.p, (size_t)s->n_jobs * 4, cudaMemcpyDeviceToHost, s->stream);
	}
	if (e == cudaSuccess)
		e = cudaStreamSynchronize(s->stream);
	if (e != cudaSuccess) {
		result_recycle(r);
		return fail(USB_ECUDA, "result download failed: %s", cudaGetErrorString(e));
	}
	{
		// staging -> result: four copies, in parallel when they are large
		struct Part {
			void *dst;
			const void *src;
			size_t n;
		} parts[4] = {{r->hits.data(), st_hits, (size_t)nh * sizeof(usb_hit)}, {r->qoff.data(), st_off, ((size_t)n_q + 1) * 8},
		              {r->runs.data(), st_runs, (size_t)nr * 4}, {r->qstat.data(), st_qstat, want_q ? (size_t)s->n_jobs * sizeof(usb_qstat) : 0}};
		size_t total = 0;
		for (const Part &p : parts)
			total += p.n;
		if (total < ((size_t)8 << 20)) {
			for (const Part &p : parts)
				if (p.n)
					memcpy(p.dst, p.src, p.n);
		} else {
			// every part in four pieces, one per thread
			std::vector<std::thread> th;
			for (int k = 0; k < 4; ++k)
				th.emplace_back([&, k]() {
					for (int q = 0; q < 4; ++q) {
						const Part &p = parts[q];
						const size_t a0 = p.n * k / 4, a1 = p.n * (k + 1) / 4;
						if (a1 > a0)
							memcpy((uint8_t *)p.dst + a0, (const uint8_t *)p.src + a0, a1 - a0);
					}
				});
			for (auto &t : th)
				t.join();
		}
	}
	if (group)
		for (uint32_t j = 0; j < s->n_jobs; ++j)
			r->qstat[j].n_cand = ncand[j];
	if (g_bt.on) {
		const double x = AppendTimers::now();
		g_bt.t[2] += x - tk;
		tk = x;
	}
	if (g_bt.on) {
		const double x = AppendTimers::now();
		g_bt.t[3] += x - tk;
		tk = x;
	}
	if (group)
		order_hits_like_hitmgr(r->hits, r->qoff, s->P.local != 0);
	if (g_bt.on)
		g_bt.t[4] += AppendTimers::now() - tk;
	*out = r;
	return 0;
}

extern "C" int usb_batch_download(usb_searcher *s, usb_result **out)
{
	if (!s || !out)
		return fail(USB_EINVAL, "null argument");
	if (!s->ran)
		return fail(USB_EINVAL, "usb_batch_download before usb_batch_run");
	CK(cudaSetDevice(s->ix->device));
	return download_result(s, s->n_q, true, out);
}

extern "C" int usb_search_batch(usb_searcher *s, const uint8_t *qseqs, const uint64_t *q_off, uint32_t n_q,
  usb_result **out)
{
	double tk = g_bt.on ? AppendTimers::now() : 0;
	int rc = upload_queries(s, qseqs, q_off, n_q);
	if (rc)
		return rc;
	if (g_bt.on) {
		cudaStreamSynchronize(s->stream);
		const double x = AppendTimers::now();
		g_bt.t[0] += x - tk;
		tk = x;
	}
	if ((rc = usb_batch_run(s, nullptr)))
		return rc;
	if (g_bt.on) {
		g_bt.t[1] += AppendTimers::now() - tk;
		++g_bt.calls;
	}
	return usb_batch_download(s, out);
}

extern "C" int usb_batch_export_hits_device(usb_searcher *s, void *dev_dst, uint64_t cap_hits, uint64_t *n_hits)
{
	if (!s || !s->ran || !n_hits)
		return fail(USB_EINVAL, "usb_batch_export_hits_device: no completed batch");
	CK(cudaSetDevice(s->ix->device));
	*n_hits = s->last_hits;
	const uint64_t n = std::min<uint64_t>(cap_hits, s->last_hits);
	if (n && dev_dst) {
		CK(cudaMemcpyAsync(dev_dst, s->d_hits.p, n * sizeof(usb_hit), cudaMemcpyDeviceToDevice, s->stream));
		CK(cudaStreamSynchronize(s->stream));
	}
	return 0;
}

// ------------------------------------------------------------------ result accessors
extern "C" uint64_t usb_result_hit_count(const usb_result *r) { return r ? r->hits.size() : 0; }
extern "C" const usb_hit *usb_result_hits(const usb_result *r) { return r ? r->hits.data() : nullptr; }
extern "C" const uint32_t *usb_result_runs(const usb_result *r, uint64_t *n_runs)
{
	if (n_runs)
		*n_runs = r ? r->runs.size() : 0;
	return r ? r->runs.data() : nullptr;
}
extern "C" const uint64_t *usb_result_query_offsets(const usb_result *r) { return r ? r->qoff.data() : nullptr; }
extern "C" const usb_qstat *usb_result_qstats(const usb_result *r) { return r ? r->qstat.data() : nullptr; }
extern "C" void usb_result_free(usb_result *r) { result_recycle(r); }

extern "C" uint32_t usb_result_path(const usb_result *r, const usb_hit *h, char *buf)
{
	static const char ops[4] = {'M', 'D', 'I', '?'};
	uint32_t n = 0;
	for (uint32_t k = 0; k < h->run_cnt; ++k) {
		const uint32_t v = r->runs[(size_t)h->run_off + k];
		const uint32_t len = v >> 2;
		memset(buf + n, ops[v & 3], len);
		n += len;
	}
	buf[n] = 0;
	return n;
}

// ------------------------------------------------------------------ stage-level entry points
extern "C" int usb_rank_batch(usb_searcher *s, const uint8_t *qseqs, const uint64_t *q_off, uint32_t n_q,
  uint32_t k_max, uint32_t *cand_t, uint32_t *cand_u, uint32_t *n_cand, uint32_t *u_out)
{
	if (k_max == 0 || k_max > RANK_KCAP)
		return fail(USB_EINVAL, "k_max must be 1..%u", RANK_KCAP);
	int rc = upload_queries(s, qseqs, q_off, n_q);
	if (rc)
		return rc;
	const uint32_t strands = s->P.strand_both ? 2 : 1, n_jobs = n_q * strands;
	const uint32_t N = s->ix->S.n();
	CK(cudaMemsetAsync(s->d_ctr.p, 0, sizeof(DevCounters), s->stream));
	if ((rc = launch_rank(s, n_jobs, strands, k_max, u_out != nullptr)))
		return rc;
	DevCounters c;
	CK(cudaMemcpyAsync(&c, s->d_ctr.p, sizeof c, cudaMemcpyDeviceToHost, s->stream));
	std::vector<uint32_t> nemit(n_jobs);
	if (n_jobs) {
		CK(cudaMemcpyAsync(cand_t, s->d_cand_t.p, (size_t)n_jobs * k_max * 4, cudaMemcpyDeviceToHost, s->stream));
		if (cand_u)
			CK(cudaMemcpyAsync(cand_u, s->d_cand_u.p, (size_t)n_jobs * k_max * 4, cudaMemcpyDeviceToHost, s->stream));
		CK(cudaMemcpyAsync(n_cand, s->d_ncand.p, (size_t)n_jobs * 4, cudaMemcpyDeviceToHost, s->stream));
		if (u_out)
			CK(cudaMemcpyAsync(u_out, s->d_uout.p, (size_t)n_jobs * N * 4, cudaMemcpyDeviceToHost, s->stream));
	}
	CK(cudaStreamSynchronize(s->stream));
	if (c.err)
		return fail(USB_ELIMIT, "%s", err_text(c.err));
	return 0;
}

extern "C" int usb_align_pairs(usb_searcher *s, const uint8_t *qseqs, const uint64_t *q_off, uint32_t n_q,
  const uint32_t *pair_q, const uint32_t *pair_t, uint32_t n_pairs, uint8_t *aligned, usb_result **out,
  uint32_t *hsp_out, uint32_t max_hsp)
{
	if (!pair_q || !pair_t || !aligned || !out)
		return fail(USB_EINVAL, "usb_align_pairs: null argument");
	const usb_index *ix = s->ix;
	for (uint32_t i = 0; i < n_pairs; ++i)
		if (pair_q[i] >= n_q || pair_t[i] >= ix->S.n())
			return fail(USB_EINVAL, "pair %u out of range", i);
	int rc = upload_queries(s, qseqs, q_off, n_q);
	if (rc)
		return rc;
	const uint32_t hsp_cap = std::max<uint32_t>(64, ix->S.max_len / 8 + 16);
	AlignGeom g;
	StageGeom sg;
	const bool staged = stage_geometry(s, s->max_ql, ix->S.max_len, hsp_cap, sg);
	if (!staged && (rc = align_geometry(s, s->max_ql, ix->S.max_len, hsp_cap, g)))
		return rc;
	DevBuf<uint32_t> d_pq, d_pt, d_hsp;
	DevBuf<uint8_t> d_al;
	const size_t hsp_words = hsp_out ? (size_t)n_pairs * (1 + 4 * max_hsp) : 0;
	uint64_t runs_cap = std::max<uint64_t>((uint64_t)1 << 20, (uint64_t)n_pairs * 64);
	StageCaps caps;
	caps.recs = (uint64_t)n_pairs + 64;
	caps.hsp_words = caps.recs * 12;
	caps.staged = caps.recs;
	auto cleanup = [&]() { d_pq.release(); d_pt.release(); d_hsp.release(); d_al.release(); };
	if ((rc = d_pq.reserve(n_pairs + 1)) || (rc = d_pt.reserve(n_pairs + 1)) || (rc = d_al.reserve(n_pairs + 1)) ||
	    (rc = d_hsp.reserve(hsp_words + 1)) || (rc = s->d_hits.reserve(n_pairs + 1)) ||
	    (!staged && (rc = s->d_slab.reserve((size_t)g.n_warps * g.slab_stride)))) {
		cleanup();
		return rc;
	}
	cudaMemcpyAsync(d_pq.p, pair_q, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, s->stream);
	cudaMemcpyAsync(d_pt.p, pair_t, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, s->stream);
	DevCounters c;
	memset(&c, 0, sizeof c);
	for (int attempt = 0; n_pairs; ++attempt) {
		if ((rc = s->d_runs.reserve(runs_cap))) {
			cleanup();
			return rc;
		}
		cudaMemsetAsync(s->d_ctr.p, 0, sizeof(DevCounters), s->stream);
		cudaMemsetAsync(d_al.p, 0, n_pairs, s->stream);
		AlignArgs a;
		if (staged) {
			memset(&a, 0, sizeof a);
			a.P = s->D;
			a.q = s->d_q.p;
			a.q_off = s->d_qoff.p;
			a.db_seq = ix->d_seqs.p;
			a.db_off = ix->d_seq_off.p;
			a.db_len = ix->d_seq_len.p;
			a.ctr = s->d_ctr.p;
		} else
			fill_align_args(s, g, hsp_cap, a);
		a.n_jobs = n_pairs;
		a.strands = 1;
		a.k_max = 1;
		a.pair_q = d_pq.p;
		a.pair_t = d_pt.p;
		a.hits = s->d_hits.p;
		a.hits_cap = n_pairs;
		a.runs = s->d_runs.p;
		a.runs_cap = (uint32_t)std::min<uint64_t>(s->d_runs.cap, 0xfffffff0ull);
		a.aligned = d_al.p;
		a.hsp_out = hsp_out ? d_hsp.p : nullptr;
		a.max_hsp = max_hsp;
		cudaError_t e = cudaSuccess;
		if (staged) {
			if ((rc = run_staged(s, sg, a, hsp_cap, caps))) {
				cleanup();
				return rc;
			}
		} else {
			e = launch_align(a, g, s->stream);
			++s->launches;
		}
		if (e == cudaSuccess)
			e = cudaMemcpyAsync(&c, s->d_ctr.p, sizeof c, cudaMemcpyDeviceToHost, s->stream);
		if (e == cudaSuccess)
			e = cudaStreamSynchronize(s->stream);
		if (e != cudaSuccess) {
			cleanup();
			return fail(USB_ECUDA, "align kernel failed: %s", cudaGetErrorString(e));
		}
		if (c.err && !(c.err & ~(ERR_RUNS_FULL | ERR_HSPARENA_FULL)) && attempt < 5) {
			if (c.err & ERR_RUNS_FULL)
				runs_cap *= 4;
			if (c.err & ERR_HSPARENA_FULL)
				caps.hsp_words *= 4;
			continue;
		}
		break;
	}
	if (c.err) {
		cleanup();
		return fail(USB_ELIMIT, "%s", err_text(c.err));
	}
	cudaMemcpyAsync(aligned, d_al.p, n_pairs, cudaMemcpyDeviceToHost, s->stream);
	if (hsp_out)
		cudaMemcpyAsync(hsp_out, d_hsp.p, hsp_words * 4, cudaMemcpyDeviceToHost, s->stream);
	s->last_hits = c.n_hits;
	s->last_runs = c.n_runs;
	s->n_jobs = 0;
	rc = download_result(s, n_pairs, false, out); // grouped by pair index (hit.rank)
	cleanup();
	return rc;
}

extern "C" int usb_local_evalue(const usb_searcher *s, int32_t raw, uint32_t ql, double *evalue, double *bits)
{
	if (!s || !s->P.local)
		return fail(USB_EINVAL, "usb_local_evalue: not a local searcher");
	EStats es;
	es.init(s->P);
	if (evalue)
		*evalue = es.raw_to_evalue((double)raw, ql);
	if (bits)
		*bits = es.raw_to_bits((double)raw);
	return 0;
}

extern "C" int usb_params_evalue(const usb_params *p, int32_t raw, uint32_t ql, double *evalue, double *bits)
{
	if (!p || p->struct_size != sizeof(usb_params) || !p->local)
		return fail(USB_EINVAL, "usb_params_evalue: not a -usearch_local parameter block");
	EStats es;
	es.init(*p);
	if (evalue)
		*evalue = es.raw_to_evalue((double)raw, ql);
	if (bits)
		*bits = es.raw_to_bits((double)raw);
	return 0;
}

extern "C" int usb_local_pairs(usb_searcher *s, const uint8_t *qseqs, const uint64_t *q_off, uint32_t n_q,
  const uint32_t *pair_q, const uint32_t *pair_t, uint32_t n_pairs, usb_result **out)
{
	if (!s || !pair_q || !pair_t || !out)
		return fail(USB_EINVAL, "usb_local_pairs: null argument");
	if (!s->P.local)
		return fail(USB_EINVAL, "usb_local_pairs: not a local searcher");
	const usb_index *ix = s->ix;
	for (uint32_t i = 0; i < n_pairs; ++i)
		if (pair_q[i] >= n_q || pair_t[i] >= ix->S.n())
			return fail(USB_EINVAL, "pair %u out of range", i);
	int rc = upload_queries(s, qseqs, q_off, n_q);
	if (rc)
		return rc;
	LocalGeom g;
	if ((rc = local_geometry(s, s->max_ql, ix->S.max_len, g)))
		return rc;
	DevBuf<uint32_t> d_pq, d_pt;
	uint64_t hits_cap = (uint64_t)n_pairs * 4 + 64;
	uint64_t runs_cap = std::max<uint64_t>((uint64_t)1 << 20, hits_cap * 64);
	auto cleanup = [&]() { d_pq.release(); d_pt.release(); };
	if ((rc = d_pq.reserve(n_pairs + 1)) || (rc = d_pt.reserve(n_pairs + 1)) ||
	    (rc = s->d_slab.reserve((size_t)g.n_warps * g.slab_stride))) {
		cleanup();
		return rc;
	}
	cudaMemcpyAsync(d_pq.p, pair_q, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, s->stream);
	cudaMemcpyAsync(d_pt.p, pair_t, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, s->stream);
	DevCounters c;
	memset(&c, 0, sizeof c);
	for (int attempt = 0; n_pairs; ++attempt) {
		if ((rc = s->d_runs.reserve(runs_cap)) || (rc = s->d_hits.reserve(hits_cap))) {
			cleanup();
			return rc;
		}
		cudaMemsetAsync(s->d_ctr.p, 0, sizeof(DevCounters), s->stream);
		LocalArgs a;
		fill_local_args(s, g, a);
		a.n_jobs = n_pairs;
		a.strands = 1;
		a.pair_q = d_pq.p;
		a.pair_t = d_pt.p;
		a.hits = s->d_hits.p;
		a.hits_cap = (uint32_t)hits_cap;
		a.runs = s->d_runs.p;
		a.runs_cap = (uint32_t)std::min<uint64_t>(s->d_runs.cap, 0xfffffff0ull);
		if ((rc = launch_local(s, a, g))) {
			cleanup();
			return rc;
		}
		cudaError_t e = cudaMemcpyAsync(&c, s->d_ctr.p, sizeof c, cudaMemcpyDeviceToHost, s->stream);
		if (e == cudaSuccess)
			e = cudaStreamSynchronize(s->stream);
		if (e != cudaSuccess) {
			cleanup();
			return fail(USB_ECUDA, "local kernel failed: %s", cudaGetErrorString(e));
		}
		if (c.err && !(c.err & ~(ERR_RUNS_FULL | ERR_HITS_FULL)) && attempt < 4) {
			if (c.err & ERR_RUNS_FULL)
				runs_cap *= 4;
			if (c.err & ERR_HITS_FULL)
				hits_cap = std::max<uint64_t>(hits_cap * 2, (uint64_t)c.n_hits + 64);
			continue;
		}
		break;
	}
	cleanup();
	if (c.err)
		return fail(USB_ELIMIT, "%s", err_text(c.err));
	s->last_hits = c.n_hits;
	s->last_runs = c.n_runs;
	s->n_jobs = 0;
	if ((rc = download_result(s, n_pairs, false, out))) // grouped by pair index (hit.rank)
		return rc;
	usb_result *r = *out;
	for (uint32_t i = 0; i < n_pairs; ++i)
		std::sort(r->hits.begin() + r->qoff[i], r->hits.begin() + r->qoff[i + 1],
		  [](const usb_hit &x, const usb_hit &y) { return x.sub < y.sub; });
	return 0;
}

extern "C" int usb_viterbi_batch(usb_searcher *s, const uint8_t *a_seq, const uint64_t *a_off, const uint8_t *b_seq,
  const uint64_t *b_off, const uint8_t *flags, uint32_t n, char *paths, const uint64_t *path_off, int32_t *score2)
{
	if (!s || !a_off || !b_off || !flags || !paths || !path_off || !score2)
		return fail(USB_EINVAL, "usb_viterbi_batch: null argument");
	if (n == 0)
		return 0;
	CK(cudaSetDevice(s->ix->device));
	uint32_t max_a = 0, max_b = 0;
	uint64_t path_total = 0;
	for (uint32_t i = 0; i < n; ++i) {
		uint64_t la = a_off[i + 1] - a_off[i], lb = b_off[i + 1] - b_off[i];
		if (la == 0 || lb == 0 || la > 65000 || lb > (1u << 24))
			return fail(USB_EINVAL, "rectangle %u: side lengths %llu x %llu unsupported", i, (unsigned long long)la,
			  (unsigned long long)lb);
		max_a = std::max<uint32_t>(max_a, (uint32_t)la);
		max_b = std::max<uint32_t>(max_b, (uint32_t)lb);
		path_total = std::max<uint64_t>(path_total, path_off[i] + la + lb + 1);
	}
	AlignGeom g;
	int rc = align_geometry(s, max_a, max_b, 64, g);
	if (rc)
		return rc;
	DevBuf<uint8_t> d_a, d_b, d_f;
	DevBuf<uint64_t> d_ao, d_bo, d_po;
	DevBuf<char> d_paths;
	DevBuf<int> d_sc;
	auto cleanup = [&]() {
		d_a.release(); d_b.release(); d_f.release(); d_ao.release(); d_bo.release(); d_po.release();
		d_paths.release(); d_sc.release();
	};
	if ((rc = d_a.reserve(a_off[n] + 1)) || (rc = d_b.reserve(b_off[n] + 1)) || (rc = d_f.reserve(n)) ||
	    (rc = d_ao.reserve(n + 1)) || (rc = d_bo.reserve(n + 1)) || (rc = d_po.reserve(n)) ||
	    (rc = d_paths.reserve(path_total)) || (rc = d_sc.reserve(n)) ||
	    (rc = s->d_slab.reserve((size_t)g.n_warps * g.slab_stride))) {
		cleanup();
		return rc;
	}
	cudaMemcpyAsync(d_a.p, a_seq, a_off[n], cudaMemcpyHostToDevice, s->stream);
	cudaMemcpyAsync(d_b.p, b_seq, b_off[n], cudaMemcpyHostToDevice, s->stream);
	cudaMemcpyAsync(d_f.p, flags, n, cudaMemcpyHostToDevice, s->stream);
	cudaMemcpyAsync(d_ao.p, a_off, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, s->stream);
	cudaMemcpyAsync(d_bo.p, b_off, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, s->stream);
	cudaMemcpyAsync(d_po.p, path_off, (size_t)n * 8, cudaMemcpyHostToDevice, s->stream);
	cudaMemsetAsync(s->d_ctr.p, 0, sizeof(DevCounters), s->stream);
	ViterbiArgs v;
	memset(&v, 0, sizeof v);
	fill_align_args(s, g, 64, v.base);
	v.a_seq = d_a.p; v.a_off = d_ao.p; v.b_seq = d_b.p; v.b_off = d_bo.p; v.flags = d_f.p;
	v.n = n; v.paths = d_paths.p; v.path_off = d_po.p; v.score2 = d_sc.p;
	cudaError_t e = launch_viterbi(v, g, s->stream);
	++s->launches;
	DevCounters c;
	memset(&c, 0, sizeof c);
	if (e == cudaSuccess)
		e = cudaMemcpyAsync(&c, s->d_ctr.p, sizeof c, cudaMemcpyDeviceToHost, s->stream);
	if (e == cudaSuccess)
		e = cudaMemcpyAsync(paths, d_paths.p, path_total, cudaMemcpyDeviceToHost, s->stream);
	if (e == cudaSuccess)
		e = cudaMemcpyAsync(score2, d_sc.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s->stream);
	if (e == cudaSuccess)
		e = cudaStreamSynchronize(s->stream);
	cleanup();
	if (e != cudaSuccess)
		return fail(USB_ECUDA, "viterbi kernel failed: %s", cudaGetErrorString(e));
	if (c.err)
		return fail(USB_ELIMIT, "%s", err_text(c.err));
	return 0;
}

#include "usb_cluster.inc"
