/*
 * usb200.h -- C ABI of the B200-native USEARCH/UCLUST hot path (libusb200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch
 * types, no exceptions.  Every entry point names the reference interface it replaces
 * (paths relative to /root/reference/src).  All functions return 0 on success or a negative
 * USB_E* code; usb_last_error() returns the message of the last failure on the calling thread.
 * The reference's own error convention is Die() = message + exit(1) (myutils.cpp:867); the host
 * shim (usearch12_b200/csrc/host) maps non-zero returns to that behaviour.
 *
 * There is NO CPU fallback behind this interface: every compute entry point runs CUDA kernels
 * built for sm_100a and fails with USB_ECUDA when no device is usable.
 */
#ifndef USB200_H
#define USB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define USB_OK 0
#define USB_EINVAL (-1)   /* bad argument / unsupported option combination (fails loudly) */
#define USB_ECUDA (-2)    /* CUDA runtime error or no device */
#define USB_ENOMEM (-3)
#define USB_ELIMIT (-4)   /* an internal capacity was exceeded (message says which) */

/* POD snapshot of the options the path reads.  Float options stay `float`
 * because the reference stores them as float and widens on read (opts.cpp:8-15,80-88); the
 * accept test compares double(ids)/double(cols) with (double)(float)id (accepter.cpp:36-38). */
typedef struct usb_params {
	uint32_t struct_size;   /* = sizeof(usb_params), ABI check */
	int32_t is_nucleo;      /* 1 = nucleotide DB; 0 = amino acid DB (only with local = 1) */
	float id;               /* -id */
	uint32_t maxaccepts;    /* -maxaccepts (terminator.cpp:23-31), default 1; 0 = no limit: with more than 1 024 targets the
	                           candidate loop then walks whole U-sorted lists (at most 2^29 query-strand x target pairs per batch) */
	uint32_t maxrejects;    /* -maxrejects, default 32 (8 for cluster_fast) */
	int32_t strand_both;    /* -strand both (searcher.cpp:144-158) */
	uint32_t word_length;   /* UDB word length, nt 8 (udbparams.cpp:246-250) */
	uint32_t big;           /* -big, 100000 (o_defaults.inc:25) */
	uint32_t bump;          /* -bump, 50 (udbusortedsearcher.cpp:269-282) */
	uint32_t stepwords;     /* -stepwords, 8 (wordparams.cpp:167-192) */
	uint32_t band;          /* -band, 16 (alnheuristics.cpp:33) */
	uint32_t minhsp;        /* -minhsp, 16 */
	uint32_t hspw;          /* HSP finder word length, nt 5 (alnheuristics.cpp:37) */
	float xdrop_nw;         /* -xdrop_nw, 8 */
	float match;            /* -match, 1 */
	float mismatch;         /* -mismatch, -2 */
	float gap_open;         /* internal gap open, nt -10 (alnparams.cpp:378-384) */
	float gap_ext;          /* internal gap extend, -1 */
	float term_gap_open;    /* terminal gap open, -0.5 */
	float term_gap_ext;     /* terminal gap extend, -0.5 */
	int32_t dbmask;         /* 1 = fastnucleo soft-masking of the DB (makeudb.cpp:11-25) */
	int32_t cluster_mode;   /* 1 = cluster_fast semantics: raw un-masked centroids, growing DB */
	int32_t fulldp;         /* -fulldp: no HSPs, full Viterbi per candidate (globalalignmem.cpp:153-157);
	                           band = 0 (-band 0) alone selects the full DP for holes only (:103-106) */
	/* -usearch_local (searchcmd.cpp:42, makedbsearcher.cpp:90-123) */
	int32_t local;          /* 1 = LocalAligner2 + X-drop gapped extension instead of GlobalAligner */
	float evalue;           /* -evalue (mandatory for usearch_local) */
	float xdrop_u;          /* -xdrop_u, 16: ungapped X-drop (alnheuristics.cpp:29) */
	float xdrop_g;          /* -xdrop_g, 32: gapped X-drop (alnheuristics.cpp:30) */
	float lopen;            /* local gap open, -10 (alnparams.cpp:362-369; -lopen/-lext defaults count as set) */
	float lext;             /* local gap extend, -1 */
	float ka_dbsize;        /* -ka_dbsize, 1e9 (o_defaults.inc:2; the default counts as set, so the
	                           letter count of the DB is never used, makedbsearcher.cpp:92-96) */
	/* Accepter / Terminator options beyond -id (accepter.cpp:41-94,145-197, terminator.cpp:66-86).
	 * accept_flags = the options that are set (the reference tests ofilled()/oget_flag()); the
	 * values are compared the way the reference compares them (double vs widened float). */
	uint32_t accept_flags;  /* USB_ACC_* */
	float maxid;            /* -maxid (default 1.0 counts as set, o_defaults.inc:7) */
	uint32_t mincols;       /* -mincols: alignment columns between the first and last M */
	uint32_t maxgaps;       /* -maxgaps: internal gap columns */
	uint32_t maxdiffs;      /* -maxdiffs: mismatches + internal gap columns */
	uint32_t mindiffs;      /* -mindiffs */
	float query_cov, max_query_cov;   /* (lastMq - firstMq + 1) / QL   (arscorer.cpp:122-137) */
	float target_cov, max_target_cov; /* (ids + mismatches) / TL        (arscorer.cpp:139-154) */
	float abskew;           /* -abskew: target size= / query size= (arscorer.cpp:809-816) */
	float min_sizeratio;    /* -min_sizeratio: same ratio, tested before the alignment */
	float minqt, maxqt;     /* QL / TL bounds */
	float minsl, maxsl;     /* shorter / longer bounds */
	float termid;           /* -termid: stop when the lowest accepted identity <= termid */
	float termidd;          /* -termidd: stop when highest - lowest accepted identity > termidd */
} usb_params;

#define USB_ACC_SELF 0x1u           /* -self: reject pairs with identical labels */
#define USB_ACC_NOTSELF 0x2u        /* -notself: reject pairs with different labels */
#define USB_ACC_SELFID 0x4u         /* -selfid: reject pairs with identical letters (global only) */
#define USB_ACC_MAXID 0x8u
#define USB_ACC_MINCOLS 0x10u
#define USB_ACC_MAXGAPS 0x20u
#define USB_ACC_QUERY_COV 0x40u
#define USB_ACC_MAX_QUERY_COV 0x80u
#define USB_ACC_TARGET_COV 0x100u
#define USB_ACC_MAX_TARGET_COV 0x200u
#define USB_ACC_MAXDIFFS 0x400u
#define USB_ACC_MINDIFFS 0x800u
#define USB_ACC_ABSKEW 0x1000u
#define USB_ACC_MIN_SIZERATIO 0x2000u
#define USB_ACC_MINQT 0x4000u
#define USB_ACC_MAXQT 0x8000u
#define USB_ACC_MINSL 0x10000u
#define USB_ACC_MAXSL 0x20000u
#define USB_ACC_TERMID 0x40000u
#define USB_ACC_TERMIDD 0x80000u
/* options that need the label identities / size= annotations of usb_index_set_attrs and
 * usb_batch_set_query_attrs */
#define USB_ACC_NEEDS_LABELS (USB_ACC_SELF | USB_ACC_NOTSELF)
#define USB_ACC_NEEDS_SIZES (USB_ACC_ABSKEW | USB_ACC_MIN_SIZERATIO)

/* Defaults of -usearch_global (cluster_fast=0) or -cluster_fast (=1). */
void usb_default_params(usb_params *p, int cluster_fast);
/* Switches a parameter block to -usearch_local on a nucleotide (nucleo=1: UDB words of 8, seed
 * words of 5) or amino acid DB (nucleo=0: UDB words of 5 over 20 letters, udbparams.cpp:236-261;
 * seed words of 3, makedbsearcher.cpp:104-123; BLOSUM62). */
void usb_set_local(usb_params *p, int nucleo, float evalue);
/* Switches a -usearch_global parameter block to an amino acid database (BASELINE config 1; the
 * reference builds a GlobalAligner for either alphabet, makedbsearcher.cpp:132-140): UDB words of 5
 * over 20 letters (udbparams.cpp:251-257), HSP words of 3 (alnheuristics.cpp:37), BLOSUM62
 * (blosum62.cpp:17-96), gap open -17 / extend -1 (alnparams.cpp:381-384), HSP identity gate
 * max(id, 0.5) (alnheuristics.cpp:56), no strands. */
void usb_set_amino(usb_params *p);

/* One accepted hit == the statistics the reference derives lazily from an AlignResult
 * (alignresult.h:17-245, arscorer.cpp:201-296 FillLo).  Coordinates are 0-based. */
typedef struct usb_hit {
	uint32_t query;        /* query index within the batch */
	uint32_t target;       /* DB target index (SeqDB index, m_Target->m_Index) */
	uint32_t strand;       /* 0 = plus, 1 = query reverse-complemented */
	uint32_t rank;         /* position of the target in the U-sorted candidate order */
	uint32_t ids, mism, intgaps, opens;
	uint32_t first_mq, first_mt, last_mq, last_mt; /* first / last M column positions */
	uint32_t first_mcol, alnlen;                   /* alnlen = cols between first and last M */
	uint32_t ql, tl;
	uint32_t run_off, run_cnt; /* path as runs in the run arena: (length << 2) | op, op 0=M 1=D 2=I */
	/* local hits (AlignResult::CreateLocalGapped, alignresult.cpp:173): the segment is
	 * [first_mq, last_mq] x [first_mt, last_mt] (m_HSP), alnlen = all columns */
	int32_t raw;           /* raw score (arscorer.cpp:87-103); 0 for global hits */
	uint32_t sub;          /* index of this AR among the ARs of its target (localmulti.cpp:93-97) */
} usb_hit;

/* Per (query,strand) search counters (diagnostics; the reference has no equivalent object). */
typedef struct usb_qstat {
	uint32_t n_cand;     /* TopOrder.Size (udbusortedsearcher.cpp:109-120) */
	uint32_t n_tried;    /* Align() calls made before the Terminator fired */
	uint32_t n_hspfail;  /* attempts rejected by the HSP identity gate (globalalignmem.cpp:171) */
	uint32_t n_dp;       /* banded Viterbi calls */
	uint32_t dp_cells;   /* DP cells computed */
	uint32_t n_accept;
	uint32_t seq_bytes;  /* sum over attempts of (query + target) letters read */
} usb_qstat;

typedef struct usb_index usb_index;       /* UDBData + SeqDB on one device (udbdata.h:15-31) */
typedef struct usb_searcher usb_searcher; /* UDBUsortedSearcher + GlobalAligner + Accepter + Terminator */
typedef struct usb_result usb_result;     /* what HitMgr holds after a batch (hitmgr.cpp:120-161) */

const char *usb_last_error(void);
/* Number of usable CUDA devices (0 when none); never fails. */
int usb_device_count(void);

/* ---- index: replaces LoadDB/LoadUDB (loaddb.cpp:100-127) = MaskDB (makeudb.cpp:11-25) +
 * UDBParams::FromCmdLine (udbparams.cpp:58) + UDBData::FromSeqDB (udbbuild.cpp:303-398).
 * seqs = concatenated target letters, seq_off[n_seq+1] byte offsets.  Sequences are copied. */
int usb_index_create(int device, const usb_params *p, const uint8_t *seqs, const uint64_t *seq_off,
  uint32_t n_seq, usb_index **out);
/* Appends n targets (indexes N..N+n-1) to a live index: UDBData::AddSIToDB_CopyData
 * (udbbuild.cpp:286 -> AddSeqNoncoded :256 -> AddWord/GrowRow :111,:74), as cluster_fast does for
 * every new centroid.  Searchers created on the index see the new targets on their next batch. */
int usb_index_append(usb_index *ix, const uint8_t *seqs, const uint64_t *seq_off, uint32_t n);
/* Capacity hint for an index that grows by appends (UDBData::AddSeq grows its SeqDB and rows on demand,
 * udbbuild.cpp:74-130, seqdb.cpp AddSeq_CopyData): room for n_seqs targets with n_letters letters in
 * all, on the host and on the device.  Optional; appends beyond the hint still work. */
int usb_index_reserve(usb_index *ix, uint32_t n_seqs, uint64_t n_letters);
void usb_index_free(usb_index *ix);
uint32_t usb_index_seq_count(const usb_index *ix);
uint64_t usb_index_posting_count(const usb_index *ix);
/* Bytes per posting of the rows as they lie in HBM now: 2 = bank-aware 2-byte layout (one static
 * segment of at most 131 070 targets, walked by the U-sort kernel only), 4 = ascending 4-byte rows
 * (UDBData::m_UDBRows, udbdata.h:15-31).  The layout is an internal choice; results are identical. */
uint32_t usb_index_posting_width(const usb_index *ix);
/* Host-side introspection for parity tests: one UDB row (m_UDBRows[word], m_Sizes[word]). */
int usb_index_row(const usb_index *ix, uint32_t word, const uint32_t **row, uint32_t *size);
/* Masked target sequence as indexed (SeqDB::GetSeq after SeqDB::Mask, seqdb.cpp:415). */
int usb_index_seq(const usb_index *ix, uint32_t target, const uint8_t **seq, uint32_t *len);

/* ---- full-length dereplication on the device: replaces DerepFull (derepfull.cpp:130-212; hashing
 * seqhash.cpp:6-51) with its -threads 1 result: sequences are equal when they have the same length
 * and the same letters ignoring case; uniq_of[i] = number of sequence i's unique, uniques numbered
 * in order of first occurrence (what DerepResult::m_ClusterCount / GetClusterIndex give,
 * derepresult.cpp:211-225,811-820); *n_uniq = their count.  Callers: -fastx_uniques and the first
 * step of -cluster_fast (clusterfast.cpp:97-100). */
int usb_derep_full(int device, const uint8_t *seqs, const uint64_t *seq_off, uint32_t n, uint32_t *uniq_of,
                   uint32_t *n_uniq);

/* ---- .udb database files (udbfile.h:17-62, udbio.cpp:242-364, seqdbio.cpp:17-258).  Host only: these
 * entry points need no device.
 * usb_udb_write = -makeudb_usearch (makeudb.cpp:27-60): MaskDB + UDBData::FromSeqDB + ToUDBFile for
 * the default index (not hashed / coded / spaced, -dbstep 1, -dbaccelpct 100); the file is byte for
 * byte the one the reference writes, so either program can load the other's databases.
 * labels[i] = NUL-terminated label of target i. */
int usb_udb_write(const char *path, const usb_params *p, const uint8_t *seqs, const uint64_t *seq_off,
  const char *const *labels, uint32_t n_seq);
/* 1 when the file starts with the .udb magic (loaddb.cpp:100-107 picks the loader the same way). */
int usb_udb_probe(const char *path);
/* UDBData::FromUDBFile (udbio.cpp:242-279): header, row sizes, rows, SeqDB.  The sequences come
 * back masked as stored (lower case = masked): pass them to usb_index_create with dbmask = 0,
 * like the reference, which does not mask a loaded .udb again (loaddb.cpp:107-118). */
typedef struct usb_udb usb_udb;
int usb_udb_read(const char *path, usb_udb **out);
void usb_udb_free(usb_udb *u);
uint32_t usb_udb_seq_count(const usb_udb *u);
int usb_udb_is_nucleo(const usb_udb *u);
uint32_t usb_udb_word_length(const usb_udb *u);
const uint8_t *usb_udb_seqs(const usb_udb *u, const uint64_t **seq_off);
const char *usb_udb_label(const usb_udb *u, uint32_t i);
/* m_UDBRows[word], m_Sizes[word] as stored in the file (parity checks against usb_index_row). */
int usb_udb_row(const usb_udb *u, uint32_t word, const uint32_t **row, uint32_t *size);

/* Host-only test hook: the 2-byte device layout (groups of 256 increment descriptors, see
 * usb_index_posting_width) of ONE row with the given ascending targets in an index of n_targets. */
int usb_debug_half_row(const uint32_t *targets, uint32_t n, uint32_t n_targets, uint16_t *out, uint32_t out_cap,
  uint32_t *groups, uint32_t *dummy0);

/* ---- searcher: replaces MakeDBSearcher (makedbsearcher.cpp:75) wiring for one device. */
int usb_searcher_create(usb_index *ix, const usb_params *p, usb_searcher **out);
void usb_searcher_free(usb_searcher *s);

/* ---- the hot path.  Replaces the per-query loop of Thread() (search.cpp:51-87):
 * Searcher::Search (searcher.cpp:122) -> UDBUsortedSearcher::SearchImpl
 * (udbusortedsearcher.cpp:122) -> GlobalAligner::Align (globalaligner.cpp:37) -> Searcher::OnAR
 * (searcher.cpp:52), for a whole batch of queries.  Host buffers in, host result out. */
int usb_search_batch(usb_searcher *s, const uint8_t *qseqs, const uint64_t *q_off, uint32_t n_q,
  usb_result **out);

/* Staged form of the same call (bench.py times the middle step with inputs resident in HBM). */
int usb_batch_upload(usb_searcher *s, const uint8_t *qseqs, const uint64_t *q_off, uint32_t n_q);
/* Runs all kernels on the uploaded batch.  ms[0]=U-sort kernel, ms[1]=align kernel, ms[2]=total
 * device time (CUDA events on the searcher's stream); ms may be NULL. */
int usb_batch_run(usb_searcher *s, float *ms);
int usb_batch_download(usb_searcher *s, usb_result **out);
/* Counters of the last usb_batch_run: out[0] = UDB postings read by the U-sort kernel,
 * out[1] = hits, out[2] = path runs, out[3] = (query,strand) jobs. */
int usb_batch_counters(const usb_searcher *s, uint64_t out[4]);
/* Device time of the last usb_batch_run by kernel family, in milliseconds (CUDA events on the
 * library stream): out[0] U-sort (UDBUsortedSearcher::SetTargetOrder, udbusortedsearcher.cpp:284-410),
 * out[1] HSP gate kernels (GlobalAlign_AllOpts up to the identity gate, globalalignmem.cpp:129-176),
 * out[2] DP kernels (AlignHSPMem / ViterbiFastBandMem + FillLo, globalalignmem.cpp:70-112),
 * out[3] work-list and commit kernels (Terminator, terminator.cpp:64-100), out[4] the number of
 * (query, target) records the gate passed on to the DP, out[5] DP cells, out[6] letters (query +
 * target) of those records, out[7] words of chained-HSP coordinates written by the gate.
 * Entries 1..7 are 0 when the candidate loop ran as one kernel (amino acids, unusual scores). */
int usb_batch_kernel_ms(const usb_searcher *s, double out[8]);
/* Label identities and size= annotations for the Accepter rules that read labels (-self, -notself:
 * accepter.cpp:150-154; -min_sizeratio, -abskew: GetSizeFromLabel, label.cpp:152-161).  label_id:
 * two sequences have the same label iff their ids are equal (the caller numbers the distinct
 * labels); size: the size= value.  Either array may be NULL when no set option needs it.
 * Targets [first, first + n) of the index; queries of the uploaded batch (call between
 * usb_batch_upload and usb_batch_run, or before usb_search_batch with queries_ahead = 1: the
 * attributes then apply to the next batch). */
int usb_index_set_attrs(usb_index *ix, uint32_t first, uint32_t n, const uint32_t *label_id, const uint32_t *size);
int usb_batch_set_query_attrs(usb_searcher *s, uint32_t n_q, const uint32_t *label_id, const uint32_t *size);
/* Number of kernel launches issued by this searcher so far. */
uint64_t usb_searcher_launch_count(const usb_searcher *s);
/* Device-resident packed hit records of the last usb_batch_run (for the NCCL gather of
 * section 8e): copies min(n_hits, cap_hits) usb_hit records into a caller-owned DEVICE buffer. */
int usb_batch_export_hits_device(usb_searcher *s, void *dev_dst, uint64_t cap_hits, uint64_t *n_hits);

/* ---- cluster_fast: one round of the greedy centroid loop.  Replaces, for a block of queries
 * in cluster order, the body of ClusterFast()'s loop (clusterfast.cpp:120-129): Searcher::Search
 * against the centroids so far, then ClusterSink::OnQueryDone (clustersink.cpp:306-330), which
 * appends a query without a hit to the database as a new centroid.  All n_q queries are searched
 * against the database as it is on entry; the longest prefix whose results provably equal the
 * sequential ones is committed (*n_committed >= 1 when n_q >= 1), its no-hit queries are appended
 * to the index in order, and cluster_idx[q] (q < *n_committed) receives the centroid index the
 * query belongs to (== its own new index when it became a centroid).  *out holds the hits of the
 * committed queries.  Call again with the remaining queries.  Needs -maxaccepts 1, -strand plus. */
int usb_cluster_round(usb_searcher *s, const uint8_t *qseqs, const uint64_t *q_off, uint32_t n_q,
  uint32_t *n_committed, uint32_t *cluster_idx, usb_result **out);

/* Karlin-Altschul statistics of a local hit (estats.cpp:73-96, gapped): E-value and bit score of
 * raw score `raw` for a query of ql letters, with the searcher's -ka_dbsize. */
int usb_local_evalue(const usb_searcher *s, int32_t raw, uint32_t ql, double *evalue, double *bits);
/* The same from a parameter block alone (usb_set_local); host only, needs no device. */
int usb_params_evalue(const usb_params *p, int32_t raw, uint32_t ql, double *evalue, double *bits);

/* a18/a19: LocalAligner2::AlignMulti for explicit (query, target) pairs (localmulti.cpp:9-118).
 * Every AR of pair i becomes a hit with rank = i and sub = its index among the pair's ARs
 * (no -id filter; the E-value gate of AlignPos applies). */
int usb_local_pairs(usb_searcher *s, const uint8_t *qseqs, const uint64_t *q_off, uint32_t n_q,
  const uint32_t *pair_q, const uint32_t *pair_t, uint32_t n_pairs, usb_result **out);

/* Result accessors.  Hits are grouped by query (ascending) and, within a query, in the
 * reference's output order (HitMgr::Sort, hitmgr.cpp:477; sort.h:63-102). */
uint64_t usb_result_hit_count(const usb_result *r);
const usb_hit *usb_result_hits(const usb_result *r);
const uint32_t *usb_result_runs(const usb_result *r, uint64_t *n_runs);
/* first_hit[q] .. first_hit[q+1] = hit range of query q (n_q+1 entries). */
const uint64_t *usb_result_query_offsets(const usb_result *r);
/* qstat[2*q + strand] */
const usb_qstat *usb_result_qstats(const usb_result *r);
void usb_result_free(usb_result *r);
/* Expands a hit's path to the reference's M/D/I string (PathInfo::GetPath, pathinfo.cpp:37-214);
 * buf must hold ql+tl+1 bytes.  Returns the length. */
uint32_t usb_result_path(const usb_result *r, const usb_hit *h, char *buf);

/* ---- stage-level entry points (each is one kernel; used by the parity tests and by bench.py's
 * per-kernel roofline).  All take host buffers.
 *
 * a1..a6: U-sort candidate ranking (udbusortedsearcher.cpp:109-120 SetTargetOrder).
 * For each query writes up to k_max candidates (target, U) in TopOrder order into
 * cand_t/cand_u[q*k_max ..] and TopOrder.Size into n_cand[q].  u_out (may be NULL) receives the
 * full U vector of each query (n_q * n_seq entries). */
int usb_rank_batch(usb_searcher *s, const uint8_t *qseqs, const uint64_t *q_off, uint32_t n_q,
  uint32_t k_max, uint32_t *cand_t, uint32_t *cand_u, uint32_t *n_cand, uint32_t *u_out);

/* a9..a15: GlobalAligner::Align for explicit (query, target) pairs (globalaligner.cpp:37-61).
 * aligned[i] = 0 when the HSP gate rejected the pair; otherwise hits[i] is filled (rank = i) and
 * the path goes to the result's run arena.  hsp_out (may be NULL): per pair 1 + 4*max_hsp words:
 * chained HSP count then {Loi, Loj, Len, 2*Score}. */
int usb_align_pairs(usb_searcher *s, const uint8_t *qseqs, const uint64_t *q_off, uint32_t n_q,
  const uint32_t *pair_q, const uint32_t *pair_t, uint32_t n_pairs, uint8_t *aligned,
  usb_result **out, uint32_t *hsp_out, uint32_t max_hsp);

/* a14/a15: banded Viterbi + traceback on explicit rectangles (viterbifastbandmem.cpp:232-253
 * ViterbiFastMainDiagMem).  flags bit0..3 = left_a, left_b, right_a, right_b terminal sides.
 * paths receives NUL-terminated M/D/I strings at path_off[i] (capacity la+lb+1 each);
 * score2[i] = 2 * alignment score. */
int usb_viterbi_batch(usb_searcher *s, const uint8_t *a, const uint64_t *a_off, const uint8_t *b,
  const uint64_t *b_off, const uint8_t *flags, uint32_t n, char *paths, const uint64_t *path_off,
  int32_t *score2);

#ifdef __cplusplus
}
#endif
#endif
