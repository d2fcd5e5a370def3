/*
 * uso.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the usearch12 USEARCH/UCLUST hot path (UDB word index, U-sort candidate
 * ranking, HSP-seeded banded global Viterbi, accept/terminate loop) used ONLY by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg as the checker for the CUDA path.
 * The product (usearch12_b200/) never links, imports or executes anything in this directory.
 *
 * Parity pinning: the oracle is checked against outputs of the UNMODIFIED reference binary
 * (oracle/_ref/usearch12, built by oracle/Makefile.ref from /root/reference/src) -- the golden
 * files under tests/golden/ were produced by that binary via tools/make_golden.py.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference/src).
 */
#ifndef USO_H
#define USO_H
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Snapshot of the options the path reads.  Floats stay `float`: the
 * reference stores every float option as float and widens on read (opts.cpp:8-15,80-88). */
typedef struct uso_params {
	int is_nucleo;       /* 1 = nt DB, 0 = amino acid DB (BLOSUM62, gap open -17, HSP words of 3 letters) */
	float id;            /* -id */
	unsigned maxaccepts; /* terminator.cpp:23-31 */
	unsigned maxrejects;
	int strand_both;     /* -strand both */
	unsigned word_length;/* UDB word length, nt 8 (udbparams.cpp:246-250) */
	unsigned big;        /* -big 100000 (o_defaults.inc) */
	unsigned bump;       /* -bump 50 */
	unsigned stepwords;  /* -stepwords 8 */
	unsigned band;       /* -band 16 */
	unsigned minhsp;     /* -minhsp 16 */
	unsigned hspw;       /* HSP finder word length, nt 5 (alnheuristics.cpp:37) */
	float xdrop_nw;      /* -xdrop_nw 8 */
	float match;         /* -match 1 */
	float mismatch;      /* -mismatch -2 */
	int dbmask_fast;     /* 1 = fastnucleo soft masking of the DB (makeudb.cpp:11-25) */
	int cluster_mode;    /* 1 = cluster_fast semantics (no DB masking, growing DB) */
	int fulldp;          /* -fulldp: no HSPs, full Viterbi on every candidate (globalalignmem.cpp:153-157) */
	int local;           /* 1 = -usearch_local: LocalAligner2 (localmulti.cpp), X-drop gapped extension */
	float evalue;        /* -evalue (mandatory for usearch_local) */
	float xdrop_u;       /* -xdrop_u 16 */
	float xdrop_g;       /* -xdrop_g 32 */
	float lopen;         /* local gap open, -10 (alnparams.cpp:362-369) */
	float lext;          /* local gap extend, -1 */
	float ka_dbsize;     /* -ka_dbsize, default 1e9 (o_defaults.inc:2); 0 = total DB letters */
} uso_params;

void uso_default_params(uso_params *p, int cluster_fast);
void uso_set_amino(uso_params *p); /* word_length 5 over 20 letters, hspw 3 */

/* One accepted hit (global alignment). */
typedef struct uso_hit {
	uint32_t query;      /* query index in the batch */
	uint32_t target;     /* DB target index */
	uint8_t strand;      /* 0 plus, 1 minus (query rev-comped) */
	uint32_t ids, mism, intgaps, opens;
	uint32_t first_mq, first_mt, last_mq, last_mt; /* 0-based first/last M positions */
	uint32_t first_mcol, alnlen; /* cols between first and last M inclusive */
	uint32_t ql, tl;
	/* local alignments only (alignresult.cpp:173): segment start/length in query and target,
	 * raw score (arscorer.cpp:87-103), E-value and bit score (estats.cpp:73-96) */
	uint32_t loi, loj, leni, lenj;
	double raw, evalue, bits;
	char *path;          /* full path over {M,D,I}, NUL-terminated, owned by the hit */
} uso_hit;

typedef struct uso_db uso_db;
typedef struct uso_searcher uso_searcher;

/* DB: sequences are copied; masking applied per params (dbmask_fast && !cluster_mode). */
uso_db *uso_db_create(const uso_params *p);
uint32_t uso_db_add(uso_db *db, const uint8_t *seq, uint32_t L, const char *label); /* udbbuild.cpp:286 */
void uso_db_free(uso_db *db);
uint32_t uso_db_seq_count(const uso_db *db);
const uint8_t *uso_db_seq(const uso_db *db, uint32_t i, uint32_t *L);
const char *uso_db_label(const uso_db *db, uint32_t i);
/* index introspection for parity tests */
uint32_t uso_db_slot_count(const uso_db *db);
const uint32_t *uso_db_row(const uso_db *db, uint32_t word, uint32_t *size);

uso_searcher *uso_searcher_create(uso_db *db, const uso_params *p);
void uso_searcher_free(uso_searcher *s);

/* Search one query (both strands if strand_both).  Hits are appended to *hits (realloc'd),
 * in HitMgr output order (hitmgr.cpp:477, sort.h:63-102).  Returns number of hits appended. */
unsigned uso_search(uso_searcher *s, uint32_t qindex, const uint8_t *q, uint32_t L,
  uso_hit **hits, unsigned *nhits, unsigned *caphits);
void uso_hits_free(uso_hit *hits, unsigned n);

/* ---- stage-level entry points for kernel parity tests ---- */
/* a1/a2: unique UDB words of a query in first-occurrence order; returns count. */
unsigned uso_query_unique_words(const uso_params *p, const uint8_t *q, uint32_t L, uint32_t *words);
/* a4..a6: U vector + ranked candidate list (small-DB path).  U must hold seq_count entries;
 * cand_t/cand_u hold up to seq_count entries.  Returns candidate count (TopOrder.Size). */
unsigned uso_rank_candidates(uso_searcher *s, const uint8_t *q, uint32_t L, uint32_t *U,
  uint32_t *cand_t, uint32_t *cand_u);
/* the same for the big-database path (udbusortedsearcherbig.cpp:82-135); cand_* sized db n */
unsigned uso_rank_candidates_big(uso_searcher *s, const uint8_t *q, uint32_t L, uint32_t *cand_t, uint32_t *cand_u);
/* a10..a12: ungapped + chained HSPs for (q, target).  Arrays hold up to max_hsp entries of
 * {Loi, Loj, Len, score*2 (int)}; returns chained count, *n_ungapped set, *hsp_fract_id set. */
unsigned uso_global_hsps(uso_searcher *s, const uint8_t *q, uint32_t LQ, const uint8_t *t, uint32_t LT,
  uint32_t *ungapped, unsigned *n_ungapped, uint32_t *chained, unsigned max_hsp, float *hsp_fract_id);
/* a14/a15: banded Viterbi on a rectangle with explicit terminal flags; writes path (cap LA+LB+1). */
float uso_viterbi_band(const uso_params *p, const uint8_t *A, uint32_t LA, const uint8_t *B, uint32_t LB,
  int left_a, int left_b, int right_a, int right_b, char *path);
/* a9: full global alignment of (q,t); returns 0 if rejected by the HSP gate, else 1 and path. */
int uso_global_align(uso_searcher *s, const uint8_t *q, uint32_t LQ, const uint8_t *t, uint32_t LT, char *path);
/* a24: FastMaskSeq (fastmask.cpp:88-158) */
void uso_fastmask(const uint8_t *seq, uint32_t L, uint8_t *out);
/* a23: reverse complement (seqinfo.cpp:292-325) */
void uso_revcomp(const uint8_t *seq, uint32_t L, uint8_t *out);
/* comppath.cpp:7-48 */
void uso_compress_path(const char *path, char *out);

/* a18-a20: local alignment of (q,t) around the seed (qpos,tpos) (localaligner.cpp:101-211);
 * returns 0 if rejected, else fills loi/loj/leni/lenj/score and path. */
int uso_local_align_pos(uso_searcher *s, const uint8_t *q, uint32_t LQ, const uint8_t *t, uint32_t LT,
  uint32_t qpos, uint32_t tpos, uint32_t *hsp4, float *score, char *path);
/* X-drop forward extension alone (xdropfwdmem.cpp:344-749): the cmd_test known answer. */
float uso_xdrop_fwd(const uso_params *p, const uint8_t *A, uint32_t LA, const uint8_t *B, uint32_t LB, float X,
  uint32_t *leni, uint32_t *lenj, char *path);
void uso_write_userout_local(FILE *f, const uso_hit *h, const char *qlabel, const char *tlabel, int nucleo);
void uso_write_blast6_local(FILE *f, const uso_hit *h, const char *qlabel, const char *tlabel); /* blast6out.cpp:27-80 */
void uso_write_uc_hit_local(FILE *f, const uso_hit *h, const char *qlabel, const char *tlabel, int nucleo);

/* ---- output formats (byte-identical to the reference's sinks) ---- */
/* userout with -userfields query+target+id+alnlen+mism+opens+qlo+qhi+tlo+thi+caln+qstrand (userout.cpp:150-215) */
void uso_write_userout(FILE *f, const uso_hit *h, const char *qlabel, const char *tlabel);
void uso_write_blast6(FILE *f, const uso_hit *h, const char *qlabel, const char *tlabel); /* blast6out.cpp:27-80 */
/* same with the strand column of an amino acid search ('.') when nucleo == 0 */
void uso_write_userout2(FILE *f, const uso_hit *h, const char *qlabel, const char *tlabel, int nucleo);
void uso_write_uc_hit2(FILE *f, const uso_hit *h, const char *qlabel, const char *tlabel, int nucleo);
void uso_write_uc_hit(FILE *f, const uso_hit *h, const char *qlabel, const char *tlabel); /* outputuc.cpp:45-69 */
void uso_write_uc_nohit(FILE *f, uint32_t ql, const char *qlabel);                        /* outputuc.cpp:19-20 */

#ifdef __cplusplus
}
#endif
#endif
