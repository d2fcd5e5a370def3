/*
 * uso.c -- CPU ORACLE (test infrastructure, NOT product code).  See uso.h.
 *
 * Plain-C restatement of the reference algorithm for the usearch_global / cluster_fast hot
 * path.  Floating-point DP is kept in `float` with the -9e9f sentinel exactly as the reference
 * does (viterbifastbandmem.cpp:34-50, mx.h:12); the CUDA path uses scaled integers and is
 * compared against this.  References are /root/reference/src/<file>:<line>.
 */
#include "uso.h"
#include <ctype.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MINUS_INF (-9e9f) /* mx.h:12 */
#define MAXREPS 8         /* hspfinder.h:10 */
#define BADWORD 0xffffffffu
#define TB_DM 1 /* tracebit.h:4-7 */
#define TB_IM 2
#define TB_MD 4
#define TB_MI 8

/* ------------------------------------------------------------------ tables */
static uint8_t g_c2l[256];       /* alpha.cpp g_CharToLetterNucleo: ACGTU/acgtu -> 0..3, else 0xff */
static uint8_t g_match[256][256];/* alpha2.cpp:220-264 g_MatchMxNucleo */
static uint8_t g_comp[256];      /* alpha.cpp g_CharToCompChar ('?' = invalid -> keep char) */
static int g_tables_done;
/* current alphabet (set on every API entry from the params): nt or aa */
static uint8_t g_c2l_aa[256];      /* alpha.cpp g_CharToLetterAmino: ACDEFGHIKLMNPQRSTVWY both cases -> 0..19 */
static uint8_t g_match_aa[256][256];/* alpha2.cpp:220-279 g_MatchMxAmino */
static float g_blosum[256][256];   /* blosum62.cpp:17-96 */
static const uint8_t *g_cur_c2l = 0;
static unsigned g_cur_A = 4;
static uint8_t (*g_cur_match)[256] = 0;

static void init_tables(void)
	{
	if (g_tables_done)
		return;
	memset(g_c2l, 0xff, sizeof g_c2l);
	const char *nt = "ACGTU";
	for (int i = 0; i < 5; ++i)
		{
		uint8_t l = (uint8_t) (i == 4 ? 3 : i);
		g_c2l[(uint8_t) nt[i]] = l;
		g_c2l[(uint8_t) tolower(nt[i])] = l;
		}

	/* alpha2.cpp:94-150: single-letter bits and IUPAC code bits */
	uint8_t bit[256], bits[256];
	memset(bit, 0, sizeof bit);
	memset(bits, 0, sizeof bits);
	const char *plain = "AaCcGgTtUu";
	const uint8_t plainbit[] = {1, 1, 2, 2, 4, 4, 8, 8, 8, 8};
	for (int i = 0; i < 10; ++i)
		{
		bit[(uint8_t) plain[i]] = plainbit[i];
		bits[(uint8_t) plain[i]] = plainbit[i];
		}
	static const struct { char code; const char *chars; } codes[] = {
		{'M', "AC"}, {'R', "AG"}, {'W', "AT"}, {'S', "CG"}, {'Y', "CT"}, {'K', "GT"},
		{'V', "ACG"}, {'H', "ACT"}, {'D', "AGT"}, {'B', "CGT"}, {'X', "GATC"}, {'N', "GATC"}};
	for (unsigned k = 0; k < sizeof codes / sizeof codes[0]; ++k)
		{
		uint8_t b = 0;
		for (const char *p = codes[k].chars; *p; ++p)
			b |= bit[(uint8_t) *p];
		bits[(uint8_t) codes[k].code] = b;
		bits[(uint8_t) tolower(codes[k].code)] = b;
		}
	/* alpha2.cpp:220-264 */
	for (unsigned i = 0; i < 256; ++i)
		for (unsigned j = 0; j < 256; ++j)
			{
			int ai = isalpha((int) i) != 0, aj = isalpha((int) j) != 0;
			uint8_t m;
			if (!ai || !aj)
				{
				int gi = (i == '-' || i == '.'), gj = (j == '-' || j == '.');
				m = (uint8_t) (gi && gj);
				}
			else if (toupper((int) i) == toupper((int) j))
				m = 1;
			else
				m = (uint8_t) (((bit[i] & bits[j]) != 0) || ((bit[j] & bits[i]) != 0));
			g_match[i][j] = m;
			}

	/* complement table (alpha.cpp g_CharToCompChar); note lower-case 'u' is absent there */
	memset(g_comp, '?', sizeof g_comp);
	const char *from = "ABCDGHKMNRSTUVWXY";
	const char *to =   "TVGHCDMKNYSAABWXR";
	for (int i = 0; from[i]; ++i)
		{
		g_comp[(uint8_t) from[i]] = (uint8_t) to[i];
		if (from[i] != 'U')
			g_comp[(uint8_t) tolower(from[i])] = (uint8_t) tolower(to[i]);
		}
	g_comp[0] = 0;

	/* amino acid tables */
	memset(g_c2l_aa, 0xff, sizeof g_c2l_aa);
	const char *aa = "ACDEFGHIKLMNPQRSTVWY";
	for (int i = 0; i < 20; ++i)
		{
		g_c2l_aa[(uint8_t) aa[i]] = (uint8_t) i;
		g_c2l_aa[(uint8_t) tolower(aa[i])] = (uint8_t) i;
		}
	for (unsigned i = 0; i < 256; ++i)
		for (unsigned j = 0; j < 256; ++j)
			{
			int ai = isalpha((int) i) != 0, aj = isalpha((int) j) != 0;
			uint8_t m;
			if (!ai || !aj)
				m = (uint8_t) ((i == '-' || i == '.') && (j == '-' || j == '.'));
			else if (toupper((int) i) == toupper((int) j))
				m = 1;
			else
				m = (uint8_t) (toupper((int) i) == 'X' || toupper((int) j) == 'X');
			g_match_aa[i][j] = m;
			}
	g_match_aa['B']['N'] = g_match_aa['N']['B'] = g_match_aa['B']['D'] = g_match_aa['D']['B'] = 1;
	g_match_aa['Z']['Q'] = g_match_aa['Q']['Z'] = g_match_aa['Z']['E'] = g_match_aa['E']['Z'] = 1;
	/* BLOSUM62 in NCBI order, 1/2-bit units; both cases; everything else 0 (blosum62.cpp:48-80) */
	static const char *bl_alpha = "ARNDCQEGHILKMFPSTWYVBZX*";
	static const signed char bl[24][24] = {
		{ 4,-1,-2,-2, 0,-1,-1, 0,-2,-1,-1,-1,-1,-2,-1, 1, 0,-3,-2, 0,-2,-1, 0,-4},
		{-1, 5, 0,-2,-3, 1, 0,-2, 0,-3,-2, 2,-1,-3,-2,-1,-1,-3,-2,-3,-1, 0,-1,-4},
		{-2, 0, 6, 1,-3, 0, 0, 0, 1,-3,-3, 0,-2,-3,-2, 1, 0,-4,-2,-3, 3, 0,-1,-4},
		{-2,-2, 1, 6,-3, 0, 2,-1,-1,-3,-4,-1,-3,-3,-1, 0,-1,-4,-3,-3, 4, 1,-1,-4},
		{ 0,-3,-3,-3, 9,-3,-4,-3,-3,-1,-1,-3,-1,-2,-3,-1,-1,-2,-2,-1,-3,-3,-2,-4},
		{-1, 1, 0, 0,-3, 5, 2,-2, 0,-3,-2, 1, 0,-3,-1, 0,-1,-2,-1,-2, 0, 3,-1,-4},
		{-1, 0, 0, 2,-4, 2, 5,-2, 0,-3,-3, 1,-2,-3,-1, 0,-1,-3,-2,-2, 1, 4,-1,-4},
		{ 0,-2, 0,-1,-3,-2,-2, 6,-2,-4,-4,-2,-3,-3,-2, 0,-2,-2,-3,-3,-1,-2,-1,-4},
		{-2, 0, 1,-1,-3, 0, 0,-2, 8,-3,-3,-1,-2,-1,-2,-1,-2,-2, 2,-3, 0, 0,-1,-4},
		{-1,-3,-3,-3,-1,-3,-3,-4,-3, 4, 2,-3, 1, 0,-3,-2,-1,-3,-1, 3,-3,-3,-1,-4},
		{-1,-2,-3,-4,-1,-2,-3,-4,-3, 2, 4,-2, 2, 0,-3,-2,-1,-2,-1, 1,-4,-3,-1,-4},
		{-1, 2, 0,-1,-3, 1, 1,-2,-1,-3,-2, 5,-1,-3,-1, 0,-1,-3,-2,-2, 0, 1,-1,-4},
		{-1,-1,-2,-3,-1, 0,-2,-3,-2, 1, 2,-1, 5, 0,-2,-1,-1,-1,-1, 1,-3,-1,-1,-4},
		{-2,-3,-3,-3,-2,-3,-3,-3,-1, 0, 0,-3, 0, 6,-4,-2,-2, 1, 3,-1,-3,-3,-1,-4},
		{-1,-2,-2,-1,-3,-1,-1,-2,-2,-3,-3,-1,-2,-4, 7,-1,-1,-4,-3,-2,-2,-1,-2,-4},
		{ 1,-1, 1, 0,-1, 0, 0, 0,-1,-2,-2, 0,-1,-2,-1, 4, 1,-3,-2,-2, 0, 0, 0,-4},
		{ 0,-1, 0,-1,-1,-1,-1,-2,-2,-1,-1,-1,-1,-2,-1, 1, 5,-2,-2, 0,-1,-1, 0,-4},
		{-3,-3,-4,-4,-2,-2,-3,-2,-2,-3,-2,-3,-1, 1,-4,-3,-2,11, 2,-3,-4,-3,-2,-4},
		{-2,-2,-2,-3,-2,-1,-2,-3, 2,-1,-1,-2,-1, 3,-3,-2,-2, 2, 7,-1,-3,-2,-1,-4},
		{ 0,-3,-3,-3,-1,-2,-2,-3,-3, 3, 1,-2, 1,-1,-2,-2, 0,-3,-1, 4,-3,-2,-1,-4},
		{-2,-1, 3, 4,-3, 0, 1,-1, 0,-3,-4, 0,-3,-3,-2, 0,-1,-4,-3,-3, 4, 1,-1,-4},
		{-1, 0, 0, 1,-3, 3, 4,-2, 0,-3,-3, 1,-1,-3,-1, 0,-1,-3,-2,-2, 1, 4,-1,-4},
		{ 0,-1,-1,-1,-2,-1,-1,-1,-1,-1,-1,-1,-1,-1,-2, 0, 0,-2,-1,-1,-1,-1,-1,-4},
		{-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4, 1}};
	memset(g_blosum, 0, sizeof g_blosum);
	for (int i = 0; i < 24; ++i)
		for (int j = 0; j < 24; ++j)
			{
			uint8_t ui = (uint8_t) bl_alpha[i], uj = (uint8_t) bl_alpha[j];
			uint8_t li = (uint8_t) tolower(ui), lj = (uint8_t) tolower(uj);
			float v = (float) bl[i][j];
			g_blosum[ui][uj] = v; g_blosum[ui][lj] = v; g_blosum[li][uj] = v; g_blosum[li][lj] = v;
			}
	g_cur_c2l = g_c2l;
	g_cur_match = g_match;
	g_tables_done = 1;
	}

/* selects the alphabet the UDB words and identity counts use (udbparams.cpp:235-261) */
static void use_alpha(const uso_params *p)
	{
	init_tables();
	if (p->is_nucleo)
		{
		g_cur_c2l = g_c2l; g_cur_A = 4; g_cur_match = g_match;
		}
	else
		{
		g_cur_c2l = g_c2l_aa; g_cur_A = 20; g_cur_match = g_match_aa;
		}
	}

static unsigned slot_count_of(const uso_params *p)
	{
	unsigned A = p->is_nucleo ? 4 : 20, n = 1;
	for (unsigned i = 0; i < p->word_length; ++i)
		n *= A;
	return n;
	}

void uso_default_params(uso_params *p, int cluster_fast)
	{
	memset(p, 0, sizeof *p);
	p->is_nucleo = 1;
	p->id = 0.97f;
	p->maxaccepts = 1;                      /* terminator.cpp:10-31 */
	p->maxrejects = cluster_fast ? 8 : 32;
	p->strand_both = 0;
	p->word_length = 8;                     /* udbparams.cpp:246-250 */
	p->big = 100000;                        /* o_defaults.inc */
	p->bump = 50;
	p->stepwords = 8;
	p->band = 16;
	p->minhsp = 16;
	p->hspw = 5;                            /* alnheuristics.cpp:37 */
	p->xdrop_nw = 8.0f;
	p->match = 1.0f;
	p->mismatch = -2.0f;
	p->dbmask_fast = 1;
	p->cluster_mode = cluster_fast;
	p->evalue = 10.0f;
	p->xdrop_u = 16.0f;
	p->xdrop_g = 32.0f;
	p->lopen = -10.0f;
	p->lext = -1.0f;
	p->ka_dbsize = 1e9f;
	}

/* amino acid defaults: UDB words of 5 letters over the 20-letter alphabet (udbparams.cpp:236-261),
 * LocalAligner2 words of 3 letters (makedbsearcher.cpp:116-120) */
void uso_set_amino(uso_params *p)
	{
	p->is_nucleo = 0;
	p->word_length = 5;
	p->hspw = 3;
	}

/* ------------------------------------------------------------------ small helpers */
static void *xrealloc(void *p, size_t n)
	{
	void *q = realloc(p, n ? n : 1);
	if (!q)
		{
		fprintf(stderr, "uso: out of memory\n");
		abort();
		}
	return q;
	}

/* seqinfo.cpp:292-325 */
void uso_revcomp(const uint8_t *seq, uint32_t L, uint8_t *out)
	{
	init_tables();
	for (uint32_t i = 0; i < L; ++i)
		{
		uint8_t c = seq[i];
		uint8_t rc = g_comp[c];
		if (rc == '?')
			rc = c;
		out[L - i - 1] = rc;
		}
	}

/* fastmask.cpp:88-158 (soft masking only; -hardmask not restated) */
void uso_fastmask(const uint8_t *seq, uint32_t L, uint8_t *out)
	{
	uint8_t *tmp = (uint8_t *) xrealloc(0, L + 1); /* allow in-place (seqdb.cpp:448 masks in place) */
	for (uint32_t i = 0; i < L; ++i)
		tmp[i] = (uint8_t) toupper(seq[i]);
	if (L >= 2)
		{
		uint8_t lastc = '?';
		unsigned start = UINT_MAX;
		for (unsigned i = 0; i < L; ++i)
			{
			uint8_t c = (uint8_t) toupper(seq[i]);
			if (c != lastc || i + 1 == L)
				{
				unsigned n1 = i - start;
				if (n1 >= 5)
					for (unsigned j = start + 2; j < i; ++j)
						tmp[j] = (uint8_t) tolower(tmp[j]);
				start = i;
				}
			lastc = c;
			}
		for (unsigned startpos = 0; startpos <= 1; ++startpos)
			{
			unsigned lastpair = UINT_MAX;
			unsigned start2 = UINT_MAX;
			for (unsigned i = startpos; i < L - 1; i += 2)
				{
				uint8_t c1 = (uint8_t) toupper(seq[i]);
				uint8_t c2 = (uint8_t) toupper(seq[i + 1]);
				unsigned pair = ((unsigned) c1 << 8) + c2;
				if (pair != lastpair)
					{
					unsigned n2 = i - start2;
					if (n2 >= 5)
						for (unsigned j = start2 + 2; j < i; ++j)
							tmp[j] = (uint8_t) tolower(tmp[j]);
					start2 = i;
					}
				lastpair = pair;
				}
			}
		}
	memcpy(out, tmp, L);
	free(tmp);
	}

/* comppath.cpp:7-48 */
void uso_compress_path(const char *path, char *out)
	{
	char *p = out;
	if (path[0] == 0)
		{
		*p = 0;
		return;
		}
	char last = path[0];
	unsigned n = 1;
	for (unsigned i = 1;; ++i)
		{
		char c = path[i];
		if (c == last && c != 0)
			{
			++n;
			continue;
			}
		if (n == 1)
			*p++ = last;
		else
			p += sprintf(p, "%u%c", n, last);
		if (c == 0)
			break;
		last = c;
		n = 1;
		}
	*p = 0;
	}

/* ------------------------------------------------------------------ UDB words */
/* udbparams.cpp:540-555 SeqToWordNoPattern */
static uint32_t seq_to_word(const uint8_t *s, unsigned w)
	{
	uint32_t word = 0;
	for (unsigned i = 0; i < w; ++i)
		{
		uint8_t c = s[i];
		if (islower(c))
			return BADWORD;
		unsigned letter = g_cur_c2l[c];
		if (letter >= g_cur_A)
			return BADWORD;
		word = word * g_cur_A + letter;
		}
	return word;
	}

/* udbsearcher.cpp:128-151 + 161-194 (query) == udbparams.cpp:644-711 (target, step 1) */
static unsigned unique_words(const uint8_t *seq, uint32_t L, unsigned w, uint8_t *found, uint32_t *words,
  uint32_t *uwords)
	{
	if (L < w)
		return 0;
	unsigned end = L - w; /* GetLastValidWordPos udbparams.cpp:578-583 */
	unsigned n = 0;
	for (unsigned pos = 0; pos <= end; ++pos)
		{
		uint32_t word = seq_to_word(seq + pos, w);
		if (word != BADWORD)
			words[n++] = word;
		}
	unsigned nu = 0;
	for (unsigned i = 0; i < n; ++i)
		if (!found[words[i]])
			{
			uwords[nu++] = words[i];
			found[words[i]] = 1;
			}
	for (unsigned i = 0; i < n; ++i)
		found[words[i]] = 0;
	return nu;
	}

unsigned uso_query_unique_words(const uso_params *p, const uint8_t *q, uint32_t L, uint32_t *words)
	{
	use_alpha(p);
	unsigned slots = slot_count_of(p);
	uint8_t *found = (uint8_t *) calloc(slots, 1);
	uint32_t *all = (uint32_t *) xrealloc(0, (L + 1) * sizeof(uint32_t));
	unsigned n = unique_words(q, L, p->word_length, found, all, words);
	free(found);
	free(all);
	return n;
	}

/* ------------------------------------------------------------------ DB + index */
struct uso_db
	{
	uso_params P;
	uint32_t n, cap;
	uint8_t **seqs;
	uint32_t *lens;
	char **labels;
	uint32_t slot_count;
	uint32_t *sizes, *caps;
	uint32_t **rows;
	uint8_t *found;
	uint32_t *tw, *tuw;
	uint32_t twcap;
	};

uso_db *uso_db_create(const uso_params *p)
	{
	init_tables();
	uso_db *db = (uso_db *) calloc(1, sizeof *db);
	db->P = *p;
	db->slot_count = slot_count_of(p); /* udbparams.cpp:72-78: AlphaSize^w, non-hashed */
	db->sizes = (uint32_t *) calloc(db->slot_count, sizeof(uint32_t));
	db->caps = (uint32_t *) calloc(db->slot_count, sizeof(uint32_t));
	db->rows = (uint32_t **) calloc(db->slot_count, sizeof(uint32_t *));
	db->found = (uint8_t *) calloc(db->slot_count, 1);
	return db;
	}

/* udbbuild.cpp:286 AddSIToDB_CopyData -> :256-284 AddSeqNoncoded -> :111 AddWord / :74 GrowRow.
 * (FromSeqDB's two-pass prealloc, udbbuild.cpp:303-398, yields the same rows.) */
uint32_t uso_db_add(uso_db *db, const uint8_t *seq, uint32_t L, const char *label)
	{
	if (db->n == db->cap)
		{
		db->cap = db->cap ? db->cap * 2 : 1024;
		db->seqs = (uint8_t **) xrealloc(db->seqs, db->cap * sizeof(uint8_t *));
		db->lens = (uint32_t *) xrealloc(db->lens, db->cap * sizeof(uint32_t));
		db->labels = (char **) xrealloc(db->labels, db->cap * sizeof(char *));
		}
	use_alpha(&db->P);
	uint32_t idx = db->n++;
	uint8_t *s = (uint8_t *) xrealloc(0, L + 1);
	memcpy(s, seq, L);
	s[L] = 0;
	/* loaddb.cpp:117-118: only the LoadUDB path masks; cluster_fast indexes raw reads */
	if (db->P.dbmask_fast && !db->P.cluster_mode)
		uso_fastmask(s, L, s);
	db->seqs[idx] = s;
	db->lens[idx] = L;
	db->labels[idx] = strdup(label ? label : "");
	if (L + 1 > db->twcap)
		{
		db->twcap = L + 1024;
		db->tw = (uint32_t *) xrealloc(db->tw, db->twcap * sizeof(uint32_t));
		db->tuw = (uint32_t *) xrealloc(db->tuw, db->twcap * sizeof(uint32_t));
		}
	unsigned nu = unique_words(s, L, db->P.word_length, db->found, db->tw, db->tuw);
	for (unsigned i = 0; i < nu; ++i)
		{
		uint32_t w = db->tuw[i];
		if (db->sizes[w] == db->caps[w])
			{
			uint32_t c = db->caps[w];
			uint32_t nc = c == 0 ? 16 : c * 2;
			db->rows[w] = (uint32_t *) xrealloc(db->rows[w], nc * sizeof(uint32_t));
			db->caps[w] = nc;
			}
		db->rows[w][db->sizes[w]++] = idx;
		}
	return idx;
	}

void uso_db_free(uso_db *db)
	{
	if (!db)
		return;
	for (uint32_t i = 0; i < db->n; ++i)
		{
		free(db->seqs[i]);
		free(db->labels[i]);
		}
	for (uint32_t w = 0; w < db->slot_count; ++w)
		free(db->rows[w]);
	free(db->seqs); free(db->lens); free(db->labels);
	free(db->sizes); free(db->caps); free(db->rows); free(db->found);
	free(db->tw); free(db->tuw);
	free(db);
	}

uint32_t uso_db_seq_count(const uso_db *db) { return db->n; }
const uint8_t *uso_db_seq(const uso_db *db, uint32_t i, uint32_t *L) { if (L) *L = db->lens[i]; return db->seqs[i]; }
const char *uso_db_label(const uso_db *db, uint32_t i) { return db->labels[i]; }
uint32_t uso_db_slot_count(const uso_db *db) { return db->slot_count; }
const uint32_t *uso_db_row(const uso_db *db, uint32_t word, uint32_t *size) { *size = db->sizes[word]; return db->rows[word]; }

/* ------------------------------------------------------------------ searcher */
typedef struct hsp { unsigned Loi, Loj, Len; float Score; } hsp;

struct uso_searcher
	{
	uso_db *db;
	uso_params P;
	float (*subst)[256]; /* setnucmx.cpp:11-99 */
	/* AlnHeuristics (alnheuristics.cpp:26-69) */
	float XDropGlobalHSP, MinGlobalHSPScore, MinGlobalHSPFractId;
	unsigned MinGlobalHSPLength, BandRadius;
	/* AlnParams global (alnparams.cpp:378-384): Init4(mx,-10,-1,-0.5,-0.5) */
	float Open, Ext, TermOpen, TermExt;
	/* UDB searcher scratch */
	uint8_t *found;
	uint32_t *qw, *quw; uint32_t qwcap;
	uint32_t *U; uint32_t Ucap; int big;
	uint32_t *TopU, *TopT, *TopOrder, *TopT2; uint32_t topcap; uint32_t ntop_prev;
	uint32_t *cs_sizes, *cs_offsets; uint32_t cscap;
	/* HSP finder */
	unsigned hsp_w, hsp_A, hsp_wordcount, hsp_hi;
	uint32_t *wordsA, *wordsB; uint32_t wAcap, wBcap;
	unsigned nwordsA, nwordsB;
	uint32_t *word2posA, *wordcountsA;
	hsp *ung; unsigned nung, ungcap;
	unsigned *chain; unsigned nchain;
	/* DP */
	float *Mrow, *Drow; uint32_t rowcap;
	uint8_t *TB; size_t tbcap;
	char *subpath; uint32_t subcap;
	/* terminator */
	unsigned acc, rej;
	/* EStats (estats.cpp) */
	double es_gl, es_ul, es_gk, es_uk, es_loggk, es_loguk, es_dbsize, es_maxevalue;
	uint64_t es_letters;
	/* LocalAligner2 (localaligner2.cpp) */
	unsigned la_w, la_A, la_dict, la_hi, la_nq;
	uint32_t *la_qcounts, *la_qcounts2, *la_base, *la_qwords, *la_qposvec, *la_twords;
	uint32_t la_qcap, la_tcap;
	float la_min_ungapped;
	/* X-drop DP (xdpmem.h) */
	float *xd_M, *xd_D; uint32_t xd_rowcap;
	uint8_t *xd_TB; size_t xd_tbcap;
	uint8_t *xd_revA, *xd_revB; uint32_t xd_seqcap;
	int xd_poison;
	};

static void set_nuc_subst(float (*mx)[256], float match, float mismatch)
	{
	/* setnucmx.cpp:33-87: ACGTU x ACGTU (both cases) match/mismatch; N and everything else 0 */
	memset(mx, 0, 256 * 256 * sizeof(float));
	const char *al = "ACGTU";
	for (int i = 0; i < 5; ++i)
		for (int j = 0; j < 5; ++j)
			{
			float v = (g_c2l[(uint8_t) al[i]] == g_c2l[(uint8_t) al[j]]) ? match : mismatch;
			uint8_t ui = (uint8_t) al[i], uj = (uint8_t) al[j];
			uint8_t li = (uint8_t) tolower(ui), lj = (uint8_t) tolower(uj);
			mx[ui][uj] = v; mx[ui][lj] = v; mx[li][uj] = v; mx[li][lj] = v;
			}
	}

static void la_init(uso_searcher *s);

uso_searcher *uso_searcher_create(uso_db *db, const uso_params *p)
	{
	init_tables();
	uso_searcher *s = (uso_searcher *) calloc(1, sizeof *s);
	s->db = db;
	s->P = *p;
	s->subst = (float (*)[256]) xrealloc(0, 256 * 256 * sizeof(float));
	if (p->is_nucleo)
		set_nuc_subst(s->subst, p->match, p->mismatch);
	else
		memcpy(s->subst, g_blosum, 256 * 256 * sizeof(float)); /* alnparams.cpp:330-349: BLOSUM62 */
	s->xd_poison = getenv("USO_XD_POISON") != 0;
	la_init(s);
	/* alnheuristics.cpp:26-61 */
	s->XDropGlobalHSP = p->xdrop_nw;
	s->BandRadius = p->band;
	s->MinGlobalHSPLength = p->minhsp;
	if (p->is_nucleo)
		{
		s->MinGlobalHSPFractId = p->id > 0.75f ? p->id : 0.75f;
		s->MinGlobalHSPScore = s->MinGlobalHSPFractId * s->MinGlobalHSPLength * p->match;
		s->Open = -10.0f;
		}
	else
		{
		/* alnheuristics.cpp:40-58: smallest diagonal score of the matrix over the 20 letters */
		float MinDiagScore = 9e9f;
		for (unsigned i = 0; i < 20; ++i)
			{
			uint8_t c = (uint8_t) "ACDEFGHIKLMNPQRSTVWY"[i];
			if (s->subst[c][c] < MinDiagScore)
				MinDiagScore = s->subst[c][c];
			}
		s->MinGlobalHSPFractId = p->id > 0.5f ? p->id : 0.5f;
		s->MinGlobalHSPScore = s->MinGlobalHSPFractId * MinDiagScore * s->MinGlobalHSPLength;
		s->Open = -17.0f; /* alnparams.cpp:381-384 */
		}
	s->Ext = -1.0f; s->TermOpen = -0.5f; s->TermExt = -0.5f;
	s->found = (uint8_t *) calloc(db->slot_count, 1);
	/* hspfinder.cpp:193-217 */
	s->hsp_w = (p->is_nucleo || !p->local) ? p->hspw : 0;
	s->hsp_A = p->is_nucleo ? 4 : 20;
	s->hsp_wordcount = 1;
	for (unsigned i = 0; i < s->hsp_w; ++i)
		s->hsp_wordcount *= s->hsp_A;
	s->hsp_hi = s->hsp_wordcount / s->hsp_A;
	s->word2posA = (uint32_t *) xrealloc(0, s->hsp_wordcount * MAXREPS * sizeof(uint32_t));
	s->wordcountsA = (uint32_t *) xrealloc(0, s->hsp_wordcount * sizeof(uint32_t));
	return s;
	}

void uso_searcher_free(uso_searcher *s)
	{
	if (!s)
		return;
	free(s->subst); free(s->found); free(s->qw); free(s->quw); free(s->U);
	free(s->TopU); free(s->TopT); free(s->TopOrder); free(s->TopT2);
	free(s->cs_sizes); free(s->cs_offsets);
	free(s->wordsA); free(s->wordsB); free(s->word2posA); free(s->wordcountsA);
	free(s->ung); free(s->chain); free(s->Mrow); free(s->Drow); free(s->TB); free(s->subpath);
	free(s->la_qcounts); free(s->la_qcounts2); free(s->la_base); free(s->la_qwords); free(s->la_qposvec);
	free(s->la_twords); free(s->xd_M); free(s->xd_D); free(s->xd_TB); free(s->xd_revA); free(s->xd_revB);
	free(s);
	}

/* ------------------------------------------------------------------ U-sort ranking */
static void cs_alloc(uso_searcher *s, unsigned n)
	{
	if (n > s->cscap)
		{
		s->cscap = n + 256;
		s->cs_sizes = (uint32_t *) xrealloc(s->cs_sizes, s->cscap * sizeof(uint32_t));
		s->cs_offsets = (uint32_t *) xrealloc(s->cs_offsets, s->cscap * sizeof(uint32_t));
		}
	}

/* countsort.cpp:6-108 CountSortOrderDesc */
static unsigned count_sort_order_desc(uso_searcher *s, const uint32_t *values, unsigned n, uint32_t *order)
	{
	unsigned maxv = 0, nextv = 0;
	for (unsigned i = 0; i < n; ++i)
		if (values[i] > maxv)
			{
			nextv = maxv;
			maxv = values[i];
			}
	unsigned minv = nextv / 2;
	cs_alloc(s, maxv + 1);
	uint32_t *sizes = s->cs_sizes, *offsets = s->cs_offsets;
	memset(sizes, 0, (maxv + 1) * sizeof(uint32_t));
	for (unsigned i = 0; i < n; ++i)
		if (values[i] >= minv)
			++sizes[values[i]];
	unsigned off = 0;
	for (int v = (int) maxv; v >= (int) minv; --v)
		{
		offsets[v] = off;
		off += sizes[v];
		}
	for (unsigned i = 0; i < n; ++i)
		if (values[i] >= minv)
			order[offsets[values[i]]++] = i;
	return offsets[minv];
	}

/* countsort.cpp:110-191 CountSortSubsetDesc */
static unsigned count_sort_subset_desc(uso_searcher *s, const uint32_t *values, unsigned n,
  const uint32_t *subset, uint32_t *result)
	{
	unsigned maxv = 0, nextv = 0;
	for (unsigned i = 0; i < n; ++i)
		{
		unsigned v = values[subset[i]];
		if (v > maxv)
			{
			nextv = maxv;
			maxv = v;
			}
		}
	unsigned minv = nextv / 2;
	cs_alloc(s, maxv + 1);
	uint32_t *sizes = s->cs_sizes, *offsets = s->cs_offsets;
	memset(sizes, 0, (maxv + 1) * sizeof(uint32_t));
	for (unsigned i = 0; i < n; ++i)
		{
		unsigned v = values[subset[i]];
		if (v >= minv)
			++sizes[v];
		}
	unsigned off = 0;
	for (int v = (int) maxv; v >= (int) minv; --v)
		{
		offsets[v] = off;
		off += sizes[v];
		}
	for (unsigned i = 0; i < n; ++i)
		{
		unsigned k = subset[i];
		unsigned v = values[k];
		if (v >= minv)
			result[offsets[v]++] = k;
		}
	return offsets[minv];
	}

static void alloc_query(uso_searcher *s, uint32_t L)
	{
	if (L + 1 > s->qwcap)
		{
		s->qwcap = L + 1024;
		s->qw = (uint32_t *) xrealloc(s->qw, s->qwcap * sizeof(uint32_t));
		s->quw = (uint32_t *) xrealloc(s->quw, s->qwcap * sizeof(uint32_t));
		}
	}

static void alloc_top(uso_searcher *s, uint32_t n)
	{
	if (n > s->topcap)
		{
		s->topcap = n + 65536;
		s->TopU = (uint32_t *) xrealloc(s->TopU, s->topcap * sizeof(uint32_t));
		s->TopT = (uint32_t *) xrealloc(s->TopT, s->topcap * sizeof(uint32_t));
		s->TopOrder = (uint32_t *) xrealloc(s->TopOrder, s->topcap * sizeof(uint32_t));
		s->TopT2 = (uint32_t *) xrealloc(s->TopT2, s->topcap * sizeof(uint32_t));
		}
	}

/* udbusortedsearcher.cpp:109-120 SetTargetOrder = words, unique words, SetU(1), SetTop(1), SortTop.
 * Returns TopOrder.Size; candidate k is TopT[TopOrder[k]]. */
static unsigned set_target_order(uso_searcher *s, const uint8_t *q, uint32_t L)
	{
	uso_db *db = s->db;
	const unsigned N = db->n;
	alloc_query(s, L);
	unsigned nu = unique_words(q, L, s->P.word_length, s->found, s->qw, s->quw);
	/* SetU_NonCoded udbusortedsearcher.cpp:375-410 */
	if (N > s->Ucap)
		{
		s->Ucap = N + 65536;
		s->U = (uint32_t *) xrealloc(s->U, s->Ucap * sizeof(uint32_t));
		}
	if (N == 0)
		return 0;
	uint32_t *U = s->U;
	memset(U, 0, N * sizeof(uint32_t));
	for (unsigned i = 0; i < nu; ++i)
		{
		uint32_t w = s->quw[i];
		const uint32_t *row = db->rows[w];
		unsigned size = db->sizes[w];
		for (unsigned j = 0; j < size; ++j)
			++U[row[j]];
		}
	/* SetTop udbusortedsearcher.cpp:269-282 */
	alloc_top(s, N);
	unsigned MinU = 1;
	unsigned top = 0;
	if (s->P.bump != 0)
		{
		/* SetTopBump :230-267 */
		double Bump = s->P.bump / 100.0;
		unsigned MaxCount = 0;
		for (unsigned t = 0; t < N; ++t)
			{
			unsigned n = U[t];
			if (n >= MinU)
				{
				if (n > MaxCount)
					{
					unsigned NewMin = (unsigned) (n * Bump);
					if (NewMin > MinU && NewMin < MaxCount)
						MinU = NewMin;
					MaxCount = n;
					}
				s->TopU[top] = n;
				s->TopT[top] = t;
				++top;
				}
			}
		}
	else
		{
		/* SetTopNoBump :205-228 */
		for (unsigned t = 0; t < N; ++t)
			if (U[t] >= MinU)
				{
				s->TopU[top] = U[t];
				s->TopT[top] = t;
				++top;
				}
		}
	/* SortTop -> CountSortTop :154-162 (-quicksort not restated) */
	return count_sort_order_desc(s, s->TopU, top, s->TopOrder);
	}

unsigned uso_rank_candidates(uso_searcher *s, const uint8_t *q, uint32_t L, uint32_t *U,
  uint32_t *cand_t, uint32_t *cand_u)
	{
	unsigned k = set_target_order(s, q, L);
	if (U)
		memcpy(U, s->U, s->db->n * sizeof(uint32_t));
	for (unsigned i = 0; i < k; ++i)
		{
		unsigned o = s->TopOrder[i];
		if (cand_t) cand_t[i] = s->TopT[o];
		if (cand_u) cand_u[i] = s->TopU[o];
		}
	return k;
	}

/* ------------------------------------------------------------------ HSP finder */
/* hspfinder.cpp:226-270 SeqToWords: rolling words, wildcard -> letter 0, never skipped */
static unsigned hsp_seq_to_words(const uso_searcher *s, const uint8_t *seq, unsigned L, uint32_t *words)
	{
	const unsigned w = s->hsp_w, A = s->hsp_A;
	const uint8_t *c2l = s->P.is_nucleo ? g_c2l : g_c2l_aa; /* hspfinder.cpp:201 m_CharToLetter */
	if (L < w)
		return 0;
	uint32_t word = 0;
	const uint8_t *front = seq, *back = seq;
	for (unsigned i = 0; i < w - 1; ++i)
		{
		unsigned l = c2l[*front++];
		if (l >= A) l = 0;
		word = word * A + l;
		}
	for (unsigned i = w - 1; i < L; ++i)
		{
		unsigned l = c2l[*front++];
		if (l >= A) l = 0;
		word = word * A + l;
		*words++ = word;
		l = c2l[*back++];
		if (l >= A) l = 0;
		word -= l * s->hsp_hi;
		}
	return L - w + 1;
	}

/* hspfinder.cpp:304-323 SetA */
static void hsp_set_a(uso_searcher *s, const uint8_t *A, unsigned LA)
	{
	if (LA + 1 > s->wAcap)
		{
		s->wAcap = LA + 512;
		s->wordsA = (uint32_t *) xrealloc(s->wordsA, s->wAcap * sizeof(uint32_t));
		}
	memset(s->wordcountsA, 0, s->hsp_wordcount * sizeof(uint32_t));
	s->nwordsA = hsp_seq_to_words(s, A, LA, s->wordsA);
	for (unsigned pos = 0; pos < s->nwordsA; ++pos)
		{
		unsigned word = s->wordsA[pos];
		unsigned n = s->wordcountsA[word];
		if (n == MAXREPS)
			continue;
		s->word2posA[word * MAXREPS + n] = pos;
		++s->wordcountsA[word];
		}
	}

/* hspfinder.cpp:325-331 SetB */
static void hsp_set_b(uso_searcher *s, const uint8_t *B, unsigned LB)
	{
	if (LB + 1 > s->wBcap)
		{
		s->wBcap = LB + 512;
		s->wordsB = (uint32_t *) xrealloc(s->wordsB, s->wBcap * sizeof(uint32_t));
		}
	s->nwordsB = hsp_seq_to_words(s, B, LB, s->wordsB);
	}

/* hspfinder.cpp:594-636 IsGlobalHSP */
static int is_global_hsp(unsigned ALo, unsigned BLo, unsigned LA, unsigned LB)
	{
	if (LA <= LB)
		{
		unsigned MaxGap = LA / 4 + 1;
		if (ALo > BLo && ALo - BLo > MaxGap)
			return 0;
		unsigned AR = LA - ALo, BR = LB - BLo;
		if (AR > BR && AR - BR > MaxGap)
			return 0;
		}
	else
		{
		unsigned MaxGap = LB / 4 + 1;
		if (BLo > ALo && BLo - ALo > MaxGap)
			return 0;
		unsigned AR = LA - ALo, BR = LB - BLo;
		if (BR > AR && BR - AR > MaxGap)
			return 0;
		}
	return 1;
	}

static void ung_push(uso_searcher *s, unsigned Loi, unsigned Loj, unsigned Len, float Score)
	{
	if (s->nung == s->ungcap)
		{
		s->ungcap = s->ungcap ? s->ungcap * 2 : 64;
		s->ung = (hsp *) xrealloc(s->ung, s->ungcap * sizeof(hsp));
		s->chain = (unsigned *) xrealloc(s->chain, s->ungcap * sizeof(unsigned));
		}
	hsp *h = &s->ung[s->nung++];
	h->Loi = Loi; h->Loj = Loj; h->Len = Len; h->Score = Score;
	}

/* ungappedblast.cpp:8-211 (StaggerOk=false on this path) */
static void ungapped_blast(uso_searcher *s, const uint8_t *A, unsigned LA, const uint8_t *B, unsigned LB,
  float X, unsigned MinLength, float MinScore)
	{
	s->nung = 0;
	const unsigned w = s->hsp_w;
	if (LB < 2 * w)
		return;
	float (*Mx)[256] = s->subst;
	unsigned BPos = 0;
	for (;;)
		{
		if (BPos >= s->nwordsB)
			break;
		unsigned Word = s->wordsB[BPos];
		unsigned NA = s->wordcountsA[Word];
		if (NA == 0)
			{
			++BPos;
			continue;
			}
		int found = 0;
		for (unsigned i = 0; i < NA; ++i)
			{
			unsigned APos = s->word2posA[Word * MAXREPS + i];
			unsigned Diag = (LA + BPos) - APos;
			unsigned BPos2 = BPos + w - 1;
			unsigned APos2 = APos + w - 1;
			if (APos2 >= LA || BPos2 >= LB)
				continue;
			float Score = 0;
			for (unsigned j = 0; j < w; ++j)
				Score += Mx[A[APos + j]][B[BPos + j]];
			float BestScore = Score;
			unsigned BestBPos2 = BPos2;
			for (;;) /* extend right */
				{
				++BPos2;
				if (BPos2 >= LB)
					break;
				++APos2;
				if (APos2 >= LA)
					break;
				Score += Mx[A[APos2]][B[BPos2]];
				if (Score > BestScore)
					{
					BestScore = Score;
					BestBPos2 = BPos2;
					}
				else if (BestScore - Score > X)
					break;
				}
			unsigned APos1 = APos, BPos1 = BPos, BestBPos1 = BPos1;
			Score = BestScore;
			for (;;) /* extend left */
				{
				if (BPos1 == 0 || APos1 == 0)
					break;
				--BPos1;
				--APos1;
				Score += Mx[A[APos1]][B[BPos1]];
				if (Score > BestScore)
					{
					BestScore = Score;
					BestBPos1 = BPos1;
					}
				else if (BestScore - Score > X)
					break;
				}
			unsigned Blo = BestBPos1, Bhi = BestBPos2;
			unsigned Length = Bhi - Blo + 1;
			unsigned Alo = (LA + BestBPos1) - Diag;
			int Ok = (Length >= MinLength && BestScore >= MinScore);
			Ok = Ok && is_global_hsp(Alo, Blo, LA, LB);
			if (Ok)
				{
				ung_push(s, Alo, Blo, Length, BestScore);
				BPos = Bhi + 1;
				found = 1;
				break;
				}
			}
		if (!found)
			++BPos;
		}
	}

/* hsp.h:102-126 IsStaggered (three of four terms clamped, literally) */
static int is_staggered(const hsp *h, unsigned LA, unsigned LB)
	{
	int Hii = (int) (h->Loi + h->Len - 1), Hij = (int) (h->Loj + h->Len - 1);
	int TermGapLeftA = (int) h->Loi - (int) h->Loj;
	int TermGapLeftB = (int) h->Loj - (int) h->Loi;
	int TermGapRightA = (int) LA - Hii - 1 - ((int) LB - Hij - 1);
	int TermGapRightB = (int) LB - Hij - 1 - ((int) LA - Hii - 1);
	if (TermGapLeftA < 0) TermGapLeftA = 0;
	if (TermGapLeftB < 0) TermGapLeftB = 0;
	if (TermGapRightB < 0) TermGapRightB = 0;
	int GapA = TermGapLeftA + TermGapRightA;
	int GapB = TermGapLeftB + TermGapRightB;
	if (GapA == 0 || GapB == 0)
		return 0;
	double r = (LA < LB ? (double) GapA / LA : (double) GapB / LB);
	return r > 0.5;
	}

/* chainer.cpp:352-500.  The "delete dominated chains" branch (:447-448) compares a score with
 * itself and never fires, so only Lo break-points matter: HSPs are visited in (Loi, stable by
 * index) order -- qsort tie order pinned as stable, see SURVEY 8c -- and each takes the
 * best-scoring earlier-visited HSP with Hii < Loi and Hij < Loj (first wins ties, :337-338). */
static void chain_hsps(uso_searcher *s, unsigned LA, unsigned LB)
	{
	const unsigned n = s->nung;
	s->nchain = 0;
	if (n == 0)
		return;
	unsigned order[n];
	float cscore[n];
	unsigned prev[n];
	/* break-point sort: Pos ascending, Lo before Hi, stable.  Lo order = stable sort on Loi. */
	for (unsigned i = 0; i < n; ++i)
		order[i] = i;
	for (unsigned i = 1; i < n; ++i) /* insertion sort: stable */
		{
		unsigned k = order[i];
		unsigned j = i;
		while (j > 0 && s->ung[order[j - 1]].Loi > s->ung[k].Loi)
			{
			order[j] = order[j - 1];
			--j;
			}
		order[j] = k;
		}
	for (unsigned oi = 0; oi < n; ++oi)
		{
		unsigned k = order[oi];
		const hsp *h = &s->ung[k];
		float best = 0;
		unsigned bestc = UINT_MAX;
		for (unsigned oj = 0; oj < oi; ++oj) /* FindBestChainLT :322-350, list in push order */
			{
			unsigned c = order[oj];
			const hsp *ch = &s->ung[c];
			unsigned cAhi = ch->Loi + ch->Len - 1, cBhi = ch->Loj + ch->Len - 1;
			if (cAhi < h->Loi && cBhi < h->Loj && (bestc == UINT_MAX || cscore[c] > best))
				{
				bestc = c;
				best = cscore[c];
				}
			}
		prev[k] = bestc;
		cscore[k] = (bestc == UINT_MAX) ? h->Score : cscore[bestc] + h->Score;
		}
	unsigned opt = 0;
	float optscore = cscore[0];
	for (unsigned k = 1; k < n; ++k) /* :470-480 strict >, lowest index wins */
		if (cscore[k] > optscore)
			{
			opt = k;
			optscore = cscore[k];
			}
	unsigned len = 0;
	for (unsigned k = opt; k != UINT_MAX; k = prev[k])
		++len;
	unsigned i = 1;
	for (unsigned k = opt; k != UINT_MAX; k = prev[k])
		s->chain[len - i++] = k;
	s->nchain = len;
	/* hspfinder.cpp:537-553: whole chain dropped if any HSP is staggered */
	for (unsigned c = 0; c < s->nchain; ++c)
		if (is_staggered(&s->ung[s->chain[c]], LA, LB))
			{
			s->nchain = 0;
			return;
			}
	}

/* getglobalhsps.cpp:9-60 */
static unsigned get_global_hsps(uso_searcher *s, const uint8_t *A, unsigned LA, const uint8_t *B, unsigned LB,
  unsigned MinLength, float *HSPFractId)
	{
	ungapped_blast(s, A, LA, B, LB, s->XDropGlobalHSP, MinLength, s->MinGlobalHSPScore);
	chain_hsps(s, LA, LB);
	unsigned TotalLength = 0, TotalSame = 0;
	for (unsigned c = 0; c < s->nchain; ++c)
		{
		const hsp *h = &s->ung[s->chain[c]];
		TotalLength += h->Len;
		for (unsigned k = 0; k < h->Len; ++k) /* GetHSPIdCount hspfinder.cpp:561-579 */
			if ((s->P.is_nucleo ? g_match : g_match_aa)[A[h->Loi + k]][B[h->Loj + k]])
				++TotalSame;
		}
	*HSPFractId = TotalLength == 0 ? 0.0f : (float) TotalSame / (float) TotalLength;
	return s->nchain;
	}

/* ------------------------------------------------------------------ banded Viterbi */
typedef struct alnp { float OpenA, OpenB, ExtA, ExtB, LOpenA, LOpenB, LExtA, LExtB, ROpenA, ROpenB, RExtA, RExtB; } alnp;

static void dp_alloc(uso_searcher *s, unsigned LA, unsigned LB)
	{
	if (LB + 3 > s->rowcap)
		{
		s->rowcap = LB + 1024;
		s->Mrow = (float *) xrealloc(s->Mrow, (s->rowcap + 2) * sizeof(float));
		s->Drow = (float *) xrealloc(s->Drow, (s->rowcap + 2) * sizeof(float));
		}
	size_t need = (size_t) (LA + 1) * (LB + 1);
	if (need > s->tbcap)
		{
		s->tbcap = need + need / 4 + 4096;
		s->TB = (uint8_t *) xrealloc(s->TB, s->tbcap);
		}
	}

/* diagbox.h:150-171 GetRange_j */
static void get_range_j(unsigned LA, unsigned LB, unsigned dlo, unsigned dhi, unsigned i, unsigned *Startj,
  unsigned *Endj)
	{
	unsigned sj = (dlo + i >= LA) ? dlo + i - LA : 0;
	if (sj >= LB)
		sj = LB - 1;
	unsigned ej = (dhi + i + 1 >= LA) ? dhi + i + 1 - LA : 0;
	if (ej > LB)
		ej = LB;
	*Startj = sj;
	*Endj = ej;
	}

/* viterbifastbandmem.cpp:12-230 + tracebackbitmem.cpp:8-73.  TB is (LA+1) x (LB+1) bytes. */
static float viterbi_band(float (*Mx)[256], float *MrowBuf, float *Drow, uint8_t *TB, const uint8_t *A,
  unsigned LA, const uint8_t *B, unsigned LB, unsigned DiagLo, unsigned DiagHi, const alnp *AP, char *path)
	{
	float *Mrow = MrowBuf + 1; /* Mrow[-1] is addressable */
	const size_t W = (size_t) LB + 1;
	float OpenA = AP->LOpenA, ExtA = AP->LExtA;
	Mrow[-1] = MINUS_INF;
	for (unsigned j = 0; j <= LB; ++j)
		{
		Mrow[j] = MINUS_INF;
		Drow[j] = MINUS_INF;
		}
	for (unsigned i = 0; i < LA; ++i)
		{
		unsigned Startj, Endj;
		get_range_j(LA, LB, DiagLo, DiagHi, i, &Startj, &Endj);
		if (Endj == 0)
			continue;
		float OpenB = Startj == 0 ? AP->LOpenB : AP->OpenB;
		float ExtB = Startj == 0 ? AP->LExtB : AP->ExtB;
		const float *MxRow = Mx[A[i]];
		float I0 = MINUS_INF;
		float M0;
		if (i == 0)
			M0 = 0;
		else
			M0 = (Startj == 0) ? MINUS_INF : Mrow[(int) Startj - 1];
		uint8_t *TBrow = TB + (size_t) i * W;
		if (Startj > 0)
			TBrow[Startj - 1] = TB_IM;
		for (unsigned j = Startj; j < Endj; ++j)
			{
			uint8_t b = B[j];
			uint8_t bits = 0;
			float SavedM0 = M0;
			float xM = M0;
			if (Drow[j] > xM)
				{
				xM = Drow[j];
				bits = TB_DM;
				}
			if (I0 > xM)
				{
				xM = I0;
				bits = TB_IM;
				}
			M0 = Mrow[j];
			Mrow[j] = xM + MxRow[b];
			float md = SavedM0 + OpenB;
			Drow[j] += ExtB;
			if (md >= Drow[j])
				{
				Drow[j] = md;
				bits |= TB_MD;
				}
			float mi = SavedM0 + OpenA;
			I0 += ExtA;
			if (mi >= I0)
				{
				I0 = mi;
				bits |= TB_MI;
				}
			OpenB = AP->OpenB;
			ExtB = AP->ExtB;
			TBrow[j] = bits;
			}
		TBrow[LB] = 0;
		float md = M0 + AP->ROpenB;
		Drow[LB] += AP->RExtB;
		if (md >= Drow[LB])
			{
			Drow[LB] = md;
			TBrow[LB] = TB_MD;
			}
		M0 = MINUS_INF;
		OpenA = AP->OpenA;
		ExtA = AP->ExtA;
		}
	unsigned Startj, Endj;
	get_range_j(LA, LB, DiagLo, DiagHi, LA - 1, &Startj, &Endj);
	uint8_t *TBrow = TB + (size_t) LA * W;
	float I1 = MINUS_INF;
	Mrow[(int) Startj - 1] = MINUS_INF;
	for (unsigned j = Startj; j < Endj; ++j)
		{
		TBrow[j] = 0;
		float mi = Mrow[(int) j - 1] + AP->ROpenA;
		I1 += AP->RExtA;
		if (mi > I1)
			{
			I1 = mi;
			TBrow[j] = TB_MI;
			}
		}
	float FinalM = Mrow[LB - 1], FinalD = Drow[LB], FinalI = I1;
	float Score = FinalM;
	char State = 'M';
	if (FinalD > Score) { Score = FinalD; State = 'D'; }
	if (FinalI > Score) { Score = FinalI; State = 'I'; }

	/* tracebackbitmem.cpp:8-73 */
	unsigned n = 0;
	size_t i = LA, j = LB;
	for (;;)
		{
		if (i == 0 && j == 0)
			break;
		path[n++] = State;
		uint8_t t;
		if (State == 'M')
			{
			t = TB[(i - 1) * W + (j - 1)];
			if (t & TB_DM) State = 'D';
			else if (t & TB_IM) State = 'I';
			else State = 'M';
			--i; --j;
			}
		else if (State == 'D')
			{
			t = TB[(i - 1) * W + j];
			State = (t & TB_MD) ? 'M' : 'D';
			--i;
			}
		else
			{
			t = TB[i * W + (j - 1)];
			State = (t & TB_MI) ? 'M' : 'I';
			--j;
			}
		}
	for (unsigned k = 0; k < n / 2; ++k)
		{
		char c = path[k];
		path[k] = path[n - 1 - k];
		path[n - 1 - k] = c;
		}
	path[n] = 0;
	return Score;
	}

/* viterbifastbandmem.cpp:232-253 ViterbiFastMainDiagMem */
static float viterbi_main_diag(float (*Mx)[256], float *Mrow, float *Drow, uint8_t *TB, const uint8_t *A,
  unsigned LA, const uint8_t *B, unsigned LB, unsigned BandRadius, const alnp *AP, char *path)
	{
	unsigned DiagLo = LA < LB ? LA : LB;
	unsigned DiagHi = LA > LB ? LA : LB;
	if (BandRadius == 0)
		{
		/* -band 0: ViterbiFastMem (viterbifastmem.cpp:9-170) = the same recurrence over the full
		 * rectangle; a band that covers every diagonal reproduces it cell for cell */
		return viterbi_band(Mx, Mrow, Drow, TB, A, LA, B, LB, 1, LA + LB - 1, AP, path);
		}
	if (DiagLo > BandRadius)
		DiagLo -= BandRadius;
	else
		DiagLo = 1;
	DiagHi += BandRadius;
	unsigned MaxDiag = LA + LB - 1;
	if (DiagHi > MaxDiag)
		DiagHi = MaxDiag;
	return viterbi_band(Mx, Mrow, Drow, TB, A, LA, B, LB, DiagLo, DiagHi, AP, path);
	}

static void global_ap(const uso_searcher *s, alnp *AP)
	{
	AP->OpenA = AP->OpenB = s->Open;
	AP->ExtA = AP->ExtB = s->Ext;
	AP->LOpenA = AP->LOpenB = AP->ROpenA = AP->ROpenB = s->TermOpen;
	AP->LExtA = AP->LExtB = AP->RExtA = AP->RExtB = s->TermExt;
	}

/* alnparams.cpp:100-152 AlnParams::Init(AP, HSP, LA, LB) */
static void local_ap(const alnp *AP, int leftA, int leftB, int rightA, int rightB, alnp *L)
	{
	*L = *AP;
	if (!leftA)  { L->LOpenA = AP->OpenA; L->LExtA = AP->ExtA; }
	if (!leftB)  { L->LOpenB = AP->OpenB; L->LExtB = AP->ExtB; }
	if (!rightA) { L->ROpenA = AP->OpenA; L->RExtA = AP->ExtA; }
	if (!rightB) { L->ROpenB = AP->OpenB; L->RExtB = AP->ExtB; }
	}

float uso_viterbi_band(const uso_params *p, const uint8_t *A, uint32_t LA, const uint8_t *B, uint32_t LB,
  int left_a, int left_b, int right_a, int right_b, char *path)
	{
	init_tables();
	float (*mx)[256] = (float (*)[256]) xrealloc(0, 256 * 256 * sizeof(float));
	set_nuc_subst(mx, p->match, p->mismatch);
	float *Mrow = (float *) xrealloc(0, (LB + 4) * sizeof(float));
	float *Drow = (float *) xrealloc(0, (LB + 4) * sizeof(float));
	uint8_t *TB = (uint8_t *) xrealloc(0, (size_t) (LA + 1) * (LB + 1));
	alnp G, L;
	G.OpenA = G.OpenB = -10.0f; G.ExtA = G.ExtB = -1.0f;
	G.LOpenA = G.LOpenB = G.ROpenA = G.ROpenB = -0.5f;
	G.LExtA = G.LExtB = G.RExtA = G.RExtB = -0.5f;
	local_ap(&G, left_a, left_b, right_a, right_b, &L);
	float sc = viterbi_main_diag(mx, Mrow, Drow, TB, A, LA, B, LB, p->band, &L, path);
	free(mx); free(Mrow); free(Drow); free(TB);
	return sc;
	}

/* globalalignmem.cpp:70-112 AlignHSPMem: appends the hole's path at path+*n */
static void align_hole(uso_searcher *s, const uint8_t *A, unsigned LA, const uint8_t *B, unsigned LB,
  unsigned Loi, unsigned Loj, unsigned Leni, unsigned Lenj, char *path, unsigned *n)
	{
	if (Leni == 0)
		{
		for (unsigned k = 0; k < Lenj; ++k)
			path[(*n)++] = 'I';
		return;
		}
	if (Lenj == 0)
		{
		for (unsigned k = 0; k < Leni; ++k)
			path[(*n)++] = 'D';
		return;
		}
	alnp G, L;
	global_ap(s, &G);
	/* hsp.h:50-68: LeftA = Loi==0, LeftB = Loj==0, RightA = Loi+Leni==LA, RightB = Loj+Lenj==LB */
	local_ap(&G, Loi == 0, Loj == 0, Loi + Leni == LA, Loj + Lenj == LB, &L);
	dp_alloc(s, Leni, Lenj);
	if (Leni + Lenj + 1 > s->subcap)
		{
		s->subcap = Leni + Lenj + 1024;
		s->subpath = (char *) xrealloc(s->subpath, s->subcap);
		}
	viterbi_main_diag(s->subst, s->Mrow, s->Drow, s->TB, A + Loi, Leni, B + Loj, Lenj, s->BandRadius, &L,
	  s->subpath);
	size_t k = strlen(s->subpath);
	memcpy(path + *n, s->subpath, k);
	*n += (unsigned) k;
	}

/* globalalignmem.cpp:129-236 GlobalAlign_AllOpts (FullDPAlways=false, FailIfNoHSPs=true).
 * HSPFinder::SetA must already have been done for A (Aligner::SetQuery). */
static int global_align(uso_searcher *s, const uint8_t *A, unsigned LA, const uint8_t *B, unsigned LB, char *path)
	{
	unsigned MinHSPLength = (s->MinGlobalHSPLength == 0 ? 32 : s->MinGlobalHSPLength);
	if (MinHSPLength > LA / 4)
		MinHSPLength = LA / 4;
	if (MinHSPLength < 16)
		MinHSPLength = 16;
	if (s->P.fulldp)
		{
		/* globalalignmem.cpp:153-157: FullDPAlways */
		alnp G;
		global_ap(s, &G);
		dp_alloc(s, LA, LB);
		viterbi_main_diag(s->subst, s->Mrow, s->Drow, s->TB, A, LA, B, LB, 0, &G, path);
		return 1;
		}
	float HSPFractId;
	hsp_set_b(s, B, LB); /* Searcher::SetTarget -> GlobalAligner::SetTargetImpl globalaligner.cpp:69-74 */
	unsigned HSPCount = get_global_hsps(s, A, LA, B, LB, MinHSPLength, &HSPFractId);
	if (HSPFractId < s->MinGlobalHSPFractId)
		return 0;
	unsigned n = 0;
	if (HSPCount == 0)
		{
		if (s->MinGlobalHSPLength > 0 && LA > 64)
			return 0;
		alnp G;
		global_ap(s, &G);
		dp_alloc(s, LA, LB);
		viterbi_main_diag(s->subst, s->Mrow, s->Drow, s->TB, A, LA, B, LB, s->BandRadius, &G, path);
		return 1;
		}
	const hsp *prev = 0;
	for (unsigned c = 0; c < HSPCount; ++c)
		{
		const hsp *h = &s->ung[s->chain[c]];
		unsigned Loi, Loj, Leni, Lenj; /* GetHole :25-68 */
		if (prev)
			{
			Loi = prev->Loi + prev->Len;
			Loj = prev->Loj + prev->Len;
			Leni = h->Loi - Loi;
			Lenj = h->Loj - Loj;
			}
		else
			{
			Loi = 0; Loj = 0; Leni = h->Loi; Lenj = h->Loj;
			}
		align_hole(s, A, LA, B, LB, Loi, Loj, Leni, Lenj, path, &n);
		for (unsigned k = 0; k < h->Len; ++k)
			path[n++] = 'M';
		prev = h;
		}
	unsigned Loi = prev->Loi + prev->Len, Loj = prev->Loj + prev->Len;
	align_hole(s, A, LA, B, LB, Loi, Loj, LA - Loi, LB - Loj, path, &n);
	path[n] = 0;
	return 1;
	}

unsigned uso_global_hsps(uso_searcher *s, const uint8_t *q, uint32_t LQ, const uint8_t *t, uint32_t LT,
  uint32_t *ungapped, unsigned *n_ungapped, uint32_t *chained, unsigned max_hsp, float *hsp_fract_id)
	{
	hsp_set_a(s, q, LQ);
	hsp_set_b(s, t, LT);
	unsigned MinHSPLength = (s->MinGlobalHSPLength == 0 ? 32 : s->MinGlobalHSPLength);
	if (MinHSPLength > LQ / 4) MinHSPLength = LQ / 4;
	if (MinHSPLength < 16) MinHSPLength = 16;
	unsigned nc = get_global_hsps(s, q, LQ, t, LT, MinHSPLength, hsp_fract_id);
	*n_ungapped = s->nung;
	for (unsigned i = 0; i < s->nung && i < max_hsp; ++i)
		{
		ungapped[4 * i] = s->ung[i].Loi; ungapped[4 * i + 1] = s->ung[i].Loj;
		ungapped[4 * i + 2] = s->ung[i].Len; ungapped[4 * i + 3] = (uint32_t) (int) (s->ung[i].Score * 2);
		}
	for (unsigned i = 0; i < nc && i < max_hsp; ++i)
		{
		const hsp *h = &s->ung[s->chain[i]];
		chained[4 * i] = h->Loi; chained[4 * i + 1] = h->Loj;
		chained[4 * i + 2] = h->Len; chained[4 * i + 3] = (uint32_t) (int) (h->Score * 2);
		}
	return nc;
	}

int uso_global_align(uso_searcher *s, const uint8_t *q, uint32_t LQ, const uint8_t *t, uint32_t LT, char *path)
	{
	hsp_set_a(s, q, LQ);
	return global_align(s, q, LQ, t, LT, path);
	}

/* ------------------------------------------------------------------ AlignResult stats */
/* arscorer.cpp:201-296 FillLo + :554-570 GetGapOpenCount (global: m_HSP spans both sequences) */
/* ------------------------------------------------------------------ local alignment (config 5)
 * EStats (estats.cpp:25-101): Karlin-Altschul statistics in double; the DB size and the maximum
 * E-value reach the constructor as floats (makedbsearcher.cpp:90-98). */
static void estats_init(uso_searcher *s)
	{
	uso_db *db = s->db;
	uint64_t letters = 0;
	for (uint32_t i = 0; i < db->n; ++i)
		letters += db->lens[i];
	/* makedbsearcher.cpp:92-96: -ka_dbsize has a default (o_defaults.inc:2, 1e9) and defaults count
	 * as "filled", so the letter count of the DB is never used */
	float DBSize = s->P.ka_dbsize > 0.0f ? s->P.ka_dbsize : (float) letters;
	s->es_dbsize = (double) DBSize;
	s->es_maxevalue = (double) (float) s->P.evalue;
	if (s->P.is_nucleo)
		{
		s->es_gl = 1.280; s->es_ul = 1.330; s->es_gk = 0.460; s->es_uk = 0.621;
		}
	else
		{
		s->es_gl = 0.267; s->es_ul = 0.311; s->es_gk = 0.0410; s->es_uk = 0.128;
		}
	s->es_loggk = log(s->es_gk);
	s->es_loguk = log(s->es_uk);
	s->es_letters = letters;
	}

/* The three EStats expressions below are written the way the reference BINARY evaluates them:
 * its own build flags are -O3 -ffast-math (src/Makefile:11-14), under which gcc simplifies the
 * source algebraically (x/Log2 -> x*(1/Log2), NM/pow(2,B) -> NM*exp2(-B), BitScore*Log2 cancels,
 * one fused multiply-add).  Checked against the disassembly of oracle/_ref/usearch12; the forms
 * matter for the last printed digit and for E-values below 1e-300 (source form: 0, binary:
 * denormal arithmetic). */
#define INV_LOG2 (1.0 / log(2.0))

/* estats.cpp:65-71 */
static double es_min_ungapped_raw(const uso_searcher *s, unsigned QL)
	{
	return ((log((double) QL * s->es_dbsize) + s->es_loguk) - log(s->es_maxevalue)) / s->es_ul;
	}

/* estats.cpp:79-85 (gapped) */
static double es_raw_to_bits(const uso_searcher *s, double raw)
	{
	return fma(raw, s->es_gl, -s->es_loggk) * INV_LOG2;
	}

/* estats.cpp:73-96 */
static double es_raw_to_evalue(const uso_searcher *s, double raw, unsigned QL)
	{
	double x = (s->es_loggk - raw * s->es_gl) * INV_LOG2;
	double e = exp2(x);
	return (double) QL * (e * s->es_dbsize);
	}

static void la_refresh_estats(uso_searcher *s)
	{
	/* the reference builds g_ES once, after the DB is loaded; the oracle's DB may still be
	 * growing when the searcher is created, so recompute when the letter count changed */
	uint64_t letters = 0;
	for (uint32_t i = 0; i < s->db->n; ++i)
		letters += s->db->lens[i];
	if (letters != s->es_letters || s->es_gl == 0.0)
		estats_init(s);
	}

/* LocalAligner2::InitImpl (localaligner2.cpp:47-62) */
static void la_init(uso_searcher *s)
	{
	s->la_A = s->P.is_nucleo ? 4 : 20;
	s->la_w = s->P.hspw; /* makedbsearcher.cpp:104-123: 5 nt / 3 aa unless -hspw */
	s->la_dict = 1;
	for (unsigned i = 0; i < s->la_w; ++i)
		s->la_dict *= s->la_A;
	s->la_hi = s->la_dict / s->la_A;
	s->la_qcounts = (uint32_t *) calloc(s->la_dict, sizeof(uint32_t));
	s->la_qcounts2 = (uint32_t *) calloc(s->la_dict, sizeof(uint32_t));
	s->la_base = (uint32_t *) calloc(s->la_dict, sizeof(uint32_t));
	}

/* rolling words, wildcards -> letter 0, nothing skipped (localaligner2.cpp:85-117, localmulti.cpp:30-62) */
static unsigned la_words(const uso_searcher *s, const uint8_t *seq, unsigned L, uint32_t *words)
	{
	const uint8_t *c2l = s->P.is_nucleo ? g_c2l : g_c2l_aa;
	const unsigned A = s->la_A, w = s->la_w;
	uint32_t Word = 0;
	const uint8_t *Front = seq, *Back = seq;
	for (unsigned i = 0; i + 1 < w; ++i)
		{
		unsigned Letter = c2l[*Front++];
		if (Letter >= A)
			Letter = 0;
		Word = Word * A + Letter;
		}
	unsigned n = 0;
	for (unsigned pos = w - 1; pos < L; ++pos)
		{
		unsigned Letter = c2l[*Front++];
		if (Letter >= A)
			Letter = 0;
		Word = Word * A + Letter;
		words[n++] = Word;
		Letter = c2l[*Back++];
		if (Letter >= A)
			Letter = 0;
		Word -= Letter * s->la_hi;
		}
	return n;
	}

/* LocalAligner2::SetQueryImpl (localaligner2.cpp:64-155) + LocalAligner::SetQueryImpl (:213-217) */
static void la_set_query(uso_searcher *s, const uint8_t *Q, unsigned QL)
	{
	la_refresh_estats(s);
	s->la_min_ungapped = (float) es_min_ungapped_raw(s, QL);
	s->la_nq = 0;
	if (QL <= s->la_w)
		return;
	if (s->la_qcap < QL + 1)
		{
		s->la_qcap = QL + 1;
		s->la_qwords = (uint32_t *) xrealloc(s->la_qwords, s->la_qcap * sizeof(uint32_t));
		s->la_qposvec = (uint32_t *) xrealloc(s->la_qposvec, s->la_qcap * sizeof(uint32_t));
		}
	unsigned n = la_words(s, Q, QL, s->la_qwords);
	s->la_nq = n;
	for (unsigned i = 0; i < n; ++i)
		++s->la_qcounts2[s->la_qwords[i]];
	unsigned Base = 0;
	for (unsigned i = 0; i < n; ++i)
		{
		uint32_t Word = s->la_qwords[i];
		unsigned c = s->la_qcounts2[Word];
		if (c == 0)
			continue;
		s->la_base[Word] = Base;
		s->la_qcounts2[Word] = 0;
		Base += c;
		}
	for (unsigned i = 0; i < n; ++i)
		{
		uint32_t Word = s->la_qwords[i];
		unsigned c = s->la_qcounts[Word];
		s->la_qcounts[Word] = c + 1;
		s->la_qposvec[s->la_base[Word] + c] = i;
		}
	}

/* LocalAligner2::OnQueryDoneImpl (localaligner2.cpp:157-168) */
static void la_query_done(uso_searcher *s)
	{
	for (unsigned i = 0; i < s->la_nq; ++i)
		{
		s->la_qcounts[s->la_qwords[i]] = 0;
		s->la_qcounts2[s->la_qwords[i]] = 0;
		}
	s->la_nq = 0;
	}

static void xd_alloc(uso_searcher *s, unsigned LA, unsigned LB)
	{
	if (s->xd_rowcap < LB + 136)
		{
		s->xd_rowcap = LB + 136;
		s->xd_M = (float *) xrealloc(s->xd_M, s->xd_rowcap * sizeof(float));
		s->xd_D = (float *) xrealloc(s->xd_D, s->xd_rowcap * sizeof(float));
		for (unsigned i = 0; i < s->xd_rowcap; ++i)
			s->xd_M[i] = s->xd_D[i] = 0.0f;
		}
	size_t need = (size_t) (LA + 2) * (LB + 2);
	if (s->xd_tbcap < need)
		{
		s->xd_tbcap = need;
		s->xd_TB = (uint8_t *) xrealloc(s->xd_TB, need);
		}
	if (s->xd_seqcap < LA + LB + 2)
		{
		s->xd_seqcap = LA + LB + 2;
		s->xd_revA = (uint8_t *) xrealloc(s->xd_revA, s->xd_seqcap);
		s->xd_revB = (uint8_t *) xrealloc(s->xd_revB, s->xd_seqcap);
		}
	}

#define XD_UNWRITTEN 0x80 /* oracle-only poison: a traceback that touches a cell this call did not write aborts */

/* XDropFwdFastMem (xdropfwdmem.cpp:344-749) + XDropFwdTraceBackBitMem (:271-342).
 * path receives the forward path (NUL-terminated, capacity >= LA+LB+1). */
static float xdrop_fwd(uso_searcher *s, const uint8_t *A, unsigned LA, const uint8_t *B, unsigned LB, float X,
  unsigned *Leni, unsigned *Lenj, char *path)
	{
	float (*Mx)[256] = s->subst;
	if (LA == 1 || LB == 1)
		{
		*Leni = 1; *Lenj = 1;
		path[0] = 'M'; path[1] = 0;
		return Mx[A[0]][B[0]];
		}
	xd_alloc(s, LA, LB);
	const float Open = s->P.lopen, Ext = s->P.lext;
	const float AbsOpen = -Open, AbsExt = -Ext;
	const size_t stride = (size_t) LB + 2;
	uint8_t *TB = s->xd_TB;
	memset(TB, XD_UNWRITTEN, (size_t) (LA + 2) * stride);
	float *Mrow = s->xd_M + 1, *Drow = s->xd_D + 1;
	if (s->xd_poison)
		for (unsigned i = 0; i < s->xd_rowcap; ++i)
			s->xd_M[i] = s->xd_D[i] = 1e30f;
	Mrow[-1] = MINUS_INF;
	Drow[0] = MINUS_INF;
	Drow[1] = MINUS_INF;

	float BestScore = Mx[A[0]][B[0]];
	unsigned Besti = 0, Bestj = 0;
	unsigned prev_jlo = 0, prev_jhi = 0, jlo = 1, jhi = 1;
	float M0 = BestScore;
	for (unsigned i = 1; i < LA; ++i)
		{
		if (jlo == prev_jlo)
			{
			Mrow[jlo - 1] = MINUS_INF;
			Drow[jlo] = MINUS_INF;
			}
		unsigned endj = prev_jhi + 1 < LB ? prev_jhi + 1 : LB;
		for (unsigned j = endj + 1; j <= (jhi + 1 < LB ? jhi + 1 : LB); ++j)
			{
			Mrow[j - 1] = MINUS_INF;
			Drow[j] = MINUS_INF;
			}
		unsigned next_jlo = UINT_MAX, next_jhi = UINT_MAX;
		const float *MxRow = Mx[A[i]];
		float I0 = MINUS_INF;
		uint8_t *TBrow = TB + (size_t) i * stride;
		float SavedM0;
		for (unsigned j = jlo; j <= jhi; ++j)
			{
			uint8_t b = B[j];
			uint8_t TraceBits = 0;
			SavedM0 = M0;
			/* MATCH */
			{
			float xM = M0;
			if (Drow[j] > xM)
				{
				xM = Drow[j];
				TraceBits = TB_DM;
				}
			if (I0 > xM)
				{
				xM = I0;
				TraceBits = TB_IM;
				}
			M0 = Mrow[j];
			float sc = xM + MxRow[b];
			Mrow[j] = sc;
			float h = sc - BestScore + X;
			if (h > 0)
				{
				if (j + 1 < next_jlo) next_jlo = j + 1;
				next_jhi = j + 1;
				}
			if (h > AbsOpen)
				if (j < next_jlo) next_jlo = j;
			if (h > AbsExt && j == jhi && jhi + 1 < LB)
				{
				++jhi;
				unsigned new_endj = jhi + 1 < LB ? jhi + 1 : LB;
				if (new_endj < endj) new_endj = endj;
				for (unsigned j2 = endj + 1; j2 <= new_endj; ++j2)
					{
					if (j2 - 1 > j)
						Mrow[j2 - 1] = MINUS_INF;
					Drow[j2] = MINUS_INF;
					}
				endj = new_endj;
				}
			if (sc >= BestScore)
				{
				BestScore = sc;
				Besti = i;
				Bestj = j;
				}
			}
			/* DELETE */
			if (j != jlo)
				{
				float md = SavedM0 + Open;
				Drow[j] += Ext;
				if (md >= Drow[j])
					{
					Drow[j] = md;
					TraceBits |= TB_MD;
					}
				float h = Drow[j] - BestScore + X;
				if (h > 0)
					{
					if (j - 1 < next_jlo) next_jlo = j - 1;
					if (j - 1 > next_jhi) next_jhi = j - 1; /* max() with the UINT_MAX start value */
					}
				}
			/* INSERT */
			{
			float mi = SavedM0 + Open;
			I0 += Ext;
			if (mi >= I0)
				{
				I0 = mi;
				TraceBits |= TB_MI;
				}
			float h = I0 - BestScore + X;
			if (h > 0)
				{
				if (j + 1 < next_jlo) next_jlo = j + 1;
				next_jhi = j + 1;
				}
			if (h > AbsExt && j == jhi && jhi + 1 < LB)
				{
				++jhi;
				unsigned new_endj = jhi + 1 < LB ? jhi + 1 : LB;
				if (new_endj < endj) new_endj = endj;
				for (unsigned j2 = endj + 1; j2 <= new_endj; ++j2)
					{
					Mrow[j2 - 1] = MINUS_INF;
					Drow[j2] = MINUS_INF;
					}
				endj = new_endj;
				}
			}
			TBrow[j] = TraceBits;
			}
		/* special case for the end of Drow */
		if (jhi < LB)
			{
			const unsigned jhi1 = jhi + 1;
			TBrow[jhi1] = 0;
			float md = M0 + Open;
			Drow[jhi1] += Ext;
			if (md >= Drow[jhi1])
				{
				Drow[jhi1] = md;
				TBrow[jhi1] = TB_MD;
				}
			}
		if (next_jlo == UINT_MAX)
			break;
		prev_jlo = jlo;
		prev_jhi = jhi;
		jlo = next_jlo;
		jhi = next_jhi;
		if (jlo >= LB) jlo = LB - 1;
		if (jhi >= LB) jhi = LB - 1;
		if (jlo == prev_jlo)
			{
			M0 = MINUS_INF;
			Drow[jlo] = MINUS_INF;
			}
		else
			M0 = Mrow[jlo - 1];
		}
	if (BestScore <= 0.0f)
		{
		*Leni = 0; *Lenj = 0;
		path[0] = 0;
		return 0.0f;
		}
	/* traceback from (Besti,Bestj) in state M (xdropfwdmem.cpp:13-49,271-342) */
	unsigned i = Besti, j = Bestj, n = 0;
	char State = 'M';
	for (;;)
		{
		path[n++] = State;
		if (i == 0 && j == 0)
			break;
		char Next;
		uint8_t c;
		if (State == 'M')
			{
			c = TB[(size_t) i * stride + j];
			Next = (c & TB_DM) ? 'D' : (c & TB_IM) ? 'I' : 'M';
			--i; --j;
			}
		else if (State == 'D')
			{
			c = TB[(size_t) i * stride + j + 1];
			Next = (c & TB_MD) ? 'M' : 'D';
			--i;
			}
		else
			{
			c = TB[(size_t) (i + 1) * stride + j];
			Next = (c & TB_MI) ? 'M' : 'I';
			--j;
			}
		if (c == XD_UNWRITTEN)
			{
			fprintf(stderr, "oracle: X-drop traceback read an unwritten trace cell\n");
			abort();
			}
		State = Next;
		}
	for (unsigned k = 0; k < n / 2; ++k)
		{
		char t = path[k]; path[k] = path[n - 1 - k]; path[n - 1 - k] = t;
		}
	path[n] = 0;
	*Leni = Besti + 1;
	*Lenj = Bestj + 1;
	return BestScore;
	}

/* XDropBwdFastMem (xdropbwdmem.cpp:23-70) */
static float xdrop_bwd(uso_searcher *s, const uint8_t *A, unsigned LA, const uint8_t *B, unsigned LB, float X,
  unsigned *Leni, unsigned *Lenj, char *path)
	{
	xd_alloc(s, LA, LB);
	uint8_t *RevA = (uint8_t *) xrealloc(0, LA + 1), *RevB = (uint8_t *) xrealloc(0, LB + 1);
	for (unsigned i = 0; i < LA; ++i) RevA[i] = A[LA - i - 1];
	for (unsigned i = 0; i < LB; ++i) RevB[i] = B[LB - i - 1];
	float Score = xdrop_fwd(s, RevA, LA, RevB, LB, X, Leni, Lenj, path);
	free(RevA); free(RevB);
	if (Score <= 0.0)
		return Score;
	size_t n = strlen(path);
	for (size_t k = 0; k < n / 2; ++k)
		{
		char t = path[k]; path[k] = path[n - 1 - k]; path[n - 1 - k] = t;
		}
	return Score;
	}

#define XD_MAXL 4096 /* xdpmem.h:6 g_MaxL */

/* xdropfwdsplit.cpp:15-22 GetSubL */
static unsigned xd_sub_l(unsigned L)
	{
	if (L <= XD_MAXL)
		return L;
	if (L < 2 * XD_MAXL)
		return L / 2;
	return XD_MAXL;
	}

/* XDropFwdSplit (xdropfwdsplit.cpp:24-91): extensions of at most g_MaxL letters, one after the other */
static float xdrop_fwd_split(uso_searcher *s, const uint8_t *A, unsigned LA, const uint8_t *B, unsigned LB, float X,
  unsigned *Leni, unsigned *Lenj, char *path)
	{
	*Leni = 0;
	*Lenj = 0;
	path[0] = 0;
	size_t n = 0;
	char *sub = (char *) xrealloc(0, (size_t) 2 * XD_MAXL + 8);
	float SumScore = 0.0f;
	for (;;)
		{
		if (*Leni == LA || *Lenj == LB)
			break;
		unsigned SubLA = xd_sub_l(LA - *Leni), SubLB = xd_sub_l(LB - *Lenj);
		unsigned SubLeni, SubLenj;
		float Score = xdrop_fwd(s, A + *Leni, SubLA, B + *Lenj, SubLB, X, &SubLeni, &SubLenj, sub);
		if (Score == 0.0f)
			break;
		SumScore += Score;
		*Leni += SubLeni;
		*Lenj += SubLenj;
		size_t m = strlen(sub);
		memcpy(path + n, sub, m);
		n += m;
		path[n] = 0;
		if (SubLeni < SubLA && SubLenj < SubLB)
			break;
		}
	free(sub);
	return SumScore;
	}

/* XDropBwdSplit (xdropbwdsplit.cpp:15-79) */
static float xdrop_bwd_split(uso_searcher *s, const uint8_t *A, unsigned LA, const uint8_t *B, unsigned LB, float X,
  unsigned *Leni, unsigned *Lenj, char *path)
	{
	*Leni = 0;
	*Lenj = 0;
	path[0] = 0;
	size_t n = 0;
	char *sub = (char *) xrealloc(0, (size_t) 2 * XD_MAXL + 8);
	float SumScore = 0.0f;
	unsigned DoneA = 0, DoneB = 0;
	for (;;)
		{
		if (DoneA == LA || DoneB == LB)
			break;
		unsigned SubLA = xd_sub_l(LA - DoneA), SubLB = xd_sub_l(LB - DoneB);
		unsigned SubLeni, SubLenj;
		float Score = xdrop_bwd(s, A + LA - DoneA - SubLA, SubLA, B + LB - DoneB - SubLB, SubLB, X, &SubLeni, &SubLenj, sub);
		if (Score == 0.0f)
			break;
		SumScore += Score;
		*Leni += SubLeni;
		*Lenj += SubLenj;
		size_t m = strlen(sub);
		memmove(path + m, path, n + 1); /* PrependPath */
		memcpy(path, sub, m);
		n += m;
		if (SubLeni < SubLA && SubLenj < SubLB)
			break;
		DoneA += SubLeni;
		DoneB += SubLenj;
		}
	free(sub);
	return SumScore;
	}

/* XDropAlignMemMaxL2 (xdropalignmem.cpp:26-216).  h4 = {Loi, Loj, Leni, Lenj}. */
static float xdrop_align(uso_searcher *s, const uint8_t *A, unsigned LA, const uint8_t *B, unsigned LB,
  unsigned AncLoi, unsigned AncLoj, unsigned AncLen, float X, uint32_t *h4, char *path)
	{
	path[0] = 0;
	if (AncLen <= 1)
		return 0.0f;
	unsigned AncHii = AncLoi + AncLen - 1, AncHij = AncLoj + AncLen - 1;
	const uint8_t *FwdA = A + AncHii, *FwdB = B + AncHij;
	unsigned FwdLA = LA - AncHii, FwdLB = LB - AncHij;
	char *bwd = (char *) xrealloc(0, (size_t) AncLoi + AncLoj + 8);
	char *fwd = (char *) xrealloc(0, (size_t) FwdLA + FwdLB + 4);
	unsigned BwdLeni, BwdLenj, FwdLeni, FwdLenj;
	/* xdropalignmem.cpp:87-139 */
	float BwdScore = (AncLoi > XD_MAXL || AncLoj > XD_MAXL)
	  ? xdrop_bwd_split(s, A, AncLoi + 1, B, AncLoj + 1, X, &BwdLeni, &BwdLenj, bwd)
	  : xdrop_bwd(s, A, AncLoi + 1, B, AncLoj + 1, X, &BwdLeni, &BwdLenj, bwd);
	float FwdScore = (FwdLA > XD_MAXL || FwdLB > XD_MAXL)
	  ? xdrop_fwd_split(s, FwdA, FwdLA, FwdB, FwdLB, X, &FwdLeni, &FwdLenj, fwd)
	  : xdrop_fwd(s, FwdA, FwdLA, FwdB, FwdLB, X, &FwdLeni, &FwdLenj, fwd);
	size_t n = 0;
	for (const char *p = bwd; *p; ++p) path[n++] = *p;
	for (unsigned k = 0; k + 2 < AncLen; ++k) path[n++] = 'M';
	for (const char *p = fwd; *p; ++p) path[n++] = *p;
	path[n] = 0;
	free(bwd); free(fwd);
	float (*Mx)[256] = s->subst;
	float AncScore = 0.0f;
	for (unsigned k = 0; k < AncLen; ++k)
		AncScore += Mx[A[AncLoi + k]][B[AncLoj + k]];
	float DupeScore = Mx[A[AncLoi]][B[AncLoj]];
	DupeScore += Mx[A[AncHii]][B[AncHij]];
	float Score = BwdScore + FwdScore + AncScore - DupeScore;
	h4[0] = AncLoi + 1 - BwdLeni;
	h4[1] = AncLoj + 1 - BwdLenj;
	h4[2] = BwdLeni + FwdLeni + AncLen - 2;
	h4[3] = BwdLenj + FwdLenj + AncLen - 2;
	return Score;
	}

/* GetAnchor (localaligner.cpp:11-64): best run of strictly positive pair scores */
static float get_anchor(float (*Mx)[256], const uint8_t *Q, const uint8_t *T, unsigned Loi, unsigned Loj, unsigned L,
  unsigned *AncLoi, unsigned *AncLoj, unsigned *AncLen)
	{
	unsigned Startk = UINT_MAX, BestStartk = UINT_MAX, Length = 0;
	float AnchorScore = 0.0f, BestScore = 0.0f;
	for (unsigned k = 0; k < L; ++k)
		{
		float Score = Mx[Q[Loi + k]][T[Loj + k]];
		if (Score > 0)
			{
			if (Startk == UINT_MAX)
				{
				Startk = k;
				AnchorScore = Score;
				}
			else
				AnchorScore += Score;
			}
		else
			{
			if (AnchorScore > BestScore)
				{
				BestScore = AnchorScore;
				BestStartk = Startk;
				Length = k - Startk;
				}
			Startk = UINT_MAX;
			}
		}
	if (AnchorScore > BestScore)
		{
		BestScore = AnchorScore;
		BestStartk = Startk;
		Length = L - Startk;
		}
	*AncLoi = Loi + BestStartk;
	*AncLoj = Loj + BestStartk;
	*AncLen = Length;
	return BestScore;
	}

/* LocalAligner::AlignPos (localaligner.cpp:101-211).  Returns 1 and fills h4/score/path when an
 * AlignResult would be created. */
static int la_align_pos(uso_searcher *s, const uint8_t *Q, unsigned QL, const uint8_t *T, unsigned TL,
  unsigned QueryPos, unsigned TargetPos, uint32_t *h4, float *score, char *path)
	{
	float (*Mx)[256] = s->subst;
	const float XDropU = s->P.xdrop_u;
	float LeftScore = 0.0f, LeftTotal = 0.0f;
	unsigned LeftLength = 0, k = 0;
	int i = (int) QueryPos, j = (int) TargetPos;
	while (i >= 0 && j >= 0)
		{
		++k;
		LeftTotal += Mx[Q[i]][T[j]];
		if (LeftTotal > LeftScore)
			{
			LeftScore = LeftTotal;
			LeftLength = k;
			}
		else if (LeftScore - LeftTotal > XDropU)
			break;
		--i; --j;
		}
	float RightScore = 0.0f, RightTotal = 0.0f;
	unsigned RightLength = 0;
	i = (int) QueryPos + 1;
	j = (int) TargetPos + 1;
	k = 0;
	while (i < (int) QL && j < (int) TL)
		{
		++k;
		RightTotal += Mx[Q[i]][T[j]];
		if (RightTotal > RightScore)
			{
			RightScore = RightTotal;
			RightLength = k;
			}
		else if (RightScore - RightTotal > XDropU)
			break;
		++i; ++j;
		}
	const float Score = LeftScore + RightScore;
	if (Score < s->la_min_ungapped)
		return 0;
	unsigned Loi = (QueryPos + 1) - LeftLength, Loj = (TargetPos + 1) - LeftLength;
	unsigned SegLength = LeftLength + RightLength;
	unsigned AncLoi, AncLoj, AncLen;
	float AncRaw = get_anchor(Mx, Q, T, Loi, Loj, SegLength, &AncLoi, &AncLoj, &AncLen);
	if (AncRaw <= 0.0f)
		return 0;
	float Gapped = xdrop_align(s, Q, QL, T, TL, AncLoi, AncLoj, AncLen, s->P.xdrop_g, h4, path);
	if (Gapped <= 0.0f)
		return 0;
	double Evalue = es_raw_to_evalue(s, Gapped, QL);
	if (Evalue > s->P.evalue)
		return 0;
	*score = Gapped;
	return 1;
	}

/* HSPData::OverlapFract (hsp.h:74-89) and LocalAligner2::LargeOverlap (localaligner2.cpp:252-258) */
static int la_large_overlap(const uint32_t *a, const uint32_t *b)
	{
	if (a[2] == 0 || a[3] == 0)
		return 0;
	unsigned aHii = a[0] + a[2] - 1, aHij = a[1] + a[3] - 1, bHii = b[0] + b[2] - 1, bHij = b[1] + b[3] - 1;
	unsigned MaxLoi = a[0] > b[0] ? a[0] : b[0], MaxLoj = a[1] > b[1] ? a[1] : b[1];
	unsigned MinHii = aHii < bHii ? aHii : bHii, MinHij = aHij < bHij ? aHij : bHij;
	unsigned Ovi = MinHii < MaxLoi ? 0 : MinHii - MaxLoi;
	unsigned Ovj = MinHij < MaxLoj ? 0 : MinHij - MaxLoj;
	double f = (double) (Ovi * Ovj) / (double) (a[2] * a[3]);
	return f > 0.5;
	}

typedef struct la_ar { uint32_t h4[4]; float score; char *path; } la_ar;

/* LocalAligner2::AlignMulti (localmulti.cpp:9-118).  Returns the number of ARs in *out (malloc'd). */
static unsigned la_align_multi(uso_searcher *s, const uint8_t *Q, unsigned QL, const uint8_t *T, unsigned TL, la_ar **out)
	{
	*out = 0;
	if (TL < 2 * s->la_w)
		return 0;
	if (s->la_tcap < TL + 1)
		{
		s->la_tcap = TL + 1;
		s->la_twords = (uint32_t *) xrealloc(s->la_twords, s->la_tcap * sizeof(uint32_t));
		}
	const unsigned TargetWordCount = la_words(s, T, TL, s->la_twords);
	la_ar *ars = 0;
	unsigned nar = 0, cap = 0;
	char *path = (char *) xrealloc(0, (size_t) QL + TL + 8);
	for (unsigned TargetPos = 0; TargetPos < TargetWordCount; )
		{
		uint32_t TargetWord = s->la_twords[TargetPos];
		unsigned N = s->la_qcounts[TargetWord];
		int skipped = 0;
		for (unsigned i = 0; i < N; ++i)
			{
			unsigned QueryPos = s->la_qposvec[s->la_base[TargetWord] + i];
			la_ar ar;
			if (!la_align_pos(s, Q, QL, T, TL, QueryPos, TargetPos, ar.h4, &ar.score, path))
				continue;
			int keep = 1;
			for (unsigned k = 0; k < nar; ++k)
				if (la_large_overlap(ar.h4, ars[k].h4))
					{
					keep = 0;
					break;
					}
			if (!keep)
				continue;
			if (nar == cap)
				{
				cap = cap ? 2 * cap : 4;
				ars = (la_ar *) xrealloc(ars, cap * sizeof(la_ar));
				}
			ar.path = strdup(path);
			ars[nar++] = ar;
			unsigned NewTargetPos = ar.h4[1] + ar.h4[3]; /* Hij + 1 */
			if (NewTargetPos > TargetPos)
				TargetPos = NewTargetPos;
			else
				++TargetPos;
			skipped = 1;
			break;
			}
		if (!skipped)
			++TargetPos;
		}
	free(path);
	*out = ars;
	return nar;
	}

float uso_xdrop_fwd(const uso_params *p, const uint8_t *A, uint32_t LA, const uint8_t *B, uint32_t LB, float X,
  uint32_t *leni, uint32_t *lenj, char *path)
	{
	uso_db *db = uso_db_create(p);
	uso_searcher *s = uso_searcher_create(db, p);
	unsigned li, lj;
	float sc = xdrop_fwd(s, A, LA, B, LB, X, &li, &lj, path);
	*leni = li; *lenj = lj;
	uso_searcher_free(s);
	uso_db_free(db);
	return sc;
	}

int uso_local_align_pos(uso_searcher *s, const uint8_t *q, uint32_t LQ, const uint8_t *t, uint32_t LT,
  uint32_t qpos, uint32_t tpos, uint32_t *hsp4, float *score, char *path)
	{
	use_alpha(&s->P);
	la_refresh_estats(s);
	s->la_min_ungapped = (float) es_min_ungapped_raw(s, LQ);
	return la_align_pos(s, q, LQ, t, LT, qpos, tpos, hsp4, score, path);
	}

/* ------------------------------------------------------------------ hit statistics */
static void fill_hit(uso_hit *h, const uint8_t *Q, const uint8_t *T, const char *path)
	{
	unsigned first = UINT_MAX, last = UINT_MAX, col = 0;
	for (; path[col]; ++col)
		if (path[col] == 'M')
			{
			if (first == UINT_MAX)
				first = col;
			last = col;
			}
	unsigned qpos = 0, tpos = 0;
	for (unsigned c = 0; c < first; ++c)
		{
		if (path[c] == 'M' || path[c] == 'D') ++qpos;
		if (path[c] == 'M' || path[c] == 'I') ++tpos;
		}
	h->first_mq = qpos;
	h->first_mt = tpos;
	h->first_mcol = first;
	h->ids = h->mism = h->intgaps = h->opens = 0;
	char lastc = 'M';
	for (unsigned c = first; c <= last; ++c)
		{
		char ch = path[c];
		if (ch == 'M')
			{
			if (g_cur_match[Q[qpos]][T[tpos]]) ++h->ids; else ++h->mism;
			++qpos; ++tpos;
			}
		else if (ch == 'D')
			{
			if (c > first) ++h->intgaps;
			++qpos;
			}
		else
			{
			if (c > first) ++h->intgaps;
			++tpos;
			}
		if (ch != 'M' && lastc == 'M')
			++h->opens;
		lastc = ch;
		}
	h->last_mq = qpos - 1;
	h->last_mt = tpos - 1;
	h->alnlen = last - first + 1;
	}

/* ------------------------------------------------------------------ search loop */
static void hits_push(uso_hit **hits, unsigned *n, unsigned *cap, const uso_hit *h)
	{
	if (*n == *cap)
		{
		*cap = *cap ? *cap * 2 : 256;
		*hits = (uso_hit *) xrealloc(*hits, *cap * sizeof(uso_hit));
		}
	(*hits)[(*n)++] = *h;
	}

/* terminator.cpp:64-100 (no -termid/-termidd) */
static int terminate(uso_searcher *s, int accept)
	{
	if (accept) ++s->acc; else ++s->rej;
	if (s->P.maxaccepts > 0 && s->acc == s->P.maxaccepts)
		return 1;
	if (s->P.maxrejects > 0 && s->rej == s->P.maxrejects)
		return 1;
	return 0;
	}

/* Searcher::SetTarget + Align + OnAR (searcher.cpp:86,26,63,52) for one candidate. */
static int try_target(uso_searcher *s, uint32_t qindex, const uint8_t *q, uint32_t L, int strand,
  unsigned t, uso_hit **hits, unsigned *nhits, unsigned *caphits, char *path)
	{
	uso_db *db = s->db;
	const uint8_t *T = db->seqs[t];
	unsigned TL = db->lens[t];
	if (s->P.local)
		{
		/* Searcher::Align, AlignMulti branch (searcher.cpp:31-49): every AR of this target goes
		 * through the Accepter; the target counts once for the Terminator */
		la_ar *ars;
		unsigned nar = la_align_multi(s, q, L, T, TL, &ars);
		int any = 0;
		for (unsigned k = 0; k < nar; ++k)
			{
			uso_hit h;
			memset(&h, 0, sizeof h);
			h.query = qindex; h.target = t; h.strand = (uint8_t) strand; h.ql = L; h.tl = TL;
			h.loi = ars[k].h4[0]; h.loj = ars[k].h4[1]; h.leni = ars[k].h4[2]; h.lenj = ars[k].h4[3];
			fill_hit(&h, q + h.loi, T + h.loj, ars[k].path);
			h.first_mq += h.loi; h.last_mq += h.loi; h.first_mt += h.loj; h.last_mt += h.loj;
			/* arscorer.cpp:87-103 raw score = the path re-scored (alnparams.cpp:447-505) */
			float raw = 0.0f;
			{
			const uint8_t *a = q + h.loi, *b = T + h.loj;
			char last = 'M';
			for (const char *pp = ars[k].path; *pp; ++pp)
				{
				if (*pp == 'M')
					raw += s->subst[toupper(*a++)][toupper(*b++)];
				else if (*pp == 'D')
					{
					raw += last == 'M' ? s->P.lopen : s->P.lext;
					++a;
					}
				else
					{
					raw += last == 'M' ? s->P.lopen : s->P.lext;
					++b;
					}
				last = *pp;
				}
			}
			if (raw != ars[k].score)
				{
				fprintf(stderr, "oracle: re-scored path %.1f != X-drop score %.1f\n", raw, ars[k].score);
				abort();
				}
			h.raw = (double) raw;
			h.evalue = es_raw_to_evalue(s, h.raw, L);
			h.bits = es_raw_to_bits(s, h.raw);
			double FractId = h.alnlen == 0 ? 0.0 : (double) h.ids / (double) h.alnlen;
			int accept = !(FractId < (double) s->P.id) && !(h.evalue > (double) s->P.evalue);
			if (accept)
				{
				any = 1;
				h.path = ars[k].path;
				hits_push(hits, nhits, caphits, &h);
				}
			else
				free(ars[k].path);
			}
		free(ars);
		return terminate(s, any);
		}
	int aligned = global_align(s, q, L, T, TL, path);
	if (!aligned)
		return terminate(s, 0);
	uso_hit h;
	memset(&h, 0, sizeof h);
	h.query = qindex; h.target = t; h.strand = (uint8_t) strand; h.ql = L; h.tl = TL;
	fill_hit(&h, q, T, path);
	/* accepter.cpp:27-38: double FractId < (double)(float)id -> reject */
	double FractId = h.alnlen == 0 ? 0.0 : (double) h.ids / (double) h.alnlen;
	int accept = !(FractId < (double) s->P.id);
	if (accept)
		{
		h.path = strdup(path);
		hits_push(hits, nhits, caphits, &h);
		}
	return terminate(s, accept);
	}

/* wordparams.cpp:125-192 (nucleo) */
static void word_counting_params(const uso_searcher *s, unsigned nuniq, unsigned *MinU, unsigned *Step)
	{
	double FractId = (double) s->P.id;
	double WordFract = 1 - (1 - FractId) * s->P.word_length;
	unsigned Thresh;
	if (WordFract < 0.0)
		Thresh = 1;
	else
		{
		WordFract *= nuniq;
		Thresh = WordFract < 1.0 ? 1 : (unsigned) WordFract;
		}
	if (s->P.stepwords == 0)
		{
		*Step = 1;
		*MinU = Thresh;
		}
	else
		{
		*Step = Thresh / s->P.stepwords;
		if (*Step == 0)
			*Step = 1;
		*MinU = Thresh < s->P.stepwords / 2 ? Thresh : s->P.stepwords / 2;
		}
	}

/* One strand: Searcher::Search body (searcher.cpp:122-160) around SearchImpl
 * (udbusortedsearcher.cpp:122-152) or UDBSearchBig (udbusortedsearcherbig.cpp:31-135). */
static void search_strand(uso_searcher *s, uint32_t qindex, const uint8_t *q, uint32_t L, int strand,
  uso_hit **hits, unsigned *nhits, unsigned *caphits)
	{
	uso_db *db = s->db;
	const unsigned N = db->n;
	/* SetQueryImpl udbusortedsearcher.cpp:39-58: big flag is sticky */
	if (!s->big && N > s->P.big)
		{
		s->big = 1;
		if (s->U)
			memset(s->U, 0, s->Ucap * sizeof(uint32_t));
		s->ntop_prev = 0;
		}
	use_alpha(&s->P);
	if (s->P.local)
		la_set_query(s, q, L); /* Aligner::SetQuery -> LocalAligner2::SetQueryImpl */
	else
		hsp_set_a(s, q, L); /* Aligner::SetQuery -> HSPFinder::SetA */
	s->acc = s->rej = 0; /* Terminator::OnNewQuery */
	char *path = (char *) xrealloc(0, (size_t) L + 60000 + 2);
	size_t pathcap = (size_t) L + 60000 + 2;
	if (!s->big)
		{
		unsigned top = set_target_order(s, q, L);
		for (unsigned k = 0; k < top; ++k)
			{
			unsigned t = s->TopT[s->TopOrder[k]];
			if ((size_t) L + db->lens[t] + 2 > pathcap)
				{
				pathcap = (size_t) L + db->lens[t] + 2;
				path = (char *) xrealloc(path, pathcap);
				}
			if (try_target(s, qindex, q, L, strand, t, hits, nhits, caphits, path))
				break;
			}
		}
	else if (N > 0)
		{
		alloc_query(s, L);
		unsigned nu = unique_words(q, L, s->P.word_length, s->found, s->qw, s->quw);
		unsigned MinU, Step;
		word_counting_params(s, nu, &MinU, &Step);
		if (s->Ucap < N)
			{
			uint32_t nc = ((N + 65536 + 65535) / 65536) * 65536;
			s->U = (uint32_t *) xrealloc(s->U, nc * sizeof(uint32_t));
			memset(s->U, 0, nc * sizeof(uint32_t));
			s->Ucap = nc;
			s->ntop_prev = 0;
			}
		alloc_top(s, N);
		uint32_t *U = s->U;
		unsigned top = 0;
		for (unsigned i = 0; i < nu; i += Step)
			{
			uint32_t w = s->quw[i];
			const uint32_t *row = db->rows[w];
			unsigned size = db->sizes[w];
			for (unsigned j = 0; j < size; ++j)
				{
				uint32_t t = row[j];
				if (U[t] == 0)
					s->TopT[top++] = t;
				++U[t];
				}
			}
		s->ntop_prev = top;
		if (top > 0)
			{
			unsigned top2 = count_sort_subset_desc(s, U, top, s->TopT, s->TopT2);
			for (unsigned k = 0; k < top2; ++k)
				{
				unsigned t = s->TopT2[k];
				if ((size_t) L + db->lens[t] + 2 > pathcap)
					{
					pathcap = (size_t) L + db->lens[t] + 2;
					path = (char *) xrealloc(path, pathcap);
					}
				if (try_target(s, qindex, q, L, strand, t, hits, nhits, caphits, path))
					break;
				}
			}
		/* OnQueryDoneImpl udbusortedsearcher.cpp:65-84 */
		for (unsigned i = 0; i < s->ntop_prev; ++i)
			U[s->TopT[i]] = 0;
		s->ntop_prev = 0;
		}
	if (s->P.local)
		la_query_done(s);
	free(path);
	}

/* Candidate order of UDBSearchBig (udbusortedsearcherbig.cpp:82-135) without the alignments:
 * sampled words, first-touch list, CountSortSubsetDesc.  For tests of the ranking kernel. */
unsigned uso_rank_candidates_big(uso_searcher *s, const uint8_t *q, uint32_t L, uint32_t *cand_t, uint32_t *cand_u)
	{
	uso_db *db = s->db;
	const unsigned N = db->n;
	if (N == 0)
		return 0;
	use_alpha(&s->P);
	alloc_query(s, L);
	unsigned nu = unique_words(q, L, s->P.word_length, s->found, s->qw, s->quw);
	unsigned MinU, Step;
	word_counting_params(s, nu, &MinU, &Step);
	if (s->Ucap < N)
		{
		uint32_t nc = ((N + 65536 + 65535) / 65536) * 65536;
		s->U = (uint32_t *) xrealloc(s->U, nc * sizeof(uint32_t));
		memset(s->U, 0, nc * sizeof(uint32_t));
		s->Ucap = nc;
		}
	else
		memset(s->U, 0, s->Ucap * sizeof(uint32_t));
	alloc_top(s, N);
	uint32_t *U = s->U;
	unsigned top = 0, top2 = 0;
	for (unsigned i = 0; i < nu; i += Step)
		{
		uint32_t w = s->quw[i];
		const uint32_t *row = db->rows[w];
		unsigned size = db->sizes[w];
		for (unsigned j = 0; j < size; ++j)
			{
			uint32_t t = row[j];
			if (U[t] == 0)
				s->TopT[top++] = t;
			++U[t];
			}
		}
	if (top > 0)
		top2 = count_sort_subset_desc(s, U, top, s->TopT, s->TopT2);
	for (unsigned k = 0; k < top2; ++k)
		{
		if (cand_t) cand_t[k] = s->TopT2[k];
		if (cand_u) cand_u[k] = U[s->TopT2[k]];
		}
	for (unsigned i = 0; i < top; ++i)
		U[s->TopT[i]] = 0;
	s->ntop_prev = 0;
	return top2;
	}

/* sort.h:63-102,132 QuickSortOrderDesc<float> -- the reference's own (unstable) quicksort,
 * restated with an explicit stack; partition scheme and recursion order (left part first)
 * are what determine the tie order. */
static void quicksort_order_desc(const float *v, unsigned n, unsigned *order)
	{
	for (unsigned i = 0; i < n; ++i)
		order[i] = i;
	if (n == 0)
		return;
	int stack[128][2];
	int sp = 0;
	stack[sp][0] = 0; stack[sp][1] = (int) n - 1; ++sp;
	while (sp > 0)
		{
		--sp;
		int left = stack[sp][0], right = stack[sp][1];
		int i = left, j = right;
		float pivot = v[order[(left + right) / 2]];
		while (i <= j)
			{
			while (v[order[i]] > pivot) i++;
			while (v[order[j]] < pivot) j--;
			if (i <= j)
				{
				unsigned tmp = order[i]; order[i] = order[j]; order[j] = tmp;
				i++; j--;
				}
			}
		/* reference recurses (left,j) then (i,right); the two ranges are disjoint so order of
		 * processing does not change the result */
		if (i < right) { stack[sp][0] = i; stack[sp][1] = right; ++sp; }
		if (left < j) { stack[sp][0] = left; stack[sp][1] = j; ++sp; }
		}
	}

unsigned uso_search(uso_searcher *s, uint32_t qindex, const uint8_t *q, uint32_t L,
  uso_hit **hits, unsigned *nhits, unsigned *caphits)
	{
	unsigned n0 = *nhits;
	search_strand(s, qindex, q, L, 0, hits, nhits, caphits);
	if (s->P.strand_both)
		{
		uint8_t *rc = (uint8_t *) xrealloc(0, L + 1);
		uso_revcomp(q, L, rc);
		search_strand(s, qindex, rc, L, 1, hits, nhits, caphits);
		free(rc);
		}
	unsigned n = *nhits - n0;
	if (n > 1)
		{
		/* HitMgr::Sort hitmgr.cpp:477: order by float score = (float) FractId (arscorer.cpp:818) */
		float *sc = (float *) xrealloc(0, n * sizeof(float));
		unsigned *ord = (unsigned *) xrealloc(0, n * sizeof(unsigned));
		uso_hit *tmp = (uso_hit *) xrealloc(0, n * sizeof(uso_hit));
		for (unsigned i = 0; i < n; ++i)
			{
			const uso_hit *h = &(*hits)[n0 + i];
			if (s->P.local)
				sc[i] = (float) h->raw; /* arscorer.cpp:818-824 */
			else
				sc[i] = (float) (h->alnlen == 0 ? 0.0 : (double) h->ids / (double) h->alnlen);
			tmp[i] = *h;
			}
		quicksort_order_desc(sc, n, ord);
		for (unsigned i = 0; i < n; ++i)
			(*hits)[n0 + i] = tmp[ord[i]];
		free(sc); free(ord); free(tmp);
		}
	return n;
	}

void uso_hits_free(uso_hit *hits, unsigned n)
	{
	for (unsigned i = 0; i < n; ++i)
		free(hits[i].path);
	free(hits);
	}

/* ------------------------------------------------------------------ output formats */
static double pct_id(const uso_hit *h)
	{
	double f = h->alnlen == 0 ? 0.0 : (double) h->ids / (double) h->alnlen;
	return 100.0 * f;
	}

/* userout.cpp:150-215 with fields query+target+id+alnlen+mism+opens+qlo+qhi+tlo+thi+caln+qstrand.
 * Global: m_HSP spans both sequences (alignresult.cpp:137-145) so qlo..thi = 1,QL,1,TL. */
void uso_write_userout2(FILE *f, const uso_hit *h, const char *qlabel, const char *tlabel, int nucleo)
	{
	char *cp = (char *) xrealloc(0, strlen(h->path) * 2 + 16);
	uso_compress_path(h->path, cp);
	/* amino acid searches have no strand: '.' (arscorer.cpp GetQueryStrand) */
	fprintf(f, "%s\t%s\t%.1f\t%u\t%u\t%u\t%u\t%u\t%u\t%u\t%s\t%c\n", qlabel, tlabel, pct_id(h), h->alnlen,
	  h->mism, h->opens, 1u, h->ql, 1u, h->tl, cp, nucleo ? (h->strand ? '-' : '+') : '.');
	free(cp);
	}

void uso_write_userout(FILE *f, const uso_hit *h, const char *qlabel, const char *tlabel)
	{
	uso_write_userout2(f, h, qlabel, tlabel, 1);
	}

/* blast6out.cpp:27-80; arscorer.cpp:748-808: target coords flip when query is rev-comped */
void uso_write_blast6(FILE *f, const uso_hit *h, const char *qlabel, const char *tlabel)
	{
	unsigned tlo = h->strand ? h->tl : 1u, thi = h->strand ? 1u : h->tl;
	fprintf(f, "%s\t%s\t%.1f\t%u\t%u\t%u\t%u\t%u\t%u\t%u\t*\t*\n", qlabel, tlabel, pct_id(h), h->alnlen,
	  h->mism, h->opens, 1u, h->ql, tlo, thi);
	}

/* outputuc.cpp:45-69 */
void uso_write_uc_hit2(FILE *f, const uso_hit *h, const char *qlabel, const char *tlabel, int nucleo)
	{
	char *cp = (char *) xrealloc(0, strlen(h->path) * 2 + 16);
	uso_compress_path(h->path, cp);
	fprintf(f, "H\t%u\t%u\t%.1f\t%c\t%u\t%u\t%s\t%s\t%s\n", h->target, h->ql, pct_id(h),
	  nucleo ? (h->strand ? '-' : '+') : '.', 0u, 0u, cp, qlabel, tlabel);
	free(cp);
	}

void uso_write_uc_hit(FILE *f, const uso_hit *h, const char *qlabel, const char *tlabel)
	{
	uso_write_uc_hit2(f, h, qlabel, tlabel, 1);
	}

/* outputuc.cpp:19-20 */
void uso_write_uc_nohit(FILE *f, uint32_t ql, const char *qlabel)
	{
	fprintf(f, "N\t*\t%u\t*\t.\t*\t*\t*\t%s\t*\n", ql, qlabel);
	}

/* local hits: coordinates of the segment, on the plus strand of the query when it was
 * reverse-complemented (arscorer.cpp:683-745 without ORFs); blast6 swaps the target ends then
 * (arscorer.cpp:748-808) */
static void local_coords(const uso_hit *h, unsigned *qlo, unsigned *qhi, unsigned *tlo, unsigned *thi)
	{
	unsigned hii = h->loi + h->leni - 1, hij = h->loj + h->lenj - 1;
	*qlo = h->strand ? h->ql - hii - 1 : h->loi;
	*qhi = h->strand ? h->ql - h->loi - 1 : hii;
	*tlo = h->loj;
	*thi = hij;
	}

/* -userfields query+target+id+alnlen+mism+opens+qlo+qhi+tlo+thi+evalue+bits+raw+caln+qstrand */
void uso_write_userout_local(FILE *f, const uso_hit *h, const char *qlabel, const char *tlabel, int nucleo)
	{
	char *cp = (char *) xrealloc(0, strlen(h->path) * 2 + 16);
	uso_compress_path(h->path, cp);
	unsigned qlo, qhi, tlo, thi;
	local_coords(h, &qlo, &qhi, &tlo, &thi);
	fprintf(f, "%s\t%s\t%.1f\t%u\t%u\t%u\t%u\t%u\t%u\t%u\t%.3g\t%.0f\t%.0f\t%s\t%c\n", qlabel, tlabel, pct_id(h),
	  h->alnlen, h->mism, h->opens, qlo + 1, qhi + 1, tlo + 1, thi + 1, h->evalue, h->bits,
	  h->raw, cp, nucleo ? (h->strand ? '-' : '+') : '.');
	free(cp);
	}

void uso_write_blast6_local(FILE *f, const uso_hit *h, const char *qlabel, const char *tlabel)
	{
	unsigned qlo, qhi, tlo, thi;
	local_coords(h, &qlo, &qhi, &tlo, &thi);
	fprintf(f, "%s\t%s\t%.1f\t%u\t%u\t%u\t%u\t%u\t%u\t%u\t%.2g\t%.1f\n", qlabel, tlabel, pct_id(h), h->alnlen,
	  h->mism, h->opens, qlo + 1, qhi + 1, h->strand ? thi + 1 : tlo + 1, h->strand ? tlo + 1 : thi + 1, h->evalue, h->bits);
	}

void uso_write_uc_hit_local(FILE *f, const uso_hit *h, const char *qlabel, const char *tlabel, int nucleo)
	{
	char *cp = (char *) xrealloc(0, strlen(h->path) * 2 + 16);
	uso_compress_path(h->path, cp);
	unsigned qlo, qhi, tlo, thi;
	local_coords(h, &qlo, &qhi, &tlo, &thi);
	fprintf(f, "H\t%u\t%u\t%.1f\t%c\t%u\t%u\t%s\t%s\t%s\n", h->target, h->ql, pct_id(h),
	  nucleo ? (h->strand ? '-' : '+') : '.', qlo, tlo, cp, qlabel, tlabel);
	free(cp);
	}
