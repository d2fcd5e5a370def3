/*
 * uso_main.c -- CLI around the CPU ORACLE (test infrastructure, NOT product code).
 *   uso_cli usearch_global QUERY.fa DB.fa ID plus|both|aa USEROUT UC B6 [maxaccepts maxrejects]
 * Writes the same three output files the reference writes for
 *   -usearch_global Q -db DB -id ID -strand S -userout .. -uc .. -blast6out ..
 *   -userfields query+target+id+alnlen+mism+opens+qlo+qhi+tlo+thi+caln+qstrand
 * Used by tools/pin_oracle.sh to diff the oracle against oracle/_ref/usearch12 at full config sizes.
 */
#include "uso.h"
#include <ctype.h>
#include <stdlib.h>
#include <string.h>

typedef struct { char **labels; uint8_t **seqs; uint32_t *lens; unsigned n, cap; } fasta;

/* fastaseqsource.cpp:25-124: label = rest of '>' line; letters = isalpha chars; empty seqs skipped */
static void read_fasta(const char *fn, fasta *F)
	{
	FILE *f = fopen(fn, "r");
	if (!f) { fprintf(stderr, "cannot open %s\n", fn); exit(1); }
	memset(F, 0, sizeof *F);
	char *line = 0; size_t lcap = 0; ssize_t len;
	char *label = 0; uint8_t *seq = 0; size_t sl = 0, scap = 0;
	for (;;)
		{
		len = getline(&line, &lcap, f);
		if (len < 0 || line[0] == '>')
			{
			if (label && sl > 0)
				{
				if (F->n == F->cap)
					{
					F->cap = F->cap ? F->cap * 2 : 1024;
					F->labels = realloc(F->labels, F->cap * sizeof(char *));
					F->seqs = realloc(F->seqs, F->cap * sizeof(uint8_t *));
					F->lens = realloc(F->lens, F->cap * sizeof(uint32_t));
					}
				F->labels[F->n] = label;
				F->seqs[F->n] = malloc(sl + 1);
				memcpy(F->seqs[F->n], seq, sl);
				F->lens[F->n] = (uint32_t) sl;
				++F->n;
				}
			else
				free(label);
			label = 0;
			if (len < 0)
				break;
			while (len > 0 && (line[len - 1] == '\n' || line[len - 1] == '\r')) line[--len] = 0;
			label = strdup(line + 1);
			sl = 0;
			continue;
			}
		for (ssize_t i = 0; i < len; ++i)
			if (isalpha((unsigned char) line[i]))
				{
				if (sl + 1 > scap) { scap = scap ? scap * 2 : 4096; seq = realloc(seq, scap); }
				seq[sl++] = (uint8_t) line[i];
				}
		}
	free(line); free(seq);
	fclose(f);
	}


/* ------------------------------------------------------------------ cluster_fast
 * clusterfast.cpp:81-133 with -threads 1: dereplicate (derepfull.cpp:130-212; equality is
 * case-insensitive, seqhash.cpp:6-51; uniques in first-occurrence order), optional -sort, then the
 * serial greedy loop: search the unique against the centroids so far (Terminator 1 accept /
 * 8 rejects, terminator.cpp:10-14); no hit -> new centroid appended to the growing UDB
 * (clustersink.cpp:306-330, udbbuild.cpp:286).  Output: .uc S/H records incl. the duplicate
 * members (outputuc.cpp:9-92), C records (clustersink.cpp:477-492), centroids FASTA in decreasing
 * cluster size order (clustersink.cpp:246-272, sort.h:63-102), 80 letters per line. */
static int seq_eq_ci(const uint8_t *a, const uint8_t *b, uint32_t L)
	{
	for (uint32_t i = 0; i < L; ++i)
		if (toupper(a[i]) != toupper(b[i]))
			return 0;
	return 1;
	}

static void qsort_order_desc_u(const unsigned *v, int left, int right, unsigned *order)
	{
	int i = left, j = right;
	unsigned pivot = v[order[(left + right) / 2]];
	while (i <= j)
		{
		while (v[order[i]] > pivot) i++;
		while (v[order[j]] < pivot) j--;
		if (i <= j)
			{
			unsigned t = order[i]; order[i] = order[j]; order[j] = t;
			i++; j--;
			}
		}
	if (left < j) qsort_order_desc_u(v, left, j, order);
	if (i < right) qsort_order_desc_u(v, i, right, order);
	}

static int cluster_fast_main(int argc, char **argv)
	{
	/* uso_cli cluster_fast READS.fa ID UC CENTROIDS [sort: none|length|size] */
	if (argc < 6)
		{
		fprintf(stderr, "usage: uso_cli cluster_fast READS.fa ID UC CENTROIDS [none|length|size]\n");
		return 2;
		}
	uso_params P;
	uso_default_params(&P, 1);
	P.id = (float) atof(argv[3]);
	const char *sortname = argc > 6 ? argv[6] : "none";
	fasta R;
	read_fasta(argv[2], &R);
	/* dereplication: open hash on upper-cased letters */
	unsigned nb = 1;
	while (nb < 2 * R.n + 16) nb <<= 1;
	int *bucket = malloc(nb * sizeof(int));
	for (unsigned i = 0; i < nb; ++i) bucket[i] = -1;
	unsigned *uniq_of = malloc((R.n + 1) * sizeof(unsigned));   /* read -> unique index */
	unsigned *first = malloc((R.n + 1) * sizeof(unsigned));     /* unique -> first read */
	unsigned *usize = calloc(R.n + 1, sizeof(unsigned));
	unsigned nu = 0;
	for (unsigned i = 0; i < R.n; ++i)
		{
		uint32_t h = 2166136261u;
		for (uint32_t k = 0; k < R.lens[i]; ++k)
			h = (h ^ (uint32_t) toupper(R.seqs[i][k])) * 16777619u;
		unsigned b = h & (nb - 1);
		for (;;)
			{
			if (bucket[b] < 0)
				{
				bucket[b] = (int) nu;
				first[nu] = i;
				uniq_of[i] = nu++;
				break;
				}
			unsigned f = first[bucket[b]];
			if (R.lens[f] == R.lens[i] && seq_eq_ci(R.seqs[f], R.seqs[i], R.lens[i]))
				{
				uniq_of[i] = (unsigned) bucket[b];
				break;
				}
			b = (b + 1) & (nb - 1);
			}
		++usize[uniq_of[i]];
		}
	unsigned *order = malloc((nu + 1) * sizeof(unsigned));
	for (unsigned u = 0; u < nu; ++u) order[u] = u;
	if (strcmp(sortname, "length") == 0 || strcmp(sortname, "size") == 0)
		{
		unsigned *v = malloc((nu + 1) * sizeof(unsigned));
		for (unsigned u = 0; u < nu; ++u)
			v[u] = sortname[0] == 'l' ? R.lens[first[u]] : usize[u];
		if (nu > 0)
			qsort_order_desc_u(v, 0, (int) nu - 1, order);
		free(v);
		}
	uso_db *db = uso_db_create(&P);
	uso_searcher *s = uso_searcher_create(db, &P);
	FILE *fc = fopen(argv[4], "w");
	unsigned *csize = calloc(nu + 1, sizeof(unsigned));
	unsigned *centroid_read = malloc((nu + 1) * sizeof(unsigned));
	unsigned nclust = 0;
	uso_hit *hits = 0; unsigned nh = 0, cap = 0;
	char *cp = malloc(1 << 20);
	for (unsigned k = 0; k < nu; ++k)
		{
		unsigned u = order[k];
		unsigned r0 = first[u];
		nh = 0;
		unsigned n = uso_search(s, u, R.seqs[r0], R.lens[r0], &hits, &nh, &cap);
		if (n == 0)
			{
			unsigned c = uso_db_add(db, R.seqs[r0], R.lens[r0], R.labels[r0]);
			centroid_read[c] = r0;
			csize[c] = usize[u];
			nclust = c + 1;
			fprintf(fc, "S\t%u\t%u\t*\t.\t*\t*\t*\t%s\t*\n", c, R.lens[r0], R.labels[r0]);
			for (unsigned i = r0 + 1; i < R.n; ++i)
				if (uniq_of[i] == u)
					fprintf(fc, "H\t%u\t%u\t100.0\t.\t0\t%u\t=\t%s\t%s\n", c, R.lens[r0], R.lens[r0], R.labels[i], R.labels[r0]);
			}
		else
			{
			const uso_hit *h = &hits[0];
			unsigned c = h->target;
			csize[c] += usize[u];
			uso_compress_path(h->path, cp);
			double pct = 100.0 * (h->alnlen == 0 ? 0.0 : (double) h->ids / (double) h->alnlen);
			for (unsigned i = r0; i < R.n; ++i)
				if (uniq_of[i] == u)
					fprintf(fc, "H\t%u\t%u\t%.1f\t%c\t0\t0\t%s\t%s\t%s\n", c, R.lens[r0], pct, h->strand ? '-' : '+', cp,
					  R.labels[i], uso_db_label(db, c));
			for (unsigned i = 0; i < n; ++i)
				free(hits[i].path);
			}
		}
	for (unsigned c = 0; c < nclust; ++c)
		fprintf(fc, "C\t%u\t%u\t*\t*\t*\t*\t*\t%s\t*\n", c, csize[c], uso_db_label(db, c));
	fclose(fc);
	FILE *ff = fopen(argv[5], "w");
	unsigned *corder = malloc((nclust + 1) * sizeof(unsigned));
	for (unsigned c = 0; c < nclust; ++c) corder[c] = c;
	if (nclust > 0)
		qsort_order_desc_u(csize, 0, (int) nclust - 1, corder);
	for (unsigned k = 0; k < nclust; ++k)
		{
		unsigned c = corder[k];
		uint32_t L;
		const uint8_t *seq = uso_db_seq(db, c, &L);
		fprintf(ff, ">%s\n", uso_db_label(db, c));
		for (uint32_t i = 0; i < L; i += 80)
			{
			fwrite(seq + i, 1, L - i < 80 ? L - i : 80, ff);
			fputc('\n', ff);
			}
		}
	fclose(ff);
	return 0;
	}

/* uso_cli usearch_local Q.fa DB.fa ID EVALUE aa|nt|ntboth USEROUT UC B6 [maxaccepts maxrejects]
 * = -usearch_local Q -db DB -id ID -evalue E [-strand plus] -userout .. -uc .. -blast6out ..
 *   -userfields query+target+id+alnlen+mism+opens+qlo+qhi+tlo+thi+evalue+bits+raw+caln+qstrand */
static int usearch_local_main(int argc, char **argv)
	{
	if (argc < 10)
		{
		fprintf(stderr, "usage: uso_cli usearch_local Q.fa DB.fa ID EVALUE aa|nt|ntboth USEROUT UC B6 [maxaccepts maxrejects]\n");
		return 2;
		}
	uso_params P;
	uso_default_params(&P, 0);
	P.local = 1;
	P.id = (float) atof(argv[4]);
	P.evalue = (float) atof(argv[5]);
	int nucleo = strcmp(argv[6], "nt") == 0 || strcmp(argv[6], "ntboth") == 0;
	P.strand_both = strcmp(argv[6], "ntboth") == 0;
	if (!nucleo)
		uso_set_amino(&P);
	if (argc > 11)
		{
		P.maxaccepts = (unsigned) atoi(argv[10]);
		P.maxrejects = (unsigned) atoi(argv[11]);
		}
	fasta Q, D;
	read_fasta(argv[2], &Q);
	read_fasta(argv[3], &D);
	uso_db *db = uso_db_create(&P);
	for (unsigned i = 0; i < D.n; ++i)
		uso_db_add(db, D.seqs[i], D.lens[i], D.labels[i]);
	uso_searcher *s = uso_searcher_create(db, &P);
	FILE *fu = fopen(argv[7], "w"), *fc = fopen(argv[8], "w"), *fb = fopen(argv[9], "w");
	uso_hit *hits = 0; unsigned nh = 0, cap = 0;
	for (unsigned i = 0; i < Q.n; ++i)
		{
		nh = 0;
		unsigned n = uso_search(s, i, Q.seqs[i], Q.lens[i], &hits, &nh, &cap);
		for (unsigned k = 0; k < n; ++k)
			{
			const char *tl = uso_db_label(db, hits[k].target);
			uso_write_userout_local(fu, &hits[k], Q.labels[i], tl, nucleo);
			uso_write_uc_hit_local(fc, &hits[k], Q.labels[i], tl, nucleo);
			uso_write_blast6_local(fb, &hits[k], Q.labels[i], tl);
			free(hits[k].path);
			}
		if (n == 0)
			uso_write_uc_nohit(fc, Q.lens[i], Q.labels[i]);
		}
	fclose(fu); fclose(fc); fclose(fb);
	return 0;
	}

int main(int argc, char **argv)
	{
	if (argc >= 2 && strcmp(argv[1], "cluster_fast") == 0)
		return cluster_fast_main(argc, argv);
	if (argc >= 2 && strcmp(argv[1], "usearch_local") == 0)
		return usearch_local_main(argc, argv);
	if (argc < 9 || strcmp(argv[1], "usearch_global") != 0)
		{
		fprintf(stderr, "usage: uso_cli usearch_global Q.fa DB.fa ID plus|both|aa USEROUT UC B6 [maxaccepts maxrejects]\n");
		return 2;
		}
	uso_params P;
	uso_default_params(&P, 0);
	P.id = (float) atof(argv[4]);
	P.strand_both = strcmp(argv[5], "both") == 0;
	if (strcmp(argv[5], "aa") == 0) /* amino acid DB: no -strand option (search.cpp:23-34) */
		uso_set_amino(&P);
	if (argc > 10)
		{
		P.maxaccepts = (unsigned) atoi(argv[9]);
		P.maxrejects = (unsigned) atoi(argv[10]);
		}
	if (argc > 11)
		P.band = (unsigned) atoi(argv[11]);
	if (argc > 12)
		P.fulldp = atoi(argv[12]);
	fasta Q, D;
	read_fasta(argv[2], &Q);
	read_fasta(argv[3], &D);
	uso_db *db = uso_db_create(&P);
	for (unsigned i = 0; i < D.n; ++i)
		uso_db_add(db, D.seqs[i], D.lens[i], D.labels[i]);
	uso_searcher *s = uso_searcher_create(db, &P);
	FILE *fu = fopen(argv[6], "w"), *fc = fopen(argv[7], "w"), *fb = fopen(argv[8], "w");
	uso_hit *hits = 0; unsigned nh = 0, cap = 0;
	for (unsigned i = 0; i < Q.n; ++i)
		{
		nh = 0;
		unsigned n = uso_search(s, i, Q.seqs[i], Q.lens[i], &hits, &nh, &cap);
		for (unsigned k = 0; k < n; ++k)
			{
			const char *tl = uso_db_label(db, hits[k].target);
			uso_write_userout2(fu, &hits[k], Q.labels[i], tl, P.is_nucleo);
			uso_write_uc_hit2(fc, &hits[k], Q.labels[i], tl, P.is_nucleo);
			uso_write_blast6(fb, &hits[k], Q.labels[i], tl);
			free(hits[k].path);
			}
		if (n == 0)
			uso_write_uc_nohit(fc, Q.lens[i], Q.labels[i]);
		}
	fclose(fu); fclose(fc); fclose(fb);
	return 0;
	}
