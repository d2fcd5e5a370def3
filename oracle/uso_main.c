/*
 * uso_main.c -- CLI around the CPU ORACLE (test infrastructure, NOT product code).
 *   uso_cli usearch_global QUERY.fa DB.fa ID [plus|both] USEROUT UC B6 [maxaccepts maxrejects]
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

int main(int argc, char **argv)
	{
	if (argc < 9 || strcmp(argv[1], "usearch_global") != 0)
		{
		fprintf(stderr, "usage: uso_cli usearch_global Q.fa DB.fa ID plus|both USEROUT UC B6 [maxaccepts maxrejects]\n");
		return 2;
		}
	uso_params P;
	uso_default_params(&P, 0);
	P.id = (float) atof(argv[4]);
	P.strand_both = strcmp(argv[5], "both") == 0;
	if (argc > 10)
		{
		P.maxaccepts = (unsigned) atoi(argv[9]);
		P.maxrejects = (unsigned) atoi(argv[10]);
		}
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
			uso_write_userout(fu, &hits[k], Q.labels[i], tl);
			uso_write_uc_hit(fc, &hits[k], Q.labels[i], tl);
			uso_write_blast6(fb, &hits[k], Q.labels[i], tl);
			free(hits[k].path);
			}
		if (n == 0)
			uso_write_uc_nohit(fc, Q.lens[i], Q.labels[i]);
		}
	fclose(fu); fclose(fc); fclose(fb);
	return 0;
	}
