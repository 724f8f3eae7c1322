/*
 * pdict_oracle.c -- CPU restatement of Biostrings' PDict matching path.
 *
 * TEST INFRASTRUCTURE ONLY: this is the parity checker (and, where the compiled
 * reference in oracle/_ref is unavailable, the "port" CPU baseline).  Only tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline legs may load it; the product
 * (biostrings_b200/, include/) never does.
 *
 * It is a plain restatement of what the reference computes, not of how:
 *   - TB preprocessing (id of first pattern per TB word + high2low for later ones,
 *     with pp_exclude):           src/match_pdict_Twobit.c:30-46,104-148,
 *                                 src/match_pdict_ACtree2.c:750-781,792-846
 *   - fixedS=TRUE TB scan (a non-base byte resets the window):
 *                                 src/match_pdict_Twobit.c:158-175, src/utils.c:124-146,
 *                                 src/match_pdict_ACtree2.c:871-891,1001-1021
 *   - fixedS=FALSE TB scan (IUPAC subject letters expand, bytes >= 16 reset):
 *                                 src/match_pdict_ACtree2.c:1028-1061,1124-1157
 *   - head/tail verification with out-of-limits = mismatch, early exit, min/max:
 *                                 src/match_pdict_utils.c:153-214,283-300,809-861,
 *                                 src/lowlevel_matching.c:26-89
 *   - reporting (end = tb_end + |tail|; counts; per-key ascending ends):
 *                                 src/match_pdict_utils.c:87-116, src/match_reporting.c:150-201
 * Pinned against the reference's golden vectors and against oracle/_ref (the compiled
 * reference itself) in tests/test_oracle_*.py.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

#define NA_INT INT_MIN

typedef struct {
	int child[4];	/* node ids, 0 = none (the root is never a child) */
	int P_id;	/* leaf: 1-based id of the first pattern with this TB word, else 0 */
} Node;

typedef struct oracle_pptb {
	int M, w, algo;		/* algo: 0 = ACtree2, 1 = Twobit */
	Node *nodes;
	long nnodes, cap;
	int *high2low;		/* NA_INT or 1-based id */
	int **low2high;		/* per 0-based key: 0-based keys sharing its TB (ascending) */
	int *low2high_n;
} oracle_pptb;

static int base2code(unsigned char c)
{
	switch (c) {	/* R/XStringCodec-class.R:114; base_codes = c(1,2,4,8) */
	case 1: return 0;
	case 2: return 1;
	case 4: return 2;
	case 8: return 3;
	}
	return -1;
}

static long new_node(oracle_pptb *x)
{
	if (x->nnodes == x->cap) {
		x->cap = x->cap ? 2 * x->cap : 1024;
		x->nodes = realloc(x->nodes, (size_t) x->cap * sizeof(Node));
	}
	memset(&x->nodes[x->nnodes], 0, sizeof(Node));
	return x->nnodes++;
}

void oracle_pptb_free(oracle_pptb *x)
{
	int i;

	if (x == NULL)
		return;
	if (x->low2high != NULL)
		for (i = 0; i < x->M; i++)
			free(x->low2high[i]);
	free(x->low2high);
	free(x->low2high_n);
	free(x->high2low);
	free(x->nodes);
	free(x);
}

#define FAIL(...) do { snprintf(err, errlen, __VA_ARGS__); oracle_pptb_free(x); return NULL; } while (0)

oracle_pptb *oracle_pptb_build(int algo, int M, const uint8_t *tb_concat,
		const int *tb_widths, const int *pp_exclude, int *high2low_out,
		char *err, int errlen)
{
	oracle_pptb *x = calloc(1, sizeof(*x));
	const uint8_t *p = tb_concat;
	int i, d, w = -1;

	x->M = M;
	x->algo = algo;
	x->high2low = malloc((size_t) (M ? M : 1) * sizeof(int));
	for (i = 0; i < M; i++)
		x->high2low[i] = NA_INT;
	if (M == 0 && algo == 0)
		FAIL("Trusted Band is empty");
	new_node(x);	/* root */
	for (i = 0; i < M; p += tb_widths[i], i++) {
		long nid = 0;

		if (pp_exclude != NULL && pp_exclude[i] != NA_INT)
			continue;
		if (algo == 1 && tb_widths[i] == 0)
			FAIL("empty trusted region for pattern %d", i + 1);
		if (w == -1) {
			if (tb_widths[i] == 0)
				FAIL("first element in Trusted Band is of length 0");
			w = tb_widths[i];
			if (algo == 1 && w > 14)
				FAIL("the width of the Trusted Band must be <= 14 "
				     "when 'type=\"Twobit\"'");
		} else if (tb_widths[i] != w) {
			if (algo == 1)
				FAIL("all the trusted regions must have the same length");
			FAIL("element %d in Trusted Band has a different length "
			     "than first element", i + 1);
		}
		for (d = 0; d < w; d++) {
			int c = base2code(p[d]);
			long nid2;

			if (c < 0) {
				if (algo == 1)
					FAIL("non-base DNA letter found in Trusted Band "
					     "for pattern %d", i + 1);
				FAIL("non base DNA letter found in Trusted Band "
				     "for pattern %d", i + 1);
			}
			nid2 = x->nodes[nid].child[c];
			if (nid2 == 0) {
				nid2 = new_node(x);
				x->nodes[nid].child[c] = (int) nid2;
			}
			nid = nid2;
		}
		if (x->nodes[nid].P_id != 0)
			x->high2low[i] = x->nodes[nid].P_id;
		else
			x->nodes[nid].P_id = i + 1;
	}
	x->w = w;
	/* low2high (IRanges H2LGrouping): ascending ids with high2low == k */
	x->low2high = calloc((size_t) (M ? M : 1), sizeof(int *));
	x->low2high_n = calloc((size_t) (M ? M : 1), sizeof(int));
	for (i = 0; i < M; i++)
		if (x->high2low[i] != NA_INT)
			x->low2high_n[x->high2low[i] - 1]++;
	for (i = 0; i < M; i++)
		if (x->low2high_n[i] != 0) {
			x->low2high[i] = malloc((size_t) x->low2high_n[i] * sizeof(int));
			x->low2high_n[i] = 0;
		}
	for (i = 0; i < M; i++)
		if (x->high2low[i] != NA_INT) {
			int k = x->high2low[i] - 1;

			x->low2high[k][x->low2high_n[k]++] = i;
		}
	if (high2low_out != NULL)
		memcpy(high2low_out, x->high2low, (size_t) M * sizeof(int));
	return x;
}

/* R/PDict-class.R:219-227: the reused whole-pattern pptb0 has its dups blanked. */
void oracle_pptb_blank_dups(oracle_pptb *x)
{
	int i;

	for (i = 0; i < x->M; i++) {
		x->high2low[i] = NA_INT;
		free(x->low2high[i]);
		x->low2high[i] = NULL;
		x->low2high_n[i] = 0;
	}
}

int oracle_pptb_width(const oracle_pptb *x) { return x->w; }
long oracle_pptb_nnodes(const oracle_pptb *x) { return x->nnodes; }

/* ---- hit buffers ------------------------------------------------------------ */

typedef struct { int *a, *b; long n, cap; } Pairs;

static void pairs_push(Pairs *p, int a, int b)
{
	if (p->n == p->cap) {
		p->cap = p->cap ? 2 * p->cap : 4096;
		p->a = realloc(p->a, (size_t) p->cap * sizeof(int));
		p->b = realloc(p->b, (size_t) p->cap * sizeof(int));
	}
	p->a[p->n] = a;
	p->b[p->n] = b;
	p->n++;
}

/* ---- TB scan ------------------------------------------------------------------ */

typedef struct { long n_states, cap_states; } Budget;

/* all TB words t with (S[i] & t[i]) != 0 for every i of the window */
static int expand(const oracle_pptb *x, const uint8_t *win, int d, long nid,
		  int n, Pairs *tbhits, long *work, long max_work)
{
	int c;

	if (d == x->w) {
		pairs_push(tbhits, x->nodes[nid].P_id - 1, n);
		return 0;
	}
	for (c = 0; c < 4; c++) {
		long nid2;

		if (!(win[d] & (1 << c)))
			continue;
		nid2 = x->nodes[nid].child[c];
		if (nid2 == 0)
			continue;
		if (++*work > max_work)
			return -1;
		if (expand(x, win, d + 1, nid2, n, tbhits, work, max_work) != 0)
			return -1;
	}
	return 0;
}

static int scan_tb(const oracle_pptb *x, const uint8_t *S, int L, int fixedS,
		   Pairs *tbhits, long max_work, char *err, int errlen)
{
	int n, run = 0, w = x->w, d;

	if (w <= 0)
		return 0;
	for (n = 1; n <= L; n++) {
		unsigned char c = S[n - 1];

		if (fixedS) {
			/* src/utils.c:128-144: a non-base byte resets the window */
			if (base2code(c) < 0) {
				run = 0;
				continue;
			}
			if (++run < w)
				continue;
			{
				long nid = 0;

				for (d = 0; d < w && nid >= 0; d++) {
					long nid2 = x->nodes[nid].child[base2code(S[n - w + d])];

					nid = nid2 == 0 ? -1 : nid2;
				}
				if (nid > 0)
					pairs_push(tbhits, x->nodes[nid].P_id - 1, n);
			}
		} else {
			long work = 0;

			/* src/match_pdict_ACtree2.c:1137-1141: bytes >= 16 reset.  Byte 0
			   (not a DNA code) is treated as a reset too: documented
			   domain restriction, see DESIGN.md. */
			if (c >= 16 || c == 0) {
				run = 0;
				continue;
			}
			if (++run < w)
				continue;
			if (expand(x, S + n - w, 0, 0, n, tbhits, &work, max_work) != 0) {
				snprintf(err, errlen,
					 "too many IUPAC ambiguity letters in 'subject'");
				return -1;
			}
		}
	}
	return 0;
}

/* ---- flanks -------------------------------------------------------------------- */

static int letters_match(unsigned char p, unsigned char s, int fixedP, int fixedS)
{
	/* src/lowlevel_matching.c:26-56 */
	if (fixedP)
		return fixedS ? p == s : (p & ~s & 0xFF) == 0;
	return fixedS ? (~p & s & 0xFF) == 0 : (p & s) != 0;
}

/* src/lowlevel_matching.c:66-89 */
static int nmismatch_at_Pshift(const uint8_t *P, int Plen, const uint8_t *S, int L,
			       int Pshift, int max_nmis, int fixedP, int fixedS)
{
	int nmis = 0, i, j;

	for (i = 0, j = Pshift; i < Plen; i++, j++) {
		if (j >= 0 && j < L && letters_match(P[i], S[j], fixedP, fixedS))
			continue;
		if (nmis++ >= max_nmis)
			break;
	}
	return nmis;
}

static int cmp_pair(const void *pa, const void *pb)
{
	const long *a = pa, *b = pb;

	return (a[0] > b[0]) - (a[0] < b[0]);
}

/*
 * One subject against one PDict3Parts (match_pdict(), src/match_pdict.c:40-77).
 * head/tail: concat + widths, or NULL widths when absent.
 * Outputs: counts[M] (always), and the hit list (0-based key, 1-based end) through
 * hit_keys / hit_ends / nhits (malloc'ed; caller frees with oracle_free) if non-NULL.
 * Per key the ends come out in ascending TB-end order.
 * Returns 0, or -1 with a message in err.
 */
int oracle_match(const oracle_pptb *x,
		const uint8_t *head_concat, const int *head_widths,
		const uint8_t *tail_concat, const int *tail_widths,
		const uint8_t *S, int L,
		int max_nmis, int min_nmis, int fixedP, int fixedS,
		long max_work,
		int *counts, int **hit_keys, int **hit_ends, long *nhits,
		char *err, int errlen)
{
	Pairs tbhits = {0}, hits = {0};
	long *hoff = NULL, *toff = NULL, *order = NULL, i, j;
	int M = x->M, w = x->w, k, rc = 0;

	memset(counts, 0, (size_t) M * sizeof(int));
	if (x->algo == 1 && !fixedS) {
		snprintf(err, errlen, "cannot treat IUPAC extended letters in the subject "
			 "as ambiguities when 'pdict' is a PDict object of the \"Twobit\" type");
		return -1;
	}
	if (scan_tb(x, S, L, fixedS, &tbhits, max_work, err, errlen) != 0) {
		rc = -1;
		goto done;
	}
	hoff = calloc((size_t) M + 1, sizeof(long));
	toff = calloc((size_t) M + 1, sizeof(long));
	for (k = 0; k < M; k++) {
		hoff[k + 1] = hoff[k] + (head_widths ? head_widths[k] : 0);
		toff[k + 1] = toff[k] + (tail_widths ? tail_widths[k] : 0);
	}
	/* group TB hits by tb id, keeping ascending n inside a group (stable) */
	order = malloc((size_t) (tbhits.n ? tbhits.n : 1) * 2 * sizeof(long));
	for (i = 0; i < tbhits.n; i++) {
		order[2 * i] = (long) tbhits.a[i] * ((long) INT_MAX + 1) + tbhits.b[i];
		order[2 * i + 1] = i;
	}
	qsort(order, (size_t) tbhits.n, 2 * sizeof(long), cmp_pair);
	for (i = 0; i < tbhits.n; i = j) {
		int key0 = tbhits.a[order[2 * i + 1]], g;

		for (j = i; j < tbhits.n && tbhits.a[order[2 * j + 1]] == key0; j++)
			;
		/* keys = {key0} U low2high[key0]   (src/match_pdict_utils.c:153-181) */
		for (g = -1; g < x->low2high_n[key0]; g++) {
			int key = g < 0 ? key0 : x->low2high[key0][g];
			int Hlen = (int) (hoff[key + 1] - hoff[key]);
			int Tlen = (int) (toff[key + 1] - toff[key]);
			const uint8_t *H = head_concat ? head_concat + hoff[key] : NULL;
			const uint8_t *T = tail_concat ? tail_concat + toff[key] : NULL;
			long q;

			for (q = i; q < j; q++) {
				int tb_end = tbhits.b[order[2 * q + 1]], nmis;

				/* src/match_pdict_utils.c:183-214 */
				nmis = nmismatch_at_Pshift(H, Hlen, S, L, tb_end - Hlen - w,
							   max_nmis, fixedP, fixedS);
				if (nmis <= max_nmis)
					nmis += nmismatch_at_Pshift(T, Tlen, S, L, tb_end,
							max_nmis - nmis, fixedP, fixedS);
				if (nmis <= max_nmis && nmis >= min_nmis) {
					counts[key]++;
					if (hit_keys != NULL)
						pairs_push(&hits, key, tb_end + Tlen);
				}
			}
		}
	}
done:
	free(order);
	free(hoff);
	free(toff);
	free(tbhits.a);
	free(tbhits.b);
	if (hit_keys != NULL && rc == 0) {
		*hit_keys = hits.a;
		*hit_ends = hits.b;
		*nhits = hits.n;
	} else {
		free(hits.a);
		free(hits.b);
	}
	return rc;
}

void oracle_free(void *p) { free(p); }
