// bsgpu_tbdir.cuh -- Trusted-Band dictionaries with heads / tails (or duplicated bands), band width
// <= 32: the first min(w, 12) letters of a window ARE the address of a directory, so a window costs
// one 8-byte gather and a directory hit one more.
//
// Replaces the hash probe of scan_tb_kernel followed by tb_hit() / verify_flanks_kernel for the
// windows made of base letters only.  With 10^6 keys over 4^12 prefixes 6 % of all windows hit a
// prefix by chance (config 3 of BASELINE: 36-mers, Trusted Band 1..12, max.mismatch 2; 12 % for the
// 2M keys of config 5), and every such hit walked bucket -> key -> group -> flank lengths -> flanks ->
// subject letters, six dependent gathers in a divergent loop (or went through a candidate buffer, a
// host read of its length and a second kernel).  Here (the layout of qscan2_kernel, bsgpu_qindex.cuh):
//   cells : 8 bytes per 16 prefixes = 16 two-bit slot counts + the index of the cell's first slot
//   slots : 8 bytes per KEY (the keys sharing a band sit side by side, src/match_pdict_utils.c:
//           153-181): key id and a 16-bit fingerprint = the next <= 8 band letters after the prefix
//           (must be equal), then the first base letters of the tail, then the last base letters of
//           the head, compared with the subject letters around the window by XOR + popcount.  The
//           popcount is a lower bound of the head + tail mismatches of src/match_pdict_utils.c:
//           183-214 (with fixedS = FALSE the subject letters that are not bases are left out of it),
//           so what it rejects cannot match; what passes is verified in full: the rest of the band
//           against the key, the flanks by XOR + popcount on the packed planes when they are
//           base-only and the subject is taken literally, letter by letter through the match tables
//           otherwise (src/lowlevel_matching.c:26-89).
// Same warp organisation as qscan2_kernel: lane l owns the windows l, l + 32, ... of a batch of
// 1024, slots go through a ring in shared memory, the ring is drained 32 x U loads at a time.
// Windows holding IUPAC ambiguity letters (fixedS = FALSE) stay with scan_tb_iupac_kernel.
#pragma once
#include "bsgpu_qindex.cuh"

#define TBD_PREFIX 12
#define TBD_MAX_W 32
#define TBD_ID_BITS 24
#define TBD_FP_SHIFT 24         // 16 bits: 8 letters = [rest of the band][tail][head]
#define TBD_NT_SHIFT 40         // 4 bits: tail letters in the fingerprint
#define TBD_NH_SHIFT 44         // 4 bits: head letters
#define TBD_NR_SHIFT 48         // 4 bits: band letters after the prefix

struct BsgpuTbDirView {
	const uint2 *cells;                     // [4^prefix / 16] {16 x 2-bit slot counts, first slot}
	const unsigned long long *slots;
	int32_t w;                              // band width
	int32_t prefix;                         // min(w, 12) letters address the directory
};

// one key that passed the fingerprint, window ending at concat index e
__device__ __noinline__ void tbdir_survivor(const BsgpuSubjectView *S, const BsgpuIndexView *I, const BsgpuReport *R,
		int key, int64_t e, int max_nmis, int min_nmis, int fixedP, int fixedS)
{
	const int w = I->w;

	if (w > TBD_PREFIX) {
		// the whole band (<= 32 letters) against the key
		int64_t a = e - w + 1, wi = a >> 5;
		int sh = 2 * (int) (a & 31);
		uint64_t c0 = __ldg(S->codes + wi), c1 = __ldg(S->codes + wi + 1);
		uint64_t x = sh ? (c0 >> sh) | (c1 << (64 - sh)) : c0;

		if (((x ^ __ldg(I->keys + key)) & I->seed_mask) != 0ull)
			return;
	}
	int seg = S->nseg == 1 ? 0 : find_segment(*S, e);
	int64_t lo = __ldg(S->seg_start + seg), L = __ldg(S->seg_len + seg);
	uint32_t fl = __ldg(I->flank_len + key);
	int nmis, tl;

	if (fixedS && (fl & 0x8000u)) {
		int hl = (int) (fl & 63u);
		ulonglong2 f = __ldg(I->flank2 + key);

		tl = (int) ((fl >> 6) & 63u);
		nmis = packed_flank_mismatches(*S, e - w - hl + 1, hl, f.x, lo, L);
		if (nmis <= max_nmis)
			nmis += packed_flank_mismatches(*S, e + 1, tl, f.y, lo, L);
	} else {
		int64_t n = e - lo + 1;                 // 1-based band end inside the segment
		const uint8_t *segp = S->bytes + lo;
		int64_t h0 = I->head ? __ldg(I->head_off + key) : 0;
		int hl = I->head ? (int) (__ldg(I->head_off + key + 1) - h0) : 0;
		int64_t t0 = I->tail ? __ldg(I->tail_off + key) : 0;

		tl = I->tail ? (int) (__ldg(I->tail_off + key + 1) - t0) : 0;
		nmis = count_mismatches(I->head + h0, hl, segp, L, n - hl - w, max_nmis, fixedP, fixedS);
		if (nmis <= max_nmis)
			nmis += count_mismatches(I->tail + t0, tl, segp, L, n, max_nmis - nmis, fixedP, fixedS);
	}
	if (nmis <= max_nmis && nmis >= min_nmis)
		report_match(*R, *S, key, seg, e - lo + 1, e, tl);
}

#define TBD_CAP 256
#define TBD_U 2
#define TBD_WARPS 2
#define TBD_GROUP 2
#define TBD_WORDS 38            // words jb-1 .. jb+35 of the batch (a window of 32 letters + 8 tail letters behind it)
static_assert(TBD_GROUP * 96 + 32 * TBD_U <= TBD_CAP, "ring too small");

// MASK_INV: fixedS = FALSE, subject letters that are not bases (IUPAC codes may match anything) are
// left out of the fingerprint comparison
template <bool MASK_INV>
__global__ void __launch_bounds__(TBD_WARPS * 32, 16)
scan_tbdir_kernel(const __grid_constant__ BsgpuSubjectView S, const __grid_constant__ BsgpuIndexView I,
		  const __grid_constant__ BsgpuTbDirView D, const __grid_constant__ BsgpuReport R,
		  int64_t start_lo, int64_t start_hi, int max_nmis, int min_nmis, int fixedP, int fixedS)
{
	const unsigned FULL = 0xFFFFFFFFu;
	__shared__ uint64_t wc_all[TBD_WARPS][TBD_WORDS];
	__shared__ uint64_t wi_all[MASK_INV ? TBD_WARPS : 1][TBD_WORDS];
	__shared__ uint32_t mw_all[TBD_WARPS][32];
	__shared__ uint2 ring_all[TBD_WARPS][TBD_CAP];  // {slot index, window in the batch}
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint64_t *wc = wc_all[warp], *wi = wi_all[MASK_INV ? warp : 0];
	uint32_t *mw = mw_all[warp];
	uint2 *ring = ring_all[warp];
	const int64_t word_lo = start_lo >> 5, word_hi = (start_hi + 31) >> 5;
	const int64_t nwords_total = S.total >> 5;      // the planes hold total / 32 (+ spare) words
	const int64_t stride = (int64_t) gridDim.x * TBD_WARPS * 32;
	const int w = D.w, prefix = D.prefix;
	const uint32_t key_mask = (1u << (2 * prefix)) - 1u;
	const uint32_t lt = (1u << lane) - 1u;
	const uint64_t keep = l2_policy_evict_last();
	uint32_t head = 0, tail = 0;

	for (int64_t jb = word_lo + ((int64_t) blockIdx.x * TBD_WARPS + warp) * 32; jb < word_hi; jb += stride) {
		const int64_t j = jb + lane, base = j << 5, base0 = jb << 5;
		uint32_t m = 0;

		__syncwarp();
		// words jb-1 .. jb+36 (beyond the buffer: zeros, never part of a window that is scanned)
		for (int k = lane; k < TBD_WORDS; k += 32) {
			const int64_t x = jb - 1 + k;
			const bool in = x >= 0 && x < nwords_total;

			wc[k] = in ? __ldcs(S.codes + x) : 0ull;
			if (MASK_INV)
				wi[k] = in ? __ldcs(S.inv2 + x) : ~0ull;
		}
		if (j < word_hi) {
			uint64_t v = (uint64_t) __ldcs(S.valid + j) | ((uint64_t) __ldcs(S.valid + j + 1) << 32);

			m = (uint32_t) all_set_run(v, w);
			if (base < start_lo)
				m &= 0xFFFFFFFFu << (int) (start_lo - base);
			if (base + 32 > start_hi)
				m &= (start_hi - base) <= 0 ? 0u : 0xFFFFFFFFu >> (int) (base + 32 - start_hi);
		}
		const bool all_on = __all_sync(FULL, m == 0xFFFFFFFFu);
		mw[lane] = m;
		__syncwarp();
#pragma unroll 1
		for (int g0 = 0; g0 <= 32; g0 += TBD_GROUP) {
			while (tail - head >= (g0 == 32 ? 1u : 32u * TBD_U)) {
				const uint32_t nround = min(tail - head, 32u * TBD_U);
				uint2 item[TBD_U];
				unsigned long long ent[TBD_U];
#pragma unroll
				for (int u = 0; u < TBD_U; u++) {
					const uint32_t k = (uint32_t) (u * 32 + lane);

					item[u] = ring[(head + k) & (TBD_CAP - 1)];
					ent[u] = k < nround ? ldg_keep_u64(D.slots + item[u].x, keep) : QS2_PTR;
				}
				head += nround;
				__syncwarp();
#pragma unroll
				for (int u = 0; u < TBD_U; u++) {
					const unsigned long long e = ent[u];
					const int pos = (int) item[u].y;

					if ((e & QS2_PTR) == 0ull) {
						const int nt = (int) (e >> TBD_NT_SHIFT) & 15, nh = (int) (e >> TBD_NH_SHIFT) & 15;
						const int nr = (int) (e >> TBD_NR_SHIFT) & 15;
						const uint32_t fp = (uint32_t) (e >> TBD_FP_SHIFT) & 0xFFFFu;
						// band letters after the prefix, tail letters from pos + w on, head letters
						// ending at pos - 1 (the batch's words start 32 letters before its first window)
						const int at_r = pos + prefix + 32, at_t = pos + w + 32, at_h = pos - nh + 32;
						uint32_t r_lo, t_lo, h_lo, hi_;

						q2_cut(wc[at_r >> 5], wc[(at_r >> 5) + 1], 2 * (at_r & 31), r_lo, hi_);
						q2_cut(wc[at_t >> 5], wc[(at_t >> 5) + 1], 2 * (at_t & 31), t_lo, hi_);
						q2_cut(wc[at_h >> 5], wc[(at_h >> 5) + 1], 2 * (at_h & 31), h_lo, hi_);
						const uint32_t dr = (r_lo ^ fp) & ((1u << (2 * nr)) - 1u);
						uint32_t dt = (t_lo ^ (fp >> (2 * nr))) & ((1u << (2 * nt)) - 1u);
						uint32_t dh = (h_lo ^ (fp >> (2 * (nr + nt)))) & ((1u << (2 * nh)) - 1u);

						dt = (dt | (dt >> 1)) & 0x5555u;
						dh = (dh | (dh >> 1)) & 0x5555u;
						if (MASK_INV) {
							uint32_t it_lo, ih_lo;

							q2_cut(wi[at_t >> 5], wi[(at_t >> 5) + 1], 2 * (at_t & 31), it_lo, hi_);
							q2_cut(wi[at_h >> 5], wi[(at_h >> 5) + 1], 2 * (at_h & 31), ih_lo, hi_);
							dt &= ~it_lo;
							dh &= ~ih_lo;
						}
						if (dr == 0u && __popc(dt) + __popc(dh) <= max_nmis)
							tbdir_survivor(&S, &I, &R, (int) ((uint32_t) e & ((1u << TBD_ID_BITS) - 1u)),
								       base0 + pos + w - 1, max_nmis, min_nmis, fixedP, fixedS);
					}
					const bool need = (e & (QS2_PTR | QS2_MORE)) != 0ull && (uint32_t) (u * 32 + lane) < nround;
					const uint32_t who = __ballot_sync(FULL, need);

					if (who != 0u) {
						if (need)
							ring[(tail + __popc(who & lt)) & (TBD_CAP - 1)] =
								make_uint2((e & QS2_PTR) ? (uint32_t) e : item[u].x + 1u, (uint32_t) pos);
						tail += __popc(who);
					}
				}
				__syncwarp();
			}
			if (g0 == 32)
				break;
			uint32_t key[TBD_GROUP];
			uint2 cell[TBD_GROUP];
			uint64_t c0 = wc[g0 + 1];
#pragma unroll
			for (int qq = 0; qq < TBD_GROUP; qq++) {
				const uint64_t c1 = wc[g0 + qq + 2];
				uint32_t w_lo, w_hi;

				q2_cut(c0, c1, 2 * lane, w_lo, w_hi);
				key[qq] = w_lo & key_mask;
				c0 = c1;
			}
#pragma unroll
			for (int qq = 0; qq < TBD_GROUP; qq++)
				cell[qq] = (all_on || ((mw[g0 + qq] >> lane) & 1u)) ? ldg_keep_u2(D.cells + (key[qq] >> 4), keep)
										     : make_uint2(0u, 0u);
#pragma unroll
			for (int qq = 0; qq < TBD_GROUP; qq++) {
				const int b2 = 2 * (int) (key[qq] & 15u);
				const uint32_t c = (cell[qq].x >> b2) & 3u;
				const uint32_t b_lo = __ballot_sync(FULL, (c & 1u) != 0u), b_hi = __ballot_sync(FULL, (c & 2u) != 0u);

				if (c != 0u) {
					const uint32_t below = cell[qq].x & ((1u << b2) - 1u);
					const uint32_t first = cell[qq].y + __popc(below & 0x55555555u) + 2 * __popc(below & 0xAAAAAAAAu);
					const uint32_t at = tail + __popc(b_lo & lt) + 2 * __popc(b_hi & lt);
					const uint32_t p = (uint32_t) ((g0 + qq) * 32 + lane);

					ring[at & (TBD_CAP - 1)] = make_uint2(first, p);
					if (c > 1u)
						ring[(at + 1u) & (TBD_CAP - 1)] = make_uint2(first + 1u, p);
					if (c > 2u)
						ring[(at + 2u) & (TBD_CAP - 1)] = make_uint2(first + 2u, p);
				}
				tail += __popc(b_lo) + 2 * __popc(b_hi);
			}
			__syncwarp();
		}
	}
}
