// bsgpu_tbdir.cuh -- Trusted Bands of <= 12 letters with heads / tails: the band's 2w bits ARE the
// address of a directory, so a window costs one 8-byte gather and a Trusted-Band hit one more.
//
// Replaces, for short bands, the hash probe of scan_tb_kernel followed by tb_hit(): with 10^6 keys
// over 4^12 bands 6 % of all windows hit a band by chance (config 3 of BASELINE: 36-mers, Trusted
// Band 1..12, max.mismatch 2), and every such hit walked bucket -> key -> group -> flank lengths ->
// packed flanks -> subject words, six dependent gathers in a divergent loop.  Here (the layout of
// qscan2_kernel, bsgpu_qindex.cuh):
//   cells : 8 bytes per 16 bands = 16 two-bit slot counts + the index of the cell's first slot
//   slots : 8 bytes per KEY (the keys sharing a band sit side by side, src/match_pdict_utils.c:
//           153-181): key id, and a 16-bit fingerprint of the key's flanks -- the first <= 8 tail
//           letters, then the last letters of the head -- compared with the subject letters around
//           the window by XOR + popcount (a lower bound of the head + tail mismatches of
//           src/match_pdict_utils.c:183-214); only what passes reads the key's packed flanks.
// Same warp organisation as qscan2_kernel: lane l owns the windows l, l + 32, ... of a batch of
// 1024, slots go through a ring in shared memory, the ring is drained 32 x U loads at a time.
#pragma once
#include "bsgpu_qindex.cuh"

#define TBD_MAX_W 12
#define TBD_ID_BITS 24
#define TBD_FP_SHIFT 24         // 16 bits: 8 letters
#define TBD_NT_SHIFT 40         // 4 bits: tail letters in the fingerprint
#define TBD_NH_SHIFT 44         // 4 bits: head letters in the fingerprint (after the tail's)

struct BsgpuTbDirView {
	const uint2 *cells;                     // [4^w / 16] {16 x 2-bit slot counts, first slot}
	const unsigned long long *slots;
	int32_t w;
};

// one key of a band that passed the fingerprint: head + tail by XOR + popcount on the planes
__device__ __noinline__ void tbdir_survivor(const BsgpuSubjectView *S, const BsgpuIndexView *I, const BsgpuReport *R,
		int key, int64_t e, int max_nmis, int min_nmis)
{
	int seg = S->nseg == 1 ? 0 : find_segment(*S, e);
	int64_t lo = __ldg(S->seg_start + seg), L = __ldg(S->seg_len + seg);
	uint32_t fl = __ldg(I->flank_len + key);
	int hl = (int) (fl & 63u), tl = (int) ((fl >> 6) & 63u);
	ulonglong2 f = __ldg(I->flank2 + key);
	int nmis = packed_flank_mismatches(*S, e - I->w - hl + 1, hl, f.x, lo, L);

	if (nmis <= max_nmis)
		nmis += packed_flank_mismatches(*S, e + 1, tl, f.y, lo, L);
	if (nmis <= max_nmis && nmis >= min_nmis)
		report_match(*R, *S, key, seg, e - lo + 1, e, tl);
}

#define TBD_CAP 256
#define TBD_U 2
#define TBD_WARPS 2
#define TBD_GROUP 2
static_assert(TBD_GROUP * 96 + 32 * TBD_U <= TBD_CAP, "ring too small");

__global__ void __launch_bounds__(TBD_WARPS * 32, 16)
scan_tbdir_kernel(const __grid_constant__ BsgpuSubjectView S, const __grid_constant__ BsgpuIndexView I,
		  const __grid_constant__ BsgpuTbDirView D, const __grid_constant__ BsgpuReport R,
		  int64_t start_lo, int64_t start_hi, int max_nmis, int min_nmis)
{
	const unsigned FULL = 0xFFFFFFFFu;
	__shared__ uint64_t wc_all[TBD_WARPS][36];    // words jb-1 .. jb+33 (+1: a cut reads two words)
	__shared__ uint32_t mw_all[TBD_WARPS][32];
	__shared__ uint2 ring_all[TBD_WARPS][TBD_CAP];  // {slot index, window in the batch}
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint64_t *wc = wc_all[warp];
	uint32_t *mw = mw_all[warp];
	uint2 *ring = ring_all[warp];
	const int64_t word_lo = start_lo >> 5, word_hi = (start_hi + 31) >> 5;
	const int64_t stride = (int64_t) gridDim.x * TBD_WARPS * 32;
	const int w = D.w;
	const uint32_t key_mask = w >= 16 ? 0xFFFFFFFFu : (1u << (2 * w)) - 1u;
	const uint32_t lt = (1u << lane) - 1u;
	const uint64_t keep = l2_policy_evict_last();
	uint32_t head = 0, tail = 0;

	for (int64_t jb = word_lo + ((int64_t) blockIdx.x * TBD_WARPS + warp) * 32; jb < word_hi; jb += stride) {
		const int64_t j = jb + lane, base = j << 5, base0 = jb << 5;
		const bool ld = j < word_hi + 2;        // the planes have spare words after the end
		uint32_t m = 0;

		__syncwarp();
		wc[lane + 1] = ld ? __ldcs(S.codes + j) : 0ull;
		if (lane == 0)
			wc[0] = jb > 0 ? __ldcs(S.codes + jb - 1) : 0ull;
		if (lane == 31) {
			wc[33] = j + 1 < word_hi + 2 ? __ldcs(S.codes + j + 1) : 0ull;
			wc[34] = 0ull;
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
						const uint32_t fp = (uint32_t) (e >> TBD_FP_SHIFT) & 0xFFFFu;
						// tail letters from pos + w on, head letters ending at pos - 1 (the batch's
						// words start 32 letters before its first window)
						const int at_t = pos + w + 32, at_h = pos - nh + 32;
						uint32_t t_lo, t_hi, h_lo, h_hi;

						q2_cut(wc[at_t >> 5], wc[(at_t >> 5) + 1], 2 * (at_t & 31), t_lo, t_hi);
						q2_cut(wc[at_h >> 5], wc[(at_h >> 5) + 1], 2 * (at_h & 31), h_lo, h_hi);
						uint32_t dt = (t_lo ^ fp) & ((1u << (2 * nt)) - 1u);
						uint32_t dh = (h_lo ^ (fp >> (2 * nt))) & ((1u << (2 * nh)) - 1u);

						dt = (dt | (dt >> 1)) & 0x5555u;
						dh = (dh | (dh >> 1)) & 0x5555u;
						if (__popc(dt) + __popc(dh) <= max_nmis)
							tbdir_survivor(&S, &I, &R, (int) ((uint32_t) e & ((1u << TBD_ID_BITS) - 1u)),
								       base0 + pos + w - 1, max_nmis, min_nmis);
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
