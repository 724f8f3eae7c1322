// bsgpu_oligo.cuh -- oligonucleotideFrequency on the packed subject planes.
//
//   oligo_freq_kernel     counts every width-letter window made of base letters only, at the
//                         0-based starts that are multiples of `step` inside its sequence
//                         (replaces update_int_oligo_freqs / update_double_oligo_freqs,
//                          src/letter_frequency.c:183-267, i.e. the rolling
//                          _shift_twobit_signature walk of src/utils.c:123-146)
//
// The reference rolls a 2-bit signature along the sequence and resets it at every non-base
// letter; all three of its step regimes (step = 1, 1 < step < width, step >= width) count the
// same thing: window [s, s + width) is counted iff s % step == 0, s + width <= length and all
// of its letters are A/C/G/T.  On the 2-bit code plane a window is a 2*width-bit field at bit
// 2*s and "all base letters" is a run of width set bits in the validity plane, so one thread
// takes one 32-letter word of the planes, finds the valid window starts in it with width - 1
// shift-ANDs and visits only those.  Separator and guard letters are not base letters, hence
// no window straddles two sequences and the sequence of a window is the one holding its start.
//
// One table over everything with step = 1 (a DNAString, or a collapsed set: the common case)
// needs neither the sequence nor the position of a window, and the kernel was bound by
// instruction issue (about 40 instructions per window with run-time 64-bit shifts), so that
// case has its own fully unrolled path: the 48 letters a word's windows can touch are three
// 32-bit words; for the "first letter most significant" order they are digit-reversed ONCE
// per thread and pre-shifted by 2*(17 - width) bits, after which window i is the 2*width-bit
// field at the compile-time offset 2*(31 - i): one funnel shift and one AND per window.
#pragma once
#include "bsgpu_kernels.cuh"

struct BsgpuOligoParams {
	int32_t w;              // window width, 1..15 (src/utils.c:101-102)
	int32_t step;           // >= 1
	int32_t invert;         // fast.moving.side = "left": first letter least significant
	int32_t per_segment;    // 1: out[segment * 4^w + code], 0: one row
	int32_t coord_step;     // chunk subjects: starts are judged in subject coordinates
	int32_t replicas;       // copies of the table in shared memory (0 = global atomics)
	int32_t slice_bits;     // > 0 (width 8): blockIdx.y owns the codes with code >> slice_bits == blockIdx.y,
				// 2^slice_bits counters in shared memory; every slice scans the whole subject
	int32_t fast;           // step == 1 and one row: the unrolled path
	int64_t scan_lo, scan_hi;       // only windows whose last letter is in [scan_lo, scan_hi)
	int64_t nwords;
};

// offset of a window in the reference's table: sum of code(letter i) * 4^(w-1-i)
// (endianness 0, src/utils.c:135-142), or * 4^i when the two-bit order is inverted.
__device__ __forceinline__ uint32_t oligo_offset(uint32_t field, int w, int invert)
{
	if (invert)
		return field;
	uint32_t r = __brev(field) >> (32 - 2 * w);             // digits reversed, bits of a digit swapped

	return ((r & 0x55555555u) << 1) | ((r >> 1) & 0x55555555u);
}

// digits (bit pairs) of a 32-bit word in reverse order
__device__ __forceinline__ uint32_t oligo_rev16(uint32_t x)
{
	uint32_t r = __brev(x);

	return ((r & 0x55555555u) << 1) | ((r >> 1) & 0x55555555u);
}

template <typename CT>
__device__ __forceinline__ void oligo_bump(uint32_t code, const BsgpuOligoParams &P, uint32_t *sh_hist,
					   CT *__restrict__ out, int tid)
{
	if (P.slice_bits > 0) {
		if ((code >> P.slice_bits) == blockIdx.y)
			atomicAdd(&sh_hist[code & ((1u << P.slice_bits) - 1u)], 1u);
	} else if (P.replicas > 0) {
		atomicAdd(&sh_hist[code * P.replicas + (tid & (P.replicas - 1))], 1u);
	} else {
		atomicAdd(out + code, (CT) 1);
	}
}

template <typename CT>
__global__ void __launch_bounds__(1024)
oligo_freq_kernel(BsgpuSubjectView S, BsgpuOligoParams P, CT *__restrict__ out)
{
	extern __shared__ uint32_t sh_hist[];
	const uint32_t ncol = 1u << (2 * P.w);
	const uint32_t cmask = ncol - 1;
	const int tid = threadIdx.x;

	const uint32_t nsh = P.slice_bits > 0 ? 1u << P.slice_bits : ncol * (uint32_t) P.replicas;   // counters in shared memory
	if (nsh > 0) {
		for (uint32_t k = tid; k < nsh; k += blockDim.x)
			sh_hist[k] = 0;
		__syncthreads();
	}
	const int64_t stride = (int64_t) gridDim.x * blockDim.x;

	for (int64_t j = (int64_t) blockIdx.x * blockDim.x + tid; j < P.nwords; j += stride) {
		uint64_t vv = (uint64_t) __ldg(S.valid + j) | ((uint64_t) __ldg(S.valid + j + 1) << 32);
		uint64_t run = vv;

		for (int k = 1; k < P.w; k++)
			run &= vv >> k;
		uint32_t starts = (uint32_t) run;       // bit i: letters i .. i+w-1 are all base letters
		if (starts == 0)
			continue;
		uint64_t lo = __ldg(S.codes + j), hi = __ldg(S.codes + j + 1);
		int64_t p0 = j * 32;

		if (P.fast) {
			// windows whose last letter is in [scan_lo, scan_hi): starts i in [a, b)
			int64_t a = P.scan_lo - p0 - P.w + 1, b = P.scan_hi - p0 - P.w + 1;

			if (a > 0)
				starts = a >= 32 ? 0u : starts & (0xFFFFFFFFu << (int) a);
			if (b < 32)
				starts = b <= 0 ? 0u : starts & (0xFFFFFFFFu >> (32 - (int) b));
			if (starts == 0)
				continue;
			uint32_t q0, q1, q2;

			if (P.invert) {         // letter m at bits 2m: window i starts at bit 2i
				q0 = (uint32_t) lo;
				q1 = (uint32_t) (lo >> 32);
				q2 = (uint32_t) hi;
			} else {                // 48 letters reversed: window i at bit 2 * (31 - i) after the pre-shift
				uint32_t r0 = oligo_rev16((uint32_t) hi), r1 = oligo_rev16((uint32_t) (lo >> 32)),
					 r2 = oligo_rev16((uint32_t) lo);
				uint32_t s0 = 2u * (17u - (uint32_t) P.w);      // 4 .. 32

				q0 = __funnelshift_rc(r0, r1, s0);
				q1 = __funnelshift_rc(r1, r2, s0);
				q2 = __funnelshift_rc(r2, 0u, s0);
			}
			// either way window number i of `starts` is now the field at bit 2 * i
			if (!P.invert)
				starts = __brev(starts);
#pragma unroll
			for (int i = 0; i < 32; i++) {
				if (!(starts & (1u << i)))
					continue;
				uint32_t f = i < 16 ? __funnelshift_r(q0, q1, 2 * (i & 15))
						    : __funnelshift_r(q1, q2, 2 * (i & 15));

				oligo_bump<CT>(f & cmask, P, sh_hist, out, tid);
			}
			continue;
		}
		int seg = 0;
		bool seg_known = S.nseg == 1;

		while (starts != 0) {
			int i = __ffs(starts) - 1;

			starts &= starts - 1;
			int64_t p = p0 + i, e = p + P.w - 1;
			if (e < P.scan_lo || e >= P.scan_hi)
				continue;
			if (!seg_known) {
				seg = find_segment(S, p);
				seg_known = true;
			} else {
				while (seg + 1 < S.nseg && __ldg(S.seg_start + seg + 1) <= p)
					seg++;
			}
			if (P.step > 1) {
				int64_t local = p - __ldg(S.seg_start + seg);

				if (P.coord_step)
					local += __ldg(S.seg_coord + seg);
				if (local % P.step != 0)
					continue;
			}
			uint64_t f = i == 0 ? lo : (lo >> (2 * i)) | (hi << (64 - 2 * i));
			uint32_t code = oligo_offset((uint32_t) f & cmask, P.w, P.invert);

			if (P.replicas > 0)
				atomicAdd(&sh_hist[code * P.replicas + (tid & (P.replicas - 1))], 1u);
			else
				atomicAdd(out + (P.per_segment ? (size_t) seg * ncol : (size_t) 0) + code, (CT) 1);
		}
	}
	if (P.slice_bits > 0) {
		__syncthreads();
		for (uint32_t c = tid; c < nsh; c += blockDim.x)
			if (sh_hist[c] != 0)
				atomicAdd(out + ((size_t) blockIdx.y << P.slice_bits) + c, (CT) sh_hist[c]);
	} else if (P.replicas > 0) {
		__syncthreads();
		for (uint32_t c = tid; c < ncol; c += blockDim.x) {
			uint32_t sum = 0;

			for (int r = 0; r < P.replicas; r++)
				sum += sh_hist[c * P.replicas + r];
			if (sum != 0)
				atomicAdd(out + c, (CT) sum);
		}
	}
}
