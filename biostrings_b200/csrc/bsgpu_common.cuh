// bsgpu_common.cuh -- shared host/device definitions of the PDict matching path.
//
// Data layout in HBM:
//   subject  : bytes[total] (Biostrings codes, segments separated by one 0x80 byte, with
//              0x80 guards at both ends), codes[total/32] (uint64, 2 bits per letter,
//              letter i of a word at bits 2i..2i+1, A=0 C=1 G=2 T=3, 0 for non-base),
//              valid[total/32] (uint32, bit i = letter is A/C/G/T), iupac[total/32]
//              (uint32, bit i = letter code in 1..15), inv2[total/32] (uint64, bit 2i =
//              letter i is not a base letter).
//   index    : bucketed hash table, 16 B per bucket = two {fingerprint, id} slots, keyed
//              by the first min(w,32) letters of the Trusted Band packed the same way.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#define BSGPU_EMPTY_ID 0x3FFFFFFFu       // id field of an empty slot (ids are < 2^30 - 1)
#define BSGPU_ID_MASK 0x3FFFFFFFu
#define BSGPU_OVERFLOW_FLAG 0x80000000u  // on slot 1: a key of this bucket lives further on
#define BSGPU_SEP_BYTE 0x80u        // never matches anything in any of the 4 match tables
#define BSGPU_GUARD 128             // guard letters in front of / behind the subject
#define BSGPU_POS_BITS 34           // low bits of a 64-bit record = position / coordinate
#define BSGPU_POS_MASK ((1ull << BSGPU_POS_BITS) - 1)
#define BSGPU_SEGKEY_BITS 30        // (segment << 30 | key) records for set subjects

// 2-bit code of a base letter (A=1,C=2,G=4,T=8 -> 0..3), src/utils.c:55-70 with
// base_codes = c(1,2,4,8) (R/PDict-class.R:98,157).  Only call on base letters.
__host__ __device__ __forceinline__ uint32_t bsgpu_base_code(uint32_t b)
{
	return (b >> 1) - (b >> 3);
}

__host__ __device__ __forceinline__ bool bsgpu_is_base(uint32_t b)
{
	return b == 1 || b == 2 || b == 4 || b == 8;
}

// Multiply-shift hash of a <=64-bit TB key: bucket = top bits of key * C mod 2^64,
// fingerprint = a 32-bit mix of the rest.  Every fingerprint match is confirmed against
// the stored key, so the fingerprint only has to be cheap and well spread.
#define BSGPU_HASH_C1 0x9E3779B1u
#define BSGPU_HASH_C2 0x85EBCA77u

__host__ __device__ __forceinline__ void bsgpu_hash(uint64_t key, uint32_t &top, uint32_t &fp)
{
	uint32_t lo = (uint32_t) key, hi = (uint32_t) (key >> 32);
	uint64_t p = (uint64_t) lo * BSGPU_HASH_C1;
	uint32_t p_lo = (uint32_t) p;
	uint32_t p_hi = (uint32_t) (p >> 32) + lo * BSGPU_HASH_C2 + hi * BSGPU_HASH_C1;
	top = p_hi;                     // bucket = top >> (32 - log2(nbuckets))
	fp = p_lo ^ hi;
}

// How a match is reported (the reference's MatchBuf / vcount drivers,
// src/match_pdict_utils.c:87-116, src/match_pdict.c:85-127,320-432).
enum {
	BSGPU_R_COUNT = 0,      // counts[key] += 1
	BSGPU_R_HIT_COORD = 1,  // record key<<34 | end coordinate       (dedupe-able)
	BSGPU_R_HIT_TBPOS = 2,  // record key<<34 | concat index of TB end (order-preserving)
	BSGPU_R_VMAT = 3,       // mat[seg*M + key] += 1                 (vcount, collapse=0)
	BSGPU_R_VC1 = 4,        // ans[key] += weight[seg]               (collapse=1, int)
	BSGPU_R_VC2 = 5,        // ans[seg] += weight[key]               (collapse=2, int)
	BSGPU_R_SEGKEY = 6      // record seg<<30 | key                  (vwhich / double weights)
};

struct BsgpuSubjectView {
	const uint8_t *bytes;
	const uint64_t *codes;
	const uint32_t *valid;
	const uint32_t *iupac;
	const uint64_t *inv2;        // bit 2i = letter i is not A/C/G/T (2-bit layout)
	const int64_t *seg_start;    // [nseg] concat index of the first letter
	const int32_t *seg_len;      // [nseg]
	const int64_t *seg_coord;    // [nseg] reported end = local end + seg_coord
	int32_t nseg;
	int64_t total;               // letters in the concat buffer (multiple of 128)
};

struct BsgpuIndexView {
	const uint4 *table;          // [nbuckets] {fp0, id0, fp1, id1 | overflow flag}
	const uint64_t *keys;        // [M] seed key of pattern id (valid for TB leaders)
	const int32_t *seed_next;    // w > 32: next TB leader sharing the 32-letter seed, or -1
	uint32_t bucket_shift;       // 32 - log2(nbuckets)
	uint32_t bucket_mask;        // nbuckets - 1
	int32_t w;                   // TB width
	int32_t seed_w;              // min(w, 32)
	uint64_t seed_mask;          // low 2*seed_w bits
	int32_t M;
	const int32_t *grp_off;      // [M+1] CSR: TB id -> keys sharing its TB (leader first)
	const int32_t *grp_keys;
	const uint8_t *tb_bytes;     // [M*w] TB letters (used when w > 32)
	const uint8_t *head;         // concat head letters
	const int64_t *head_off;     // [M+1]
	const uint8_t *tail;
	const int64_t *tail_off;     // [M+1]
	int32_t all_singletons;      // every TB group has exactly one key
	// head / tail of every key 2-bit packed (letter i at bits 2i) for the XOR + popcount check;
	// flank_len = |head| | |tail| << 6 | (1 << 15 if both are base-only and <= 32 letters)
	const ulonglong2 *flank2;    // [M] {head, tail}
	const uint16_t *flank_len;   // [M]
};

struct BsgpuReport {
	int32_t mode;
	int32_t M;
	int32_t *counts;                 // R_COUNT / R_VMAT / R_VC1 / R_VC2 target
	const int32_t *weight;           // R_VC1 / R_VC2
	unsigned long long *records;     // record buffer
	unsigned long long *n_records;   // append cursor (keeps counting past capacity)
	unsigned long long capacity;
};
