// Host-side construction of the "pair table" of the full scan (xxm_scan5.cuh).
//
// The scan has to decide, for every window, whether its left k-mer (match columns 0..len0-1 of the window, at base offsets
// coloff[]) is in the p-value table of the current PWM (pValueHash[0][id] != 0, branch_and_bound-inl.h:386-392).  When
// those columns span at most 8 bases, the 8 bases split into the 4-mer x at offsets 0..3 and the 4-mer y at offsets 4..7,
// and membership is a 256 x 256 bit matrix M[x][y].  For a table enumerated by a score threshold (every table
// set_kmer_to_Pvalue_hash builds, :464-577) the rows of M are nested, i.e. M[x][y] = (rank(x) >= need(y)), and then two
// byte tables a[x], b[y] with  M[x][y] <=> a[x] + b[y] >= 128  represent it EXACTLY.  build_pair_table derives a and b from
// the enumerated ids alone, verifies the identity for all 65536 (x, y) (no assumption about how the ids were enumerated),
// and returns false when it does not hold, when more than 127 levels are needed or when the span exceeds 8 bases; the
// caller then uses the generic scan kernel.
//
// Table layout: entry q (10 bits = 5 consecutive bases, base 0 least significant) describes the two 4-mers that start at
// base 0 (x0 = q & 255) and base 1 (x1 = (q >> 2) & 255):  E[q] = a[x0] << 24 | b[x0] << 16 | a[x1] << 8 | b[x1].
#pragma once
#include <cstdint>
#include <vector>

namespace xxm {

constexpr int kPairEntries = 1024;

// ids: enumerated left k-mer ids that count as members (column c in bits [2c+1:2c]); coloff[c]: base offset of column c
bool build_pair_table(const uint32_t* ids, size_t nids, const int* coloff, int len0, uint32_t* table /*[kPairEntries]*/, int* levels);

}  // namespace xxm
