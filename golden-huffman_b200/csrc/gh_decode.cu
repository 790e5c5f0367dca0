// K5-K7 -- decode of the reference's unindexed serial bit string (replaces reference
// include/canonical_huff_encoder.cc:377-419 CanonicalHuffDecoder::decode_file; result-identical to the Fast
// and Table decoders :422-461, :519-568).
//
// The stream has no block boundaries, so codeword boundaries are *discovered* by self-synchronisation
// (Klein & Wiseman 2003; Weissenberger & Schmidt 2018), arranged here as a fixed-point iteration:
//
//   payload = subsequences of S bytes, one thread each (S a multiple of 128, chosen so the grid fills the GPU).
//   state[i] = (entry_i, count_i, exit_i, eof_i): decoding subsequence i from bit `entry_i` yields count_i symbols
//              and leaves the subsequence `exit_i` bits into subsequence i+1 (or hits the end mark).
//   K5a  speculate : every thread decodes its subsequence from entry 0 (entry_0 is the true one).
//   K5b  sync round: thread i compares entry_i with exit_{i-1}. If they differ it walks BOTH paths in lockstep
//                    (always stepping the one that is behind) until they meet -- from there on the old result
//                    is still valid, so only the few symbols up to the meeting point are re-decoded; if they
//                    never meet inside the subsequence its exit changes and the round is marked dirty.
//                    Rounds repeat until one is clean; the fixed point is exactly the serial decoder's path
//                    because entry_0 is exact and every other entry equals its predecessor's exit.
//                    Round k makes subsequences 0..k exact whatever the data, so termination and exactness are
//                    unconditional; codes that synchronise slowly (near-fixed-length ones) just take more rounds,
//                    each touching fewer subsequences.
//   K6   offsets   : the first end mark on the synchronised path truncates the counts (one thread walks that one
//                    subsequence to count the symbols before it); exclusive scan -> output offsets.
//                    Codes that need about a whole subsequence to re-synchronise take re-walk rounds over
//                    dense work lists instead (dec_worklist_kernel), and the 8/9-bit code of equally frequent
//                    bytes takes no rounds at all: K5c, the phase walk (transfer functions composed by a scan).
//   K7   write     : every thread re-decodes its subsequence from its now-exact entry, up to 2 codewords per
//                    12-bit table lookup (long codes: left-justified first_code search, as the reference's Fast
//                    decoder does), and stores symbols 16 at a time with 128-bit stores.
// K5a and K7 share the "cursor reader" (below): table entries are addends for one cursor word, lanes read their
// subsequences in 32-byte units (one sector per lane and request) in a word-synchronous loop.
// The lookup tables are expanded from the header's canonical tables by a kernel (dec_build_luts_kernel).
//
// Algorithmic traffic: C bytes read + N bytes written; this implementation reads the payload twice
// (speculate + write), which the roofline accounting in bench.py does NOT credit.
#include "gh_common.cuh"

namespace gh {

// lutW entry: 64 bits (addend, 4 symbols) or, with GH_LUTW_U32, 32 bits (addend << 16 | 2 symbols): half the
// shared-memory wavefronts per lookup for slightly fewer symbols per lookup
#ifndef GH_LUTW_U64
#define GH_LUTW_U32 1
#endif
#ifdef GH_LUTW_U32
typedef uint32_t LutWEntry;
constexpr int kLutWEntryShift = 2;
constexpr int kLutWSyms = 2;
#else
typedef uint2 LutWEntry;
constexpr int kLutWEntryShift = 3;
constexpr int kLutWSyms = kLutWMaxSyms;
#endif

constexpr int kDecThreads = 256;   // threads per block = subsequences per tile, all decode kernels
#ifndef GH_DEC_S_BLOCKS
#define GH_DEC_S_BLOCKS 6
#endif
constexpr u32 kNoEof = 0xffffffffu;
constexpr u32 kEofPosUnknown = 0xfffffffeu;
constexpr u32 kMinSubBytes = 128;
constexpr u32 kMaxSubBytes = 1u << 27;

// per-subsequence state, one 64-bit word so it is read and written atomically
//   [31:0] count   [47:32] entry   [55:48] exit   [56] has_eof
// A path is the codeword chain that starts at bit `entry` of the subsequence. It runs THROUGH end marks (a
// mis-phased path meets bit patterns that look like the end mark all the time; stopping there would throw away
// the exit it needs to synchronise). `count` is the number of codewords on the path, end marks included, and a
// second array holds how many of them are end marks: both are additive, so when a corrected path meets the
// stored one the stored tail can always be reused. `exit` is where the path leaves the subsequence.
__host__ __device__ inline u64 pack_state(u32 count, u32 entry, u32 exit, u32 eof) {
  return u64(count) | (u64(entry & 0xffffu) << 32) | (u64(exit & 0xffu) << 48) | (u64(eof & 1u) << 56);
}
__host__ __device__ inline u32 st_count(u64 s) { return u32(s); }
__host__ __device__ inline u32 st_entry(u64 s) { return u32(s >> 32) & 0xffffu; }
__host__ __device__ inline u32 st_exit(u64 s) { return u32(s >> 48) & 0xffu; }
__host__ __device__ inline u32 st_eof(u64 s) { return u32(s >> 56) & 1u; }

struct DecControl {  // device-resident, copied back to the host after each round
  u32 changed;
  u32 eof_index;  // first subsequence (in order) whose path contains the end mark
  u64 total;      // symbols before that end mark
  u32 exit_bit;
  u32 eof_found;
  u32 sub_bytes;
  u32 exits_changed;  // subsequences whose exit moved in the last round
  u64 n_sub;
  u32 eof_prefix;     // symbols before the first end mark inside subsequence eof_index
  u32 fine;           // 1: fine-grained pipeline (2 KiB segments, piece states valid)
  u32 work_count;     // entries appended to the next re-walk list in this round
  u32 phase;          // 1: synchronised by the phase walk (8/9-bit codes)
};

struct DecWorkspace {
  const DecodeTables* tables;  // small canonical tables (host-built)
  const uint16_t* lut1;        // [2^12]  device-built, see gh_internal.h
  const uint16_t* lutC;        // [2^kLutCBits]
  const LutWEntry* lutW;       // [2^kLutWBits]
  const u32* lutP;             // [2^12]
  DecControl* ctl;
  u64* sub;        // [n_sub]
  u32* neof;       // [n_sub]  end-mark codewords on each subsequence's path
  u32* eofpos;     // [n_sub]  symbols before the path's first end mark, kEofPosUnknown if it was not observed
  u64* out_off;    // [n_sub]  output offset of each subsequence's first symbol
  u32* pieces;     // [32 * n_sub] fine pipeline only: state of each 64-byte piece
  u32* work[2];    // [n_sub] each: subsequences to re-walk in this / the next round (near-fixed-length codes)
  // phase walk (8/9-bit codes), subsequences of >= kPhaseMinSub bytes: per subsequence the transfer function
  // entry -> exit (9 x 4 bits) and, per entry, codeword count / first end mark / end marks
  u64* ph_fn;
  u32* ph_cnt;
  u32* ph_first;
  u32* ph_neof;
  u64* ph_tile_fn;   // composition over kPhaseTile subsequences
  u32* ph_tile_entry;
  u64* tile_sum;   // [n_tiles]
  u64* tile_base;  // [n_tiles]
};

struct DecGeometry {
  const uint8_t* payload;
  u64 slice_bits;  // bits that belong to this slice
  u64 readable;    // bytes that may be read from `payload`
  u32 sub_bytes;
  u64 n_sub;
  u32 entry0;
};

// ---- canonical decode rule on the small tables ---------------------------------------------------------------
struct SmemCanon {
  u32 first_code_lj[34];
  u32 start_pos[34];
  uint16_t symbol[GH_NSYM + 3];
  u32 min_len, max_len;
};

__device__ __forceinline__ void load_canon(SmemCanon& s, const DecodeTables* __restrict__ g) {
  for (unsigned i = threadIdx.x; i < 34; i += blockDim.x) {
    s.first_code_lj[i] = g->first_code_lj[i];
    s.start_pos[i] = g->start_pos[i];
  }
  for (unsigned i = threadIdx.x; i < GH_NSYM + 3; i += blockDim.x) s.symbol[i] = g->symbol[i];
  if (threadIdx.x == 0) {
    s.min_len = g->min_len;
    s.max_len = g->max_len;
  }
}

// Smallest len in [from, max_len] with window >= first_code[len] << (32 - len): the reference's rule
// (include/canonical_huff_encoder.cc:396-402) on the left-justified table (its Fast decoder, :437-453 / cfind).
__device__ __forceinline__ void canon_search(const SmemCanon& s, u32 window, u32 from, u32& sym, u32& len) {
  len = from;
  const u32 max_len = s.max_len;
  while (len < max_len && window < s.first_code_lj[len]) ++len;
  u32 idx = s.start_pos[len] + ((window - s.first_code_lj[len]) >> (32 - len));
  idx = idx < u32(GH_NSYM) ? idx : u32(GH_NSYM - 1);
  sym = s.symbol[idx];
}

// One codeword from the next 32 stream bits (left-justified in `window`) through the 12-bit table.
__device__ __forceinline__ void decode_one(const SmemCanon& s, const uint16_t* lut1, u32 window, u32& sym, u32& len) {
  const u32 e = lut1[window >> (32 - kLut1Bits)];
  len = e & 63u;
  sym = e >> 6;
  if (len == 0) canon_search(s, window, kLut1Bits + 1, sym, len);
}

// out-of-line variant for the bulk loops' miss paths (rare; keeps the unrolled loop bodies small): (symbol << 8) | length
__device__ __noinline__ u32 decode_one_packed(const SmemCanon* s, const uint16_t* lut1, u32 window) {
  u32 sym, len;
  decode_one(*s, lut1, window, sym, len);
  return (sym << 8) | len;
}

// ---- K5 prologue: expand the canonical tables into the lookup tables, on the device -------------------------
// thread w handles window value w of each table it is in range for
__global__ void __launch_bounds__(256)
dec_build_luts_kernel(const DecodeTables* __restrict__ tables, uint16_t* __restrict__ lut1, uint16_t* __restrict__ lutC,
                      LutWEntry* __restrict__ lutW, u32* __restrict__ lutP) {
  __shared__ SmemCanon s;
  load_canon(s, tables);
  __syncthreads();
  const u32 w = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 min_len = s.min_len;
  // walk whole codewords inside a k-bit window value v (left-justified), stopping at the end mark
  auto walk = [&](u32 v, int k, int max_syms, u32& total_len, u32& nsym, u32& packed) {
    total_len = 0, nsym = 0, packed = 0;
    u32 win = v << (32 - k);
    int avail = k;
    while (int(nsym) < max_syms) {
      if (avail < int(min_len)) break;
      u32 sym, len;
      canon_search(s, win, min_len, sym, len);
      if (int(len) > avail || sym == u32(GH_EOF_SYMBOL)) break;
      if (nsym < 4) packed |= (sym & 0xffu) << (8 * nsym);
      ++nsym;
      total_len += len;
      win = len < 32 ? win << len : 0u;
      avail -= int(len);
    }
  };
  u32 tl, ns, pk;
  if (w < (1u << kLutCBits)) {
    walk(w, kLutCBits, 64, tl, ns, pk);
    lutC[w] = uint16_t(ns ? ((ns << kCurShift) - tl) : kLutMiss);
  }
  if (w < (1u << kLutWBits)) {
    walk(w, kLutWBits, kLutWSyms, tl, ns, pk);
#ifdef GH_LUTW_U32
    lutW[w] = ns ? ((((ns << (kCurShift + 3)) - tl) << 16) | pk) : (kLutMiss << 16);
#else
    lutW[w] = ns ? make_uint2((ns << (kCurShift + 3)) - tl, pk) : make_uint2(kLutMiss, 0u);
#endif
  }
  if (w < (1u << kLutPBits)) {
    walk(w, kLutPBits, 2, tl, ns, pk);
    u32 s1, l1 = 0;
    if (ns) canon_search(s, w << (32 - kLutPBits), min_len, s1, l1);
    lutP[w] = ns ? (tl | (ns << 4) | (l1 << 6) | ((pk & 0xffffu) << 16)) : 0u;
  }
  if (w < (1u << kLut1Bits)) {
    // single codeword, end mark included (as symbol 256); 0 when it does not fit in 12 bits
    u32 sym, len;
    canon_search(s, w << (32 - kLut1Bits), min_len, sym, len);
    lut1[w] = uint16_t(len <= u32(kLut1Bits) ? ((sym << 6) | len) : 0u);
  }
}

// ---- MSB-first bit reader over global memory --------------------------------------------------------------
// Each lane streams its own subsequence with 128-bit loads. The four words of the current vector sit in
// registers and are rotated as they are consumed, so the refill is straight-line code (a few predicated
// instructions) instead of a chain of branches: with 32 lanes refilling at different symbols, some lane needs
// a refill on almost every iteration, so whatever the refill costs is paid by the whole warp every time.
__device__ __noinline__ uint4 fetch_tail(const uint8_t* bytes, u64 readable, u64 v) {
  u32 w[4];
  for (int k = 0; k < 4; ++k) {
    u32 x = 0;
    for (int b = 0; b < 4; ++b) {
      const u64 idx = v * 16 + u64(k * 4 + b);
      const u32 byte = idx < readable ? u32(bytes[idx]) : 0xffu;  // past the end: the reference's 1-padding
      x |= byte << (8 * b);
    }
    w[k] = x;
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

struct BitReader {
  const uint8_t* bytes;
  u64 readable;
  u64 full_vecs;  // vectors that lie entirely inside the readable range
  u64 vi;         // index of the vector held in `ahead`
  uint4 ahead;    // the vector after the one being consumed, requested one vector early to cover the latency
  u32 w0, w1, w2, w3;  // upcoming 32-bit words (stream order), w0 first
  u32 left;            // how many of them are still unread (1..4)
  u64 buf;             // next `avail` stream bits, left-justified
  int avail;           // kept >= 32 between symbols

  __device__ __forceinline__ uint4 fetch(u64 v) const {
    if (v < full_vecs) return ldg128(reinterpret_cast<const uint4*>(bytes) + v);
    return fetch_tail(bytes, readable, v);
  }
  __device__ __forceinline__ void load_next() {
    w0 = ahead.x, w1 = ahead.y, w2 = ahead.z, w3 = ahead.w;
    left = 4;
    ++vi;
    ahead = fetch(vi);
  }
  __device__ __forceinline__ void push() {
    buf |= u64(be32(w0)) << (32 - avail);
    avail += 32;
    w0 = w1, w1 = w2, w2 = w3;
    if (--left == 0) load_next();
  }
  __device__ __forceinline__ void seek(const uint8_t* base, u64 readable_bytes, u64 bitpos) {
    bytes = base;
    readable = readable_bytes;
    full_vecs = readable_bytes >> 4;
    vi = bitpos >> 7;
    ahead = fetch(vi);
    load_next();
    const u32 skip = u32(bitpos >> 5) & 3u;  // words of this vector that lie before bitpos
    for (u32 k = 0; k < skip; ++k) w0 = w1, w1 = w2, w2 = w3;
    left = 4 - skip;
    buf = 0;
    avail = 0;
    push();
    push();
    const u32 drop = u32(bitpos) & 31u;
    buf <<= drop;
    avail -= int(drop);
  }
  __device__ __forceinline__ u32 window() const { return u32(buf >> 32); }
  __device__ __forceinline__ void consume(u32 len) {
    buf <<= len;
    avail -= int(len);
    if (avail < 32) push();
  }
};

__device__ __forceinline__ u64 sub_end_bits(const DecGeometry& g, u64 i) {
  const u64 start = i * u64(g.sub_bytes) * 8;
  const u64 left = g.slice_bits - start;
  const u64 full = u64(g.sub_bytes) * 8;
  return left < full ? left : full;
}

// ---- cursor reader: the bulk loops of K5a and K7 -------------------------------------------------------------
// A lane holds two consecutive big-endian stream words (hi:lo) and ONE cursor word `acc` whose low ten bits F say
// where the next codeword starts:  F = 512 + s, where s = S0 - (bits of hi:lo already consumed) is exactly the
// funnel-shift distance that brings the next K-bit window, pre-scaled by the table's entry size, to the bottom of
// a register (SHF.R.W takes its distance modulo 32, so `acc` itself is the shift operand). A lookup is
//     x = (hi:lo) >> acc;   e = table[x & mask];   acc += e;                        (SHF, LOP3, LDS, IADD)
// because a table entry is stored as the addend  (what the kernel counts << 10) - (bits consumed).  Lookups are
// allowed while s is in [0, 32), i.e. F in [512, 544), i.e. (acc & 0x1E0) == 0 -- one LOP3 -- and the loop around
// them is word-synchronous: every lane takes the next word of its 32-byte unit (one LDG.256 per lane) from a
// statically indexed register (F += 32), then looks up while allowed. F in [192, 512) means "window not inside
// hi:lo yet": that is the state after a lookup crossed into the next word, and it is also how a lane starts in
// the middle of a unit -- F starts up to nine words low and the first rotations are dummies, so the loop needs no
// per-word bounds tests. F never leaves [192, 576), so nothing ever borrows from or carries into the bits above.
// A table miss (first codeword longer than K bits, or the end mark) is the entry kLutMiss = 32: F lands in
// [544, 576) (bit 9 and bit 5 set), the lookup loop ends, and the codeword is resolved with the canonical search
// (reference include/canonical_huff_encoder.cc:437-453) on a 32-bit window built from hi, lo and the next word.
constexpr u32 kCurBase = 512u;
constexpr u32 kCurBusy = 0x1E0u;     // any of these bits set: no lookup now
constexpr u32 kCurMissBit = 0x200u;  // ... and this one set as well: the last lookup missed
constexpr u32 kCurFieldMask = 0x3ffu;
constexpr int kUnitWords = 8;        // a lane's reads are 32-byte units
constexpr int kLutMaxBits = kLutCBits > kLutWBits ? (kLutCBits > 12 ? kLutCBits : 12) : (kLutWBits > 12 ? kLutWBits : 12);  // builder grid: one thread per entry of the largest table

template <int K, int SC>  // K index bits, entries of (1 << SC) bytes
struct CursorGeom {
  static constexpr int S0 = 64 - K - SC;  // shift distance of the last window position inside hi:lo
  static constexpr int CMIN = S0 - 31;    // consumed bits of hi:lo at the first one
  static constexpr u32 kMask = ((1u << K) - 1u) << SC;
};

// Plans the bulk part of a walk from absolute bit `bit0` that must not take a K-bit window beyond `end_abs`:
// units u0..ulast are consumed whole; afterwards the position is cursor_position().
template <int K, int SC>
__device__ __forceinline__ bool cursor_plan(u64 bit0, u64 end_abs, u64 full_units, u64& u0, u64& ulast, u32& f0) {
  typedef CursorGeom<K, SC> G;
  if (end_abs < u64(K + G::S0) + 32 * kUnitWords || full_units < 2) return false;
  const u64 jmax = (end_abs - u64(K + G::S0)) >> 5;  // last word that may be `hi` while lookups are taken
  u64 ul = (jmax - (kUnitWords - 2)) / kUnitWords;   // its unit must be complete: word 8 * ul + 7 <= jmax + 1
  if (ul > full_units - 2) ul = full_units - 2;      // and the unit after it readable (miss path, prefetch)
  u0 = bit0 >> 8;
  if (ul < u0) return false;
  const int f = int(u32(bit0) & 255u);
  const int jrel = (f - G::CMIN) >> 5;  // word of u0 that is `hi` at the first lookup (-1: the word before u0)
  const int c = f - 32 * jrel;          // consumed bits of hi:lo at that point, CMIN .. CMIN + 31
  f0 = u32(int(kCurBase) + G::S0 - c - 32 * (jrel + 2));
  ulast = ul;
  return true;
}

template <int K, int SC>
__device__ __forceinline__ u64 cursor_position(u64 ulast, u32 acc) {
  return 32ull * (u64(kUnitWords) * ulast + u64(kUnitWords - 2)) +
         u64(CursorGeom<K, SC>::S0 + int(kCurBase) - int(acc & kCurFieldMask));
}

// the 32 stream bits at the cursor of a lane whose last lookup missed (F already reduced by kLutMiss)
template <int K, int SC>
__device__ __forceinline__ u32 cursor_window32(u32 hi, u32 lo, u32 next_be, u32 acc) {
  const u32 c = u32(CursorGeom<K, SC>::S0 + int(kCurBase) - int(acc & kCurFieldMask));  // CMIN .. S0
  return c < 32u ? __funnelshift_l(lo, hi, c) : __funnelshift_l(next_be, lo, c - 32u);
}

// All eight words of a unit to stream order at once, BEFORE the next unit's load is issued: ptxas may give both
// loads the same scoreboard, and then the first read of a unit's registers issued after the next load would wait
// for that load as well -- a full memory latency per unit (profiles/r2h: 12 % of the writer's stall samples).
__device__ __forceinline__ void unit_to_stream_order(Unit8& u) {
#pragma unroll
  for (int k = 0; k < kUnitWords; ++k) u.w[k] = be32(u.w[k]);
}

__device__ __forceinline__ bool lutc_hit(u32 e, u32& len, u32& cnt) {  // for the walks outside the bulk loops
  cnt = (e + ((1u << kCurShift) - 1u)) >> kCurShift;
  len = (cnt << kCurShift) - e;
  return e != kLutMiss;
}

// ---- K5a: speculative decode of every subsequence from bit 0 (subsequence 0: from the true entry) ---------
// Only counts are needed here, so the walk takes as many whole codewords per lookup as fit in 14 bits (kLutCBits).
struct SmemSpeculate {
  SmemCanon canon;
  uint16_t lut1[1 << kLut1Bits];
  uint16_t lutC[1 << kLutCBits];
};

__device__ __forceinline__ void load_speculate_tables(SmemSpeculate& s, const DecWorkspace& ws) {
  load_canon(s.canon, ws.tables);
  for (unsigned i = threadIdx.x; i < (1u << kLut1Bits) / 8; i += kDecThreads)
    reinterpret_cast<uint4*>(s.lut1)[i] = reinterpret_cast<const uint4*>(ws.lut1)[i];
  for (unsigned i = threadIdx.x; i < (1u << kLutCBits) / 8; i += kDecThreads)
    reinterpret_cast<uint4*>(s.lutC)[i] = reinterpret_cast<const uint4*>(ws.lutC)[i];
}

// Walks subsequence i from its entry (first pass: bit 0, or the true entry for subsequence 0; re-walk: the left
// neighbour's current exit) and stores its state. Returns whether a re-walk moved the exit.
__device__ __forceinline__ bool speculate_subsequence(SmemSpeculate& s, const DecGeometry& g, const DecWorkspace& ws, u64 i,
                                                      int respeculate) {

  const u64 start = i * u64(g.sub_bytes) * 8;
  const u32 end = u32(sub_end_bits(g, i));
  u32 entry = (i == 0) ? g.entry0 : 0u;
  u32 old_exit = 0xffffffffu;
  if (respeculate) {
    // synchronisation round by re-walking: for codes that need about a whole subsequence to re-synchronise the
    // lockstep walk of K5b buys nothing (the paths rarely meet early) and costs two slow walks; walking the
    // corrected path once with this kernel's bulk loop is several times cheaper. Same fixed-point iteration.
    const u64 mine = ld_volatile_u64(ws.sub + i);
    if (i > 0) entry = st_exit(ld_volatile_u64(ws.sub + i - 1));
    if (entry == st_entry(mine)) return false;
    old_exit = st_exit(mine);
  }
  u32 pos = entry, count = 0, neof = 0, first_eof = kNoEof;
  const bool bulk = end >= u32(kLutCBits);
  const u32 last = end - u32(kLutCBits);  // multi-codeword steps are allowed while pos <= last (only used if bulk)
  // (a) bulk: the cursor reader above. Nothing taken from the kLutCBits-bit table can cross `end`.
  {
    typedef CursorGeom<kLutCBits, 1> G;
    u64 u, ulast;
    u32 acc;
    if (bulk && pos < end && cursor_plan<kLutCBits, 1>(start + pos, start + end, g.readable >> 5, u, ulast, acc)) {
      const bool aligned32 = (reinterpret_cast<uintptr_t>(g.payload) & 31) == 0;
      const u64 umax = (g.readable >> 5) - 1;
      const smem_addr_t lut = smem_addr(s.lutC);
      u32 hi, lo = 0;
      // one 32-byte unit: eight word steps
      auto walk_unit = [&](const Unit8& cu, u32 next_unit_word0) {  // cu in stream order, the next unit's word raw
#pragma unroll
        for (int k = 0; k < kUnitWords; ++k) {
          hi = lo;
          lo = be32(cu.w[k]);
          acc += 32u;
          while ((acc & kCurBusy) == 0u) {
            do {
              const u32 x = __funnelshift_r(lo, hi, acc);
              acc += lds_u16(lut, x & G::kMask);
            } while ((acc & kCurBusy) == 0u);
            if ((acc & kCurMissBit) == 0u) break;
            // first codeword longer than the table window, or the end mark
            acc -= kLutMiss;
            const u32 next_be = be32(k + 1 < kUnitWords ? cu.w[(k + 1) % kUnitWords] : next_unit_word0);
            const u32 sl = decode_one_packed(&s.canon, s.lut1, cursor_window32<kLutCBits, 1>(hi, lo, next_be, acc));
            if ((sl >> 8) == u32(GH_EOF_SYMBOL)) {
              if (!neof) first_eof = count + (acc >> kCurShift);
              ++neof;
            }
            acc += (1u << kCurShift) - (sl & 0xffu);
          }
        }
        count += acc >> kCurShift;
        acc &= kCurFieldMask;
      };
      // two unit buffers that swap roles, so the unit in flight is never copied (a copy would wait for the load)
      Unit8 ua = ldg_unit(g.payload + 32 * u, aligned32), ub;
      while (true) {
        ub = ldg_unit(g.payload + 32 * (u + 1 < umax ? u + 1 : umax), aligned32);
        walk_unit(ua, ub.w[0]);
        if (u == ulast) break;
        ++u;
        ua = ldg_unit(g.payload + 32 * (u + 1 < umax ? u + 1 : umax), aligned32);
        walk_unit(ub, ua.w[0]);
        if (u == ulast) break;
        ++u;
      }
      pos = u32(cursor_position<kLutCBits, 1>(ulast, acc) - start);
    }
  }
  // (a') whatever the bulk loop left (payload tail), same steps through the bounds-checked reader
  BitReader r;
  r.seek(g.payload, g.readable, start + pos);
  while (bulk && pos <= last) {
    const u32 win = r.window();
    u32 len, cnt;
    if (lutc_hit(s.lutC[win >> (32 - kLutCBits)], len, cnt)) {
      count += cnt;
    } else {
      u32 sym;
      decode_one(s.canon, s.lut1, win, sym, len);
      if (sym == u32(GH_EOF_SYMBOL)) {
        if (!neof) first_eof = count;
        ++neof;
      }
      ++count;
    }
    pos += len;
    r.consume(len);
  }
  // (b) single codewords up to the crossing
  while (pos < end) {
    u32 sym, len;
    decode_one(s.canon, s.lut1, r.window(), sym, len);
    if (sym == u32(GH_EOF_SYMBOL)) {
      if (!neof) first_eof = count;
      ++neof;
    }
    ++count;
    pos += len;
    r.consume(len);
  }
  ws.neof[i] = neof;
  ws.eofpos[i] = first_eof;  // codewords before the first end mark are all symbols
  st_volatile_u64(ws.sub + i, pack_state(count, entry, pos - end, neof != 0));
  return respeculate && (pos - end) != old_exit;
}



__global__ void __launch_bounds__(kDecThreads, GH_DEC_S_BLOCKS)
dec_speculate_kernel(DecGeometry g, DecWorkspace ws, int respeculate, const u32* __restrict__ work_in, u32 work_n,
                     u32* __restrict__ work_out) {
  __shared__ SmemSpeculate s;
  load_speculate_tables(s, ws);
  __syncthreads();
  // first pass: thread = subsequence. Re-walk rounds: thread = entry of the dense list of subsequences whose left
  // neighbour's exit moved in the previous round (so the warps of a round are full whatever fraction is stale).
  const u64 slot = u64(blockIdx.x) * kDecThreads + threadIdx.x;
  const bool live = respeculate ? slot < u64(work_n) : slot < g.n_sub;
  const u64 i = live ? (respeculate ? u64(work_in[slot]) : slot) : 0;
  bool moved = false;
  if (live) moved = speculate_subsequence(s, g, ws, i, respeculate);
  if (!respeculate) return;
  // append i + 1 to the next round's list (one atomic per warp)
  const bool push = moved && i + 1 < g.n_sub;
  const unsigned mask = __ballot_sync(0xffffffffu, push);
  if (mask) {
    const unsigned lane = threadIdx.x & 31;
    u32 base = 0;
    if (lane == 0) base = atomicAdd(&ws.ctl->work_count, u32(__popc(mask)));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (push) work_out[base + __popc(mask & ((1u << lane) - 1u))] = u32(i + 1);
    if (lane == 0) {
      ws.ctl->changed = 1u;
      atomicAdd(&ws.ctl->exits_changed, u32(__popc(mask)));
    }
  }
}

// first list of a re-walk phase: every subsequence whose assumed entry is not its left neighbour's exit
__global__ void __launch_bounds__(kDecThreads)
dec_worklist_kernel(DecGeometry g, DecWorkspace ws, u32* __restrict__ work_out) {
  const u64 i = u64(blockIdx.x) * kDecThreads + threadIdx.x;
  bool push = false;
  if (i < g.n_sub) {
    const u32 want = (i == 0) ? g.entry0 : st_exit(ws.sub[i - 1]);
    push = want != st_entry(ws.sub[i]);
  }
  const unsigned mask = __ballot_sync(0xffffffffu, push);
  if (mask) {
    const unsigned lane = threadIdx.x & 31;
    u32 base = 0;
    if (lane == 0) base = atomicAdd(&ws.ctl->work_count, u32(__popc(mask)));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (push) work_out[base + __popc(mask & ((1u << lane) - 1u))] = u32(i);
  }
}

// ---- K5b: one synchronisation round ------------------------------------------------------------------------

__global__ void __launch_bounds__(kDecThreads)
dec_sync_kernel(DecGeometry g, DecWorkspace ws) {
  __shared__ SmemSpeculate s;
  load_speculate_tables(s, ws);
  __syncthreads();
  const u64 i = u64(blockIdx.x) * kDecThreads + threadIdx.x;
  u64 mine = 0;
  u32 want = 0;
  bool stale = false;
  if (i < g.n_sub) {
    mine = ld_volatile_u64(ws.sub + i);
    want = (i == 0) ? g.entry0 : st_exit(ld_volatile_u64(ws.sub + i - 1));
    stale = want != st_entry(mine);
  }
  bool merged = true, exit_moved = false;
  if (stale) {
    const u64 start = i * u64(g.sub_bytes) * 8;
    const u32 end = u32(sub_end_bits(g, i));
    // path A = the stored one (from st_entry(mine)), path B = the one we now believe in (from `want`).
    // Always step the one that is behind; where they meet, the rest of A's stored result is B's.
    u32 pos_a = st_entry(mine), pos_b = want;
    u32 steps_a = 0, steps_b = 0, eofs_a = 0, eofs_b = 0, first_b = kNoEof;
    merged = false;
    BitReader ra, rb;
    ra.seek(g.payload, g.readable, start + pos_a);
    rb.seek(g.payload, g.readable, start + pos_b);
    // Both paths take the same kind of step from a given position (as many whole codewords as fit in kLutCBits bits
    // while that window lies inside the subsequence, else one codeword), so equal positions mean equal futures.
    // A boundary that one path steps over can be a stepping point of the other: they then meet a little later,
    // or not at all inside the subsequence -- in which case B has simply been walked to its end, which is exact.
    const bool bulk = end >= u32(kLutCBits);
    const u32 last = end - u32(kLutCBits);
    while (pos_b < end) {
      if (pos_a == pos_b) {
        merged = true;
        break;
      }
      u32 sym, len, cnt;
      if (pos_a < pos_b) {
        const u32 win = ra.window();
        if (bulk && pos_a <= last && lutc_hit(s.lutC[win >> (32 - kLutCBits)], len, cnt)) {
          steps_a += cnt;
        } else {
          decode_one(s.canon, s.lut1, win, sym, len);
          ++steps_a;
          eofs_a += sym == u32(GH_EOF_SYMBOL);
        }
        pos_a += len;
        ra.consume(len);
      } else {
        const u32 win = rb.window();
        if (bulk && pos_b <= last && lutc_hit(s.lutC[win >> (32 - kLutCBits)], len, cnt)) {
          steps_b += cnt;
        } else {
          decode_one(s.canon, s.lut1, win, sym, len);
          if (sym == u32(GH_EOF_SYMBOL)) {
            if (!eofs_b) first_b = steps_b;
            ++eofs_b;
          }
          ++steps_b;
        }
        pos_b += len;
        rb.consume(len);
      }
    }
    u32 count, exit, neof, first = first_b;
    if (merged) {  // codeword and end-mark counts are additive: B's tail is A's tail
      exit = st_exit(mine);
      count = steps_b + (st_count(mine) - steps_a);
      const u32 tail_eofs = ws.neof[i] - eofs_a;
      neof = eofs_b + tail_eofs;
      if (!eofs_b && tail_eofs) {
        // the first end mark lies in A's tail: known only if it is also A's first and A had observed it
        const u32 a_first = ws.eofpos[i];
        first = (eofs_a == 0 && a_first < kEofPosUnknown) ? steps_b + (a_first - steps_a) : kEofPosUnknown;
      }
    } else {
      exit = pos_b - end;
      count = steps_b;
      neof = eofs_b;
    }
    ws.neof[i] = neof;
    ws.eofpos[i] = first;
    st_volatile_u64(ws.sub + i, pack_state(count, want, exit, neof != 0));
    exit_moved = exit != st_exit(mine);
    if (exit_moved) ws.ctl->changed = 1u;
  }
  // how many exits moved: after the first round the host uses the fraction to decide whether this code needs
  // coarser subsequences (an exit that depends on the entry means the paths did not meet inside the subsequence)
  const unsigned moved = __ballot_sync(0xffffffffu, exit_moved);
  if ((threadIdx.x & 31) == 0 && moved) atomicAdd(&ws.ctl->exits_changed, u32(__popc(moved)));
}

// ---- K5c: phase walk for codes of 8 and 9 bits only (uniform-looking bytes) ---------------------------------
// With lengths {8, 9} and the 9-bit codewords being exactly those whose first 8 bits are zero (first_code_[8] == 1:
// the reference's canonical convention makes the longest codes the numerically smallest), a path advances 8 bits
// per codeword and changes phase only where 8 zero bits start ON its phase. Such paths take thousands of symbols to
// meet (profiles: ~5.7k), so the fixed-point rounds of K5a/K5b degenerate -- ten rounds, each one serial walk of a
// whole subsequence. Here a thread instead finds the (sparse) positions of 8 zero bits in its subsequence with a
// few logic ops per 32-bit word, replays ALL nine possible entries 0..8 over those events at once, and stores the
// subsequence's transfer function entry -> exit (9 x 4 bits) with the per-entry counts. The functions are then
// composed by a scan (they are associative), which yields every subsequence's true entry with no rounds at all.
constexpr int kPhasePaths = 9;
constexpr u32 kPhaseMinSub = 1024;
constexpr int kPhaseTile = 64;
constexpr int kPhaseScanThreads = 1024;

__host__ __device__ inline u32 fn_get(u64 f, u32 e) { return u32(f >> (4 * e)) & 15u; }
__host__ __device__ inline u64 fn_identity() {
  u64 f = 0;
  for (u32 e = 0; e < u32(kPhasePaths); ++e) f |= u64(e) << (4 * e);
  return f;
}
__host__ __device__ inline u64 fn_then(u64 f, u64 g) {  // e -> g(f(e))
  u64 r = 0;
  for (u32 e = 0; e < u32(kPhasePaths); ++e) r |= u64(fn_get(g, fn_get(f, e))) << (4 * e);
  return r;
}

// Words that contain events (positions where eight zero bits start) are first collected as records in a per-thread
// list in shared memory and replayed over the nine paths in batches, all lanes together: handled where they are
// found, the events cost the whole warp the nine-path update -- or even just a find-first-set loop -- whenever ANY
// lane has one in its current word (almost always), lane after lane (measured: 1 of 32 lanes active there).
constexpr int kPhaseListLen = 12;  // records buffered per thread: 12 x 3 words x 256 threads x 4 B = 36 KiB per block

__global__ void __launch_bounds__(kDecThreads)
dec_phase_walk_kernel(DecGeometry g, DecWorkspace ws, u32 eof_v) {
  __shared__ u32 s_list[kPhaseListLen][3][kDecThreads];  // [record][field][thread]: bank = thread % 32
  const u64 i = u64(blockIdx.x) * kDecThreads + threadIdx.x;
  const bool live = i < g.n_sub;  // every lane stays for the warp-wide votes below
  const unsigned t = threadIdx.x;
  const u64 start = i * u64(g.sub_bytes) * 8;
  const u32 end = live ? u32(sub_end_bits(g, i)) : 0u;
  const bool collapsed = i == 0 && g.entry0 > 8u;  // the image's first codeword may sit far into its first sector
  u32 ready[kPhasePaths], nlong[kPhasePaths], neof[kPhasePaths], first[kPhasePaths];
#pragma unroll
  for (int p = 0; p < kPhasePaths; ++p) {
    ready[p] = collapsed ? g.entry0 : u32(p);
    nlong[p] = 0, neof[p] = 0, first[p] = kNoEof;
  }
  u32 n_rec = 0;
  auto replay = [&]() {
    for (u32 k = 0; k < n_rec; ++k) {
      u32 m = s_list[k][0][t];               // event positions of the word, stream position b = bit 31 - b
      const u32 ninth = s_list[k][1][t];     // stream positions 8 .. 39 of the word's window, same numbering
      const u32 base = s_list[k][2][t];
      while (m) {
        const u32 b = u32(__clz(int(m)));
        m &= ~(0x80000000u >> b);
        const u32 x = base + b;
        const u32 eofbit = (ninth >> (31u - b)) & 1u;  // the ninth bit of the codeword
#pragma unroll
        for (int p = 0; p < kPhasePaths; ++p) {
          if (((ready[p] ^ x) & 7u) == 0u && ready[p] <= x) {  // x is a codeword start of this path
            if (eofbit == eof_v) {
              if (!neof[p]) first[p] = (x - (collapsed ? g.entry0 : u32(p)) - nlong[p]) >> 3;
              ++neof[p];
            }
            ++nlong[p];
            ready[p] = x + 9u;
          }
        }
      }
    }
    n_rec = 0;
  };
  const u64 full_vecs = g.readable >> 4;
  const u64 v0 = start >> 7;
  const u32 nwords = (end + 31u) >> 5;
  u32 nwords_warp = nwords;  // warp-uniform loop bound: the replay below is entered by all lanes together
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const u32 other = __shfl_xor_sync(0xffffffffu, nwords_warp, d);
    nwords_warp = other > nwords_warp ? other : nwords_warp;
  }
  auto load_vec = [&](u64 v) -> uint4 {
    return v < full_vecs ? ldg128(reinterpret_cast<const uint4*>(g.payload) + v) : fetch_tail(g.payload, g.readable, v);
  };
  uint4 cur = load_vec(v0), nxt = load_vec(v0 + 1);
  for (u32 j0 = 0; j0 < nwords_warp; j0 += 4) {
    const uint4 ahead = load_vec(v0 + (j0 >> 2) + 2);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const u32 j = j0 + u32(k);
      if (j < nwords) {
        const u32 w_hi = be32(k == 0 ? cur.x : k == 1 ? cur.y : k == 2 ? cur.z : cur.w);
        const u32 w_lo = be32(k == 0 ? cur.y : k == 1 ? cur.z : k == 2 ? cur.w : nxt.x);
        const u64 V = (u64(w_hi) << 32) | w_lo;
        u64 r = ~V;         // ones where the stream has zeros; stream position b of this word is bit 63 - b
        r &= r << 1;        // ... and the next position too
        r &= r << 2;
        r &= r << 4;        // bit 63 - b set: positions b .. b + 7 are all zero
        u32 m = u32(r >> 32);
        const u32 base = j << 5;
        if (base + 32u > end) m &= ~(0xffffffffu >> (end - base));  // only codewords that start before `end`
        if (m) {
          s_list[n_rec][0][t] = m;
          s_list[n_rec][1][t] = u32(V >> 24);
          s_list[n_rec][2][t] = base;
          ++n_rec;
        }
      }
      // replay when ANY lane's list is full, all lanes together: entered lane by lane, the replay loop would run
      // once per lane instead of once per warp
      if (__any_sync(0xffffffffu, n_rec >= u32(kPhaseListLen))) replay();
    }
    cur = nxt;
    nxt = ahead;
  }
  replay();
  if (!live) return;
  u64 fn = 0;
#pragma unroll
  for (int p = 0; p < kPhasePaths; ++p) {
    u32 q = ready[p];
    if (q < end) q += ((end - q + 7u) >> 3) << 3;
    const u32 e_p = collapsed ? g.entry0 : u32(p);
    fn |= u64(q - end) << (4 * p);
    ws.ph_cnt[i * kPhasePaths + p] = q > e_p ? (q - e_p - nlong[p]) >> 3 : 0u;
    ws.ph_first[i * kPhasePaths + p] = first[p];
    ws.ph_neof[i * kPhasePaths + p] = neof[p];
  }
  ws.ph_fn[i] = fn;
}

__global__ void __launch_bounds__(kDecThreads)
dec_phase_tiles_kernel(DecGeometry g, DecWorkspace ws) {
  const u64 t = u64(blockIdx.x) * kDecThreads + threadIdx.x;
  const u64 first = t * kPhaseTile;
  if (first >= g.n_sub) return;
  u64 f = fn_identity();
  for (u64 i = first; i < first + kPhaseTile && i < g.n_sub; ++i) f = fn_then(f, ws.ph_fn[i]);
  ws.ph_tile_fn[t] = f;
}

// one block: scan of the tile functions -> entry of every tile
__global__ void __launch_bounds__(kPhaseScanThreads)
dec_phase_scan_kernel(DecGeometry g, DecWorkspace ws) {
  __shared__ u64 s_f[kPhaseScanThreads];
  const unsigned t = threadIdx.x;
  const u64 n_tiles = (g.n_sub + kPhaseTile - 1) / kPhaseTile;
  const u64 per = (n_tiles + kPhaseScanThreads - 1) / kPhaseScanThreads;
  const u64 lo = u64(t) * per, hi = lo + per < n_tiles ? lo + per : n_tiles;
  u64 f = fn_identity();
  for (u64 k = lo; k < hi; ++k) f = fn_then(f, ws.ph_tile_fn[k]);
  s_f[t] = f;
  __syncthreads();
  for (unsigned d = 1; d < unsigned(kPhaseScanThreads); d <<= 1) {  // inclusive scan, earlier functions first
    const u64 left = t >= d ? s_f[t - d] : fn_identity();
    __syncthreads();
    if (t >= d) s_f[t] = fn_then(left, s_f[t]);
    __syncthreads();
  }
  const u64 before = t ? s_f[t - 1] : fn_identity();
  u32 e = fn_get(before, g.entry0 > 8u ? 0u : g.entry0);
  for (u64 k = lo; k < hi; ++k) {
    ws.ph_tile_entry[k] = e;
    e = fn_get(ws.ph_tile_fn[k], e);
  }
}

// every subsequence's entry: thread = tile, a serial walk over its (independent, batched) function loads
__global__ void __launch_bounds__(kDecThreads)
dec_phase_entries_kernel(DecGeometry g, DecWorkspace ws, u32* __restrict__ entry_out) {
  const u64 t = u64(blockIdx.x) * kDecThreads + threadIdx.x;
  const u64 first = t * kPhaseTile;
  if (first >= g.n_sub) return;
  u32 e = ws.ph_tile_entry[t];
  for (u64 i0 = first; i0 < first + kPhaseTile && i0 < g.n_sub; i0 += 8) {
    u64 f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = i0 + k < g.n_sub ? ws.ph_fn[i0 + k] : fn_identity();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (i0 + k < g.n_sub) entry_out[i0 + k] = e;
      e = fn_get(f[k], e);
    }
  }
}

// the chosen entry's results become the subsequence states that K6/K7 work from
__global__ void __launch_bounds__(kDecThreads)
dec_phase_finish_kernel(DecGeometry g, DecWorkspace ws, const u32* __restrict__ entry_in) {
  const u64 i = u64(blockIdx.x) * kDecThreads + threadIdx.x;
  if (i >= g.n_sub) return;
  const u32 e = entry_in[i];
  const u64 k = i * kPhasePaths + e;
  const u32 neof = ws.ph_neof[k];
  ws.neof[i] = neof;
  ws.eofpos[i] = ws.ph_first[k];
  ws.sub[i] = pack_state(ws.ph_cnt[k], i == 0 ? g.entry0 : e, fn_get(ws.ph_fn[i], e), neof != 0);
}

// ---- K6: truncate at the first end mark, turn counts into output offsets -----------------------------------
__global__ void __launch_bounds__(kDecThreads)
dec_tile_sum_kernel(DecGeometry g, DecWorkspace ws) {
  __shared__ u32 s_warp[kDecThreads / 32];
  const u64 i = u64(blockIdx.x) * kDecThreads + threadIdx.x;
  const u64 st = i < g.n_sub ? ws.sub[i] : 0ull;
  if (i < g.n_sub && st_eof(st)) atomicMin(&ws.ctl->eof_index, u32(i));
  const u32 sum = warp_sum(st_count(st));
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    u64 total = 0;
    for (int k = 0; k < kDecThreads / 32; ++k) total += s_warp[k];
    ws.tile_sum[blockIdx.x] = total;
  }
}

// The first subsequence (in stream order) whose synchronised path contains an end mark ends the stream. The number
// of symbols before that mark is normally already known (eofpos, kept by the walks above); when the mark lay in a
// reused tail whose own first end mark was a false one, that one subsequence is walked again by a single thread
// (tables staged in shared memory by the whole block, multi-codeword steps).
__global__ void __launch_bounds__(kDecThreads)
dec_locate_eof_kernel(DecGeometry g, DecWorkspace ws) {
  __shared__ SmemSpeculate s;
  const u32 i = ws.ctl->eof_index;
  if (i == kNoEof) return;
  const u32 known = ws.eofpos[i];
  if (known < kEofPosUnknown) {
    if (threadIdx.x == 0) ws.ctl->eof_prefix = known;
    return;
  }
  load_speculate_tables(s, ws);
  __syncthreads();
  if (threadIdx.x != 0) return;
  const u64 st = ws.sub[i];
  const u32 end = u32(sub_end_bits(g, i));
  u32 pos = st_entry(st), count = 0;
  BitReader r;
  r.seek(g.payload, g.readable, u64(i) * u64(g.sub_bytes) * 8 + pos);
  while (pos < end) {
    const u32 win = r.window();
    u32 len, cnt;
    if (lutc_hit(s.lutC[win >> (32 - kLutCBits)], len, cnt)) {  // whole codewords, never the end mark
      count += cnt;
    } else {
      u32 sym;
      decode_one(s.canon, s.lut1, win, sym, len);
      if (sym == u32(GH_EOF_SYMBOL)) break;
      ++count;
    }
    pos += len;
    r.consume(len);
  }
  ws.ctl->eof_prefix = count;
}

constexpr int kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads)
dec_offsets_kernel(DecGeometry g, DecWorkspace ws) {
  __shared__ u64 s_warp[kScanThreads / 32];
  __shared__ u64 s_partial;
  const unsigned t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const u64 n_tiles = (g.n_sub + kDecThreads - 1) / kDecThreads;
  const u32 eof_index = ws.ctl->eof_index;
  const u64 eof_tile = eof_index == kNoEof ? n_tiles : u64(eof_index) / kDecThreads;

  // the tile holding the first end mark: whole subsequences before it, then the symbols before the mark itself
  u64 part = 0;
  if (eof_index != kNoEof && t < unsigned(kDecThreads)) {
    const u64 i = eof_tile * kDecThreads + t;
    if (i < u64(eof_index)) part = st_count(ws.sub[i]);
    else if (i == u64(eof_index)) part = ws.ctl->eof_prefix;
  }
  part = warp_sum64(part);
  if (lane == 0) s_warp[warp] = part;
  __syncthreads();
  if (t == 0) {
    u64 p = 0;
    for (int k = 0; k < kScanThreads / 32; ++k) p += s_warp[k];
    s_partial = p;
  }
  __syncthreads();
  const u64 partial = s_partial;

  u64 carry = 0;
  for (u64 chunk = 0; chunk < n_tiles; chunk += kScanThreads) {
    const u64 tile = chunk + t;
    u64 v = 0;
    if (tile < n_tiles) v = tile < eof_tile ? ws.tile_sum[tile] : (tile == eof_tile ? partial : 0ull);
    // block-wide exclusive scan of v
    u64 incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const u64 up = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= unsigned(d)) incl += up;
    }
    __syncthreads();  // s_warp reuse
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    u64 warp_base = 0, chunk_total = 0;
    for (int k = 0; k < kScanThreads / 32; ++k) {
      const u64 wt = s_warp[k];
      if (unsigned(k) < warp) warp_base += wt;
      chunk_total += wt;
    }
    if (tile < n_tiles) ws.tile_base[tile] = carry + warp_base + incl - v;
    carry += chunk_total;
  }
  if (t == 0) {
    ws.ctl->total = carry;
    ws.ctl->eof_found = eof_index != kNoEof;
    ws.ctl->exit_bit = st_exit(ws.sub[g.n_sub - 1]);
  }
}

// ---- K7: final decode from the exact entries ------------------------------------------------------------------
// The cursor reader with the table lutW: one LDS yields the cursor/fill addend and up to 2 (4) symbols. `acc` carries,
// above the cursor field, the number of output BITS produced so far (modulo 2^22), so the same add that moves the
// cursor also advances the output fill; the symbols are shifted to the fill position (SHF takes acc >> 10 modulo 32)
// and OR-ed into the word being assembled. A word is complete when bit 5 of the fill flips -- one LOP3 on
// acc ^ acc_before -- and then goes to the lane's private ring in shared memory (one predicated STS and one add).
// The lookup loop is bound by its instruction stream (profiles/rnd2_notes.md), so everything else about the output
// -- 16-byte grouping, the 64-bit destination, the global stores -- happens outside it: between two word steps, when
// any lane of the warp may run out of ring space during the next step, every lane copies its ring to its own
// destination, four words per 128-bit store (a short loop in which all lanes work).
// Two instances: rings of 32 words, three blocks per SM (the default), and rings of 16 words, four blocks per SM, for codes
// without short codewords (min_len >= kSmallRingMinLen): those complete at most three words per step, and their lanes
// run in lockstep -- measured on uniform bytes, the small rings' shorter store bursts and the fourth block win
// (1.32 against 1.41 ms), while everything else prefers the large ones (Zipf 1.15 against 1.25 ms).
// A word step completes at most (32 + kLutWBits - 1) / min_len symbols (its lookups consume at most that many bits)
// plus the symbol of a table miss.
__host__ __device__ constexpr u32 ring_step_words(u32 min_len) { return (u32(32 + kLutWBits - 1) / min_len + 3u) / 4u + 1u; }
constexpr u32 kSmallRingMinLen = 7;
// Word w of the ring of lane l of warp v sits at ring[(v * kWords + w) * 32 + l]: bank = lane for every access to the
// rings, whatever the lanes' fill levels (rows per lane -- 128-bit loads in the copy-out, but 4 lanes per bank group --
// cost 1-2 %, and a row stride of exactly 32 words 27 %).
constexpr u32 kRingWordStep = 128;  // bytes from one word of a lane's ring to the next
template <int kRingWords>
struct SmemWriteT {
  static constexpr int kWords = kRingWords;       // usable words of a lane's ring
  SmemCanon canon;
  uint16_t lut1[1 << kLut1Bits];
  LutWEntry lutW[1 << kLutWBits];
  u32 warp_total[kDecThreads / 32];
  alignas(16) u32 ring[kDecThreads * kRingWords];
};
#ifndef GH_RING_LARGE
#define GH_RING_LARGE 32  // tuning builds: other ring sizes / blocks per SM for the default instance
#define GH_RING_LARGE_BLOCKS 3
#endif
struct WriteLarge {
  typedef SmemWriteT<GH_RING_LARGE> Smem;
  static constexpr int kBlocksPerSm = GH_RING_LARGE_BLOCKS;
  static constexpr bool kFlushOutOfLine = true;
};
struct WriteSmall {
  typedef SmemWriteT<16> Smem;
  static constexpr int kBlocksPerSm = 4;
  static constexpr bool kFlushOutOfLine = false;  // at 64 registers the call's saves and restores spill
};
static_assert(ring_step_words(1) < GH_RING_LARGE && ring_step_words(kSmallRingMinLen) < 16, "ring too small for one word step");

// `merged` becomes the open word; if `word_done`, it goes to the ring and `spill` opens the next one. Predicated PTX:
// as C++ branches the compiler turns these few moves into divergent control flow that every warp then walks on almost
// every iteration.
__device__ __forceinline__ void ring_push(smem_addr_t& ring_at, u32& part, u32 merged, u32 spill, u32 word_done) {
#ifdef GH_EMUL
  if (word_done) {
    *reinterpret_cast<u32*>(const_cast<char*>(ring_at)) = merged;
    ring_at += kRingWordStep;
    part = spill;
  } else {
    part = merged;
  }
#else
  part = merged;
  asm volatile(
      "{\n"
      " .reg .pred pw;\n"
      " setp.ne.u32 pw, %4, 0;\n"
      " @pw st.shared.u32 [%0], %2;\n"
      " @pw add.u32 %0, %0, %5;\n"
      " @pw mov.u32 %1, %3;\n"
      "}\n"
      : "+r"(ring_at), "+r"(part)
      : "r"(merged), "r"(spill), "r"(word_done), "n"(kRingWordStep)
      : "memory");
#endif
}

__device__ __forceinline__ u32 ring_load(smem_addr_t at) {
#ifdef GH_EMUL
  return *reinterpret_cast<const u32*>(at);
#else
  u32 v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(at));
  return v;
#endif
}
__device__ __forceinline__ void ring_store(smem_addr_t at, u32 v) {
#ifdef GH_EMUL
  *reinterpret_cast<u32*>(const_cast<char*>(at)) = v;
#else
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(at), "r"(v) : "memory");
#endif
}

// The copy-out of one lane's ring: whole 16-byte groups leave for `gaddr` (16-byte aligned), up to three words stay and
// move to the front. Out of line where registers allow: it runs once per ~15 word steps, and inlined at all sixteen
// word steps of the unrolled unit pair it is a quarter of the kernel's code.
__device__ __forceinline__ void ring_flush_body(smem_addr_t ring0, u32 words, u64 gaddr) {
  const u32 n4 = words >> 2;
#pragma unroll 1
  for (u32 k = 0; k < n4; ++k) {
    const smem_addr_t at = ring0 + 4 * kRingWordStep * k;
    const uint4 v = make_uint4(ring_load(at), ring_load(at + kRingWordStep), ring_load(at + 2 * kRingWordStep),
                               ring_load(at + 3 * kRingWordStep));
#ifdef GH_EMUL
    *reinterpret_cast<uint4*>(gaddr + 16ull * k) = v;
#else
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(gaddr + 16ull * k), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#endif
  }
  const u32 rest = words & 3u;
  if (n4) {
#pragma unroll 1
    for (u32 k = 0; k < rest; ++k) ring_store(ring0 + kRingWordStep * k, ring_load(ring0 + kRingWordStep * (4 * n4 + k)));
  }
}
__device__ __noinline__ void ring_flush_call(smem_addr_t ring0, u32 words, u64 gaddr) { ring_flush_body(ring0, words, gaddr); }

template <class Cfg>
__global__ void __launch_bounds__(kDecThreads, Cfg::kBlocksPerSm)
dec_write_kernel(DecGeometry g, uint8_t* __restrict__ out, u64 out_cap, DecWorkspace ws) {
  typedef typename Cfg::Smem SmemWrite;
  GH_DYNAMIC_SMEM(smem_raw);
  SmemWrite& s = *reinterpret_cast<SmemWrite*>(smem_raw);
  load_canon(s.canon, ws.tables);
  for (unsigned i = threadIdx.x; i < (1u << kLut1Bits) / 8; i += kDecThreads)
    reinterpret_cast<uint4*>(s.lut1)[i] = reinterpret_cast<const uint4*>(ws.lut1)[i];
  for (unsigned i = threadIdx.x; i < (sizeof(LutWEntry) << kLutWBits) / 16; i += kDecThreads)
    reinterpret_cast<uint4*>(s.lutW)[i] = reinterpret_cast<const uint4*>(ws.lutW)[i];
  const unsigned t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const u64 i = u64(blockIdx.x) * kDecThreads + t;
  const u32 eof_index = ws.ctl->eof_index;
  u64 st = 0;
  u32 count = 0;
  if (i < g.n_sub && i <= u64(eof_index)) {
    st = ws.sub[i];
    count = i == u64(eof_index) ? ws.ctl->eof_prefix : st_count(st);  // nothing is emitted past the end mark
  }
  const u32 incl = warp_inclusive_scan(count, lane);
  if (lane == 31) s.warp_total[warp] = incl;
  __syncthreads();
  u32 warp_base = 0;
#pragma unroll
  for (int k = 0; k < kDecThreads / 32; ++k)
    if (unsigned(k) < warp) warp_base += s.warp_total[k];
  const u64 o = ws.tile_base[blockIdx.x] + warp_base + (incl - count);
  // no early exit: the bulk loop below is warp-synchronous (its ring flushes are decided by a vote)
  const bool live = count != 0 && o < out_cap;
  u32 remaining = 0, end = 0, pos = 0;
  u64 start = 0;
  uint8_t* dst = out;
  u32 sym, len;
  if (live) {
    remaining = count;
    if (u64(remaining) > out_cap - o) remaining = u32(out_cap - o);
    start = i * u64(g.sub_bytes) * 8;
    end = u32(sub_end_bits(g, i));
    pos = st_entry(st);  // bits of the subsequence consumed so far
    dst = out + o;
    // head: single symbols up to the first 16-byte boundary of the output (the bytes before it belong to the
    // previous lane), so that the bulk loop below only ever issues whole aligned 128-bit stores
    if (reinterpret_cast<uintptr_t>(dst) & 15) {
      BitReader r;
      r.seek(g.payload, g.readable, start + pos);
      while (remaining && (reinterpret_cast<uintptr_t>(dst) & 15)) {
        decode_one(s.canon, s.lut1, r.window(), sym, len);
        r.consume(len);
        pos += len;
        *dst++ = uint8_t(sym);
        --remaining;
      }
    }
  }
  // bulk: every codeword that starts before `end` belongs to this lane (that is what `count` counted), so the loop
  // is bounded by the bit position alone; a lane whose output was clipped by out_cap takes the slow path only.
  {
    typedef CursorGeom<kLutWBits, kLutWEntryShift> G;
    u64 u = 0, ulast = 0;
    u32 acc = 0;
    const bool bulk = live && remaining && u64(count) <= out_cap - o && pos < end &&
                      cursor_plan<kLutWBits, kLutWEntryShift>(start + pos, start + end, g.readable >> 5, u, ulast, acc);
    const u32 n_units = bulk ? u32(ulast - u) + 1u : 0u;
    __syncwarp();
    u32 n_units_warp = n_units;  // the warp walks together: lanes with fewer units idle through the rest
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const u32 other = __shfl_xor_sync(0xffffffffu, n_units_warp, d);
      n_units_warp = other > n_units_warp ? other : n_units_warp;
    }
    if (n_units_warp) {
      constexpr u32 kIdle = 192u;  // a cursor field that stays below the allowed range for a whole unit (8 x 32 more)
      const bool aligned32 = (reinterpret_cast<uintptr_t>(g.payload) & 31) == 0;
      const u64 umax = (g.readable >> 5) - 1;
      const smem_addr_t lut = smem_addr(s.lutW);
      const smem_addr_t ring0 = smem_addr(s.ring + warp * 32 * SmemWrite::kWords + lane);
      // at or beyond `ring_full`, flush before the next step
      const smem_addr_t ring_full = ring0 + kRingWordStep * (u32(SmemWrite::kWords) - ring_step_words(s.canon.min_len));
      smem_addr_t ring_at = ring0;
      u64 gaddr = u64(reinterpret_cast<uintptr_t>(dst));  // destination of the ring's first word (16-byte aligned)
      u32 part = 0;
      u32 hi, lo = 0;
      bool stop = false;      // the end mark was met (the subsequence that ends the stream)
      u32 acc_end = acc;      // the cursor after this lane's last unit
      u32 units_left = n_units;
      auto append = [&](u32 addend, u32 syms) {
        const u32 fill = __umulhi(acc, 1u << (32 - kCurShift));  // acc >> kCurShift on the FMA pipe; SHF uses it modulo 32
        const u32 merged = part | __funnelshift_l(0u, syms, fill);
        const u32 spill = __funnelshift_l(syms, 0u, fill);
        const u32 next = acc + addend;
        const u32 flipped = next ^ acc;
        acc = next;
        ring_push(ring_at, part, merged, spill, flipped & (32u << kCurShift));  // bit 5 of the fill flipped: word complete
      };
      auto flush = [&]() {
        const u32 words = u32(ring_at - ring0) / kRingWordStep;
        if constexpr (Cfg::kFlushOutOfLine) ring_flush_call(ring0, words, gaddr);
        else ring_flush_body(ring0, words, gaddr);
        gaddr += 4ull * (words & ~3u);
        ring_at = ring0 + kRingWordStep * (words & 3u);
      };
      auto walk_unit = [&](const Unit8& cu, u32 next_unit_word0) {  // cu in stream order, the next unit's word raw
#pragma unroll
        for (int k = 0; k < kUnitWords; ++k) {
          hi = lo;
          lo = cu.w[k];
          acc += 32u;
          while ((acc & kCurBusy) == 0u) {
            do {
              const u32 x = __funnelshift_r(lo, hi, acc);
#ifdef GH_LUTW_U32
              const u32 e = lds_u32(lut, x & G::kMask);
              append(__umulhi(e, 1u << 16), e & 0xffffu);
#else
              const uint2 e = lds_v2(lut, x & G::kMask);
              append(e.x, e.y);
#endif
            } while ((acc & kCurBusy) == 0u);
            if ((acc & kCurMissBit) == 0u) break;
            // codeword longer than the table window, or the end mark (only in the subsequence that ends the stream)
            acc -= kLutMiss;
            const u32 next_be = k + 1 < kUnitWords ? cu.w[(k + 1) % kUnitWords] : be32(next_unit_word0);
            const u32 sl = decode_one_packed(&s.canon, s.lut1, cursor_window32<kLutWBits, kLutWEntryShift>(hi, lo, next_be, acc));
            if ((sl >> 8) == u32(GH_EOF_SYMBOL)) {
              stop = true;
              acc_end = acc;
              acc = (acc & ~kCurFieldMask) | kIdle;
              break;
            }
            append((8u << kCurShift) - (sl & 0xffu), sl >> 8);
          }
          if (__any_sync(0xffffffffu, ring_at >= ring_full)) flush();
        }
      };
      // One unit: a lane that has units left walks it, the others idle through it (their cursor field never reaches
      // the allowed range). Two unit buffers swap roles, so the unit in flight is never copied (a copy would wait for
      // the load).
      auto begin_unit = [&]() -> bool {
        const bool mine = units_left != 0u && !stop;
        if (!mine) acc = (acc & ~kCurFieldMask) | kIdle;
        return mine;
      };
      auto end_unit = [&](bool mine) {
        if (mine && !stop) {
          --units_left;
          if (units_left == 0u) acc_end = acc;
          else ++u;
        }
      };
      if (!bulk) acc = kIdle;
      Unit8 ua = ldg_unit(g.payload + 32 * (u < umax ? u : umax), aligned32), ub;
      for (u32 it = 0; it < n_units_warp; it += 2) {
        unit_to_stream_order(ua);
        ub = ldg_unit(g.payload + 32 * (u + 1 < umax ? u + 1 : umax), aligned32);
        bool mine = begin_unit();
        walk_unit(ua, ub.w[0]);
        end_unit(mine);
        if (it + 1 >= n_units_warp) break;
        unit_to_stream_order(ub);
        ua = ldg_unit(g.payload + 32 * (u + 1 < umax ? u + 1 : umax), aligned32);
        mine = begin_unit();
        walk_unit(ub, ua.w[0]);
        end_unit(mine);
      }
      if (bulk) {
        // drain: the ring's words, then the bytes of the open word
        const u32 words = u32(ring_at - ring0) / kRingWordStep;
        for (u32 k = 0; k < words; ++k) *reinterpret_cast<u32*>(gaddr + 4ull * k) = ring_load(ring0 + kRingWordStep * k);
        uint8_t* tail = reinterpret_cast<uint8_t*>(gaddr + 4ull * words);
        const u32 open_bytes = ((acc >> kCurShift) & 31u) >> 3;
        for (u32 k = 0; k < open_bytes; ++k) tail[k] = uint8_t(part >> (8 * k));
        const u64 emitted = u64(tail - dst) + open_bytes;
        dst += emitted;
        remaining = stop ? 0u : remaining - u32(emitted);
        pos = u32(cursor_position<kLutWBits, kLutWEntryShift>(ulast, acc_end) - start);
      }
    }
  }
  // what the bulk loop left (the last few bits of the subsequence, the payload tail, clipped output): one codeword
  // at a time through the bounds-checked reader from `pos`
  if (remaining) {
    BitReader r;
    r.seek(g.payload, g.readable, start + pos);
    while (remaining) {
      decode_one(s.canon, s.lut1, r.window(), sym, len);
      r.consume(len);
      *dst++ = uint8_t(sym);
      --remaining;
    }
  }
}

// ---- K6b: per-subsequence output offsets (for the warp-cooperative writer) -----------------------------------
__global__ void __launch_bounds__(kDecThreads)
dec_sub_offsets_kernel(DecGeometry g, DecWorkspace ws) {
  __shared__ u32 s_warp[kDecThreads / 32];
  const unsigned t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const u64 i = u64(blockIdx.x) * kDecThreads + t;
  const u32 eof_index = ws.ctl->eof_index;
  u32 count = 0;
  if (i < g.n_sub && i <= u64(eof_index)) count = i == u64(eof_index) ? ws.ctl->eof_prefix : st_count(ws.sub[i]);
  const u32 incl = warp_inclusive_scan(count, lane);
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  u32 warp_base = 0;
#pragma unroll
  for (int k = 0; k < kDecThreads / 32; ++k)
    if (unsigned(k) < warp) warp_base += s_warp[k];
  if (i < g.n_sub) ws.out_off[i] = ws.tile_base[blockIdx.x] + warp_base + (incl - count);
}

#ifndef GH_EXPERIMENTS
constexpr u32 kFineSubBytes = 2048;  // layout only (the fine pipeline itself is an experiment build)
#endif
#ifdef GH_EXPERIMENTS  // not part of the product build: measured slower than the coarse pipeline (profiles/README.md, r1i)
// ---- fine-grained pipeline: one warp per 2 KiB segment, 64-byte pieces, shared-memory staging -------------------
// For codes that re-synchronise within a few codewords (anything but near-fixed-length codes) the subsequence is
// fixed at 2 KiB and handled by a warp:
//   stage   coalesced 128-bit loads of the segment into shared memory (stream order, one pad word per 16 so that
//           the 32 lanes' pieces start in 32 different banks); every later bit access is a bit-addressed window
//           out of shared memory -- no refill state, so lanes stay in step;
//   K5a'    lane l walks 64-byte piece l from its first bit (lane 0 from the segment's entry) counting codewords;
//           lanes then pass their exits to the right (shuffle) and a lane whose assumed entry was wrong walks old
//           and new path in lockstep until they meet (a few codewords) and patches its count -- repeated until no
//           exit moves: the fixed-point argument of K5b inside a warp, without touching memory. The piece states
//           (count, exit, end marks) are stored, 4 bytes per 64-byte piece, plus the usual per-subsequence state,
//           so K5b/K6 run unchanged on the segments.
//   K7'     re-stages the segment, re-validates the stored piece states against the segment's now-exact entry (lane
//           0 walks a few codewords in the common case), scans the counts, and every lane decodes its piece ONCE,
//           storing symbols as bytes into a shared output window that leaves as coalesced 128-bit stores.
// (A writer that re-derived the piece entries itself -- 2.5 decode passes per piece -- measured 2x slower than the
//  thread-per-subsequence writer: profiles/r1f. Storing the piece states removes those passes.)
constexpr int kFineWarps = 16;
constexpr u32 kFineSubBytes = 2048;
constexpr u32 kPieceBits = 512;
constexpr u32 kSegBits = 32 * kPieceBits;
constexpr u32 kSegStageVecs = kSegBits / 128 + 3;  // + look-ahead for windows that start in the last piece
constexpr u32 kSegStageWords = kSegStageVecs * 4;
constexpr u32 kSegPaddedWords = kSegStageWords + kSegStageWords / 16 + 1;
constexpr u32 kOutWindow = 3072;  // symbols per copy-out window (multiple of 16)

// piece state: [9:0] codewords  [14:10] exit  [24:15] end marks among the codewords
__device__ __forceinline__ u32 pack_piece(u32 cnt, u32 exit, u32 neof) { return cnt | (exit << 10) | (neof << 15); }
__device__ __forceinline__ u32 pc_count(u32 p) { return p & 1023u; }
__device__ __forceinline__ u32 pc_exit(u32 p) { return (p >> 10) & 31u; }
__device__ __forceinline__ u32 pc_neof(u32 p) { return (p >> 15) & 1023u; }

struct SmemFineSpec {
  SmemCanon canon;
  u32 lutP[1 << kLutPBits];
  u32 in[kFineWarps][kSegPaddedWords];
};
struct SmemFineWrite {
  SmemCanon canon;
  u32 lutP[1 << kLutPBits];
  u32 in[kFineWarps][kSegPaddedWords];
  __align__(16) uint8_t out[kFineWarps][kOutWindow + 32];
};

__device__ __forceinline__ u32 seg_window(const u32* sin, u32 pos) {
  const u32 j = pos >> 5;
  const u32 a = j + (j >> 4), b = (j + 1) + ((j + 1) >> 4);
  return __funnelshift_l(sin[b], sin[a], pos & 31);
}

// coalesced copy of a segment's vectors into the padded shared layout, stream order
__device__ __forceinline__ void stage_segment(const DecGeometry& g, u64 vec0, u32* sin, unsigned lane) {
  const u64 full_vecs = g.readable >> 4;
  for (u32 k = lane; k < kSegStageVecs; k += 32) {
    const u64 v = vec0 + k;
    const uint4 q = v < full_vecs ? ldg128(reinterpret_cast<const uint4*>(g.payload) + v)
                                  : fetch_tail(g.payload, g.readable, v);
    const u32 j = 4 * k, base = j + (j >> 4);
    sin[base] = be32(q.x);
    sin[base + 1] = be32(q.y);
    sin[base + 2] = be32(q.z);
    sin[base + 3] = be32(q.w);
  }
}

// one step of a path inside a piece: one or two whole codewords (never a second one once the first reaches p_end)
__device__ __forceinline__ void piece_step(const SmemCanon& canon, const u32* lutP, const u32* sin, u32 p_end, u32& pos,
                                           u32& cnt, u32& neof) {
  const u32 win = seg_window(sin, pos);
  const u32 e = lutP[win >> (32 - kLutPBits)];
  if (e) {
    const u32 len1 = (e >> 6) & 15u;
    const bool both = ((e >> 4) & 3u) == 2u && pos + len1 < p_end;
    pos += both ? (e & 15u) : len1;
    cnt += both ? 2u : 1u;
  } else {  // longer than 12 bits, or the end mark (a path runs through it like through any codeword)
    u32 sym, len;
    canon_search(canon, win, canon.min_len, sym, len);
    pos += len;
    ++cnt;
    neof += sym == u32(GH_EOF_SYMBOL);
  }
}

// Warp-level fixed point over the 32 pieces of a staged segment. On entry every lane holds the state (entry, cnt,
// exit, neof) of the path it last walked; lane 0's entry must become `seg_entry`, lane l's the left exit.
__device__ __forceinline__ void fix_pieces(const SmemCanon& canon, const u32* lutP, const u32* sin, unsigned lane,
                                           u32 p_begin, u32 p_end, u32 seg_entry, u32& entry, u32& cnt, u32& exit,
                                           u32& neof) {
  while (true) {
    const u32 left_exit = __shfl_up_sync(0xffffffffu, exit, 1);
    const u32 want = lane == 0 ? seg_entry : p_begin + left_exit;
    bool moved = false;
    if (want != entry) {
      u32 pa = entry, pb = want, sa = 0, sb = 0, ea = 0, eb = 0;
      bool merged = false;
      while (pb < p_end) {
        if (pa == pb) {
          merged = true;
          break;
        }
        if (pa < pb) piece_step(canon, lutP, sin, p_end, pa, sa, ea);
        else piece_step(canon, lutP, sin, p_end, pb, sb, eb);
      }
      if (merged) {  // counts are additive: the rest of the stored path is the rest of the new one
        cnt = sb + (cnt - sa);
        neof = eb + (neof - ea);
      } else {
        const u32 new_exit = pb >= p_end ? pb - p_end : 0u;
        cnt = sb;
        neof = eb;
        moved = new_exit != exit;
        exit = new_exit;
      }
      entry = want;
    }
    if (!__any_sync(0xffffffffu, moved)) break;
  }
}

// ---- K5a' ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFineWarps * 32)
dec_fine_speculate_kernel(DecGeometry g, DecWorkspace ws) {
  GH_DYNAMIC_SMEM(smem_raw);
  SmemFineSpec& s = *reinterpret_cast<SmemFineSpec*>(smem_raw);
  load_canon(s.canon, ws.tables);
  for (unsigned k = threadIdx.x; k < (1u << kLutPBits) / 4; k += kFineWarps * 32)
    reinterpret_cast<uint4*>(s.lutP)[k] = reinterpret_cast<const uint4*>(ws.lutP)[k];
  __syncthreads();
  const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const u64 i = u64(blockIdx.x) * kFineWarps + wib;
  if (i >= g.n_sub) return;  // warp-uniform
  u32* sin = s.in[wib];
  const u32 seg_end = u32(sub_end_bits(g, i));  // < kSegBits only for the last segment
  stage_segment(g, i * (kFineSubBytes / 16), sin, lane);
  __syncwarp();
  const u32 seg_entry = (i == 0) ? g.entry0 : 0u;
  const u32 p_begin = lane * kPieceBits < seg_end ? lane * kPieceBits : seg_end;
  const u32 p_end = (lane + 1) * kPieceBits < seg_end ? (lane + 1) * kPieceBits : seg_end;
  u32 entry = lane == 0 ? seg_entry : p_begin;
  u32 pos = entry, cnt = 0, neof = 0;
  while (pos < p_end) piece_step(s.canon, s.lutP, sin, p_end, pos, cnt, neof);
  u32 exit = pos >= p_end ? pos - p_end : 0u;
  fix_pieces(s.canon, s.lutP, sin, lane, p_begin, p_end, seg_entry, entry, cnt, exit, neof);
  ws.pieces[i * 32 + lane] = pack_piece(cnt, exit, neof);
  const u32 seg_cnt = warp_sum(cnt), seg_neof = warp_sum(neof);
  if (lane == 31) {
    ws.sub[i] = pack_state(seg_cnt, seg_entry, exit, seg_neof != 0);
    ws.neof[i] = seg_neof;
    ws.eofpos[i] = seg_neof ? kEofPosUnknown : kNoEof;  // K6 locates the mark if this segment turns out to end the stream
  }
}

// ---- K7' ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFineWarps * 32)
dec_fine_write_kernel(DecGeometry g, uint8_t* __restrict__ out, u64 out_cap, DecWorkspace ws) {
  GH_DYNAMIC_SMEM(smem_raw);
  SmemFineWrite& s = *reinterpret_cast<SmemFineWrite*>(smem_raw);
  load_canon(s.canon, ws.tables);
  for (unsigned k = threadIdx.x; k < (1u << kLutPBits) / 4; k += kFineWarps * 32)
    reinterpret_cast<uint4*>(s.lutP)[k] = reinterpret_cast<const uint4*>(ws.lutP)[k];
  __syncthreads();
  const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const u64 i = u64(blockIdx.x) * kFineWarps + wib;
  if (i >= g.n_sub) return;  // warp-uniform
  const u32 eof_index = ws.ctl->eof_index;
  if (i > u64(eof_index)) return;
  const u64 st = ws.sub[i];
  u32 remaining = i == u64(eof_index) ? ws.ctl->eof_prefix : st_count(st);
  const u64 o = ws.out_off[i];
  if (o >= out_cap) return;
  if (u64(remaining) > out_cap - o) remaining = u32(out_cap - o);
  if (remaining == 0) return;
  u32* sin = s.in[wib];
  uint8_t* sout = s.out[wib];
  const u32 seg_end = u32(sub_end_bits(g, i));
  stage_segment(g, i * (kFineSubBytes / 16), sin, lane);
  // stored piece states describe the paths K5a' settled on with the speculative segment entry
  const u32 ps = ws.pieces[i * 32 + lane];
  u32 cnt = pc_count(ps), exit = pc_exit(ps), neof = pc_neof(ps);
  const u32 p_begin = lane * kPieceBits < seg_end ? lane * kPieceBits : seg_end;
  const u32 p_end = (lane + 1) * kPieceBits < seg_end ? (lane + 1) * kPieceBits : seg_end;
  const u32 left_exit = __shfl_up_sync(0xffffffffu, exit, 1);
  u32 entry = lane == 0 ? ((i == 0) ? g.entry0 : 0u) : p_begin + left_exit;
  __syncwarp();
  // re-validate against the exact segment entry (usually lane 0 walks a few codewords and nothing else moves)
  fix_pieces(s.canon, s.lutP, sin, lane, p_begin, p_end, st_entry(st), entry, cnt, exit, neof);
  const u32 incl = warp_inclusive_scan(cnt, lane);
  const u32 prefix = incl - cnt;
  const u32 total = __shfl_sync(0xffffffffu, incl, 31);
  const u32 seg_syms = total < remaining ? total : remaining;
  const u32 my_end = prefix >= seg_syms ? prefix : (incl < seg_syms ? incl : seg_syms);  // my symbols: [prefix, my_end)
  const u32 a0 = u32(reinterpret_cast<uintptr_t>(out + o) & 15);  // window phase == destination phase
  u32 idx = prefix;  // segment-relative index of my next symbol
  u32 pos = entry;
  for (u32 wbase = 0; wbase < seg_syms; wbase += kOutWindow) {
    const u32 wend = wbase + kOutWindow < seg_syms ? wbase + kOutWindow : seg_syms;
    const u32 lim = my_end < wend ? my_end : wend;
    uint8_t* slot = sout + a0 - wbase;  // slot[idx] is where symbol idx of the segment goes
    while (idx < lim) {
      const u32 win = seg_window(sin, pos);
      const u32 e = s.lutP[win >> (32 - kLutPBits)];
      if (e) {
        const bool both = ((e >> 4) & 3u) == 2u && idx + 1 < lim;
        slot[idx] = uint8_t(e >> 16);
        if (both) slot[idx + 1] = uint8_t(e >> 24);
        pos += both ? (e & 15u) : ((e >> 6) & 15u);
        idx += both ? 2u : 1u;
      } else {
        u32 sym, len;
        canon_search(s.canon, win, s.canon.min_len, sym, len);
        slot[idx] = uint8_t(sym);
        pos += len;
        ++idx;
      }
    }
    __syncwarp();
    const u32 m = wend - wbase;  // bytes in this window, at sout[a0 .. a0 + m)
    uint8_t* dst = out + o + wbase;
    const u32 head = a0 ? ((16 - a0) < m ? (16 - a0) : m) : 0u;
    if (lane < head) dst[lane] = sout[a0 + lane];
    const u32 nvec = (m - head) >> 4;
    const uint4* src4 = reinterpret_cast<const uint4*>(sout + a0 + head);
    uint4* dst4 = reinterpret_cast<uint4*>(dst + head);
    for (u32 k = lane; k < nvec; k += 32) dst4[k] = src4[k];
    const u32 done = head + (nvec << 4);
    if (lane < m - done) dst[done + lane] = sout[a0 + done + lane];
    __syncwarp();
  }
}

#endif  // GH_EXPERIMENTS

// ---- host orchestration -------------------------------------------------------------------------------
// Pipeline choice: 0 = automatic (currently always the coarse one), 1 = always coarse (one thread per
// subsequence), 2 = always fine (one warp per 2 KiB segment). gh_debug_select_writer / GH_DECODE_PIPELINE
// ("coarse" / "fine") override the automatic choice for A/B runs; the output is identical either way.
static int g_pipeline = 0;
static int pipeline_choice() { return g_pipeline; }
static bool g_no_phase_walk = false;  // gh_debug_disable_phase_walk: tests force the re-walk rounds on 8/9-bit codes

// Kernels that need more than the 48 KB of dynamic shared memory a kernel gets without opting in. The attribute is
// per device, so it is set on every call (it costs nothing) rather than once per process.
static int set_write_attrs() {
  GH_CUDA_TRY(cudaFuncSetAttribute(dec_write_kernel<WriteLarge>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(WriteLarge::Smem))));
  GH_CUDA_TRY(cudaFuncSetAttribute(dec_write_kernel<WriteSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(WriteSmall::Smem))));
  return GH_OK;
}

#ifdef GH_EXPERIMENTS
static int set_fine_attrs() {
  GH_CUDA_TRY(cudaFuncSetAttribute(dec_fine_speculate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   int(sizeof(SmemFineSpec))));
  GH_CUDA_TRY(cudaFuncSetAttribute(dec_fine_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   int(sizeof(SmemFineWrite))));
  return GH_OK;
}
#endif

struct DecLayout {
  size_t off_tables, off_lut1, off_lutC, off_lutW, off_lutP, off_ctl, off_sub, off_neof, off_eofpos, off_out_off, off_pieces, off_work0, off_work1, off_ph_fn, off_ph_cnt, off_ph_first, off_ph_neof, off_ph_tile_fn, off_ph_tile_entry, off_tile_sum, off_tile_base, total;
};

static DecLayout dec_layout(u64 slice_bytes) {
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const u64 max_sub = slice_bytes / kMinSubBytes + 2;
  const u64 max_tiles = max_sub / kDecThreads + 2;
  DecLayout L;
  L.off_tables = 0;
  L.off_lut1 = up(sizeof(DecodeTables));
  L.off_lutC = L.off_lut1 + up(sizeof(uint16_t) << kLut1Bits);
  L.off_lutW = L.off_lutC + up(sizeof(uint16_t) << kLutCBits);
  L.off_lutP = L.off_lutW + up(sizeof(LutWEntry) << kLutWBits);
  L.off_ctl = L.off_lutP + up(sizeof(u32) << kLutPBits);
  L.off_sub = L.off_ctl + 256;
  L.off_neof = L.off_sub + up(size_t(max_sub) * 8);
  L.off_eofpos = L.off_neof + up(size_t(max_sub) * 4);
  L.off_out_off = L.off_eofpos + up(size_t(max_sub) * 4);
  L.off_pieces = L.off_out_off + up(size_t(max_sub) * 8);
  L.off_work0 = L.off_pieces + up((size_t(slice_bytes) / kFineSubBytes + 2) * 32 * 4);
  L.off_work1 = L.off_work0 + up(size_t(max_sub) * 4);
  const u64 ph_sub = slice_bytes / kPhaseMinSub + 2, ph_tiles = ph_sub / kPhaseTile + 2;
  L.off_ph_fn = L.off_work1 + up(size_t(max_sub) * 4);
  L.off_ph_cnt = L.off_ph_fn + up(size_t(ph_sub) * 8);
  L.off_ph_first = L.off_ph_cnt + up(size_t(ph_sub) * kPhasePaths * 4);
  L.off_ph_neof = L.off_ph_first + up(size_t(ph_sub) * kPhasePaths * 4);
  L.off_ph_tile_fn = L.off_ph_neof + up(size_t(ph_sub) * kPhasePaths * 4);
  L.off_ph_tile_entry = L.off_ph_tile_fn + up(size_t(ph_tiles) * 8);
  L.off_tile_sum = L.off_ph_tile_entry + up(size_t(ph_tiles) * 4);
  L.off_tile_base = L.off_tile_sum + up(size_t(max_tiles) * 8);
  L.total = L.off_tile_base + up(size_t(max_tiles) * 8);
  return L;
}

static DecWorkspace dec_bind(void* d_ws, const DecLayout& L) {
  uint8_t* p = static_cast<uint8_t*>(d_ws);
  DecWorkspace w;
  w.tables = reinterpret_cast<const DecodeTables*>(p + L.off_tables);
  w.lut1 = reinterpret_cast<const uint16_t*>(p + L.off_lut1);
  w.lutC = reinterpret_cast<const uint16_t*>(p + L.off_lutC);
  w.lutW = reinterpret_cast<const LutWEntry*>(p + L.off_lutW);
  w.lutP = reinterpret_cast<const u32*>(p + L.off_lutP);
  w.ctl = reinterpret_cast<DecControl*>(p + L.off_ctl);
  w.sub = reinterpret_cast<u64*>(p + L.off_sub);
  w.neof = reinterpret_cast<u32*>(p + L.off_neof);
  w.eofpos = reinterpret_cast<u32*>(p + L.off_eofpos);
  w.out_off = reinterpret_cast<u64*>(p + L.off_out_off);
  w.pieces = reinterpret_cast<u32*>(p + L.off_pieces);
  w.work[0] = reinterpret_cast<u32*>(p + L.off_work0);
  w.work[1] = reinterpret_cast<u32*>(p + L.off_work1);
  w.ph_fn = reinterpret_cast<u64*>(p + L.off_ph_fn);
  w.ph_cnt = reinterpret_cast<u32*>(p + L.off_ph_cnt);
  w.ph_first = reinterpret_cast<u32*>(p + L.off_ph_first);
  w.ph_neof = reinterpret_cast<u32*>(p + L.off_ph_neof);
  w.ph_tile_fn = reinterpret_cast<u64*>(p + L.off_ph_tile_fn);
  w.ph_tile_entry = reinterpret_cast<u32*>(p + L.off_ph_tile_entry);
  w.tile_sum = reinterpret_cast<u64*>(p + L.off_tile_sum);
  w.tile_base = reinterpret_cast<u64*>(p + L.off_tile_base);
  return w;
}

static u32 choose_sub_bytes(u64 slice_bytes) {
  const u64 target = u64(sm_count() > 0 ? sm_count() : 1) * 2048 * 3 / 2;  // ~1.5 resident waves of threads
  u64 s = (slice_bytes + target - 1) / target;
  s = (s + kMinSubBytes - 1) / kMinSubBytes * kMinSubBytes;
  if (s < kMinSubBytes) s = kMinSubBytes;
  if (s > kMaxSubBytes) s = kMaxSubBytes;
  return u32(s);
}

static int dec_finish_launch(const DecGeometry& g, const DecWorkspace& ws, cudaStream_t stream, bool fine) {
  const unsigned tiles = unsigned((g.n_sub + kDecThreads - 1) / kDecThreads);
  GH_LAUNCH(dec_tile_sum_kernel, tiles, kDecThreads, 0, stream, g, ws);
  GH_LAUNCH(dec_locate_eof_kernel, 1, kDecThreads, 0, stream, g, ws);
  GH_LAUNCH(dec_offsets_kernel, 1, kScanThreads, 0, stream, g, ws);
  if (fine) GH_LAUNCH(dec_sub_offsets_kernel, tiles, kDecThreads, 0, stream, g, ws);  // only the fine writer reads out_off
  return check_launch();
}

static int dec_read_ctl(const DecWorkspace& ws, DecControl* h_ctl, cudaStream_t stream) {
  GH_CUDA_TRY(cudaMemcpyAsync(h_ctl, ws.ctl, sizeof(DecControl), cudaMemcpyDeviceToHost, stream));
  GH_CUDA_TRY(cudaStreamSynchronize(stream));
  return GH_OK;
}

static int dec_finish(const DecGeometry& g, const DecWorkspace& ws, DecControl* h_ctl, cudaStream_t stream, bool fine) {
  int rc = dec_finish_launch(g, ws, stream, fine);
  if (rc != GH_OK) return rc;
  return dec_read_ctl(ws, h_ctl, stream);
}

// `deferred` (whole-image decode only): when non-null, the common case enqueues one synchronisation round and the
// offset kernels WITHOUT reading the control block back -- the caller enqueues the writer behind them and reads the
// control block once, after everything; *deferred tells it that `result` is still to be filled from that read. If the
// round turns out not to have been clean (an exit moved: rare, the paths of a self-synchronising code meet within a
// few codewords), the caller repeats the decode with deferred == nullptr.
static int decode_sync_impl(const uint8_t* d_payload, u64 slice_bytes, u64 readable, const gh_code* code, u32 entry_bit,
                            int first_call, gh_shard_sync* result, void* d_ws, size_t ws_bytes, cudaStream_t stream,
                            DecGeometry* geom_out, bool* fine_out = nullptr, bool* deferred = nullptr) {
  if (deferred) *deferred = false;
  if (!d_payload || !code || !d_ws) return GH_ERR_ARG;
  if (slice_bytes == 0) return GH_ERR_NO_EOF;
  if ((reinterpret_cast<uintptr_t>(d_payload) & 15) || (reinterpret_cast<uintptr_t>(d_ws) & 255)) return GH_ERR_ARG;
  if (readable < slice_bytes || entry_bit > 0xffffu) return GH_ERR_ARG;
  const DecLayout L = dec_layout(slice_bytes);
  if (ws_bytes < L.total) return GH_ERR_SPACE;
  DecWorkspace ws = dec_bind(d_ws, L);
  DecControl h_ctl;
  DecGeometry g;
  g.payload = d_payload;
  g.slice_bits = slice_bytes * 8;
  g.readable = readable;
  g.entry0 = entry_bit;
  bool fine = false;

  if (first_call) {
    DecodeTables tables;
    int rc = build_decode_tables(code, &tables);
    if (rc != GH_OK) return rc;
    GH_CUDA_TRY(cudaMemcpyAsync(const_cast<DecodeTables*>(ws.tables), &tables, sizeof(tables), cudaMemcpyHostToDevice, stream));
    GH_LAUNCH(dec_build_luts_kernel, (1u << kLutMaxBits) / 256, 256, 0, stream, ws.tables, const_cast<uint16_t*>(ws.lut1),
              const_cast<uint16_t*>(ws.lutC), const_cast<LutWEntry*>(ws.lutW), const_cast<u32*>(ws.lutP));
    const bool slow_code = code->max_len - code->min_len <= 1;
    // measured (profiles/r1i): the fine pipeline is not yet faster than the coarse one, so it is opt-in
    fine = pipeline_choice() == 2;
    g.sub_bytes = choose_sub_bytes(slice_bytes);
    // near-fixed-length codes (all lengths within one bit: uniform-looking bytes) re-synchronise only when one of
    // the rare longer codewords shifts the phase; start them 4x coarser (measured on uniform bytes: the paths need
    // ~5.7k symbols on average to meet)
    if (slow_code && u64(g.sub_bytes) * 4 <= kMaxSubBytes) g.sub_bytes *= 4;
    if (fine) g.sub_bytes = kFineSubBytes;
  } else {
    GH_CUDA_TRY(cudaMemcpyAsync(&h_ctl, ws.ctl, sizeof(h_ctl), cudaMemcpyDeviceToHost, stream));
    GH_CUDA_TRY(cudaStreamSynchronize(stream));
    g.sub_bytes = h_ctl.sub_bytes;
    fine = h_ctl.fine != 0;
    if (g.sub_bytes < kMinSubBytes || (g.sub_bytes % kMinSubBytes)) return GH_ERR_ARG;
  }
  // Codes of 8 and 9 bits only whose 9-bit codewords are those with eight leading zeros: phase walk (K5c), no rounds.
  bool phase = false;
  u32 eof_v = 2;  // value of the end mark's ninth bit (2: the end mark is not a 9-bit codeword)
  if (code->min_len == 8 && code->max_len == 9 && code->first_code[8] == 1 && code->first_code[9] == 0 &&
      slice_bytes >= 4 * kPhaseMinSub && pipeline_choice() != 2 && !g_no_phase_walk) {
    phase = first_call ? true : h_ctl.phase != 0;
    for (u32 v = 0; v < 2; ++v) {
      const u32 idx = code->start_pos[9] + v;
      if (idx < u32(GH_NSYM) && code->symbol[idx] == u32(GH_EOF_SYMBOL)) eof_v = v;
    }
  }
  if (phase) {
    if (first_call) {  // no reason for coarse subsequences here: the phase walk needs no path to meet another
      g.sub_bytes = choose_sub_bytes(slice_bytes);
      if (g.sub_bytes < kPhaseMinSub) g.sub_bytes = kPhaseMinSub;
      fine = false;
    }
    g.n_sub = (slice_bytes + g.sub_bytes - 1) / g.sub_bytes;
    const unsigned blocks = unsigned((g.n_sub + kDecThreads - 1) / kDecThreads);
    const u64 n_tiles = (g.n_sub + kPhaseTile - 1) / kPhaseTile;
    const unsigned tile_blocks = unsigned((n_tiles + kDecThreads - 1) / kDecThreads);
    h_ctl = DecControl();
    h_ctl.eof_index = kNoEof;
    h_ctl.sub_bytes = g.sub_bytes;
    h_ctl.n_sub = g.n_sub;
    h_ctl.phase = 1u;
    GH_CUDA_TRY(cudaMemcpyAsync(ws.ctl, &h_ctl, sizeof(h_ctl), cudaMemcpyHostToDevice, stream));
    GH_LAUNCH(dec_phase_walk_kernel, blocks, kDecThreads, 0, stream, g, ws, eof_v);
    GH_LAUNCH(dec_phase_tiles_kernel, tile_blocks, kDecThreads, 0, stream, g, ws);
    GH_LAUNCH(dec_phase_scan_kernel, 1, kPhaseScanThreads, 0, stream, g, ws);
    GH_LAUNCH(dec_phase_entries_kernel, tile_blocks, kDecThreads, 0, stream, g, ws, ws.work[0]);
    GH_LAUNCH(dec_phase_finish_kernel, blocks, kDecThreads, 0, stream, g, ws, (const u32*)ws.work[0]);
    int rc = check_launch();
    if (rc != GH_OK) return rc;
    if (geom_out) *geom_out = g;
    if (fine_out) *fine_out = false;
    if (deferred) {  // the phase walk needs no rounds: nothing to confirm, only the totals to read later
      *deferred = true;
      return dec_finish_launch(g, ws, stream, false);
    }
    rc = dec_finish(g, ws, &h_ctl, stream, false);
    if (rc != GH_OK) return rc;
    if (result) {
      result->n_symbols = h_ctl.total;
      result->exit_bit = h_ctl.exit_bit;
      result->eof_found = h_ctl.eof_found;
      result->rounds = 1;
      result->sub_bytes = g.sub_bytes;
    }
    return GH_OK;
  }
  // Synchronisation rounds until a clean one. Round k makes subsequences 0..k exact whatever the data, so this
  // terminates after at most n_sub rounds; with codes that self-synchronise it takes two or three.
  // A code that synchronises slowly relative to the subsequence size (near-fixed-length codes: uniform bytes)
  // shows up in the first round as many exits that move (the corrected path never met the speculated one inside
  // the subsequence); the expected number of rounds is then log(n_sub) / log(1 / that fraction), so the
  // subsequences are made 4x coarser (fraction -> fraction^4) and the speculation is redone -- one extra pass
  // instead of dozens of rounds.
  u32 rounds = 0;
  bool speculate = first_call != 0, coarsened = false;
  bool rewalk = code->max_len - code->min_len <= 1;  // near-fixed-length codes: see dec_speculate_kernel
  while (true) {
    g.n_sub = (slice_bytes + g.sub_bytes - 1) / g.sub_bytes;
    const unsigned blocks = unsigned((g.n_sub + kDecThreads - 1) / kDecThreads);
    if (speculate) {
#ifdef GH_EXPERIMENTS
      if (fine) {
        int rc = set_fine_attrs();
        if (rc != GH_OK) return rc;
        GH_LAUNCH(dec_fine_speculate_kernel, unsigned((g.n_sub + kFineWarps - 1) / kFineWarps), kFineWarps * 32,
                  sizeof(SmemFineSpec), stream, g, ws);
      } else
#endif
        GH_LAUNCH(dec_speculate_kernel, blocks, kDecThreads, 0, stream, g, ws, 0, (const u32*)nullptr, 0u, (u32*)nullptr);
      int rc = check_launch();
      if (rc != GH_OK) return rc;
    }
    if (!rewalk && !fine && !coarsened) {
      // Optimistic first round: one synchronisation round and the offset kernels are enqueued together and the control
      // block is read back ONCE (or not at all here, when the caller defers it). A self-synchronising code is at its
      // fixed point after that round; if an exit did move, the rounds below take over.
      h_ctl = DecControl();
      h_ctl.eof_index = kNoEof;
      h_ctl.sub_bytes = g.sub_bytes;
      h_ctl.n_sub = g.n_sub;
      GH_CUDA_TRY(cudaMemcpyAsync(ws.ctl, &h_ctl, sizeof(h_ctl), cudaMemcpyHostToDevice, stream));
      // re-entry with a corrected first-codeword position: staleness only travels rightwards through exits that
      // move, so the first round needs the first block only
      const unsigned first_blocks = (!first_call && !speculate) ? 1u : blocks;
      GH_LAUNCH(dec_sync_kernel, first_blocks, kDecThreads, 0, stream, g, ws);
      int rc = check_launch();
      if (rc == GH_OK) rc = dec_finish_launch(g, ws, stream, false);
      if (rc != GH_OK) return rc;
      if (geom_out) *geom_out = g;
      if (fine_out) *fine_out = false;
      if (deferred) {
        *deferred = true;
        return GH_OK;
      }
      rc = dec_read_ctl(ws, &h_ctl, stream);
      if (rc != GH_OK) return rc;
      ++rounds;
      if (!h_ctl.changed) {
        if (result) {
          result->n_symbols = h_ctl.total;
          result->exit_bit = h_ctl.exit_bit;
          result->eof_found = h_ctl.eof_found;
          result->rounds = rounds;
          result->sub_bytes = g.sub_bytes;
        }
        return GH_OK;
      }
      if (u64(h_ctl.exits_changed) * 10 > g.n_sub) rewalk = true;  // the paths do not meet early with this code
    }
    bool coarsen = false;
    bool have_list = false;  // re-walk rounds: work[cur] lists the subsequences to walk, work_n of them
    u32 cur = 0, work_n = 0;
    auto reset_ctl = [&]() -> int {
      h_ctl = DecControl();
      h_ctl.eof_index = kNoEof;
      h_ctl.sub_bytes = g.sub_bytes;
      h_ctl.n_sub = g.n_sub;
      h_ctl.fine = fine ? 1u : 0u;
      GH_CUDA_TRY(cudaMemcpyAsync(ws.ctl, &h_ctl, sizeof(h_ctl), cudaMemcpyHostToDevice, stream));
      return GH_OK;
    };
    auto read_ctl = [&]() -> int {
      GH_CUDA_TRY(cudaMemcpyAsync(&h_ctl, ws.ctl, sizeof(h_ctl), cudaMemcpyDeviceToHost, stream));
      GH_CUDA_TRY(cudaStreamSynchronize(stream));
      return GH_OK;
    };
    for (u32 level_round = 0;; ++level_round) {
      int rc = reset_ctl();
      if (rc != GH_OK) return rc;
      if (rewalk && !fine) {
        // Re-walk rounds run over a dense list of the stale subsequences: a round costs what it re-walks, not a
        // pass over every subsequence with mostly idle lanes (uniform bytes: ten rounds, the later ones tiny).
        if (!have_list) {
          GH_LAUNCH(dec_worklist_kernel, blocks, kDecThreads, 0, stream, g, ws, ws.work[cur]);
          rc = check_launch();
          if (rc == GH_OK) rc = read_ctl();
          if (rc != GH_OK) return rc;
          work_n = h_ctl.work_count;
          have_list = true;
          rc = reset_ctl();
          if (rc != GH_OK) return rc;
        }
        if (work_n) {
          GH_LAUNCH(dec_speculate_kernel, (work_n + kDecThreads - 1) / kDecThreads, kDecThreads, 0, stream, g, ws, 1,
                    (const u32*)ws.work[cur], work_n, ws.work[cur ^ 1]);
        }
      } else {
        // Re-entry with a corrected first-codeword position (sharded decode): everything but subsequence 0 was at
        // the fixed point, and staleness only travels rightwards through exits that move, so the first round needs
        // the first block only; a moved exit at its end sets `changed` and the following rounds are full ones.
        const unsigned sync_blocks = (!first_call && !speculate && level_round == 0) ? 1u : blocks;
        GH_LAUNCH(dec_sync_kernel, sync_blocks, kDecThreads, 0, stream, g, ws);
        have_list = false;
      }
      rc = check_launch();
      if (rc != GH_OK) return rc;
      rc = read_ctl();
      if (rc != GH_OK) return rc;
      if (rewalk && !fine) {
        cur ^= 1;
        work_n = h_ctl.work_count;
      }
      ++rounds;
      if (!h_ctl.changed) break;
      // more than a tenth of the exits moved: the paths do not meet early with this code, stop trying to
      if (u64(h_ctl.exits_changed) * 10 > g.n_sub) {
        rewalk = true;
        if (fine && pipeline_choice() == 0 && level_round == 0 && speculate) {  // misjudged: redo it coarsely
          fine = false;
          coarsen = true;
          break;
        }
      }
      // coarsen (once, and only while the grid still fills the GPU) when more than half of the exits moved in
      // the first round: ~log(n_sub)/log(1/fraction) rounds are ahead, and a round costs one subsequence walk
      if (level_round == 0 && speculate && !coarsened && u64(h_ctl.exits_changed) * 2 > g.n_sub &&
          g.n_sub / 4 >= u64(sm_count()) * 512 && u64(g.sub_bytes) * 4 <= kMaxSubBytes) {
        coarsen = true;
        break;
      }
      if (u64(level_round) > g.n_sub + 2) return GH_ERR_FORMAT;  // cannot happen: see above
    }
    if (!coarsen) break;
    coarsened = true;
    if (g.sub_bytes == kFineSubBytes && !fine) g.sub_bytes = choose_sub_bytes(slice_bytes);
    u64 bigger = u64(g.sub_bytes) * 4;
    g.sub_bytes = u32(bigger > kMaxSubBytes ? kMaxSubBytes : bigger);
    speculate = true;
  }

  int rc = dec_finish(g, ws, &h_ctl, stream, fine);
  if (rc != GH_OK) return rc;
  if (result) {
    result->n_symbols = h_ctl.total;
    result->exit_bit = h_ctl.exit_bit;
    result->eof_found = h_ctl.eof_found;
    result->rounds = rounds;
    result->sub_bytes = g.sub_bytes;
  }
  if (geom_out) *geom_out = g;
  if (fine_out) *fine_out = fine;
  return GH_OK;
}

static int decode_write_impl(const DecGeometry& g, u32 min_len, bool fine, uint8_t* d_out, u64 out_cap, void* d_ws,
                             cudaStream_t stream) {
  const DecLayout L = dec_layout(g.slice_bits / 8);
  DecWorkspace ws = dec_bind(d_ws, L);
#ifdef GH_EXPERIMENTS
  if (fine) {
    int rc = set_fine_attrs();
    if (rc != GH_OK) return rc;
    const unsigned blocks = unsigned((g.n_sub + kFineWarps - 1) / kFineWarps);
    GH_LAUNCH(dec_fine_write_kernel, blocks, kFineWarps * 32, sizeof(SmemFineWrite), stream, g, d_out, out_cap, ws);
    return check_launch();
  }
#endif
  (void)fine;
  int rc = set_write_attrs();
  if (rc != GH_OK) return rc;
  const unsigned tiles = unsigned((g.n_sub + kDecThreads - 1) / kDecThreads);
  if (min_len >= kSmallRingMinLen)
    GH_LAUNCH(dec_write_kernel<WriteSmall>, tiles, kDecThreads, sizeof(WriteSmall::Smem), stream, g, d_out, out_cap, ws);
  else
    GH_LAUNCH(dec_write_kernel<WriteLarge>, tiles, kDecThreads, sizeof(WriteLarge::Smem), stream, g, d_out, out_cap, ws);
  return check_launch();
}

}  // namespace gh

extern "C" {

size_t gh_decode_workspace_bytes(uint64_t payload_bytes) { return gh::dec_layout(payload_bytes).total; }

// test hooks (explicit calls; the library reads no environment variables)
void gh_debug_select_writer(int pipeline) {
#ifdef GH_EXPERIMENTS
  gh::g_pipeline = pipeline < 0 || pipeline > 2 ? 0 : pipeline;
#else
  gh::g_pipeline = pipeline == 1 ? 1 : 0;  // the fine pipeline (2) exists in experiment builds only
#endif
}
void gh_debug_disable_phase_walk(int off) { gh::g_no_phase_walk = off != 0; }

int gh_decode_sync(const uint8_t* d_payload, uint64_t slice_bytes, uint64_t readable_bytes, const gh_code* code,
                   uint32_t entry_bit, int first_call, gh_shard_sync* result, void* d_workspace,
                   size_t workspace_bytes, void* stream) {
  return gh::decode_sync_impl(d_payload, slice_bytes, readable_bytes, code, entry_bit, first_call, result, d_workspace,
                              workspace_bytes, (cudaStream_t)stream, nullptr);
}

int gh_decode_write(const uint8_t* d_payload, uint64_t slice_bytes, uint64_t readable_bytes, const gh_code* code,
                    uint8_t* d_out, uint64_t out_cap, void* d_workspace, size_t workspace_bytes, void* stream) {
  using namespace gh;
  (void)code;
  if (!d_payload || !d_workspace || (!d_out && out_cap)) return GH_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(d_payload) & 15) || (reinterpret_cast<uintptr_t>(d_workspace) & 255)) return GH_ERR_ARG;
  const DecLayout L = dec_layout(slice_bytes);
  if (workspace_bytes < L.total) return GH_ERR_SPACE;
  DecWorkspace ws = dec_bind(d_workspace, L);
  DecControl h_ctl;
  GH_CUDA_TRY(cudaMemcpyAsync(&h_ctl, ws.ctl, sizeof(h_ctl), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  GH_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  DecGeometry g;
  g.payload = d_payload;
  g.slice_bits = slice_bytes * 8;
  g.readable = readable_bytes;
  g.sub_bytes = h_ctl.sub_bytes;
  g.n_sub = h_ctl.n_sub;
  g.entry0 = 0;
  if (g.sub_bytes < kMinSubBytes || g.n_sub != (slice_bytes + g.sub_bytes - 1) / g.sub_bytes) return GH_ERR_ARG;
  return decode_write_impl(g, code->min_len, h_ctl.fine != 0, d_out, out_cap, d_workspace, (cudaStream_t)stream);
}

int gh_decode(const uint8_t* d_payload, uint64_t payload_bytes, const gh_code* code, uint8_t* d_out,
              uint64_t out_cap, uint64_t* n_out, void* d_workspace, size_t workspace_bytes, void* stream) {
  return gh::decode_full(d_payload, payload_bytes, code, 0, d_out, out_cap, n_out, d_workspace, workspace_bytes, stream);
}

}  // extern "C"

namespace gh {

int decode_full(const uint8_t* d_payload, uint64_t payload_bytes, const gh_code* code, uint32_t entry_bit,
                uint8_t* d_out, uint64_t out_cap, uint64_t* n_out, void* d_workspace, size_t workspace_bytes,
                void* stream) {
  if (!d_out && out_cap) return GH_ERR_ARG;
  gh_shard_sync res;
  DecGeometry g;
  bool fine = false, deferred = false;
  // first attempt: everything enqueued back to back, the control block read once at the end
  int rc = decode_sync_impl(d_payload, payload_bytes, payload_bytes, code, entry_bit, 1, &res, d_workspace,
                            workspace_bytes, (cudaStream_t)stream, &g, &fine, &deferred);
  if (rc != GH_OK) return rc;
  rc = decode_write_impl(g, code->min_len, fine, d_out, out_cap, d_workspace, (cudaStream_t)stream);
  if (rc != GH_OK) return rc;
  if (deferred) {
    const DecLayout L = dec_layout(payload_bytes);
    DecWorkspace ws = dec_bind(d_workspace, L);
    DecControl h_ctl;
    rc = dec_read_ctl(ws, &h_ctl, (cudaStream_t)stream);
    if (rc != GH_OK) return rc;
    if (h_ctl.changed) {
      // the single round was not clean: what the writer stored is not final. Again, with as many rounds as it takes.
      rc = decode_sync_impl(d_payload, payload_bytes, payload_bytes, code, entry_bit, 1, &res, d_workspace, workspace_bytes,
                            (cudaStream_t)stream, &g, &fine, nullptr);
      if (rc != GH_OK) return rc;
      rc = decode_write_impl(g, code->min_len, fine, d_out, out_cap, d_workspace, (cudaStream_t)stream);
      if (rc != GH_OK) return rc;
      GH_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    } else {
      res.n_symbols = h_ctl.total;
      res.eof_found = h_ctl.eof_found;
    }
  } else {
    GH_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  }
  if (n_out) *n_out = res.n_symbols;
  if (!res.eof_found) return GH_ERR_NO_EOF;
  if (res.n_symbols > out_cap) return GH_ERR_SPACE;
  return GH_OK;
}

}  // namespace gh
