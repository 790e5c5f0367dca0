// K2-K4 -- LUT gather, device-wide bit-offset scan and bit packing in ONE pass over the input
// (replaces reference include/canonical_huff_encoder.cc:245-285 encode_file/encode_each_byte and the
// per-bit writer utils/include/buffer.h:241-248,277-295).
//
// Algorithmic traffic: N bytes read + C bytes written (the scan is fused by decoupled look-back, so the
// input is not read a second time to learn the bit offsets, and every input byte is looked up ONCE).
//
// One persistent CTA of 1024 threads per SM, split into four independent GROUPS of 256 threads that share one
// lookup table and synchronise among themselves with named barriers only. A group takes TILES of input from an
// atomic ticket; a tile is kSubTiles sub-tiles of 8 KiB, a thread owns 32 consecutive bytes of a sub-tile.
//   1. gather + concatenate: per byte one PRMT (which forms the complete shared-memory address: the table has one
//      64-bit entry (code << 16 | len, 2^len) per (byte value, lane) at 0x10000 + value * 256 + lane * 8, so the
//      gathers of a warp can never conflict) and one LDS.64; four codewords are concatenated into a 64-bit chunk by
//      multiply-adds with the powers of two (FMA pipe), the chunk length is the sum of the entries' low halves;
//   2. per sub-tile: warp-shuffle scan of the thread bit counts, the eight warp totals cross one named barrier;
//   3. staging: every chunk is OR-ed into the group's zeroed staging buffer at its TILE-relative bit position with
//      shared-memory atomics (neighbouring chunks share words). This needs no global offset, so it runs before
//      the tile's look-back has finished;
//   4. the tile's bit count is published as (AGGREGATE | bits) as soon as the last sub-tile is counted; after
//      staging, the group's first warp resolves the tile's start bit G by a decoupled look-back (Merrill & Garland)
//      over 96 predecessors per L2 round trip -- with 32 per round trip the chain of prefixes, not the SMs, bounds a
//      kernel this fast (16 KiB x 32 / 0.37 us = 1.4 TB/s) -- and publishes (PREFIX | end bit);
//   5. copy-out: global word (G/32 + i) = funnel shift of staged words i-1, i by (G mod 32), byte-swapped to
//      stream order, coalesced. The tile's first word, when shared with the previous tile, is NOT stored: its
//      bits go to head[tile] and encode_stitch_kernel ORs them into the word the previous tile wrote -- every
//      output word has exactly one writer per kernel, no global atomics on the payload. The last tile adds the
//      1-padding (reference include/canonical_huff_encoder.cc:255-257); the staging buffer is cleared again.
// Codewords longer than 16 bits cannot take part in step 1 (four of them do not fit a 64-bit chunk). Their table
// entries carry a flag that survives the length sum; a warp whose 1 KiB slice of the sub-tile contains one -- or the
// ragged end of the input, or the byte that the end mark follows -- handles that slice codeword by codeword
// (count_slow / stage_slow). A Huffman code gives such lengths only to symbols rarer than 2^-16 or so.
// When the code has such codewords, tiles are one sub-tile (8 KiB), so that even a tile of nothing but 32-bit
// codewords fits the staging buffer.
#include "gh_common.cuh"

namespace gh {

constexpr int kEncGroupThreads = 256;
constexpr int kEncGroups = 4;
constexpr int kEncThreads = kEncGroupThreads * kEncGroups;  // one CTA per SM
constexpr int kEncGroupWarps = kEncGroupThreads / 32;
constexpr int kEncBytesPerThread = 32;
constexpr int kEncChunks = kEncBytesPerThread / 4;                        // 64-bit chunks of four codewords
constexpr int kEncSubTileBytes = kEncGroupThreads * kEncBytesPerThread;  // 8 KiB
constexpr int kEncRowBytes = 32 * kEncBytesPerThread;                    // a warp's slice of a sub-tile
constexpr int kEncMaxSubTiles = 2;
constexpr int kEncMinTileBytes = kEncSubTileBytes;
constexpr int kEncLookDepth = 3;    // predecessors per lane and look-back round
constexpr u32 kEncLongFlag = 0x1000u;  // in the low half of a table entry: codeword longer than 16 bits

// staging buffer of a group: worst case 16384 codewords of 16 bits, or 8192 of 32 bits, plus the end mark; four zero
// words in front (staging ORs up to two words below a chunk's last word, copy-out reads word -1; 16-byte aligned
// clearing) and slack behind
constexpr int kEncStageFront = 4;
constexpr int kEncStageWords = kEncMaxSubTiles * kEncSubTileBytes * 16 / 32 + 2 + 6;
constexpr int kEncStageBytes = (kEncStageFront + kEncStageWords) * 4;

// Shared memory. The replicated table must sit at the shared-window address 0x10000 so that a single PRMT can
// assemble an entry's address from (0x01, byte value, lane * 8); the rest is laid out around it.
constexpr u32 kEncLutAddr = 0x10000u;
constexpr u32 kEncLutBytes = 256u * 256u;
struct EncGroupCtl {
  u32 wtot[2][kEncGroupWarps];  // warp totals of a sub-tile, double-buffered
  u64 tile_start;
  u32 next_tile;
  u32 pad[3];
};
struct EncLowSmem {  // at the start of dynamic shared memory
  EncGroupCtl ctl[kEncGroups];
  uint2 long_table[GH_NSYM + 1];  // (codeword, length) of every symbol, single copy: slow path and end mark
};

constexpr u64 kFlagMask = 3ull << 62;
constexpr u64 kFlagAggregate = 1ull << 62;  // value = bits of this tile only
constexpr u64 kFlagPrefix = 2ull << 62;     // value = global bit offset just after this tile

struct EncWorkspace {
  u64* tile_state;  // [ntiles]
  u32* head;        // [ntiles]
  u32* ticket;      // tile dispenser (tiles are handed out in the order groups ask for them)
};

__host__ __device__ inline u64 enc_num_tiles(u64 n, u32 tile_bytes) { return (n + tile_bytes - 1) / tile_bytes; }

// ---- named barrier of one group -----------------------------------------------------------------------------------
__device__ __forceinline__ void group_barrier(unsigned group) {
#ifdef GH_EMUL
  gh_emul::named_barrier(group + 1, kEncGroupThreads);
#else
  asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kEncGroupThreads) : "memory");
#endif
}

// ---- table gather --------------------------------------------------------------------------------------------------
#ifdef GH_EMUL
typedef const unsigned char* enc_lut_t;  // the table's base plus this lane's column
__device__ __forceinline__ uint2 enc_lut_entry(enc_lut_t lut_lane, u32 word, int k) {
  return *reinterpret_cast<const uint2*>(lut_lane + (((word >> (8 * k)) & 0xffu) << 8));
}
#else
typedef u32 enc_lut_t;  // kEncLutAddr | lane * 8
__device__ __forceinline__ uint2 enc_lut_entry(enc_lut_t lut_lane, u32 word, int k) {
  // address = 0x00 0x01 <byte k of word> <lane * 8>: bytes 3 and 2 and 0 from lut_lane, byte 1 from the input word
  const u32 addr = __byte_perm(word, lut_lane, 0x7604u | (u32(k) << 4));
  uint2 v;
  asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
#endif

// ---- staging -------------------------------------------------------------------------------------------------------
// The staging buffer is addressed by its shared-window address (device) so that the ORs are RED.OR [reg + imm] without
// generic-pointer arithmetic.
#ifdef GH_EMUL
typedef u32* enc_stage_t;
__device__ __forceinline__ enc_stage_t enc_stage_handle(u32* p) { return p; }
__device__ __forceinline__ void enc_stage_or(enc_stage_t st, u32 word, int delta, u32 v) { atomicOr(st + int(word) + delta, v); }
__device__ __forceinline__ void enc_stage_or_nz(enc_stage_t st, u32 word, int delta, u32 v) {
  if (v) atomicOr(st + int(word) + delta, v);
}
__device__ __forceinline__ u32 enc_stage_ld(enc_stage_t st, u32 word, int delta) { return st[int(word) + delta]; }
#else
typedef u32 enc_stage_t;
__device__ __forceinline__ enc_stage_t enc_stage_handle(u32* p) { return u32(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void enc_stage_or(enc_stage_t st, u32 word, int delta, u32 v) {
  asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(st + word * 4u + u32(delta * 4)), "r"(v) : "memory");
}
__device__ __forceinline__ void enc_stage_or_nz(enc_stage_t st, u32 word, int delta, u32 v) {
  asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %1, 0;\n @p red.shared.or.b32 [%0], %1;\n}" ::"r"(st + word * 4u + u32(delta * 4)), "r"(v)
               : "memory");
}
__device__ __forceinline__ u32 enc_stage_ld(enc_stage_t st, u32 word, int delta) {
  u32 v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(st + word * 4u + u32(delta * 4)) : "memory");
  return v;
}
#endif

// OR the low `len` bits of hi:lo (len in 1..64, higher bits zero) into the bit string so that they end just before
// bit `end` (= position + len). With W = end / 32 and e = end % 32, word W takes the value's last e bits at its top
// ((lo:0) >> e), word W-1 the 32 bits before them ((hi:lo) >> e) and word W-2 the rest (hi >> e): three funnel shifts
// by `end` itself (SHF takes its distance modulo 32), no other arithmetic. Words that receive nothing are OR-ed with
// zero (W, W-1) or skipped (W-2, rarely non-zero): the buffer has spare words on both sides.
__device__ __forceinline__ void stage_chunk(enc_stage_t stage, u32 end, u32 lo, u32 hi) {
  const u32 w = end >> 5;
  enc_stage_or(stage, w, 0, __funnelshift_r(0u, lo, end));
  enc_stage_or(stage, w, -1, __funnelshift_r(lo, hi, end));
  enc_stage_or_nz(stage, w, -2, __funnelshift_r(hi, 0u, end));
}

struct EncVec {
  u32 w[8];
};

__device__ __forceinline__ EncVec enc_load_vec(const uint8_t* p, bool aligned32) {
  const Unit8 u = ldg_unit(p, aligned32);
  EncVec v;
#pragma unroll
  for (int k = 0; k < 8; ++k) v.w[k] = u.w[k];
  return v;
}

// slow path: bit count / staging of the `cnt` bytes at p, codeword by codeword (bytes re-read from global memory)
__device__ __noinline__ u32 count_slow(const uint2* long_table, const uint8_t* p, int cnt) {
  u32 bits = 0;
  for (int k = 0; k < cnt; ++k) bits += long_table[p[k]].y;
  return bits;
}
__device__ __noinline__ u32 stage_slow(enc_stage_t stage, u32 pos, const uint2* long_table, const uint8_t* p, int cnt) {
  for (int k = 0; k < cnt; ++k) {
    const uint2 e = long_table[p[k]];
    pos += e.y;
    if (e.y) stage_chunk(stage, pos, e.x, 0u);
  }
  return pos;
}

// Decoupled look-back by one whole warp. The tile's own count was published as an AGGREGATE before (tiles > 0);
// this adds up the predecessors' counts back to the nearest tile that already knows its start, publishes this
// tile's end bit and returns its start bit (all 32 lanes must call it together).
__device__ __forceinline__ u64 tile_start_lookback(const EncWorkspace& ws, u64 tile, u32 tile_bits, u64 start_bit, unsigned lane) {
  u64 exclusive = start_bit;
  if (tile == 0) {
    if (lane == 0) st_volatile_u64(ws.tile_state, kFlagPrefix | (start_bit + tile_bits));
    return exclusive;
  }
  exclusive = 0;
  long long look = (long long)tile - 1;
  bool done = false;
  while (!done) {
    // One round trip covers 32 x kEncLookDepth predecessors: lane l owns the consecutive tiles
    // look - l * kEncLookDepth - r (r = 0 nearest), loads all of them at once, folds them locally (sum of
    // aggregates up to and including its nearest PREFIX) and the warp then needs ONE ballot + ONE sum.
    const long long first_idx = look - (long long)lane * kEncLookDepth;
    u64 st[kEncLookDepth];
#pragma unroll
    for (int r = 0; r < kEncLookDepth; ++r) {
      const long long idx = first_idx - r;
      st[r] = idx >= 0 ? ld_volatile_u64(ws.tile_state + idx) : kFlagPrefix;  // virtual tiles before tile 0 add nothing
    }
    u64 local = 0;
    bool local_prefix = false;
#pragma unroll
    for (int r = 0; r < kEncLookDepth; ++r) {
      const long long idx = first_idx - r;
      while ((st[r] & kFlagMask) == 0) st[r] = ld_volatile_u64(ws.tile_state + idx);  // not published yet
      if (!local_prefix) local += st[r] & ~kFlagMask;
      local_prefix = local_prefix || (st[r] & kFlagMask) == kFlagPrefix;
    }
    const unsigned has_prefix = __ballot_sync(0xffffffffu, local_prefix);
    // lanes up to and including the nearest one that found a prefix contribute
    const unsigned first = unsigned(__ffs(int(has_prefix))) - 1u;
    const u64 contrib = (has_prefix == 0 || lane <= first) ? local : 0ull;
    exclusive += warp_sum64(contrib);
    done = has_prefix != 0;
    look -= 32 * kEncLookDepth;
  }
  if (lane == 0) st_volatile_u64(ws.tile_state + tile, kFlagPrefix | (exclusive + tile_bits));
  return exclusive;
}

template <int kSubTiles>
__global__ void __launch_bounds__(kEncThreads, 1)
encode_kernel(const uint8_t* __restrict__ in, u64 n, const EncodeTable table, u64 start_bit, int append_eof,
              u32* __restrict__ out_words, u64 out_word_cap, u64* __restrict__ end_bit_out, EncWorkspace ws,
              u32 smem_bytes) {
  constexpr u32 kTileBytes = u32(kSubTiles) * kEncSubTileBytes;
  GH_DYNAMIC_SMEM(smem_raw);
  const unsigned group = threadIdx.x / kEncGroupThreads, tg = threadIdx.x % kEncGroupThreads;
  const unsigned lane = tg & 31, wg = tg >> 5;

  // ---- carve shared memory around the table's fixed address ------------------------------------------------------
  EncLowSmem& low = *reinterpret_cast<EncLowSmem*>(smem_raw);
#ifdef GH_EMUL
  const u32 lut_off = kEncLutAddr - 0x400u;
#else
  const u32 smem_base = u32(__cvta_generic_to_shared(smem_raw));
  const u32 lut_off = kEncLutAddr - smem_base;
  if (smem_base > kEncLutAddr - u32(sizeof(EncLowSmem)) - u32(kEncStageBytes) - 16u ||
      lut_off + kEncLutBytes + 3u * u32(kEncStageBytes) > smem_bytes)
    __trap();  // the launch did not provide the window this layout needs
#endif
  (void)smem_bytes;
  unsigned char* const lut_ptr = smem_raw + lut_off;
  // group 0 stages below the table, groups 1..3 above it
  u32* const stage_base = group == 0
                              ? reinterpret_cast<u32*>(smem_raw + ((sizeof(EncLowSmem) + 15) & ~size_t(15)))
                              : reinterpret_cast<u32*>(lut_ptr + kEncLutBytes + size_t(group - 1) * kEncStageBytes);
  u32* const stage_ptr = stage_base + kEncStageFront;
  const enc_stage_t stage = enc_stage_handle(stage_ptr);
  EncGroupCtl& ctl = low.ctl[group];

  // ---- tables (once per CTA) ---------------------------------------------------------------------------------------
  for (unsigned i = threadIdx.x; i < 256u * 32u; i += kEncThreads) {
    const unsigned sym = i >> 5, col = i & 31;
    const u32 len = table.length[sym];
    uint2 e;
    if (len == 0) e = make_uint2(0u, 1u);                                     // byte value that does not occur
    else if (len <= 16) e = make_uint2((table.codeword[sym] << 16) | len, 1u << len);
    else e = make_uint2(kEncLongFlag, 1u);                                    // handled by the slow path
    *reinterpret_cast<uint2*>(lut_ptr + sym * 256u + col * 8u) = e;
  }
  for (unsigned i = threadIdx.x; i < unsigned(GH_NSYM); i += kEncThreads)
    low.long_table[i] = make_uint2(table.codeword[i], u32(table.length[i]));
  for (unsigned i = tg; i < unsigned(kEncStageFront + kEncStageWords); i += kEncGroupThreads) stage_base[i] = 0;
  if (tg == 0) ctl.next_tile = atomicAdd(ws.ticket, 1u);
  __syncthreads();
#ifdef GH_EMUL
  const enc_lut_t lut_lane = lut_ptr + lane * 8u;
#else
  const enc_lut_t lut_lane = kEncLutAddr | (lane * 8u);
#endif
  const uint2* const long_table = low.long_table;
  const u64 ntiles = enc_num_tiles(n, kTileBytes);
  const bool aligned32 = (reinterpret_cast<uintptr_t>(in) & 31) == 0;
  const u32 eof_code = table.codeword[GH_EOF_SYMBOL];
  const u32 eof_len = append_eof ? u32(table.length[GH_EOF_SYMBOL]) : 0u;
  const u32 toff = tg * kEncBytesPerThread;           // this thread's slice inside a sub-tile
  const u32 roff = (tg & ~31u) * kEncBytesPerThread;  // its warp's row
  u32 wpar = 0;  // slot of wtot the next sub-tile uses

  u64 tile = ctl.next_tile;
  while (tile < ntiles) {
    const u64 tile_base = tile * kTileBytes;
    const uint8_t* const tin = in + tile_base;
    const u64 left = n - tile_base;
    const u32 rem = left < u64(kTileBytes) ? u32(left) : kTileBytes;  // bytes of this tile (short only for the last one)
    const bool last_tile = (tile + 1 == ntiles);
    // all of the tile's loads up front
    EncVec vec[kSubTiles];
#pragma unroll
    for (int j = 0; j < kSubTiles; ++j) {
      const u32 off = u32(j) * kEncSubTileBytes + toff;
      if (off + kEncBytesPerThread <= rem) {
        vec[j] = enc_load_vec(tin + off, aligned32);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) vec[j].w[k] = 0;
      }
    }
    u32 sub_base = 0;  // bits of this tile in earlier sub-tiles
    u32 tile_bits = 0;
#pragma unroll
    for (int j = 0; j < kSubTiles; ++j) {
      const u32 off = u32(j) * kEncSubTileBytes + toff;
      const u32 row = u32(j) * kEncSubTileBytes + roff;
      // ---- 1. gather + concatenate ---------------------------------------------------------------------------
      u32 c_lo[kEncChunks], c_hi[kEncChunks], c_end[kEncChunks];
      u32 bits = 0, flags = 0;
#pragma unroll
      for (int c = 0; c < kEncChunks; ++c) {
        const u32 word = vec[j].w[c];
        const uint2 e0 = enc_lut_entry(lut_lane, word, 0);
        const uint2 e1 = enc_lut_entry(lut_lane, word, 1);
        const uint2 e2 = enc_lut_entry(lut_lane, word, 2);
        const uint2 e3 = enc_lut_entry(lut_lane, word, 3);
        // acc = ((c0 * 2^l1 + c1) * 2^l2 * 2^l3) + (c2 * 2^l3 + c3): products on the FMA pipe; the additions cannot
        // carry because the products' low bits are zero
        const u32 a01 = __umulhi(e0.x, 1u << 16) * e1.y + __umulhi(e1.x, 1u << 16);  // <= 32 bits
        const u32 a23 = __umulhi(e2.x, 1u << 16) * e3.y + __umulhi(e3.x, 1u << 16);  // <= 32 bits
        const u64 t = u64(a01) * e2.y;                                               // <= 48 bits
        const u64 acc = t * e3.y + a23;
        c_lo[c] = u32(acc);
        c_hi[c] = u32(acc >> 32);
        const u32 l = (e0.x + e1.x + e2.x + e3.x) & 0xffffu;  // lengths (and long-codeword flags) add up in the low half
        flags |= l;
        bits += l;
        c_end[c] = bits;  // end of this chunk relative to the thread's first bit
      }
      // slices that need the slow path: a long codeword, the ragged end, or the byte the end mark follows
      const bool row_live = row < rem;
      const bool row_end = last_tile && row + kEncRowBytes >= rem;
      bool slow = false;
      int cnt = 0;
      bool owns_end = false;
      if (row_live) {
        slow = row_end || __any_sync(0xffffffffu, (flags & 0xf000u) != 0u);
        if (slow) {
          cnt = off >= rem ? 0 : (rem - off < u32(kEncBytesPerThread) ? int(rem - off) : kEncBytesPerThread);
          bits = count_slow(long_table, tin + off, cnt);
          owns_end = last_tile && cnt > 0 && off + u32(cnt) == rem;
          if (owns_end) bits += eof_len;
        }
      } else {
        bits = 0;
      }
      // ---- 2. scan ---------------------------------------------------------------------------------------------
      const u32 incl = warp_inclusive_scan(bits, lane);
      if (lane == 31) ctl.wtot[wpar][wg] = incl;
      group_barrier(group);
      u32 wprefix, sub_bits;
      {
        // the eight warp totals, scanned by every warp for itself
        u32 v = lane < u32(kEncGroupWarps) ? ctl.wtot[wpar][lane] : 0u;
#pragma unroll
        for (int d = 1; d < kEncGroupWarps; d <<= 1) {
          const u32 up = __shfl_up_sync(0xffffffffu, v, d);
          if (lane >= unsigned(d)) v += up;
        }
        sub_bits = __shfl_sync(0xffffffffu, v, kEncGroupWarps - 1);
        const u32 mine = __shfl_sync(0xffffffffu, v, int(wg));
        wprefix = mine - __shfl_sync(0xffffffffu, incl, 31);
      }
      wpar ^= 1;
      if (j == kSubTiles - 1) {
        tile_bits = sub_base + sub_bits;
        if (tg == 0 && tile > 0) st_volatile_u64(ws.tile_state + tile, kFlagAggregate | u64(tile_bits));
      }
      // ---- 3. staging at tile-relative positions ----------------------------------------------------------------
      const u32 pos = sub_base + wprefix + incl - bits;
      if (!slow) {
        if (row_live) {
#pragma unroll
          for (int c = 0; c < kEncChunks; ++c) stage_chunk(stage, pos + c_end[c], c_lo[c], c_hi[c]);
        }
      } else {
        const u32 p2 = stage_slow(stage, pos, long_table, tin + off, cnt);
        if (owns_end && eof_len) stage_chunk(stage, p2 + eof_len, eof_code, 0u);
      }
      sub_base += sub_bits;
    }

    // ---- 4. the tile's start bit ---------------------------------------------------------------------------------
    if (wg == 0) {
      const u64 exclusive = tile_start_lookback(ws, tile, tile_bits, start_bit, lane);
      if (lane == 0) ctl.tile_start = exclusive;
    }
    group_barrier(group);  // staging complete, tile start known
    // The next tile is drawn only now: a tile number that is held while its holder still waits for its own
    // predecessors delays every later tile (they need this one's bit count), so it is held as briefly as possible.
    if (tg == 0) ctl.next_tile = atomicAdd(ws.ticket, 1u);

    // ---- 5. copy-out ------------------------------------------------------------------------------------------------
    const u64 G = ctl.tile_start;
    const u32 phase = u32(G) & 31u;
    const u64 end_bit = G + tile_bits;
    const u64 word0 = G >> 5;
    const u32 nwords = (phase + tile_bits + 31u) >> 5;  // tile_bits > 0: every tile holds at least one codeword
    if (last_tile && tg == 0 && end_bit_out) *end_bit_out = end_bit;
    {
      u32* const dst = out_words + word0;
      const u64 room = out_word_cap > word0 ? out_word_cap - word0 : 0;
      const u32 lim = u64(nwords - 1u) < room ? nwords - 1u : u32(room);  // interior words: 1 .. nwords - 2
      for (u32 i = tg + 1u; i < lim; i += kEncGroupThreads)
        dst[i] = be32(__funnelshift_r(enc_stage_ld(stage, i, 0), enc_stage_ld(stage, i, -1), phase));
      // first and last word of the tile
      if (tg < 2u && (tg == 0 || nwords > 1u)) {
        const u32 i = tg == 0 ? 0u : nwords - 1u;
        u32 v = __funnelshift_r(enc_stage_ld(stage, i, 0), enc_stage_ld(stage, i, -1), phase);
        if (append_eof && last_tile && i == nwords - 1u) {
          const u32 pad = u32((8 - (end_bit & 7)) & 7);  // flush_bits(): 1s up to the byte boundary
          if (pad) v |= ((1u << pad) - 1u) << (32u - (u32(end_bit & 31) + pad));
        }
        if (i == 0 && tile > 0 && phase != 0) ws.head[tile] = v;  // shared with the previous tile: stitched later
        else if (u64(i) < room) dst[i] = be32(v);
      }
    }
    group_barrier(group);  // every staged word has been read, the next tile's number is there
    tile = ctl.next_tile;
    {
      uint4* const z = reinterpret_cast<uint4*>(stage_ptr);
      for (u32 i = tg; i < (tile_bits >> 7) + 1u; i += kEncGroupThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

// Second kernel: OR each tile's deferred head bits into the word its predecessor stored.
// Thread = tile. The first tile (in order) holding head bits for a given word merges the whole run.
__global__ void __launch_bounds__(256)
encode_stitch_kernel(u64 ntiles, u32* __restrict__ out_words, u64 out_word_cap, EncWorkspace ws) {
  const u64 tile = u64(blockIdx.x) * blockDim.x + threadIdx.x;
  if (tile == 0 || tile >= ntiles) return;
  const u64 g = ws.tile_state[tile - 1] & ~kFlagMask;  // global start bit of `tile`
  if ((g & 31) == 0) return;                            // starts on a word boundary: nothing deferred
  const u64 word = g >> 5;
  if (tile >= 2) {
    const u64 gp = ws.tile_state[tile - 2] & ~kFlagMask;  // start of the previous tile
    if ((gp >> 5) == word && (gp & 31) != 0) return;       // the previous tile deferred into this word too: it leads
  }
  u32 bits = ws.head[tile];
  for (u64 nx = tile + 1; nx < ntiles; ++nx) {  // only ever iterates for degenerate sub-word tiles
    const u64 gn = ws.tile_state[nx - 1] & ~kFlagMask;
    if ((gn >> 5) != word) break;
    bits |= ws.head[nx];
  }
  if (word < out_word_cap) out_words[word] |= be32(bits);
}

inline size_t enc_ws_bytes(u64 n) {
  const u64 nt = enc_num_tiles(n, kEncMinTileBytes) + 1;
  return size_t(nt * 8 + ((nt * 4 + 7) / 8) * 8 + 256);
}

}  // namespace gh

extern "C" {

size_t gh_encode_workspace_bytes(uint64_t n) { return gh::enc_ws_bytes(n); }

uint64_t gh_encode_payload_capacity(uint64_t n, const gh_code* code, uint64_t start_bit) {
  const uint64_t max_len = code ? code->max_len : 32;
  const uint64_t bits = start_bit + n * max_len + 32 + 7;
  return ((bits + 127) / 128) * 16 + 16;
}

int gh_encode(const uint8_t* d_in, uint64_t n, const gh_code* code, uint64_t start_bit, int append_eof,
              uint8_t* d_payload, uint64_t payload_cap, uint64_t* d_end_bit, void* d_workspace,
              size_t workspace_bytes, void* stream) {
  if (code && payload_cap < gh_encode_payload_capacity(n, code, start_bit)) return GH_ERR_SPACE;
  return gh::encode_unchecked(d_in, n, code, start_bit, append_eof, d_payload, payload_cap, d_end_bit, d_workspace,
                              workspace_bytes, stream);
}

}  // extern "C"

namespace gh {

int encode_unchecked(const uint8_t* d_in, uint64_t n, const gh_code* code, uint64_t start_bit, int append_eof,
                     uint8_t* d_payload, uint64_t payload_cap, uint64_t* d_end_bit, void* d_workspace,
                     size_t workspace_bytes, void* stream) {
  if (!code || !d_payload || !d_workspace) return GH_ERR_ARG;
  if (n == 0) return GH_ERR_EMPTY;
  if (!d_in) return GH_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(d_in) & 15) || (reinterpret_cast<uintptr_t>(d_payload) & 15) ||
      (reinterpret_cast<uintptr_t>(d_workspace) & 7))
    return GH_ERR_ARG;
  if (workspace_bytes < enc_ws_bytes(n)) return GH_ERR_SPACE;
  EncodeTable table;
  int rc = build_encode_table(code, &table);
  if (rc != GH_OK) return rc;
  // codewords of byte values longer than 16 bits take the slow path and halve the tile (the end mark's own length
  // does not matter: it is staged once, by the slow path)
  bool long_codes = false;
  for (int s = 0; s < 256; ++s) long_codes = long_codes || table.length[s] > 16;
  const u32 tile_bytes = long_codes ? kEncSubTileBytes : 2 * kEncSubTileBytes;
  const u64 ntiles = enc_num_tiles(n, tile_bytes);
  if (ntiles > 0x7ffffff0ull) return GH_ERR_ARG;

  EncWorkspace ws;
  uint8_t* p = static_cast<uint8_t*>(d_workspace);
  ws.tile_state = reinterpret_cast<u64*>(p);
  p += (ntiles + 1) * 8;
  ws.head = reinterpret_cast<u32*>(p);
  p += (((ntiles + 1) * 4 + 7) / 8) * 8;
  ws.ticket = reinterpret_cast<u32*>(p);
  const size_t used = size_t(p - static_cast<uint8_t*>(d_workspace)) + 8;
  GH_CUDA_TRY(cudaMemsetAsync(d_workspace, 0, used, (cudaStream_t)stream));

  // all of the SM's shared memory: the table's fixed window address decides the layout (see the kernel)
  int dev = 0, smem_max = 0;
  GH_CUDA_TRY(cudaGetDevice(&dev));
  GH_CUDA_TRY(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  GH_CUDA_TRY(cudaFuncSetAttribute(encode_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
  GH_CUDA_TRY(cudaFuncSetAttribute(encode_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
  const u64 out_word_cap = payload_cap / 4;
  // persistent: one CTA per SM, four tile-taking groups each
  u64 blocks = u64(sm_count() > 0 ? sm_count() : 1);
  const u64 want = (ntiles + kEncGroups - 1) / kEncGroups;
  if (blocks > want) blocks = want;
  if (long_codes) {
    GH_LAUNCH(encode_kernel<1>, unsigned(blocks), kEncThreads, size_t(smem_max), stream, d_in, (u64)n, table, (u64)start_bit,
              append_eof, reinterpret_cast<u32*>(d_payload), out_word_cap, reinterpret_cast<u64*>(d_end_bit), ws, u32(smem_max));
  } else {
    GH_LAUNCH(encode_kernel<2>, unsigned(blocks), kEncThreads, size_t(smem_max), stream, d_in, (u64)n, table, (u64)start_bit,
              append_eof, reinterpret_cast<u32*>(d_payload), out_word_cap, reinterpret_cast<u64*>(d_end_bit), ws, u32(smem_max));
  }
  rc = check_launch();
  if (rc != GH_OK) return rc;
  if (ntiles > 1) {
    GH_LAUNCH(encode_stitch_kernel, unsigned((ntiles + 255) / 256), 256, 0, stream, ntiles,
              reinterpret_cast<u32*>(d_payload), out_word_cap, ws);
    rc = check_launch();
  }
  return rc;
}

}  // namespace gh
