// K2-K4 -- LUT gather, device-wide bit-offset scan and bit packing in ONE pass over the input
// (replaces reference include/canonical_huff_encoder.cc:245-285 encode_file/encode_each_byte and the
// per-bit writer utils/include/buffer.h:241-248,277-295).
//
// Algorithmic traffic: N bytes read + C bytes written (the scan is fused by decoupled look-back, so the
// input is not read a second time to learn the bit offsets).
//
// Persistent blocks take tiles of kEncSubTiles x 4 KiB of input from an atomic ticket. Per tile:
//   1. every thread issues kEncSubTiles coalesced 128-bit loads up front (one 16-byte vector per 4 KiB sub-tile,
//      kept in registers) and sums the code lengths of its bytes from the shared-memory LUT, which is replicated
//      across the banks so that the gathers are (nearly) conflict-free (see EncSmem);
//   2. the per-thread bit counts are scanned per sub-tile (warp shuffles + one barrier); warp 0 publishes the
//      tile total and resolves the tile's global bit offset G by a warp-wide decoupled look-back over the
//      predecessors' (flag | value) words. One look-back per 16 KiB: the prefix can only travel 32 tiles per
//      L2 round trip, so with 4 KiB tiles that chain, not the SMs, set the pace (measured: profiles/r1b);
//   3. sub-tile by sub-tile: codewords are concatenated in registers into 64-bit chunks (4 per chunk when no
//      code is longer than 16 bits -- by multiply-adds with 2^len from the table --, else 2 by shifts) and OR-ed
//      into a zeroed shared staging buffer at sub-tile-relative bit positions (shared-memory atomics: neighbours
//      share words). Packing needs no global offset, so the other warps pack while warp 0 is still looking back.
//      The short-code variant rotates three staging buffers and needs one barrier per sub-tile;
//   4. copy-out: global word (G'/32 + i) = funnel-shift of staged words i-1, i by (G' mod 32), G' the sub-tile's
//      global bit offset -- phase alignment costs one SHF per output word -- byte-swapped to stream order,
//      coalesced. A sub-tile's trailing partial word is carried (shared memory) into the next sub-tile's first
//      word. The tile's first word, when shared with the previous tile, is NOT stored: its bits go to
//      head[tile] and a tiny second kernel ORs them into the word the previous tile wrote -- every output
//      word has exactly one writer per kernel, no global atomics on the payload. The last tile adds the
//      end-mark codeword and the 1-padding (reference include/canonical_huff_encoder.cc:255-257).
#include <stdlib.h>

#include "gh_common.cuh"

namespace gh {

constexpr int kEncThreads = 256;
constexpr int kEncBytesPerThread = 16;
constexpr int kEncSubTileBytes = kEncThreads * kEncBytesPerThread;  // 4 KiB
#ifndef GH_ENC_SUBTILES
#define GH_ENC_SUBTILES 4
#endif
#ifndef GH_ENC_BLOCKS_PER_SM
#define GH_ENC_BLOCKS_PER_SM 5
#endif
constexpr int kEncSubTiles = GH_ENC_SUBTILES;
constexpr int kEncTileBytes = kEncSubTileBytes * kEncSubTiles;      // 16 KiB per look-back
constexpr int kEncBlocksPerSm = GH_ENC_BLOCKS_PER_SM;
// lean interior copy-out loop: for which variants (measured r3b: helps the short-code variant, not the long-code one)
#ifndef GH_ENC_LEAN_COPY
#define GH_ENC_LEAN_COPY(syms_per_chunk) ((syms_per_chunk) == 4)
#endif
// fused gather (short-code variant): the chunks are built once, before the scans, and kept in registers (their
// lengths give the bit counts), instead of a length gather before the scans and a code gather after them
#ifndef GH_ENC_FUSED
#define GH_ENC_FUSED 0
#endif
#ifndef GH_ENC_LOOK_DEPTH
#define GH_ENC_LOOK_DEPTH 1
#endif
#ifndef GH_ENC_TICKET_SUBTILE
#define GH_ENC_TICKET_SUBTILE 3
#endif
#ifndef GH_ENC_POLL_SLEEP_NS
#define GH_ENC_POLL_SLEEP_NS 0
#endif
constexpr unsigned kEncPollSleepNs = GH_ENC_POLL_SLEEP_NS;  // pause between two polls of an unpublished tile state
constexpr int kEncLookDepth = GH_ENC_LOOK_DEPTH;        // look-back rounds whose loads are in flight together
constexpr int kEncTicketSubTile = GH_ENC_TICKET_SUBTILE;  // sub-tile during which the block draws its next tile

// Shared memory of the encode kernel, per variant (kSymsPerChunk = 4: no code longer than 16 bits; 2: up to 32).
//  * The (codeword, length) table is REPLICATED across the banks so that the per-byte gather is (nearly) free of
//    bank conflicts: with one copy, 32 lanes looking up 32 random bytes cost ~4 wavefronts per LDS and the L1 data
//    pipe was as busy as the ALUs (profiles/r2b: 262 M shared-load wavefronts for 67 M lookups).
//      variant 4: 64-bit entries  (code << 16 | len, 1 << len), 8 copies: slot (sym * 8 + lane % 8)     -- 16 KiB
//                 (the power of two lets the packer concatenate with multiply-adds on the FMA pipe instead of
//                 funnel shifts on the ALU pipe, which is the pipe this kernel saturates)
//      variant 2: 64-bit entries (len << 32) | code,  4 copies: slot (sym * 4 + lane % 4)              --  8 KiB
//    Lanes that share a copy are served by one broadcast when their bytes are equal, else serially.
//  * staging: worst case per sub-tile = 4096 symbols x max code length (+ end mark, + slack for the funnel shift).
//    Variant 4 rotates THREE staging buffers so that one barrier per sub-tile is enough (see the kernel).
template <int kSymsPerChunk>
struct EncSmem {
  static constexpr int kMaxLen = kSymsPerChunk == 4 ? 16 : 32;
  static constexpr int kStageWords = (kEncSubTileBytes * kMaxLen + 32 + 31) / 32 + 2;
  static constexpr int kBuffers = kSymsPerChunk == 4 ? 3 : 2;
  static constexpr int kCopies = kSymsPerChunk == 4 ? 8 : 4;
  static constexpr int kEntryBytes = 8;
  static constexpr int kSymStride = kCopies * kEntryBytes;  // bytes between consecutive symbols' entries
  static constexpr int kLutWords = 256 * kSymStride / 4;
  u32 lut[kLutWords];
  u32 stage[kBuffers][kStageWords];
  u32 warp_total[kEncSubTiles][kEncThreads / 32];
  u32 carry[2];
  u32 tile;
  u64 tile_start;
};

constexpr u64 kFlagMask = 3ull << 62;
constexpr u64 kFlagAggregate = 1ull << 62;  // value = bits of this tile only
constexpr u64 kFlagPrefix = 2ull << 62;     // value = global bit offset just after this tile

struct EncWorkspace {
  u64* tile_state;  // [ntiles]
  u32* head;        // [ntiles]
  u32* ticket;      // tile dispenser (tiles are handed out in the order blocks ask for them)
};

__host__ __device__ inline u64 enc_num_tiles(u64 n) { return (n + kEncTileBytes - 1) / kEncTileBytes; }

// the one partial vector at the end of the input (kept out of line: it runs once per launch)
__device__ __noinline__ uint4 load_ragged(const uint8_t* p, int cnt) {
  u32 w[4] = {0, 0, 0, 0};
  for (int k = 0; k < cnt; ++k) w[k >> 2] |= u32(p[k]) << (8 * (k & 3));
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// OR the low `len` bits of `acc` (len in 1..64, higher bits of acc zero) into the staging bit string at bit
// position `pos`. Worked from the value's last bit: it is shifted left by the free bits r that remain after it in
// its last word, which gives the (up to) three words directly -- three shifts, no 64-bit left-justification.
__device__ __forceinline__ void stage_bits(u32* stage, u32 pos, u64 acc, u32 len) {
#ifdef GH_PROBE_NO_STAGE  // tuning probe (wrong output): one plain store instead of up to three atomics
  stage[pos >> 5] = u32(acc) + len;
  return;
#endif
  const u32 end = pos + len;           // one past the last bit
  const u32 w_first = pos >> 5;
  const u32 w_last = (end - 1) >> 5;
  const u32 r = (0u - end) & 31u;      // free bits after the value inside word w_last
  const u32 lo = u32(acc), hi = u32(acc >> 32);
  atomicOr(stage + w_last, lo << r);
  if (w_last > w_first) atomicOr(stage + w_last - 1, __funnelshift_l(lo, hi, r));
  if (w_last > w_first + 1) atomicOr(stage + w_last - 2, __funnelshift_l(hi, 0u, r));  // hi >> (32 - r), 0 when r == 0
}

// byte k of a 16-byte vector, zero-extended: one PRMT (the scaling to a table offset is then an IMAD on the FMA pipe;
// as shift + mask the extraction costs two instructions on the ALU pipe, which is the pipe this kernel saturates)
__device__ __forceinline__ u32 vec_byte(const uint4& v, int k) {
  const u32 w = (k >> 2) == 0 ? v.x : (k >> 2) == 1 ? v.y : (k >> 2) == 2 ? v.z : v.w;
  return __byte_perm(w, 0u, 0x4440u | u32(k & 3));
}

// Decoupled look-back (Merrill & Garland) by one whole warp: publishes this tile's bit count, adds up the
// predecessors' counts back to the nearest tile that already knows its start, publishes this tile's end bit and
// returns its start bit (all 32 lanes must call it together).
__device__ __forceinline__ u64 tile_start_lookback(const EncWorkspace& ws, u64 tile, u32 tile_bits, u64 start_bit, unsigned lane) {
  u64 exclusive = start_bit;
  if (tile == 0) {
    if (lane == 0) st_volatile_u64(ws.tile_state, kFlagPrefix | (start_bit + tile_bits));
  } else {
    if (lane == 0) st_volatile_u64(ws.tile_state + tile, kFlagAggregate | u64(tile_bits));
    exclusive = 0;
    long long look = (long long)tile - 1;
    bool done = false;
#ifdef GH_PROBE_NO_LOOKBACK  // tuning probe (wrong output): pretend every tile is 6 bits per byte
    exclusive = start_bit + tile * u64(kEncTileBytes) * 6;
    done = true;
#endif
    while (!done) {
      // One round trip covers 32 x kEncLookDepth predecessors: lane l owns the kEncLookDepth consecutive tiles
      // look - l * kEncLookDepth - r (r = 0 nearest), loads all of them at once, folds them locally (sum of
      // aggregates up to and including its nearest PREFIX) and the warp then needs ONE ballot + ONE sum.
      // The kernel's throughput is capped at (tiles covered per round) / (round time): tiles that cannot find
      // a prefix in a round queue up behind those that can, so the window per round is what has to be wide.
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
        while ((st[r] & kFlagMask) == 0) {  // not published yet
          if (kEncPollSleepNs) __nanosleep(kEncPollSleepNs);
          st[r] = ld_volatile_u64(ws.tile_state + idx);
        }
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
  }
  return exclusive;

}

// gather from this lane's copy of the table (see EncSmem)
template <int kSymsPerChunk>
__device__ __forceinline__ u32 lut_word(smem_addr_t lut_lane, u32 byte) {  // variant 4: code << 16 | len; variant 2: len
  typedef EncSmem<kSymsPerChunk> Smem;
  return lds_u32(lut_lane, byte * u32(Smem::kSymStride) + (kSymsPerChunk == 4 ? 0u : 4u));
}
template <int kSymsPerChunk>
__device__ __forceinline__ void lut_entry(smem_addr_t lut_lane, u32 byte, u32& code, u32& len) {
  typedef EncSmem<kSymsPerChunk> Smem;
  if (kSymsPerChunk == 4) {
    const u32 e = lds_u32(lut_lane, byte * u32(Smem::kSymStride));
    code = e >> 16;
    len = e & 0xffffu;
  } else {
    const uint2 e = lds_v2(lut_lane, byte * u32(Smem::kSymStride));
    code = e.x;
    len = e.y;
  }
}

// one 64-bit chunk of the short-code variant from four table entries (code << 16 | len, 1 << len): acc = acc * 2^len
// + code on the FMA pipe (see the kernel); returns the chunk's bit length
__device__ __forceinline__ u32 build_chunk4(const uint2& e0, const uint2& e1, const uint2& e2, const uint2& e3, u32& lo, u32& hi) {
  const u32 a01 = __umulhi(e0.x, 1u << 16) * e1.y + __umulhi(e1.x, 1u << 16);  // <= 32 bits
  const u64 p2 = u64(a01) * e2.y;                                               // <= 48 bits
  const u32 lo2 = u32(p2) + __umulhi(e2.x, 1u << 16), hi2 = u32(p2 >> 32);
  const u64 p3 = u64(lo2) * e3.y;
  lo = u32(p3) + __umulhi(e3.x, 1u << 16);
  hi = hi2 * e3.y + u32(p3 >> 32);
  return (e0.x + e1.x + e2.x + e3.x) & 0xffffu;
}

template <int kSymsPerChunk>
__global__ void __launch_bounds__(kEncThreads, kEncBlocksPerSm)
encode_kernel(const uint8_t* __restrict__ in, u64 n, const EncodeTable table, u64 start_bit, int append_eof,
              u32* __restrict__ out_words, u64 out_word_cap, u64* __restrict__ end_bit_out, EncWorkspace ws) {
  constexpr int kChunks = kEncBytesPerThread / kSymsPerChunk;
  typedef EncSmem<kSymsPerChunk> Smem;
  GH_DYNAMIC_SMEM(smem_raw);
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  u32(&s_stage)[Smem::kBuffers][Smem::kStageWords] = sm.stage;
  u32(&s_warp_total)[kEncSubTiles][kEncThreads / 32] = sm.warp_total;
  u32(&s_carry)[2] = sm.carry;
  u32& s_tile = sm.tile;
  u64& s_tile_start = sm.tile_start;

  const unsigned t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) s_tile = atomicAdd(ws.ticket, 1u);
  for (unsigned i = t; i < 256u * Smem::kCopies; i += kEncThreads) {
    const unsigned sym = i / Smem::kCopies;
    if (kSymsPerChunk == 4) {
      sm.lut[2 * i] = (table.codeword[sym] << 16) | table.length[sym];
      sm.lut[2 * i + 1] = 1u << table.length[sym];
    } else {
      sm.lut[2 * i] = table.codeword[sym];
      sm.lut[2 * i + 1] = table.length[sym];
    }
  }
  for (unsigned i = t; i < unsigned(Smem::kBuffers) * Smem::kStageWords; i += kEncThreads) (&s_stage[0][0])[i] = 0;
  __syncthreads();
  // this lane's copy of the table: entry of byte b at lut_lane + b * kSymStride
  const smem_addr_t lut_lane = smem_addr(sm.lut) + (lane & u32(Smem::kCopies - 1)) * u32(Smem::kEntryBytes);
  u32 buf = 0;        // staging buffer of the current sub-tile (rotates through kBuffers)
  u32 prev_words = 0; // variant 4: words of the previous sub-tile's buffer that still have to be cleared
  const u64 ntiles = enc_num_tiles(n);
  const u32 eof_code = table.codeword[GH_EOF_SYMBOL];
  const u32 eof_len_all = table.length[GH_EOF_SYMBOL];

  for (u64 tile = s_tile; tile < ntiles; tile = s_tile) {
    const bool last_tile = (tile + 1 == ntiles);
    // ---- 1. all loads of the tile up front, bit count per sub-tile ---------------------------------------
    uint4 raw[kEncSubTiles];
    int cnt[kEncSubTiles];
#pragma unroll
    for (int j = 0; j < kEncSubTiles; ++j) {
      const u64 base = tile * kEncTileBytes + u64(j) * kEncSubTileBytes + u64(t) * kEncBytesPerThread;
      raw[j] = make_uint4(0, 0, 0, 0);
      cnt[j] = 0;
      if (base + kEncBytesPerThread <= n) {
        raw[j] = ldg128(reinterpret_cast<const uint4*>(in + base));
        cnt[j] = kEncBytesPerThread;
      } else if (base < n) {
        cnt[j] = int(n - base);
        raw[j] = load_ragged(in + base, cnt[j]);
      }
    }
    u32 bits[kEncSubTiles];
    int end_sub = -1;  // the sub-tile in which this thread owns the last input byte (it carries the end mark)
    constexpr bool kFused = GH_ENC_FUSED && kSymsPerChunk == 4;
    u32 ch_lo[kFused ? kEncSubTiles : 1][4], ch_hi[kFused ? kEncSubTiles : 1][4], ch_len[kFused ? kEncSubTiles : 1];
#pragma unroll
    for (int j = 0; j < kEncSubTiles; ++j) {
      u32 b = 0;
      if (kFused) {
        // missing symbols of the ragged last vector count as (code 0, length 0, 2^0)
        const bool whole = cnt[j] == kEncBytesPerThread;  // all but the one ragged vector at the end of the input
        auto entry = [&](int k) -> uint2 {
          return (whole || k < cnt[j]) ? lds_v2(lut_lane, vec_byte(raw[j], k) * u32(Smem::kSymStride)) : make_uint2(0u, 1u);
        };
        u32 lens = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const u32 clen = build_chunk4(entry(4 * c), entry(4 * c + 1), entry(4 * c + 2), entry(4 * c + 3),
                                        ch_lo[kFused ? j : 0][c], ch_hi[kFused ? j : 0][c]);
          lens |= clen << (8 * c);  // <= 64 each
          b += clen;
        }
        ch_len[kFused ? j : 0] = lens;
      } else
#ifdef GH_PROBE_NO_COUNT  // tuning probe (wrong output): no length gather
      b = 6u * u32(cnt[j]) + (raw[j].x & 1u);
      if (false)
#endif
      // variant 4 sums whole entries: the lengths (<= 16 x 16) add up in the low half, the codes above them
      if (cnt[j] == kEncBytesPerThread) {
#pragma unroll
        for (int k = 0; k < kEncBytesPerThread; ++k) b += lut_word<kSymsPerChunk>(lut_lane, vec_byte(raw[j], k));
      } else {
        for (int k = 0; k < cnt[j]; ++k) b += lut_word<kSymsPerChunk>(lut_lane, vec_byte(raw[j], k));
      }
      if (kSymsPerChunk == 4) b &= 0xffffu;
      const u64 base = tile * kEncTileBytes + u64(j) * kEncSubTileBytes + u64(t) * kEncBytesPerThread;
      if (append_eof && last_tile && base < n && base + kEncBytesPerThread >= n) {
        end_sub = j;
        b += eof_len_all;
      }
      bits[j] = b;
    }

    // ---- 2. per-sub-tile block scans; warp 0 resolves the global offset while the others start packing ----
    u32 incl[kEncSubTiles];
#pragma unroll
    for (int j = 0; j < kEncSubTiles; ++j) {
      incl[j] = warp_inclusive_scan(bits[j], lane);
      if (lane == 31) s_warp_total[j][warp] = incl[j];
    }
    __syncthreads();  // (a)
    u32 sub_bits[kEncSubTiles];  // bits of each sub-tile
    u32 pos0[kEncSubTiles];      // sub-tile-relative bit position of this thread's first bit
    u32 tile_bits = 0;
#pragma unroll
    for (int j = 0; j < kEncSubTiles; ++j) {
      u32 wb = 0, tot = 0;
#pragma unroll
      for (int k = 0; k < kEncThreads / 32; ++k) {
        const u32 wt = s_warp_total[j][k];
        if (unsigned(k) < warp) wb += wt;
        tot += wt;
      }
      sub_bits[j] = tot;
      pos0[j] = wb + incl[j] - bits[j];
      tile_bits += tot;
    }

    if (warp == 0) {
      const u64 exclusive = tile_start_lookback(ws, tile, tile_bits, start_bit, lane);
      if (lane == 0) s_tile_start = exclusive;
    }

    // ---- 3 + 4. sub-tile by sub-tile: pack (tile-relative), then copy out with the phase shift ------------
    u32 before = 0;  // bits of this tile in earlier sub-tiles
#pragma unroll
    for (int j = 0; j < kEncSubTiles; ++j) {
      u32* stage = s_stage[buf];
      {
        u32 pos = pos0[j];
        if (kFused) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const u32 clen = (ch_len[kFused ? j : 0] >> (8 * c)) & 0xffu;
            if (clen) stage_bits(stage, pos, (u64(ch_hi[kFused ? j : 0][c]) << 32) | ch_lo[kFused ? j : 0][c], clen);
            pos += clen;
          }
        } else
#ifdef GH_PROBE_NO_PACK  // tuning probe (wrong output): no code gather, no concatenation, no staging
        if (false)
#endif
        if (cnt[j] == kEncBytesPerThread) {
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
            u64 acc = 0;
            u32 clen = 0;
            if (kSymsPerChunk == 4) {
              // entries are (code << 16 | len, 1 << len) with len <= 16. acc = acc * 2^len + code, all on the FMA
              // pipe: the code is the high half of the entry (IMAD.HI), the products are IMAD / IMAD.WIDE, and adding
              // the code cannot carry because the product's low len bits are zero. Lengths are summed as whole
              // entries (they are the low halves and cannot carry into the codes) and masked once.
              const uint2 e0 = lds_v2(lut_lane, vec_byte(raw[j], c * 4 + 0) * u32(Smem::kSymStride));
              const uint2 e1 = lds_v2(lut_lane, vec_byte(raw[j], c * 4 + 1) * u32(Smem::kSymStride));
              const uint2 e2 = lds_v2(lut_lane, vec_byte(raw[j], c * 4 + 2) * u32(Smem::kSymStride));
              const uint2 e3 = lds_v2(lut_lane, vec_byte(raw[j], c * 4 + 3) * u32(Smem::kSymStride));
              const u32 a01 = __umulhi(e0.x, 1u << 16) * e1.y + __umulhi(e1.x, 1u << 16);  // <= 32 bits
              const u64 p2 = u64(a01) * e2.y;                                               // <= 48 bits
              const u32 lo2 = u32(p2) + __umulhi(e2.x, 1u << 16), hi2 = u32(p2 >> 32);
              const u64 p3 = u64(lo2) * e3.y;
              const u32 lo3 = u32(p3) + __umulhi(e3.x, 1u << 16);
              const u32 hi3 = hi2 * e3.y + u32(p3 >> 32);
              acc = (u64(hi3) << 32) | lo3;
              clen = (e0.x + e1.x + e2.x + e3.x) & 0xffffu;
            } else {
#pragma unroll
              for (int q = 0; q < kSymsPerChunk; ++q) {
                u32 code, len;
                lut_entry<kSymsPerChunk>(lut_lane, vec_byte(raw[j], c * kSymsPerChunk + q), code, len);
                acc = (acc << len) | code;  // len <= 32: at most 64 bits per chunk
                clen += len;
              }
            }
            stage_bits(stage, pos, acc, clen);  // every byte value that occurs has a code: clen >= kSymsPerChunk
            pos += clen;
          }
        } else {  // the ragged last vector of the input: symbol by symbol
          for (int k = 0; k < cnt[j]; ++k) {
            u32 code, len;
            lut_entry<kSymsPerChunk>(lut_lane, vec_byte(raw[j], k), code, len);
            stage_bits(stage, pos, code, len);
            pos += len;
          }
        }
        if (end_sub == j && eof_len_all) stage_bits(stage, pos, eof_code, eof_len_all);
      }
      __syncthreads();  // (c_j) staging of sub-tile j complete; for j == 0 also: s_tile_start written
      if (j == kEncTicketSubTile && t == 0) s_tile = atomicAdd(ws.ticket, 1u);  // next tile (read after a later barrier)

      const u64 G = s_tile_start + before;  // global bit offset of this sub-tile
      const u32 nbits = sub_bits[j];
      const u32 phase = u32(G & 31);
      const u64 end_bit = G + nbits;
      const bool stream_end = last_tile && (before + nbits == tile_bits);  // no bits follow in the whole stream
      const bool more_in_tile = before + nbits < tile_bits;                // a later sub-tile continues this word
      u32 pad = 0;
      if (append_eof && stream_end && nbits) pad = u32((8 - (end_bit & 7)) & 7);  // flush_bits(): 1s to the byte
      if (stream_end && nbits && t == 0 && end_bit_out) *end_bit_out = end_bit;
      const u64 word0 = G >> 5;
      const u32 nwords = nbits ? u32((u64(phase) + nbits + pad + 31) >> 5) : 0u;
      const bool partial_end = ((phase + nbits + pad) & 31) != 0;
      const bool shared_head = (j == 0) && (tile > 0) && (phase != 0);
#ifdef GH_PROBE_NO_COPYLOOP  // tuning probe (wrong output): no copy-out at all
      if (false)
#endif
      {
        // interior words (neither the first nor the last of the sub-tile) need none of the special cases: a lean
        // loop (two LDS, one funnel shift, one byte swap, one store). Measured with the probe builds: the copy-out
        // loop with all its cases inline cost 0.30 ms of the kernel's 1.73, its global stores almost nothing.
        const bool fits = GH_ENC_LEAN_COPY(kSymsPerChunk) && word0 + nwords <= out_word_cap;
        if (fits && nwords > 2u) {
          u32* const dst = out_words + word0;
          for (u32 i = t + 1u; i < nwords - 1u; i += kEncThreads) dst[i] = be32(__funnelshift_r(stage[i], stage[i - 1], phase));
        }
        // first and last word, and everything when the output might not fit
        for (u32 i = t; i < nwords; i += kEncThreads) {
          if (fits && i != 0u && i != nwords - 1u) continue;
          u32 v = __funnelshift_r(stage[i], i ? stage[i - 1] : 0u, phase);
          const u64 gw = word0 + i;
          if (pad && gw == (end_bit >> 5)) v |= ((1u << pad) - 1u) << (32 - (u32(end_bit & 31) + pad));
          if (i == 0 && j > 0 && phase != 0) v |= s_carry[(j - 1) & 1];  // tail of the previous sub-tile
          if (i == nwords - 1 && partial_end && more_in_tile) s_carry[j & 1] = v;  // continued by the next sub-tile
          else if (i == 0 && shared_head) ws.head[tile] = v;
#ifdef GH_PROBE_NO_COPYOUT  // tuning probe (wrong output): only one word in 64 is stored
          else if (gw < out_word_cap && (i & 63u) == 0u) out_words[gw] = be32(v);
#else
          else if (gw < out_word_cap) out_words[gw] = be32(v);
#endif
        }
      }
      if (Smem::kBuffers == 3) {
        // One barrier per sub-tile: the buffer of the PREVIOUS sub-tile is cleared now -- a thread passes barrier
        // (c_j) only after it has finished reading that buffer -- and it is packed into again after barrier
        // (c_j+1), which no thread passes before every thread has finished this clearing.
        u32* prev = s_stage[buf == 0 ? 2 : buf - 1];
        for (u32 i = t; i < prev_words; i += kEncThreads) prev[i] = 0;
        prev_words = (nbits >> 5) + 3;
        buf = buf == 2 ? 0 : buf + 1;
      } else {
        __syncthreads();  // (d_j) staged words consumed
        for (u32 i = t; i < (nbits >> 5) + 3; i += kEncThreads) stage[i] = 0;  // clean for sub-tile j + 2
        buf ^= 1;
      }
      before += nbits;
    }
    // the next tile's number was written after barrier (c_3) when it is drawn during the last sub-tile: one more
    // barrier before it is read (the two-buffer variant has (d_3) for that)
    if (Smem::kBuffers == 3 && kEncTicketSubTile == kEncSubTiles - 1) __syncthreads();
  }
}

// ---- experimental: encoder with warp-independent phases (short codes only, GH_ENCODE_KERNEL=warp) -----------------
// The probe builds (tools/enc_probe.py) show that the phases of encode_kernel add up: its barrier-separated phases
// overlap too little. Here every warp owns a 2 KiB slice of the tile (kEncSubTiles rows of 32 lanes x 16 bytes) and a
// staging area of its own, so that between the block-wide barriers the warps run their phases independently:
//   count + warp scans -> barrier 1 (the eight warp totals) -> warp 0 looks back while every warp packs all its rows
//   at warp-relative positions -> barrier 2 (the tile's start bit) -> every warp copies its own range out with its own
//   phase; the first / last word of a warp's range, when shared with a neighbour, goes to a small edge array ->
//   barrier 3 -> one thread merges the edges (and hands the tile's shared first word to head[], as encode_kernel does).
// Same output, same workspace protocol (tile_state, head[], encode_stitch_kernel). Measured once (r3i, tools/
// enc_warp_check.py): bit-identical to encode_kernel on the B200 and 1.71 ms vs 1.67 ms for 1 GiB Zipf at 4 blocks per
// SM -- no gain as it stands, so it is not the default; kept as the starting point for the next round's experiments.
#ifndef GH_ENC_WARP_BLOCKS
#define GH_ENC_WARP_BLOCKS 4
#endif
struct EncWarpSmem {
  static constexpr int kWarps = kEncThreads / 32;
  static constexpr int kRows = kEncSubTiles;
  static constexpr int kWarpBytes = kRows * 32 * kEncBytesPerThread;  // 2 KiB
  static constexpr int kStageWords = (kWarpBytes * 16 + 32 + 31) / 32 + 2;
  static constexpr int kCopies = 8;
  static constexpr int kSymStride = kCopies * 8;
  u32 lut[256 * kCopies * 2];
  u32 stage[kWarps][kStageWords];
  u32 wtot[kWarps];
  u64 edge_word[kWarps][2];
  u32 edge_val[kWarps][2];
  u32 edge_n[kWarps];
  u32 tile;
  u64 tile_start;
};
static_assert(EncWarpSmem::kWarps * EncWarpSmem::kWarpBytes == kEncTileBytes, "a tile is one slice per warp");

__global__ void __launch_bounds__(kEncThreads, GH_ENC_WARP_BLOCKS)
encode_warp_kernel(const uint8_t* __restrict__ in, u64 n, const EncodeTable table, u64 start_bit, int append_eof,
                   u32* __restrict__ out_words, u64 out_word_cap, u64* __restrict__ end_bit_out, EncWorkspace ws) {
  typedef EncWarpSmem Smem;
  GH_DYNAMIC_SMEM(smem_raw);
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const unsigned t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) sm.tile = atomicAdd(ws.ticket, 1u);
  for (unsigned i = t; i < 256u * Smem::kCopies; i += kEncThreads) {
    const unsigned sym = i / Smem::kCopies;
    sm.lut[2 * i] = (table.codeword[sym] << 16) | table.length[sym];
    sm.lut[2 * i + 1] = 1u << table.length[sym];
  }
  for (unsigned i = t; i < unsigned(Smem::kWarps) * Smem::kStageWords; i += kEncThreads) (&sm.stage[0][0])[i] = 0;
  __syncthreads();
  const smem_addr_t lut_lane = smem_addr(sm.lut) + (lane & u32(Smem::kCopies - 1)) * 8u;
  u32* const stage = sm.stage[warp];
  const u64 ntiles = enc_num_tiles(n);
  const u32 eof_code = table.codeword[GH_EOF_SYMBOL];
  const u32 eof_len = table.length[GH_EOF_SYMBOL];

  for (u64 tile = sm.tile; tile < ntiles; tile = sm.tile) {
    const bool last_tile = (tile + 1 == ntiles);
    // ---- 1. loads and bit counts of this lane's kRows vectors -----------------------------------------------
    uint4 raw[Smem::kRows];
    int cnt[Smem::kRows];
    u32 bits[Smem::kRows];
    int end_row = -1;  // the row in which this lane owns the last input byte (it carries the end mark)
#pragma unroll
    for (int r = 0; r < Smem::kRows; ++r) {
      const u64 base = tile * kEncTileBytes + u64(warp) * Smem::kWarpBytes + u64(r) * (32 * kEncBytesPerThread) +
                       u64(lane) * kEncBytesPerThread;
      raw[r] = make_uint4(0, 0, 0, 0);
      cnt[r] = 0;
      if (base + kEncBytesPerThread <= n) {
        raw[r] = ldg128(reinterpret_cast<const uint4*>(in + base));
        cnt[r] = kEncBytesPerThread;
      } else if (base < n) {
        cnt[r] = int(n - base);
        raw[r] = load_ragged(in + base, cnt[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < Smem::kRows; ++r) {
      u32 b = 0;
      if (cnt[r] == kEncBytesPerThread) {
#pragma unroll
        for (int k = 0; k < kEncBytesPerThread; ++k) b += lds_u32(lut_lane, vec_byte(raw[r], k) * u32(Smem::kSymStride));
      } else {
        for (int k = 0; k < cnt[r]; ++k) b += lds_u32(lut_lane, vec_byte(raw[r], k) * u32(Smem::kSymStride));
      }
      b &= 0xffffu;  // whole entries were summed: the lengths are their low halves
      const u64 base = tile * kEncTileBytes + u64(warp) * Smem::kWarpBytes + u64(r) * (32 * kEncBytesPerThread) +
                       u64(lane) * kEncBytesPerThread;
      if (append_eof && last_tile && base < n && base + kEncBytesPerThread >= n) {
        end_row = r;
        b += eof_len;
      }
      bits[r] = b;
    }
    // ---- 2. warp-relative positions; the warp totals are all the block has to exchange -----------------------
    u32 pos[Smem::kRows];
    u32 wbits = 0;
#pragma unroll
    for (int r = 0; r < Smem::kRows; ++r) {
      const u32 incl = warp_inclusive_scan(bits[r], lane);
      pos[r] = wbits + incl - bits[r];
      wbits += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) sm.wtot[warp] = wbits;
    __syncthreads();  // (1)
    u32 woff = 0, tile_bits = 0;
#pragma unroll
    for (int k = 0; k < Smem::kWarps; ++k) {
      const u32 v = sm.wtot[k];
      if (unsigned(k) < warp) woff += v;
      tile_bits += v;
    }
    if (warp == 0) {
      const u64 exclusive = tile_start_lookback(ws, tile, tile_bits, start_bit, lane);
      if (lane == 0) sm.tile_start = exclusive;
    }
    // ---- 3. pack all rows into the warp's own staging area ---------------------------------------------------
#pragma unroll
    for (int r = 0; r < Smem::kRows; ++r) {
      u32 p = pos[r];
      if (cnt[r] == kEncBytesPerThread) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          u32 lo, hi;
          const u32 clen = build_chunk4(lds_v2(lut_lane, vec_byte(raw[r], 4 * c + 0) * u32(Smem::kSymStride)),
                                        lds_v2(lut_lane, vec_byte(raw[r], 4 * c + 1) * u32(Smem::kSymStride)),
                                        lds_v2(lut_lane, vec_byte(raw[r], 4 * c + 2) * u32(Smem::kSymStride)),
                                        lds_v2(lut_lane, vec_byte(raw[r], 4 * c + 3) * u32(Smem::kSymStride)), lo, hi);
          stage_bits(stage, p, (u64(hi) << 32) | lo, clen);
          p += clen;
        }
      } else {  // the ragged last vector of the input: symbol by symbol
        for (int k = 0; k < cnt[r]; ++k) {
          const u32 e = lds_u32(lut_lane, vec_byte(raw[r], k) * u32(Smem::kSymStride));
          stage_bits(stage, p, e >> 16, e & 0xffffu);
          p += e & 0xffffu;
        }
      }
      if (end_row == r && eof_len) stage_bits(stage, p, eof_code, eof_len);
    }
    __syncthreads();  // (2) staging complete, tile start known
    if (t == 0) sm.tile = atomicAdd(ws.ticket, 1u);  // next tile (read after barrier 3)
    // ---- 4. copy-out of this warp's range --------------------------------------------------------------------
    const u64 tile_G = sm.tile_start;
    const u64 G = tile_G + woff;
    const u32 nbits = wbits;
    const u32 phase = u32(G & 31);
    const u64 end_bit = G + nbits;
    const bool stream_end = last_tile && nbits && (woff + nbits == tile_bits);  // no bits follow in the whole stream
    u32 pad = 0;
    if (append_eof && stream_end) pad = u32((8 - (end_bit & 7)) & 7);  // flush_bits(): 1s to the byte
    if (stream_end && lane == 0 && end_bit_out) *end_bit_out = end_bit;
    const u32 padmask = pad ? ((1u << pad) - 1u) << (32 - (u32(end_bit & 31) + pad)) : 0u;
    const u64 word0 = G >> 5;
    const u32 nwords = nbits ? u32((u64(phase) + nbits + pad + 31) >> 5) : 0u;
    const bool head_partial = nwords && phase != 0;
    const bool tail_partial = nwords && ((phase + nbits + pad) & 31) != 0 && !(nwords == 1 && head_partial);
    const u32 lo_i = head_partial ? 1u : 0u, hi_i = nwords - (tail_partial ? 1u : 0u);
    auto word_at = [&](u32 i) -> u32 {
      u32 v = __funnelshift_r(stage[i], i ? stage[i - 1] : 0u, phase);
      if (i == nwords - 1) v |= padmask;
      return v;
    };
    for (u32 i = lo_i + lane; i < hi_i; i += 32) {
      if (word0 + i < out_word_cap) out_words[word0 + i] = be32(word_at(i));
    }
    if (lane == 0) {  // words this warp shares with its neighbours
      u32 ne = 0;
      if (head_partial) sm.edge_word[warp][ne] = word0, sm.edge_val[warp][ne] = word_at(0), ++ne;
      if (tail_partial) sm.edge_word[warp][ne] = word0 + nwords - 1, sm.edge_val[warp][ne] = word_at(nwords - 1), ++ne;
      sm.edge_n[warp] = ne;
    }
    __syncwarp();
    for (u32 i = lane; i < (nbits >> 5) + 3; i += 32) stage[i] = 0;  // clean for the next tile
    __syncthreads();  // (3) edges written
    if (t == 0) {
      // contributions in stream order; equal word indices are merged; the tile's first word, when shared with the
      // previous tile, goes to head[] (encode_stitch_kernel ORs it into what that tile stored)
      const bool shared_head = tile > 0 && (tile_G & 31) != 0;
      const u64 tile_word0 = tile_G >> 5;
      u64 cur_word = ~0ull;
      u32 cur_val = 0;
      for (int w = 0; w <= Smem::kWarps; ++w) {
        const u32 ne = w < Smem::kWarps ? sm.edge_n[w] : 1u;
        for (u32 k = 0; k < ne; ++k) {
          const u64 wd = w < Smem::kWarps ? sm.edge_word[w][k] : ~0ull;  // the extra round flushes the last word
          const u32 v = w < Smem::kWarps ? sm.edge_val[w][k] : 0u;
          if (wd != cur_word) {
            if (cur_word != ~0ull) {
              if (shared_head && cur_word == tile_word0) ws.head[tile] = cur_val;
              else if (cur_word < out_word_cap) out_words[cur_word] = be32(cur_val);
            }
            cur_word = wd;
            cur_val = v;
          } else {
            cur_val |= v;
          }
        }
      }
    }
  }
}

// Second kernel: OR each tile's deferred head bits into the word its predecessor stored.
// Thread = tile. The first tile (in order) holding head bits for a given word merges the whole run.
__global__ void __launch_bounds__(256)
encode_stitch_kernel(u64 ntiles, u64 start_bit, u32* __restrict__ out_words, u64 out_word_cap, EncWorkspace ws) {
  const u64 tile = u64(blockIdx.x) * blockDim.x + threadIdx.x;
  if (tile == 0 || tile >= ntiles) return;
  const u64 g = ws.tile_state[tile - 1] & ~kFlagMask;  // global start bit of `tile`
  if ((g & 31) == 0) return;                            // starts on a word boundary: nothing deferred
  const u64 word = g >> 5;
  if (tile >= 2) {
    const u64 gp = ws.tile_state[tile - 2] & ~kFlagMask;  // start of the previous tile
    if ((gp >> 5) == word && (gp & 31) != 0) return;       // the previous tile deferred into this word too: it leads
  }
  (void)start_bit;
  u32 bits = ws.head[tile];
  for (u64 nx = tile + 1; nx < ntiles; ++nx) {  // only ever iterates for degenerate sub-word tiles
    const u64 gn = ws.tile_state[nx - 1] & ~kFlagMask;
    if ((gn >> 5) != word) break;
    bits |= ws.head[nx];
  }
  if (word < out_word_cap) out_words[word] |= be32(bits);
}

inline size_t enc_ws_bytes(u64 n) {
  const u64 nt = enc_num_tiles(n) + 1;
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
  const u64 ntiles = enc_num_tiles(n);
  if (ntiles > 0x7fffffffull) return GH_ERR_ARG;

  EncWorkspace ws;
  uint8_t* p = static_cast<uint8_t*>(d_workspace);
  ws.tile_state = reinterpret_cast<u64*>(p);
  p += (ntiles + 1) * 8;
  ws.head = reinterpret_cast<u32*>(p);
  p += (((ntiles + 1) * 4 + 7) / 8) * 8;
  ws.ticket = reinterpret_cast<u32*>(p);
  const size_t used = size_t(p - static_cast<uint8_t*>(d_workspace)) + 8;
  GH_CUDA_TRY(cudaMemsetAsync(d_workspace, 0, used, (cudaStream_t)stream));

  static bool attrs_set = false;
  if (!attrs_set) {  // both variants exceed the 48 KB a kernel gets without opting in
    GH_CUDA_TRY(cudaFuncSetAttribute(encode_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(EncSmem<4>))));
    GH_CUDA_TRY(cudaFuncSetAttribute(encode_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(EncSmem<2>))));
    attrs_set = true;
  }
  const u64 out_word_cap = payload_cap / 4;
  // persistent blocks: as many as are resident at once (shared memory allows 4 of variant 4, 3 of variant 2)
  const bool short_codes = code->max_len <= 16;
  int per_sm = 0;
#ifndef GH_EMUL
  if (short_codes) GH_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, encode_kernel<4>, kEncThreads, sizeof(EncSmem<4>)));
  else GH_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, encode_kernel<2>, kEncThreads, sizeof(EncSmem<2>)));
#endif
  if (per_sm < 1) per_sm = kEncBlocksPerSm;
  u64 blocks = u64(sm_count() > 0 ? sm_count() : 1) * u64(per_sm);
  if (blocks > ntiles) blocks = ntiles;
  const char* which = getenv("GH_ENCODE_KERNEL");
  if (short_codes && which && which[0] == 'w') {  // experimental warp-independent kernel, see encode_warp_kernel
    static bool warp_attr = false;
    if (!warp_attr) {
      GH_CUDA_TRY(cudaFuncSetAttribute(encode_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(EncWarpSmem))));
      warp_attr = true;
    }
    int wp = 0;
#ifndef GH_EMUL
    GH_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&wp, encode_warp_kernel, kEncThreads, sizeof(EncWarpSmem)));
#endif
    if (wp < 1) wp = GH_ENC_WARP_BLOCKS;
    u64 wblocks = u64(sm_count() > 0 ? sm_count() : 1) * u64(wp);
    if (wblocks > ntiles) wblocks = ntiles;
    GH_LAUNCH(encode_warp_kernel, unsigned(wblocks), kEncThreads, sizeof(EncWarpSmem), stream, d_in, (u64)n, table, (u64)start_bit,
              append_eof, reinterpret_cast<u32*>(d_payload), out_word_cap, reinterpret_cast<u64*>(d_end_bit), ws);
  } else if (short_codes) {
    GH_LAUNCH(encode_kernel<4>, unsigned(blocks), kEncThreads, sizeof(EncSmem<4>), stream, d_in, (u64)n, table, (u64)start_bit, append_eof,
              reinterpret_cast<u32*>(d_payload), out_word_cap, reinterpret_cast<u64*>(d_end_bit), ws);
  } else {
    GH_LAUNCH(encode_kernel<2>, unsigned(blocks), kEncThreads, sizeof(EncSmem<2>), stream, d_in, (u64)n, table, (u64)start_bit, append_eof,
              reinterpret_cast<u32*>(d_payload), out_word_cap, reinterpret_cast<u64*>(d_end_bit), ws);
  }
  rc = check_launch();
  if (rc != GH_OK) return rc;
  if (ntiles > 1) {
    GH_LAUNCH(encode_stitch_kernel, unsigned((ntiles + 255) / 256), 256, 0, stream, ntiles, (u64)start_bit,
              reinterpret_cast<u32*>(d_payload), out_word_cap, ws);
    rc = check_launch();
  }
  return rc;
}

}  // namespace gh
